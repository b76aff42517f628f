"""Where the end-to-end ply-step goes (run on a GPU box): kernel time of a half population, host numpy time,
and the pipelined step time for 2 / 4 lanes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bench import init_weights, GAMES_PER_GPU as G
from bgx.engine import BatchEngine
from bgx.synth import START_BOARD
eng = BatchEngine(0); eng.set_weights(*init_weights())
eng.selfplay_init(G, first_mover=1)
for _ in range(8): eng.selfplay_step(16, want_stats=False)
rec, _, _ = eng.selfplay_read(); rec[:, 31] = 0
rng = np.random.default_rng(0); rec[:, 29:31] = rng.integers(1, 7, (G, 2))
for n in (() if os.environ.get('QUICK') else (G, G // 2, G // 4)):
    q = torch.from_numpy(rec[:n]).cuda(); chosen = torch.zeros((n, 32), dtype=torch.int8, device="cuda")
    ms = []
    for _ in range(10):
        eng.select_moves(q, chosen=chosen); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
    print(f"k_select({n}): {np.mean(ms[2:]):.3f} ms", flush=True)
from bgx import host as bgx_host
q = torch.zeros((G, 32), dtype=torch.int8).pin_memory().numpy(); q[:] = rec
ch = torch.zeros((G, 32), dtype=torch.int8).pin_memory().numpy()
_, h_ply, h_gid = eng.selfplay_read()
h_win = np.zeros(G, np.int8)
for lanes in [int(x) for x in os.environ.get("LANES", "2,3,4").split(",")]:
    parts = [(i * G // lanes, (i + 1) * G // lanes) for i in range(lanes)]
    def submit(h):
        lo, hi = parts[h]
        eng.select_moves_host_async(h, q[lo:hi], {"chosen": ch[lo:hi]})
    def advance(h):
        lo, hi = parts[h]
        h_ply[lo:hi] += 1
        bgx_host.advance(ch[lo:hi], q[lo:hi], 1, h_ply[lo:hi], h_gid[lo:hi], h_win[lo:hi])
        done = np.flatnonzero(h_win[lo:hi] >= 0)
        if done.size:
            idx = done + lo
            h_gid[idx] += G; h_ply[idx] = 0
            fresh = np.zeros((idx.size, 32), np.int8); fresh[:, :24] = START_BOARD; fresh[:, 28] = (h_gid[idx] & 1) ^ 1
            q[idx] = bgx_host.advance(fresh, fresh, 1, h_ply[idx], h_gid[idx])
    for h in range(lanes): submit(h)
    for it in range(-3, 60):
        if it == 0: t0 = time.perf_counter(); th = 0.0; tw = 0.0
        for h in range(lanes):
            a0 = time.perf_counter(); eng.wait(h); a = time.perf_counter(); advance(h); submit(h); b = time.perf_counter()
            if it >= 0: th += b - a; tw += a - a0
    for h in range(lanes): eng.wait(h)
    dt = time.perf_counter() - t0
    print(f"{lanes} lanes: {dt/60*1e3:.3f} ms per ply-step ({G*60/dt/1e6:.1f} M plies/s), host advance+submit {th/60*1e3:.3f} ms, waiting {tw/60*1e3:.3f} ms per step", flush=True)
