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
pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory().numpy()
FUSED = os.environ.get("FUSED", "1") == "1"
q = pin((G, 32), torch.int8); q[:] = rec
ch = pin((G, 32), torch.int8)
bufs = [q, ch]
win = pin((G,), torch.int8); nxt_ply = pin((G,), torch.int32); h_gid = pin((G,), torch.int64)
_, h_ply, gid0 = eng.selfplay_read(); h_gid[:] = gid0
for lanes in [int(x) for x in os.environ.get("LANES", "2,3,4").split(",")]:
    parts = [(i * G // lanes, (i + 1) * G // lanes) for i in range(lanes)]
    cur = [0] * lanes
    bufs[0][:] = rec
    def submit(h):
        lo, hi = parts[h]
        if FUSED:
            nxt_ply[lo:hi] = h_ply[lo:hi] + 1
            eng.play_ply_host_async(h, bufs[cur[h]][lo:hi], nxt_ply[lo:hi], h_gid[lo:hi], bufs[1 - cur[h]][lo:hi], win[lo:hi], dice_seed=1)
        else:
            eng.select_moves_host_async(h, bufs[0][lo:hi], {"chosen": bufs[1][lo:hi]})
    def advance(h):
        lo, hi = parts[h]
        h_ply[lo:hi] += 1
        if FUSED:
            cur[h] ^= 1
        else:
            bgx_host.advance(bufs[1][lo:hi], bufs[0][lo:hi], 1, h_ply[lo:hi], h_gid[lo:hi], win[lo:hi])
        out = bufs[cur[h]] if FUSED else bufs[0]
        done = np.flatnonzero(win[lo:hi] >= 0)
        if done.size:
            idx = done + lo
            h_gid[idx] += G; h_ply[idx] = 0
            fresh = np.zeros((idx.size, 32), np.int8); fresh[:, :24] = START_BOARD; fresh[:, 28] = (h_gid[idx] & 1) ^ 1
            out[idx] = bgx_host.advance(fresh, fresh, 1, h_ply[idx], h_gid[idx])
    for h in range(lanes): submit(h)
    for it in range(-3, 60):
        if it == 0: t0 = time.perf_counter(); th = 0.0; tw = 0.0
        for h in range(lanes):
            a0 = time.perf_counter(); eng.wait(h); a = time.perf_counter(); advance(h); submit(h); b = time.perf_counter()
            if it >= 0: th += b - a; tw += a - a0
    for h in range(lanes): eng.wait(h)
    dt = time.perf_counter() - t0
    print(f"{'fused' if FUSED else 'split'} {lanes} lanes: {dt/60*1e3:.3f} ms per ply-step ({G*60/dt/1e6:.1f} M plies/s), host advance+submit {th/60*1e3:.3f} ms, waiting {tw/60*1e3:.3f} ms per step", flush=True)
