"""Where the end-to-end ply goes: kernel vs copies vs host loop (run on a GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bench import init_weights, GAMES_PER_GPU as G
from bgx.engine import BatchEngine
from bgx.synth import START_BOARD
eng = BatchEngine(0); eng.set_weights(*init_weights())
q = torch.zeros((G, 32), dtype=torch.int8).pin_memory().numpy()
out = {"chosen": torch.zeros((G, 32), dtype=torch.int8).pin_memory().numpy(), "value": torch.zeros(G).pin_memory().numpy(),
       "moves": None, "moves_len": None, "n_seq": None, "n_scored": None}
eng.selfplay_init(G, first_mover=1)
for _ in range(8): eng.selfplay_step(16, want_stats=False)
rec, _, _ = eng.selfplay_read(); q[:] = rec; q[:, 31] = 0
rng = np.random.default_rng(1)
tk = tc = th = 0.0; n = 0
for it in range(-3, 120):
    t0 = time.perf_counter()
    q[:, 29:31] = rng.integers(1, 7, (G, 2), dtype=np.int8)
    t1 = time.perf_counter()
    o = eng.select_moves_host(q, out=dict(out))
    t2 = time.perf_counter()
    ch = o["chosen"]; over = (ch[:, 26] == 15) | (ch[:, 27] == 15)
    q[:, :28] = ch[:, :28]; q[:, 28] ^= 1
    if over.any(): q[over, :24] = START_BOARD; q[over, 24:28] = 0
    t3 = time.perf_counter()
    if it >= 0:
        tk += eng.last_kernel_ms(); tc += (t2 - t1) * 1e3; th += (t1 - t0 + t3 - t2) * 1e3; n += 1
print(f"per ply-step of {G} games: call {tc/n:.3f} ms (kernel {tk/n:.3f} ms), host numpy {th/n:.3f} ms -> {G/((tc+th)/n)*1e3/1e6:.2f} M plies/s")
