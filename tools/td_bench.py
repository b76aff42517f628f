"""Time k_td_replay on one self-play round (run on the B200 box): python tools/td_bench.py [n_games] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "backgammon-engine_b200"))
from bench import init_weights  # noqa: E402
from bgx.engine import BatchEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tag = sys.argv[3] if len(sys.argv) > 3 else "rand"
eng = BatchEngine(0)
if tag == "rand":
    eng.set_weights(*init_weights())
else:
    with np.load(os.path.join(ROOT, "tests", "golden", "model.npz")) as z:
        eng.set_weights(*(z[f"trained_{k}"] for k in ("W1", "b1", "w2", "b2")))
eng.selfplay_init(n, seed=0x5EED2027, traj_cap=2048)
t0 = time.time()
st = eng.selfplay_round()
print("play", st["plies"], "plies", f"{time.time() - t0:.3f} s")
delta = torch.zeros(25604, device="cuda")
best = None
for _ in range(reps):
    td = eng.td_replay(0.1, 0.9, delta)
    ms = eng.last_kernel_ms()
    best = ms if best is None else min(best, ms)
print(f"td_replay {os.environ.get('BGX_TD_DENSE') and 'dense' or 'sparse'}: {best:.2f} ms, {td['td_steps'] / best / 1e3:.1f} M steps/s, "
      f"lazy row-steps per step {td['td_lazy_row_steps'] / max(td['td_steps'], 1):.1f}, games {td['games_finished']}, sq {td['td_sq_error']:.6g}, "
      f"|delta| {float(delta.abs().sum()):.6g}")
if os.environ.get("BGX_TD_PROFILE"):
    eng.td_profile(True)
    eng.td_replay(0.1, 0.9, delta)
    pc = eng.td_profile(False).astype(float)
    names = ["z store", "(bar A)", "hidden", "window", "(bar B)", "grad", "row pass", "end of game", "values", "c", "ring", "release-after-last", "wait-for-last"]
    print("cycles per step on CTA 0, per warp: " + " | ".join(names))
    for w in range(8):
        n = max(pc[w, 15], 1.0)
        print(f"  warp {w}: " + " ".join(f"{pc[w, i] / n:7.0f}" for i in range(13)) + f"   sum {pc[w, :11].sum() / n:.0f}")
