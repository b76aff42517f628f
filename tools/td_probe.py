"""One TD(lambda) replay launch over a played round: ms and TD steps/s (run on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bench import init_weights
from bgx import lib as L
if os.environ.get('BGX_LIB'): L.load(os.environ['BGX_LIB'])
from bgx.engine import BatchEngine
G = int(os.environ.get("TD_GAMES", "65536"))
eng = BatchEngine(0); eng.set_weights(*init_weights())
eng.selfplay_init(G, seed=7, traj_cap=2048)
eng.selfplay_round(0.0)
delta = torch.zeros(25604, dtype=torch.float32, device="cuda")
ms = []
for _ in range(3):
    st = eng.td_replay(0.1, 0.9, delta); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
print(f"k_td_replay: {min(ms):.1f} ms, {st['td_steps'] / min(ms) / 1e3:.1f} M TD steps/s, |delta| sum {float(delta.abs().sum()):.9g}", flush=True)
