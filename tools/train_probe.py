"""Does the GPU round trainer learn?  Random-init net, R rounds of G games, win rate against the random policy.

    python tools/train_probe.py G SCALE ROUNDS      # weights += SCALE * mean of the round's per-game TD(lambda) weight changes

Measured on a B200: `1 1 16000` (the reference's sequential semantics) 99.7 % after 10,000 games / 10 s;
`256 2 10000` 99.8 % after 2,500 rounds / 6 s; SCALE = G (the plain sum of stale per-game changes) saturates the net."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bgx.model import TDLGammonModel
from bgx.train import GpuTrainer
from bgx.evaluate import Arena, play_vs_random
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 64.0
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 12
torch.manual_seed(0)
m = TDLGammonModel()
arena = Arena(0)
print("before:", play_vs_random(arena, m.weights_np(), 2048), flush=True)
tr = GpuTrainer(m, G, delta_scale=scale / G)
t0 = time.time()
for r in range(rounds):
    st = tr.round(epsilon=0.0)
    if r % max(1, rounds // 8) == max(1, rounds // 8) - 1 or r == rounds - 1:
        tr.sync_model()
        print(r + 1, "rounds", f"{time.time()-t0:.1f}s", "plies", st["plies"], "p1 wins", st["p1_wins"] / G,
              "vs random:", play_vs_random(arena, m.weights_np(), 2048), flush=True)
