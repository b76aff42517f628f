"""Time k_selfplay on the bench's headline workload for several kernel configurations (run on the B200 box):

    python tools/ply_probe.py [warps ...]      e.g. python tools/ply_probe.py 16 20 24 32

65,536 games, 16 plies per launch, greedy, random-init weights: the same launch bench.py times; prints M plies/s per setting."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "backgammon-engine_b200"))
from bench import GAMES_PER_GPU, PLIES_PER_STEP, SEED, init_weights  # noqa: E402
from bgx.lib import FIRST_PARITY  # noqa: E402
from bgx.engine import BatchEngine  # noqa: E402

warps = [int(a) for a in sys.argv[1:]] or [24]
for w in warps:
    eng = BatchEngine(0)
    eng.set_weights(*init_weights())
    eng.set_option("selfplay_warps", w)
    eng.selfplay_init(GAMES_PER_GPU, first_id=0, id_stride=GAMES_PER_GPU, seed=SEED, first_mover=FIRST_PARITY, traj_cap=0)
    for _ in range(3):
        eng.selfplay_step(PLIES_PER_STEP, want_stats=False)
    ms, plies, edges = [], 0, 0
    for _ in range(10):
        st = eng.selfplay_step(PLIES_PER_STEP)
        ms.append(eng.last_kernel_ms())
        plies += st["plies"]
        edges += st["tree_edges"]
    print(f"selfplay_warps {w}: {plies / sum(ms) / 1e3:.1f} M plies/s (kernel {np.mean(ms):.3f} ms per {PLIES_PER_STEP} plies), {edges / plies:.2f} tree edges per ply", flush=True)
    del eng
