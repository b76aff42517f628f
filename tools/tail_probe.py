"""Launch-granularity probe: one ply per launch vs 16 plies per launch (run on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bench import init_weights, GAMES_PER_GPU as G
from bgx.engine import BatchEngine
eng = BatchEngine(0); eng.set_weights(*init_weights())
eng.selfplay_init(G, first_mover=1)
for _ in range(8): eng.selfplay_step(16, want_stats=False)
for n in (1, 2, 4, 16):
    ms = []
    for _ in range(10):
        eng.selfplay_step(n, want_stats=True); ms.append(eng.last_kernel_ms())
    print(f"selfplay_step({n}): {np.mean(ms):.3f} ms per launch, {np.mean(ms)/n:.3f} ms per ply-step of {G}")
rec, _, _ = eng.selfplay_read(); rec[:, 31] = 0
rng = np.random.default_rng(0); rec[:, 29:31] = rng.integers(1, 7, (G, 2))
q = torch.from_numpy(rec).cuda(); chosen = torch.zeros((G, 32), dtype=torch.int8, device="cuda"); val = torch.zeros(G, device="cuda")
nseq = torch.zeros(G, dtype=torch.int32, device="cuda")
ms = []
for _ in range(10):
    eng.select_moves(q, chosen=chosen, value=val, n_seq=nseq); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
print(f"k_select on {G} device-resident queries: {np.mean(ms):.3f} ms; max n_seq {int(nseq.max())}, mean {float(nseq.float().mean()):.1f}")
order = torch.argsort(nseq, descending=True)
qs = q[order].contiguous()
ms = []
for _ in range(10):
    eng.select_moves(qs, chosen=chosen, value=val, n_seq=nseq); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
print(f"k_select, queries sorted by n_seq descending: {np.mean(ms):.3f} ms")
