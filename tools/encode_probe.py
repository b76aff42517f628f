"""k_encode alone on 4 M rows (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bgx.engine import BatchEngine
from bgx.synth import make_queries
eng = BatchEngine(0)
q, _ = make_queries(250000, seed=3)
qd = torch.from_numpy(q).cuda().repeat(16, 1)
X = torch.empty((qd.shape[0], 198), dtype=torch.float32, device="cuda")
ms = []
for _ in range(6):
    eng.encode(qd, X); torch.cuda.synchronize(); ms.append(eng.last_kernel_ms())
print("k_encode", qd.shape[0], "rows:", min(ms), "ms", qd.shape[0] * 824 / min(ms) / 1e6, "GB/s")
