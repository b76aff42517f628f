"""Margin of the TD(lambda) replay against the reference's apply_td_updates on the golden games:
max |dw - dw_ref| / tolerance per game and tensor (<= 1 passes test_td_replay_host_golden).  Run on a GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from bgx import lib as L
if os.environ.get('BGX_LIB'): L.load(os.environ['BGX_LIB'])
from bgx.engine import BatchEngine
from conftest import golden_weights
from test_gpu_parity import records_from, td_tol
g = np.load(os.path.join(ROOT, "tests/golden/games.npz"))
gm = np.load(os.path.join(ROOT, "tests/golden/model.npz"))
eng = BatchEngine(0)
worst = 0.0
for name in g["names"]:
    name = str(name)
    w0 = golden_weights(gm, "rand" if name.startswith("rand") else "trained")
    eng.set_weights(*w0)
    rec = records_from(g[f"{name}.pre"], g[f"{name}.player"])
    new, sq = eng.td_replay_host(rec, int(g[f"{name}.winner"]) == 0, float(g[f"{name}.lr"]), float(g[f"{name}.lam"]))
    out = []
    for a, b, k in zip(new, w0, ("W1", "b1", "w2", "b2")):
        ref_new = g[f"{name}.new_{k}"].reshape(-1)
        dref = ref_new - np.asarray(b).reshape(-1)
        dgot = np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)
        r = float(np.max(np.abs(dgot - dref)) / td_tol(dref, ref_new))
        worst = max(worst, r)
        out.append(f"{k} {r:.2f}")
    print(name, len(rec), " ".join(out), f"loss {np.max(np.abs(np.sqrt(sq) - np.sqrt(g[f'{name}.losses']))):.1e}", flush=True)
print(f"worst err/tol {worst:.2f}", flush=True)
