"""Relative error of the summed TD(lambda) round delta against the per-game oracle replays (run on a GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from bgx import lib as L
if os.environ.get('BGX_LIB'): L.load(os.environ['BGX_LIB'])
from bgx.engine import BatchEngine
from oracle.oracle import Oracle
gm = np.load(os.path.join(ROOT, "tests/golden/model.npz"))
orc = Oracle(); eng = BatchEngine(0)
for tag in ("rand", "trained"):
    w0 = tuple(gm[f"{tag}_{k}"] for k in ("W1", "b1", "w2", "b2"))
    for n, first in ((12, 77), (48, 1000)):
        eng.set_weights(*w0)
        eng.selfplay_init(n, first_id=first, id_stride=n, seed=0x5EED2026, traj_cap=2048)
        eng.selfplay_round()
        rec, ply, gid = eng.selfplay_read()
        delta = torch.zeros(25604, device="cuda", dtype=torch.float32)
        eng.td_replay(0.1, 0.9, delta); torch.cuda.synchronize()
        got = delta.cpu().numpy()
        flat0 = np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in w0])
        want = np.zeros(25601)
        for slot in range(n):
            pre, _ = eng.export_trajectory(slot)
            X = np.concatenate([orc.encode(pre[t:t + 1, :28].astype(np.int32), int(pre[t, 28])) for t in range(len(pre))])
            new, sq = orc.td_replay(w0, X, rec[slot, 31] == 1, 0.1, 0.9)
            want += np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in new]).astype(np.float64) - flat0
        out = []
        for lo, hi, k in ((0, 25344, "W1"), (25344, 25472, "b1"), (25472, 25600, "w2"), (25600, 25601, "b2")):
            out.append(f"{k} {np.max(np.abs(got[lo:hi] - want[lo:hi])) / np.max(np.abs(want[lo:hi])):.2e}")
        print(tag, n, int(ply.sum()), " ".join(out), flush=True)
