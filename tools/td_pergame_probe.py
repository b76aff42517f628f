"""Per-game TD(lambda) parity at scale: N self-play trajectories replayed one by one on the GPU
(bgx_td_replay_host) and by the oracle from the same snapshot; distribution of max|dw - dw_ref| / max|dw_ref| per tensor.
Run on a GPU box:  TD_N=1024 python tools/td_pergame_probe.py"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from bgx import lib as L
if os.environ.get('BGX_LIB'): L.load(os.environ['BGX_LIB'])
from bgx.engine import BatchEngine
from oracle.oracle import Oracle, td_replay_f64 as replay_f64
N = int(os.environ.get("TD_N", "1024"))

gm = np.load(os.path.join(ROOT, "tests/golden/model.npz"))
orc = Oracle(); eng = BatchEngine(0)
for tag in ("rand", "trained"):
    w0 = tuple(gm[f"{tag}_{k}"] for k in ("W1", "b1", "w2", "b2"))
    eng.set_weights(*w0)
    eng.selfplay_init(N, first_id=5000, id_stride=N, seed=0x5EED2026, traj_cap=2048)
    eng.selfplay_round()
    rec, ply, gid = eng.selfplay_read()
    trajs = [eng.export_trajectory(s)[0].copy() for s in range(N)]
    won = [bool(rec[s, 31] == 1) for s in range(N)]

    def ref(s):
        pre = trajs[s]
        X = np.concatenate([orc.encode(pre[t:t + 1, :28].astype(np.int32), int(pre[t, 28])) for t in range(len(pre))])
        return orc.td_replay(w0, X, won[s], 0.1, 0.9) + (replay_f64(w0, X, won[s], 0.1, 0.9),)
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        refs = list(ex.map(ref, range(N)))
    err = np.zeros((N, 4)); tolr = np.zeros((N, 4)); sqe = np.zeros(N); e_gpu = np.zeros((N, 4)); e_orc = np.zeros((N, 4))
    for s in range(N):
        new, sq = eng.td_replay_host(trajs[s], won[s], 0.1, 0.9)
        sqe[s] = np.max(np.abs(np.sqrt(sq) - np.sqrt(refs[s][1]))) if len(sq) else 0.0
        for k in range(4):
            b = np.asarray(w0[k], np.float32).reshape(-1)
            dref = np.asarray(refs[s][0][k]).reshape(-1).astype(np.float64) - b
            dgot = np.asarray(new[k]).reshape(-1).astype(np.float64) - b
            m = np.max(np.abs(dref))
            d64 = np.asarray(refs[s][2][k]).reshape(-1) - b.astype(np.float64)
            e_gpu[s, k] = np.max(np.abs(dgot - d64)) / np.max(np.abs(d64))
            e_orc[s, k] = np.max(np.abs(dref - d64)) / np.max(np.abs(d64))
            err[s, k] = np.max(np.abs(dgot - dref)) / m
            tolr[s, k] = np.max(np.abs(dgot - dref)) / (1e-5 * m + np.spacing(np.float32(np.max(np.abs(b)))))
    q = lambda a: " ".join(f"{np.quantile(a, p):.2e}" for p in (0.5, 0.9, 0.99, 1.0))
    print(tag, "games", N, "steps", int(ply.sum()), "| rel err p50 p90 p99 max:", flush=True)
    for k, name in enumerate(("W1", "b1", "w2", "b2")):
        print(f"  {name}: {q(err[:, k])}   err/tol: {q(tolr[:, k])}   within tol: {np.mean(tolr[:, k] <= 1):.3f}", flush=True)
    print(f"  |sqrt(sq) - ref| max {sqe.max():.2e}", flush=True)
    print("  against the float64 replay (p50 p90 p99 max), GPU | oracle fp32:", flush=True)
    for k, name in enumerate(("W1", "b1", "w2", "b2")):
        print(f"  {name}: {q(e_gpu[:, k])} | {q(e_orc[:, k])}   GPU closer in {np.mean(e_gpu[:, k] <= e_orc[:, k]):.3f} of the games", flush=True)
