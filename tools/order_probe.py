"""Queue-order / help-policy probe for k_select on one ply of the self-play population (run on a GPU box)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "backgammon-engine_b200")]
import numpy as np, torch
from bench import init_weights, GAMES_PER_GPU as G
from bgx.engine import BatchEngine
from bgx import host as H
w = init_weights()
eng = BatchEngine(0); eng.set_weights(*w)
eng.selfplay_init(G, first_mover=1)
for _ in range(8): eng.selfplay_step(16, want_stats=False)
rec, _, _ = eng.selfplay_read(); rec[:, 31] = 0
rng = np.random.default_rng(0); rec[:, 29:31] = rng.integers(1, 7, (G, 2))
q = torch.from_numpy(rec).cuda()
nseq = torch.zeros(G, dtype=torch.int32, device="cuda")
chosen = torch.zeros((G, 32), dtype=torch.int8, device="cuda")
eng.select_moves(q, chosen=chosen, n_seq=nseq); torch.cuda.synchronize()
ns = nseq.cpu().numpy()
dbl = rec[:, 29] == rec[:, 30]
kids = np.zeros(G, np.int32)
for i in np.flatnonzero(dbl):
    kids[i] = len(H.legal_moves(rec[i, :28].astype(np.int32), int(rec[i, 28]), int(rec[i, 29])))
print("doubles", dbl.sum(), "kids>=4", (kids >= 4).sum(), ">=8", (kids >= 8).sum(), ">=10", (kids >= 10).sum(),
      "seq share kids>=8: %.3f, kids>=4: %.3f" % (ns[kids >= 8].sum() / ns.sum(), ns[kids >= 4].sum() / ns.sum()), flush=True)
k1 = np.zeros(G, np.int64); k2 = np.zeros(G, np.int64)
for i in range(G):
    st = rec[i, :28].astype(np.int32)
    k1[i] = len(H.legal_moves(st, int(rec[i, 28]), int(rec[i, 29])))
    k2[i] = k1[i] if dbl[i] else len(H.legal_moves(st, int(rec[i, 28]), int(rec[i, 30])))
bar = np.where(rec[:, 28] == 0, rec[:, 24], rec[:, 25]).astype(np.int64)
kf1 = k1.copy(); kf2 = k2.copy()
for i in np.flatnonzero(bar > 0):
    st = rec[i, :28].astype(np.int32); st[24 + int(rec[i, 28])] = 0
    kf1[i] = len(H.legal_moves(st, int(rec[i, 28]), int(rec[i, 29]))) + 1
    kf2[i] = kf1[i] if dbl[i] else len(H.legal_moves(st, int(rec[i, 28]), int(rec[i, 30]))) + 1
est0 = np.where(dbl, k1 ** 4, 2 * k1 * k2 + k1 + k2)
estb = np.where(dbl, np.where(bar >= 4, k1, k1 * kf1 ** np.clip(4 - bar, 0, 4)),
                np.where(bar == 1, k1 * kf2 + k2 * kf1 + k1 + k2, k1 * k2))
est = np.where(bar > 0, estb, est0)
lg = np.floor(np.log2(est + 1)).astype(np.int64)
for b in range(int(lg.max()) + 1):
    m = lg == b
    if m.any(): print(f"bucket {b}: {m.sum()} queries, n_seq mean {ns[m].mean():.1f} max {ns[m].max()}", flush=True)
lg2 = np.floor(2 * np.log2(est + 1)).astype(np.int64)
nat = np.arange(G)
def first(mask_list):
    used = np.zeros(G, bool); parts = []
    for m in mask_list:
        parts.append(nat[m & ~used]); used |= m
    parts.append(nat[~used])
    return np.concatenate(parts)
orders = {
    "natural": nat,
    "nseq_desc": np.argsort(-ns, kind="stable"),
    "est_desc": np.argsort(-est, kind="stable"),
    "log2_buckets": np.argsort(-lg, kind="stable"),
    "half_log2": np.argsort(-lg2, kind="stable"),
    "kids_desc": np.argsort(-kids, kind="stable"),
}
eng.close()
policies = [(8, 10, 0), (7, 10, 0), (6, 10, 0), (5, 10, 0), (7, 8, 0)]
for n in (G, G // 2):
    for (umin, gmin, pct) in policies:
        e = BatchEngine(0); e.set_weights(*w)
        for key, val in (("select_urgent_min", umin), ("select_giant_min", gmin), ("select_urgent_from_pct", pct)):
            e.set_option(key, val)
        row = []
        for name, order in orders.items():
            o = order[order < n] if n < G else order
            qs = q[torch.from_numpy(o).cuda()].contiguous()
            ch = torch.zeros((n, 32), dtype=torch.int8, device="cuda")
            ms = []
            for _ in range(7):
                e.select_moves(qs, chosen=ch); torch.cuda.synchronize(); ms.append(e.last_kernel_ms())
            row.append(f"{name} {np.mean(ms[2:]):.3f}")
        print(f"n={n} urgent_min={umin} giant_min={gmin} from={pct}%: " + " | ".join(row), flush=True)
        e.close()
