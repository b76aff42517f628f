"""Helpers around tests/golden/td_parity.npz: the per-game weight change of the reference's UNMODIFIED
apply_td_updates (train.py:124-172) on 1,024 GPU-exported self-play trajectories per weight set
(tests/golden/td_traj.npz; generator: tests/golden/make_golden.py td_parity).

Metric: per game and tensor, max|dw - dw_ref| / max|dw_ref|, the maximum taken over the fixture's 320
coordinates of that game (its 64 largest |dw_ref|, the 64 largest |w0|, 192 fixed ones covering all four tensors) and
normalised by the game's full-tensor max|dw_ref| stored in the fixture."""
import numpy as np

BOUNDS = (0, 25344, 25472, 25600, 25601)
TENSORS = ("W1", "b1", "w2", "b2")


class TdFixture:
    def __init__(self, golden, tag):
        tr, fx, gm = golden("td_traj.npz"), golden("td_parity.npz"), golden("model.npz")
        self.tag = tag
        self.w0 = tuple(gm[f"{tag}_{k}"] for k in TENSORS)
        self.w0_flat = np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in self.w0])
        self.records, self.offsets, self.p1_won = tr[f"{tag}.records"], tr[f"{tag}.offsets"], tr[f"{tag}.p1_won"]
        self.n = len(self.p1_won)
        self.dmax, self.top_idx, self.new_at = fx[f"{tag}.dmax"], fx[f"{tag}.top_idx"], fx[f"{tag}.new_at"]
        self.tail_idx = np.concatenate([fx[f"{tag}.big_w_idx"], fx[f"{tag}.fixed_idx"]])
        self.losses, self.loss_offsets = fx[f"{tag}.losses"], fx[f"{tag}.loss_offsets"]
        self.lr, self.lam = float(fx["lr"]), float(fx["lam"])

    def trajectory(self, g):
        return self.records[self.offsets[g]:self.offsets[g + 1]]

    def coords(self, g):
        return np.concatenate([self.top_idx[g], self.tail_idx])

    def ref_losses(self, g):
        return self.losses[self.loss_offsets[g]:self.loss_offsets[g + 1]]

    def rel_error(self, g, new_flat):
        """-> float[4]: max|new - new_ref| over the game's fixture coordinates of each tensor / max|dw_ref| of the tensor"""
        idx = self.coords(g)
        d = np.abs(np.asarray(new_flat, np.float64).reshape(-1)[idx] - self.new_at[g].astype(np.float64))
        out = np.zeros(4)
        for k in range(4):
            sel = (idx >= BOUNDS[k]) & (idx < BOUNDS[k + 1])
            out[k] = d[sel].max() / float(self.dmax[g, k])
        return out


def quantiles(err):
    """err float[n, 4] -> {tensor: (p50, p99, max)}"""
    return {name: (float(np.median(err[:, k])), float(np.quantile(err[:, k], 0.99)), float(err[:, k].max()))
            for k, name in enumerate(TENSORS)}


def flat(weights):
    return np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in weights])


def engine_errors(eng, fx, games=None):
    """bgx_td_replay_host on the fixture's trajectories, each from the snapshot fx.w0 -> (err float[n,4] against the
    reference's torch results, worst |sqrt(sq) - sqrt(sq_ref)| over all steps, list of flat new weights)"""
    eng.set_weights(*fx.w0)
    games = range(fx.n) if games is None else games
    err, news, worst = [], [], 0.0
    for g in games:
        new, sq = eng.td_replay_host(fx.trajectory(g), bool(fx.p1_won[g]), fx.lr, fx.lam)
        nf = flat(new)
        err.append(fx.rel_error(g, nf))
        news.append(nf)
        if len(sq):
            worst = max(worst, float(np.max(np.abs(np.sqrt(sq) - np.sqrt(fx.ref_losses(g))))))
    return np.array(err), worst, news
