"""TD(lambda) per-game parity of the GPU kernel against the reference's own apply_td_updates (run on the B200 box).

    python tools/td_parity_probe.py [out.json] [--live-torch N] [--f64 N]

For both weight sets: bgx_td_replay_host on the 1,024 fixture trajectories against tests/golden/td_parity.npz (the
reference's torch results) -> p50 / p99 / max of max|dw - dw_ref| / max|dw_ref| per tensor; optionally the same against
a live torch replay over ALL 25,601 coordinates (tests/ref_td.py, process pool) and against the float64 replay.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "backgammon-engine_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from bgx.engine import BatchEngine  # noqa: E402
from conftest import load_golden  # noqa: E402
from td_fixture import BOUNDS, TENSORS, TdFixture, engine_errors, quantiles  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
    n_live = int(sys.argv[sys.argv.index("--live-torch") + 1]) if "--live-torch" in sys.argv else 0
    n_f64 = int(sys.argv[sys.argv.index("--f64") + 1]) if "--f64" in sys.argv else 0
    eng = BatchEngine(0)
    report = {}
    for tag in ("rand", "trained"):
        fx = TdFixture(load_golden, tag)
        t0 = time.time()
        err, worst_sq, news = engine_errors(eng, fx)
        rep = {"games": fx.n, "gpu_vs_reference": quantiles(err), "worst_step_td_error_diff": worst_sq, "gpu_seconds": time.time() - t0}
        if n_live or n_f64:
            from oracle.oracle import Oracle, td_replay_f64
            orc = Oracle()
            enc = lambda r: np.concatenate([orc.encode(r[t:t + 1, :28].astype(np.int32), int(r[t, 28])) for t in range(len(r))])
        if n_live:
            from ref_td import replay_many
            games = list(range(0, fx.n, max(1, fx.n // n_live)))[:n_live]
            t0 = time.time()
            res = replay_many([(fx.w0, enc(fx.trajectory(g)), int(fx.p1_won[g]), fx.lr, fx.lam) for g in games])
            full = np.zeros((len(games), 4)); samp = np.zeros((len(games), 4)); live_fix = np.zeros((len(games), 4))
            for i, (g, (new, sq)) in enumerate(zip(games, res)):
                d_ref = new.astype(np.float64) - fx.w0_flat
                d = np.abs(news[g].astype(np.float64) - new.astype(np.float64))
                for k in range(4):
                    full[i, k] = d[BOUNDS[k]:BOUNDS[k + 1]].max() / np.abs(d_ref[BOUNDS[k]:BOUNDS[k + 1]]).max()
                samp[i] = err[g]
                live_fix[i] = fx.rel_error(g, new)
            rep["gpu_vs_live_torch_all_coordinates"] = quantiles(full)
            rep["same_games_fixture_coordinates"] = quantiles(samp)
            rep["live_torch_vs_fixture"] = quantiles(live_fix)
            rep["live_torch_seconds"] = time.time() - t0
        if n_f64:
            games = list(range(0, fx.n, max(1, fx.n // n_f64)))[:n_f64]
            e_gpu = np.zeros((len(games), 4)); e_ref = np.zeros((len(games), 4))
            for i, g in enumerate(games):
                new64 = np.concatenate([np.asarray(a).reshape(-1) for a in td_replay_f64(fx.w0, enc(fx.trajectory(g)), fx.p1_won[g], fx.lr, fx.lam)])
                idx = fx.coords(g)
                for k in range(4):
                    sel = (idx >= BOUNDS[k]) & (idx < BOUNDS[k + 1])
                    e_gpu[i, k] = np.abs(news[g][idx].astype(np.float64) - new64[idx])[sel].max() / fx.dmax[g, k]
                    e_ref[i, k] = np.abs(fx.new_at[g].astype(np.float64) - new64[idx])[sel].max() / fx.dmax[g, k]
            rep["gpu_vs_float64"] = quantiles(e_gpu)
            rep["reference_vs_float64"] = quantiles(e_ref)
        report[tag] = rep
        print(tag, json.dumps(rep, indent=1))
    if out_path:
        json.dump(report, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
