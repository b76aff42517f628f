"""Export self-play trajectories from the GPU engine for the TD(lambda) parity fixture (run on the B200 box).

    python tools/export_td_trajectories.py gpurun_out/td_traj.npz [n_games]

Per weight set of tests/golden/model.npz (random-init and trained): one greedy self-play round of n_games games
(k_selfplay, Philox dice, roll-off first mover, game ids 5000..), every pre-move record exported with
bgx_export_trajectory.  tests/golden/make_golden.py then runs the reference's unmodified apply_td_updates
(train.py:124-172) over exactly these trajectories and commits the results as tests/golden/td_parity.npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "backgammon-engine_b200"))

from bgx.engine import BatchEngine  # noqa: E402

SEED = 0x5EED2026


def main():
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    with np.load(os.path.join(ROOT, "tests", "golden", "model.npz")) as z:
        weights = {tag: tuple(z[f"{tag}_{k}"] for k in ("W1", "b1", "w2", "b2")) for tag in ("rand", "trained")}
    eng = BatchEngine(0)
    save = {"seed": np.uint64(SEED), "first_id": np.int64(5000)}
    for tag, w in weights.items():
        eng.set_weights(*w)
        eng.selfplay_init(n, first_id=5000, id_stride=n, seed=SEED, traj_cap=2048)
        st = eng.selfplay_round()
        assert st["truncated"] == 0 and st["games_finished"] == n, st
        rec, ply, gid = eng.selfplay_read()
        trajs = [eng.export_trajectory(s)[0] for s in range(n)]
        assert [len(t) for t in trajs] == ply.tolist()
        save[f"{tag}.records"] = np.concatenate(trajs)
        save[f"{tag}.offsets"] = np.concatenate([[0], np.cumsum(ply)]).astype(np.int64)
        save[f"{tag}.p1_won"] = (rec[:, 31] == 1).astype(np.int8)
        save[f"{tag}.game_id"] = gid
        print(tag, "games", n, "plies", int(ply.sum()), "P1 wins", int((rec[:, 31] == 1).sum()))
    np.savez_compressed(out, **save)
    print(out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
