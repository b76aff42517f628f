"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref built by
`make -C oracle ref`):

    python tests/golden/make_golden.py

Sources of truth used here, none of them ours:
  * oracle/_ref/libref_harness.so   the unmodified reference engine (cppsrc/game.cpp …)
  * oracle/_ref/backgammon_env*.so  the reference's own pybind11 module
  * /root/reference/pysrc/TD(λ) model/{model,train}.py imported as they are (torch CPU, 1 thread)

The files written are small .npz fixtures; the tests compare the C restatement
(oracle/), the host library and the CUDA kernels against them.  The GPU box has
no /root/reference: nothing at test time reads it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("BGX_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "backgammon-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))          # the reference's backgammon_env
sys.path.insert(0, os.path.join(REF, "pysrc", "TD(λ) model"))

import torch  # noqa: E402

torch.set_num_threads(1)

import backgammon_env as bg  # noqa: E402  (reference module)
import model as ref_model  # noqa: E402  (reference model.py)
import train as ref_train  # noqa: E402  (reference train.py)

from bgx.synth import make_queries, start_record  # noqa: E402
from oracle.oracle import Oracle, RefHarness  # noqa: E402

assert "oracle/_ref" in bg.__file__.replace(os.sep, "/"), bg.__file__

M64 = (1 << 64) - 1


def mix64(x):
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def digest_of(moves, lens, states):
    """Independent (pure Python) statement of the enumeration digest, DESIGN.md."""
    dg = 0
    for k in range(len(lens)):
        st = [int(v) for v in states[k]]
        w = [0] * 5
        for i, v in enumerate(st):
            mag = abs(v)
            for b in range(4):
                w[b] |= ((mag >> b) & 1) << i
            if v < 0:
                w[4] |= 1 << i
        m = int(lens[k]) << 40
        for j in range(int(lens[k])):
            m |= (int(moves[k, j, 0]) | (int(moves[k, j, 1]) << 5)) << (10 * j)
        h = mix64(w[0] | (w[1] << 32))
        h = mix64(h ^ (w[2] | (w[3] << 32)))
        h = mix64(h ^ w[4])
        h = mix64(h ^ m)
        dg = (dg * 0x9E3779B97F4A7C15 + h) & M64
    return dg


def philox_dice(orc, seed, game, ply):
    x = orc.philox(seed, ply, game & 0xFFFFFFFF, game >> 32, 0)
    return orc.die(x[0]), orc.die(x[1])


def rolloff_first_player(orc, seed, game):
    k = 0
    while True:
        x = orc.philox(seed, k, game & 0xFFFFFFFF, game >> 32, 1)
        s1 = orc.die(x[0]) + orc.die(x[1])
        s2 = orc.die(x[2]) + orc.die(x[3])
        if s1 != s2:
            return 0 if s1 > s2 else 1
        k += 1


def game_row(game):
    row = np.zeros(28, np.int32)
    row[:24] = game.getGameBoard()
    row[24] = game.getJailedCount(0)
    row[25] = game.getJailedCount(1)
    row[26] = game.getBornOffCount(0)
    row[27] = game.getBornOffCount(1)
    return row


def weights_of(m):
    sd = m.state_dict()
    return (sd["fc1.weight"].numpy().copy(), sd["fc1.bias"].numpy().copy(),
            sd["fc2.weight"].numpy().copy(), sd["fc2.bias"].numpy().copy())


def main():
    orc = Oracle()
    ref = RefHarness()

    # ---------------------------------------------------------------- A: full enumerations
    q, cls = make_queries(900, seed=11)
    opening = []
    for pl in (0, 1):
        for d1 in range(1, 7):
            for d2 in range(1, 7):
                opening.append(start_record(pl, d1, d2))
    q = np.concatenate([np.array(opening, np.int8), q])
    offs, mv_all, ln_all, st_all = [0], [], [], []
    for r in q:
        mv, ln, st = ref.turn_sequences(r[:28].astype(np.int32), r[28], r[29], r[30])
        offs.append(offs[-1] + len(ln))
        mv_all.append(mv.reshape(-1, 8))
        ln_all.append(ln)
        st_all.append(st.astype(np.int8))
    np.savez_compressed(os.path.join(HERE, "enum_full.npz"), queries=q, offsets=np.array(offs, np.int64),
                        moves=np.concatenate(mv_all), lens=np.concatenate(ln_all), states=np.concatenate(st_all))
    print("enum_full:", len(q), "queries,", offs[-1], "sequences")

    # ---------------------------------------------------------------- B: summaries (N, U, digest)
    q2, _ = make_queries(20000, seed=12)
    N = np.zeros(len(q2), np.int64)
    U = np.zeros(len(q2), np.int64)
    D = np.zeros(len(q2), np.uint64)
    for i, r in enumerate(q2):
        mv, ln, st = ref.turn_sequences(r[:28].astype(np.int32), r[28], r[29], r[30])
        N[i] = len(ln)
        U[i] = len({bytes(x) for x in st.astype(np.int8)})
        D[i] = digest_of(mv, ln, st)
    np.savez_compressed(os.path.join(HERE, "enum_summary.npz"), queries=q2, n_seq=N, n_unique=U, digest=D)
    print("enum_summary:", len(q2), "queries,", int(N.sum()), "sequences,", int(U.sum()), "unique")

    # ---------------------------------------------------------------- F: legalMoves / tryMove
    rng = np.random.default_rng(13)
    q3, _ = make_queries(6000, seed=13)
    lm_n = np.zeros((len(q3), 6), np.int8)
    lm = np.zeros((len(q3), 6, 26, 2), np.int8)
    for i, r in enumerate(q3):
        for die in range(1, 7):
            mvs = ref.legal_moves(r[:28].astype(np.int32), r[28], die)
            lm_n[i, die - 1] = len(mvs)
            for j, (o, d) in enumerate(mvs):
                lm[i, die - 1, j] = (o, d)
    tm_in = np.zeros((len(q3), 4), np.int8)      # player, dice, origin, dest
    tm_ok = np.zeros(len(q3), np.int8)
    tm_err = []
    tm_out = np.zeros((len(q3), 28), np.int8)
    for i, r in enumerate(q3):
        s = r[:28].astype(np.int32)
        pl = int(rng.integers(0, 2))
        dice = int(rng.integers(1, 7))
        mode = rng.random()
        legal = ref.legal_moves(s, pl, dice)
        if mode < 0.4 and legal:                    # a legal move
            o, d = legal[int(rng.integers(0, len(legal)))]
        elif mode < 0.5:                            # unchecked bear-off path (quirk Q7)
            o = int(rng.integers(0, 26))
            d = int(rng.choice([0, 25]))
        elif mode < 0.75:                           # near-legal move
            o = int(rng.integers(0, 26))
            d = o + (dice if pl == 0 else -dice) + int(rng.integers(-1, 2)) * (rng.random() < 0.2)
            d = int(np.clip(d, -1, 26)) if rng.random() < 0.1 else int(np.clip(d, 0, 25))
        else:                                       # anything
            o = int(rng.integers(-2, 28))
            d = int(rng.integers(-2, 28))
        ok, err, out = ref.try_move(s, pl, dice, o, d)
        tm_in[i] = (pl, dice, o, d)
        tm_ok[i] = ok
        tm_err.append(err)
        tm_out[i] = out
    np.savez_compressed(os.path.join(HERE, "moves.npz"), queries=q3, legal_n=lm_n, legal=lm,
                        try_in=tm_in, try_ok=tm_ok, try_err=np.array(tm_err), try_out=tm_out)
    print("moves:", len(q3), "positions;", int(tm_ok.sum()), "legal tryMoves")

    # ---------------------------------------------------------------- C: encoding + values
    torch.manual_seed(0)
    m_rand = ref_model.TDLGammonModel()
    m_trained = ref_model.TDLGammonModel()
    m_trained.load_state_dict(torch.load(os.path.join(REF, "models", "tdgammonNEW100k.pth"),
                                         map_location="cpu", weights_only=True))
    w_rand, w_trained = weights_of(m_rand), weights_of(m_trained)
    all_states = np.concatenate(st_all)
    pick = np.random.default_rng(14).choice(len(all_states), 4000, replace=False)
    enc_states = all_states[pick].astype(np.int32)
    enc_turn = (np.arange(len(pick)) % 2).astype(np.int8)
    X = np.zeros((len(pick), 198), np.float32)
    for t in (0, 1):
        sel = enc_turn == t
        X[sel] = m_rand._encode_states_np(enc_states[sel], t)
    with torch.inference_mode():
        v_rand = m_rand(torch.from_numpy(X)).squeeze(1).numpy()
        v_trained = m_trained(torch.from_numpy(X)).squeeze(1).numpy()
    np.savez_compressed(os.path.join(HERE, "model.npz"), states=enc_states.astype(np.int8), turn=enc_turn, X=X,
                        v_rand=v_rand, v_trained=v_trained,
                        **{f"rand_{k}": a for k, a in zip(("W1", "b1", "w2", "b2"), w_rand)},
                        **{f"trained_{k}": a for k, a in zip(("W1", "b1", "w2", "b2"), w_trained)})
    print("model:", len(pick), "encodings")

    # ---------------------------------------------------------------- D/E: greedy games + TD replay
    SEED = 0x5EED2026
    games = {}
    for tag, mdl, gids in (("rand", m_rand, (3, 4)), ("trained", m_trained, (5, 6, 7))):
        for gid in gids:
            game = bg.Game(0)
            p1 = bg.Player("White", bg.PlayerType.PLAYER1)
            p2 = bg.Player("Black", bg.PlayerType.PLAYER2)
            game.setPlayers(p1, p2)
            game.setTurn(rolloff_first_player(orc, SEED, gid))
            pre, players, dice, chosen, chosen_len, after, nseq, enc = [], [], [], [], [], [], [], []
            ply = 0
            while True:
                enc.append(mdl.encode_state_np(game))                 # train.py:105-106
                pre.append(game_row(game))
                players.append(game.getTurn())
                d1, d2 = philox_dice(orc, SEED, gid, ply)
                game.setDice(d1, d2)
                dice.append((d1, d2))
                n_here = len(game.legalTurnSequences(game.getTurn(), d1, d2))
                seq = mdl.make_move(game, gid, epsilon=0.0)           # model.py:180-222
                nseq.append(n_here)
                c = np.zeros((4, 2), np.int8)
                for j, (o, d) in enumerate(seq):
                    c[j] = (o, d)
                chosen.append(c)
                chosen_len.append(len(seq))
                after.append(game_row(game))
                over, winner = game.is_game_over()
                if over:
                    break
                game.setTurn(1 - game.getTurn())
                ply += 1
            # TD(λ) replay through the reference's apply_td_updates from a fresh copy of the weights
            m2 = ref_model.TDLGammonModel()
            m2.load_state_dict(mdl.state_dict())
            m2.update_learning_params(1)                              # lr 0.1, lambda 0.9
            for name in m2.eligibility_traces:
                m2.eligibility_traces[name].zero_()
            opt = torch.optim.SGD(m2.parameters(), lr=0.1)
            m2.train()
            losses = ref_train.apply_td_updates(m2, opt, enc, winner == 0)
            nw = weights_of(m2)
            games[f"{tag}{gid}"] = dict(
                pre=np.array(pre, np.int8), player=np.array(players, np.int8), dice=np.array(dice, np.int8),
                chosen=np.array(chosen, np.int8), chosen_len=np.array(chosen_len, np.int8),
                after=np.array(after, np.int8), nseq=np.array(nseq, np.int64), winner=np.int8(winner),
                enc=np.array(enc, np.float32), losses=np.array(losses, np.float64),
                lr=np.float64(m2.learning_rate), lam=np.float64(m2.lambda_decay),
                new_W1=nw[0], new_b1=nw[1], new_w2=nw[2], new_b2=nw[3])
            print(f"game {tag}{gid}: {len(pre)} plies, winner {winner}, lr {m2.learning_rate}, lambda {m2.lambda_decay}")
    flat = {"seed": np.uint64(SEED), "names": np.array(sorted(games))}
    for name, g in games.items():
        for k, a in g.items():
            flat[f"{name}.{k}"] = a
    np.savez_compressed(os.path.join(HERE, "games.npz"), **flat)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")




# ---------------------------------------------------------------- G: TD(lambda) per-game parity fixture
#   python tests/golden/make_golden.py td_parity
# Input: tests/golden/td_traj.npz = 1,024 greedy self-play trajectories per weight set EXPORTED FROM THE GPU ENGINE
# (tools/export_td_trajectories.py, run on the B200 box).  Every trajectory is replayed here by the reference's UNMODIFIED
# apply_td_updates (train.py:124-172) from the round snapshot, exactly as train.py:538-542 calls it (update_learning_params(1):
# lr 0.1, lambda 0.9; traces zeroed).  The full per-game weight change is 25,601 floats; the fixture keeps, per game and
# tensor, max|dw|, and dw at 320 coordinates: the 64 largest |dw| of that game (where a relative error is decided), the 64
# largest |w0| (where fp32 rounding of w is coarsest) and 192 fixed random ones; plus every per-step squared TD error.

TD_FIXED = 192
TD_TOP = 64


def _td_one(args):
    tag, sd, X, p1_won = args
    torch.set_num_threads(1)
    m = ref_model.TDLGammonModel()
    m.load_state_dict(sd)
    m.update_learning_params(1)
    for name in m.eligibility_traces:
        m.eligibility_traces[name].zero_()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    m.train()
    losses = ref_train.apply_td_updates(m, opt, [x for x in X], bool(p1_won))
    new = np.concatenate([a.reshape(-1) for a in weights_of(m)])
    return new, np.array(losses, np.float64)


def td_coords(w0_flat):
    rng = np.random.default_rng(20261018)
    # every tensor is represented: 128 coordinates of W1, 31 of b1, 32 of w2, and b2
    fixed = np.concatenate([np.sort(rng.choice(25344, 128, replace=False)), 25344 + np.sort(rng.choice(128, 31, replace=False)),
                            25472 + np.sort(rng.choice(128, 32, replace=False)), [25600]]).astype(np.int32)
    assert fixed.size == TD_FIXED
    big_w = np.argsort(-np.abs(w0_flat), kind="stable")[:TD_TOP].astype(np.int32)
    return fixed, big_w


def td_parity():
    import multiprocessing as mp
    src = os.path.join(HERE, "td_traj.npz")
    with np.load(src) as z:
        traj = {k: z[k] for k in z.files}
    with np.load(os.path.join(HERE, "model.npz")) as z:
        gm = {k: z[k] for k in z.files}
    enc = ref_model.TDLGammonModel()
    out = {}
    for tag in ("rand", "trained"):
        w0 = tuple(gm[f"{tag}_{k}"] for k in ("W1", "b1", "w2", "b2"))
        w0_flat = np.concatenate([a.reshape(-1) for a in w0]).astype(np.float32)
        sd = {"fc1.weight": torch.from_numpy(w0[0].copy()), "fc1.bias": torch.from_numpy(w0[1].copy()),
              "fc2.weight": torch.from_numpy(w0[2].copy()), "fc2.bias": torch.from_numpy(w0[3].copy())}
        rec, offs, won = traj[f"{tag}.records"], traj[f"{tag}.offsets"], traj[f"{tag}.p1_won"]
        n = len(won)
        jobs = []
        for g in range(n):
            r = rec[offs[g]:offs[g + 1]]
            X = np.zeros((len(r), 198), np.float32)
            for t in (0, 1):                                     # the reference's own encoder, mover's flag (train.py:105-106)
                sel = r[:, 28] == t
                if sel.any():
                    X[sel] = enc._encode_states_np(r[sel, :28].astype(np.int32), t)
            jobs.append((tag, sd, X, int(won[g])))
        with mp.get_context("fork").Pool(os.cpu_count()) as pool:
            res = pool.map(_td_one, jobs, chunksize=8)
        fixed, big_w = td_coords(w0_flat)
        bounds = (0, 25344, 25472, 25600, 25601)
        dmax = np.zeros((n, 4), np.float32)
        top_idx = np.zeros((n, TD_TOP), np.int32)
        vals = np.zeros((n, TD_TOP + TD_TOP + TD_FIXED), np.float32)   # new weights at [top |dw| of the game, top |w0|, fixed]
        losses, loff = [], [0]
        for g, (new, ls) in enumerate(res):
            dw = new.astype(np.float64) - w0_flat
            for k in range(4):
                dmax[g, k] = np.max(np.abs(dw[bounds[k]:bounds[k + 1]]))
            top_idx[g] = np.argsort(-np.abs(dw), kind="stable")[:TD_TOP]
            vals[g] = new[np.concatenate([top_idx[g], big_w, fixed])]
            losses.append(ls)
            loff.append(loff[-1] + len(ls))
        out.update({f"{tag}.dmax": dmax, f"{tag}.top_idx": top_idx, f"{tag}.new_at": vals, f"{tag}.fixed_idx": fixed,
                    f"{tag}.big_w_idx": big_w, f"{tag}.losses": np.concatenate(losses), f"{tag}.loss_offsets": np.array(loff, np.int64)})
        print(tag, n, "games,", int(offs[-1]), "TD steps; median max|dW1|", float(np.median(dmax[:, 0])))
    out["lr"] = np.float64(0.1)
    out["lam"] = np.float64(0.9)
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "td_parity.npz"), **out)
    print("td_parity.npz", os.path.getsize(os.path.join(HERE, "td_parity.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "td_parity":
        td_parity()
    else:
        main()
