"""Pins the CPU oracle (oracle/bgx_oracle.c) to the reference.

Three anchors: the known answers of the reference's own cppsrc/tests.cpp, the
golden vectors generated from the reference itself (tests/golden/make_golden.py),
and — where oracle/_ref is present — the live reference engine on fresh seeds.
"""
import os

import numpy as np
import pytest

from conftest import golden_weights

START = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]


def state(board, jail=(0, 0), off=(0, 0)):
    return np.array(list(board) + list(jail) + list(off), np.int32)


# ------------------------------------------------------------------ known answers, cppsrc/tests.cpp

def test_legal_moves_start_known_answers(orc):
    s = state(START)
    assert orc.legal_moves(s, 0, 1) == [(1, 2), (17, 18), (19, 20)]          # tests.cpp:287-298
    assert orc.legal_moves(s, 1, 1) == [(6, 5), (8, 7), (24, 23)]            # tests.cpp:300-311
    assert orc.legal_moves(s, 0, 5) == [(12, 17), (17, 22)]                  # tests.cpp:313-323
    assert orc.legal_moves(state(START, jail=(1, 0)), 0, 6) == []            # tests.cpp:325-333
    assert orc.legal_moves(state(START, jail=(1, 0)), 0, 5) == [(0, 5)]      # tests.cpp:335-344


def test_try_move_known_answers(orc):
    b = list(START)
    b[1] = -1
    ok, err, out = orc.try_move(state(b), 0, 1, 1, 2)                         # tests.cpp:77-97 capture
    assert ok and out[1] == 1 and out[25] == 1
    ok, err, _ = orc.try_move(state(START), 0, 5, 1, 6)                       # tests.cpp:99-111
    assert not ok and err == "Invalid destination."
    ok, err, _ = orc.try_move(state(START, jail=(1, 0)), 0, 5, 12, 7)         # tests.cpp:113-127
    assert not ok and err == "Invalid origin"
    b = list(START)
    b[4] = 0
    ok, err, out = orc.try_move(state(b, jail=(1, 0)), 0, 5, 0, 5)
    assert ok and out[24] == 0 and out[4] == 1
    b = [0] * 24
    b[18] = 1
    ok, err, out = orc.try_move(state(b), 0, 6, 19, 25)                       # tests.cpp:218-243
    assert ok and out[26] == 1
    b = [-8] + [0] * 20 + [4, 8, 4]
    ok, _, _ = orc.try_move(state(b), 0, 5, 22, 25)                           # tests.cpp:411-420 (unchecked)
    assert ok


def test_game_over_known_answers(orc):
    assert orc.game_over(state(START)) == -1
    assert orc.game_over(state(START, off=(0, 15))) == 1                      # tests.cpp:189-202
    assert orc.game_over(state(START, off=(15, 0))) == 0                      # tests.cpp:246-258
    assert orc.game_over(state(START, off=(15, 15))) == 0                     # P1 checked first, game.cpp:393


def test_turn_sequences_known_answers(orc):
    s = state(START)
    seqs, _ = orc.sequences_as_lists(s, 0, 1, 2)                              # tests.cpp:346-355
    assert [(1, 2), (2, 4)] in seqs and [(1, 3), (3, 4)] in seqs
    seqs, _ = orc.sequences_as_lists(s, 1, 1, 2)                              # tests.cpp:357-365
    assert [(6, 5), (5, 3)] in seqs and [(6, 4), (4, 3)] in seqs
    seqs, _ = orc.sequences_as_lists(s, 0, 1, 1)                              # tests.cpp:367-388
    assert 81 < len(seqs) < 1000 and len(seqs) == 245
    assert [(19, 20)] * 4 in seqs and all(len(q) == 4 for q in seqs)
    lone = state([-1] + [0] * 23)
    assert orc.legal_moves(lone, 1, 3) == [(1, 0)]                            # tests.cpp:433-461
    seqs, _ = orc.sequences_as_lists(lone, 1, 3, 2)
    assert seqs == [[(1, 0)], [(1, 0)]]                                       # tests.cpp:462-488
    seqs, _ = orc.sequences_as_lists(lone, 1, 1, 1)                           # tests.cpp:492-514
    assert len(seqs) > 0 and all(q == [(1, 0)] for q in seqs)
    two = state([-1, -1] + [0] * 22)
    seqs, _ = orc.sequences_as_lists(two, 1, 2, 2)                            # tests.cpp:518-546
    assert [(2, 0), (1, 0)] in seqs or [(1, 0), (2, 0)] in seqs
    fr = state([-5, -4, -4] + [0] * 19 + [4, 5])
    seqs, _ = orc.sequences_as_lists(fr, 0, 6, 5)                             # tests.cpp:550-573
    assert len(seqs) <= 2
    b = state([-8] + [0] * 21 + [1, 4])
    assert len(orc.sequences_as_lists(b, 0, 3, 1)[0]) != 0                    # tests.cpp:400-410


def test_opening_counts_survey_8c(orc):
    """sequences/unique afterstates of the opening position, measured on the reference."""
    expect = {(1, 1): (245, 42), (1, 2): (30, 15), (1, 3): (31, 16), (1, 4): (27, 14), (1, 5): (15, 8),
              (1, 6): (19, 10), (2, 2): (538, 75), (2, 3): (35, 17), (2, 4): (37, 18), (2, 5): (17, 8),
              (2, 6): (28, 14), (3, 3): (536, 73), (3, 4): (34, 17), (3, 5): (18, 9), (3, 6): (28, 14),
              (4, 4): (411, 52), (4, 5): (18, 9), (4, 6): (28, 14), (5, 5): (15, 4), (5, 6): (14, 7), (6, 6): (71, 11)}
    s = state(START)
    for (a, b), (n, u) in expect.items():
        for pl in (0, 1):
            for d1, d2 in ((a, b), (b, a)):
                got_n, got_u, _ = orc.turn_summary(s, pl, d1, d2)
                assert (got_n, got_u) == (n, u), (pl, d1, d2)


def test_quirks_survey_a3(orc):
    # Q1: no global max-dice rule
    b = [0] * 24
    b[0] = 1; b[9] = 1; b[5] = -2; b[12] = -2; b[14] = -2
    seqs, _ = orc.sequences_as_lists(state(b, off=(13, 9)), 0, 2, 3)
    assert seqs == [[(1, 3)], [(10, 12), (1, 4)], [(1, 4), (10, 12)]]
    # Q3: P1 may only over-bear from its highest occupied point
    b = [0] * 24
    b[19] = 1; b[23] = 1
    assert orc.legal_moves(state(b, off=(13, 15)), 0, 6) == [(24, 25)]
    # Q4: an opponent checker on points origin+1..7 blocks P2's over-bear
    for p1_point, expect in ((5, []), (7, []), (8, [(2, 0)])):
        b = [0] * 24
        b[1] = -1; b[p1_point - 1] = 1
        assert orc.legal_moves(state(b, off=(14, 14)), 1, 4) == expect
    # Q5: no legal move -> [] for non-doubles, [[]] for doubles
    blocked = [0] * 24
    for i in range(6):
        blocked[i] = -2
    s = state(blocked, jail=(1, 0), off=(14, 3))
    assert orc.sequences_as_lists(s, 0, 3, 4)[0] == []
    seqs, st = orc.sequences_as_lists(s, 0, 3, 3)
    assert seqs == [[]] and np.array_equal(st[0], s)


# ------------------------------------------------------------------ golden vectors from the reference

def test_enumeration_full_golden(orc, golden):
    g = golden("enum_full.npz")
    offs = g["offsets"]
    for i, r in enumerate(g["queries"]):
        mv, ln, st = orc.turn_sequences(r[:28].astype(np.int32), r[28], r[29], r[30])
        a, b = offs[i], offs[i + 1]
        assert len(ln) == b - a, i
        assert np.array_equal(mv.reshape(-1, 8), g["moves"][a:b]), i
        assert np.array_equal(ln, g["lens"][a:b]), i
        assert np.array_equal(st.astype(np.int8), g["states"][a:b]), i


def test_enumeration_summary_golden(orc, golden):
    g = golden("enum_summary.npz")
    for i, r in enumerate(g["queries"]):
        n, u, d = orc.turn_summary(r[:28].astype(np.int32), r[28], r[29], r[30])
        assert (n, u, d) == (int(g["n_seq"][i]), int(g["n_unique"][i]), int(g["digest"][i])), i


def test_enumeration_summary_batch_golden(orc, golden):
    """The threaded batch form used by the 10^6-position GPU sweep gives the golden answers too."""
    g = golden("enum_summary.npz")
    for threads in (1, 3):
        n, u, d = orc.turn_summary_batch(g["queries"], threads=threads)
        assert np.array_equal(n, g["n_seq"].astype(np.int64))
        assert np.array_equal(u, g["n_unique"].astype(np.int64))
        assert np.array_equal(d, g["digest"].astype(np.uint64))
    n, u, d = orc.turn_summary_batch(np.zeros((0, 32), np.int8))
    assert len(n) == 0


def test_moves_golden(orc, golden):
    g = golden("moves.npz")
    for i, r in enumerate(g["queries"]):
        s = r[:28].astype(np.int32)
        for die in range(1, 7):
            n = g["legal_n"][i, die - 1]
            exp = [tuple(int(x) for x in p) for p in g["legal"][i, die - 1, :n]]
            assert orc.legal_moves(s, r[28], die) == exp, (i, die)
        pl, dice, o, d = (int(x) for x in g["try_in"][i])
        ok, err, out = orc.try_move(s, pl, dice, o, d)
        assert ok == bool(g["try_ok"][i]) and err == str(g["try_err"][i]), (i, err)
        assert np.array_equal(out.astype(np.int8), g["try_out"][i]), i


def test_encoding_golden_bit_exact(orc, golden):
    g = golden("model.npz")
    st, turn = g["states"].astype(np.int32), g["turn"]
    for t in (0, 1):
        X = orc.encode(st[turn == t], t)
        assert np.array_equal(X.view(np.uint32), g["X"][turn == t].view(np.uint32))


@pytest.mark.parametrize("tag", ["rand", "trained"])
def test_values_golden(orc, golden, tag):
    """fp32 value parity with torch's forward: <= 1e-5 relative (BASELINE.json north_star)."""
    g = golden("model.npz")
    V = orc.forward(golden_weights(g, tag), g["X"])
    ref = g[f"v_{tag}"]
    assert np.max(np.abs(V - ref) / np.abs(ref)) <= 1e-5


def test_td_replay_golden(orc, golden):
    """apply_td_updates parity: max|dw - dw_ref| / max|dw_ref| <= 1e-5 per tensor."""
    g = golden("games.npz")
    gm = golden("model.npz")
    for name in g["names"]:
        name = str(name)
        w0 = golden_weights(gm, "rand" if name.startswith("rand") else "trained")
        new, sq = orc.td_replay(w0, g[f"{name}.enc"], int(g[f"{name}.winner"]) == 0,
                                float(g[f"{name}.lr"]), float(g[f"{name}.lam"]))
        for a, b, k in zip(new, w0, ("W1", "b1", "w2", "b2")):
            dref = g[f"{name}.new_{k}"].reshape(-1) - np.asarray(b).reshape(-1)
            dgot = np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)
            # + one fp32 quantum of the stored weight: trained |w| reaches 6, whose ulp alone is
            # 5e-7, above 1e-5 * max|dw| (see DESIGN.md "TD parity metric")
            tol = 1e-5 * np.max(np.abs(dref)) + np.spacing(np.float32(np.max(np.abs(g[f"{name}.new_{k}"]))))
            assert np.max(np.abs(dgot - dref)) <= tol, (name, k)
        ref_l = g[f"{name}.losses"]
        # delta = v' - v cancels ~4 digits, so compare |delta| absolutely at the value tolerance
        assert np.max(np.abs(np.sqrt(sq) - np.sqrt(ref_l))) <= 1e-5, name


def test_td_replay_f64_yardstick(orc, golden):
    """The float64 replay (oracle.td_replay_f64) is the same algorithm: torch's own results (the golden vectors) and the
    fp32 C oracle both scatter around it at the fp32 noise level, and neither is systematically closer."""
    from oracle.oracle import td_replay_f64
    g = golden("games.npz")
    gm = golden("model.npz")
    for name in g["names"]:
        name = str(name)
        rand = name.startswith("rand")
        w0 = golden_weights(gm, "rand" if rand else "trained")
        args = (g[f"{name}.enc"], int(g[f"{name}.winner"]) == 0, float(g[f"{name}.lr"]), float(g[f"{name}.lam"]))
        new32, _ = orc.td_replay(w0, *args)
        new64 = td_replay_f64(w0, *args)
        for a32, a64, b, k in zip(new32, new64, w0, ("W1", "b1", "w2", "b2")):
            b = np.asarray(b, np.float64).reshape(-1)
            d64 = np.asarray(a64).reshape(-1) - b
            e_torch = np.max(np.abs(g[f"{name}.new_{k}"].reshape(-1) - b - d64)) / np.max(np.abs(d64))
            e_orc = np.max(np.abs(np.asarray(a32, np.float64).reshape(-1) - b - d64)) / np.max(np.abs(d64))
            bound = 1e-4 if rand else 2e-3          # trained |w| reaches 6: the weight's own fp32 quantum dominates
            assert e_torch <= bound and e_orc <= bound, (name, k, e_torch, e_orc)
            if k == "W1":                             # 25,344 entries: a stable statistic (b2 is a single number)
                assert e_orc <= 3 * e_torch and e_torch <= 3 * e_orc, (name, k, e_torch, e_orc)


def test_greedy_games_golden(orc, golden):
    """Replay the reference's greedy games ply by ply: dice spec, legal set, chosen afterstate."""
    g = golden("games.npz")
    gm = golden("model.npz")
    seed = int(g["seed"])
    for name in g["names"]:
        name = str(name)
        gid = int(name.lstrip("randtie"))
        w = golden_weights(gm, "rand" if name.startswith("rand") else "trained")
        pre, after = g[f"{name}.pre"].astype(np.int32), g[f"{name}.after"].astype(np.int32)
        soft = 0
        for t in range(len(pre)):
            x = orc.philox(seed, t, gid, 0, 0)
            assert (orc.die(x[0]), orc.die(x[1])) == tuple(int(v) for v in g[f"{name}.dice"][t])
            pl = int(g[f"{name}.player"][t])
            idx, out, v, n = orc.greedy_ply(w, pre[t], pl, *g[f"{name}.dice"][t])
            assert n == int(g[f"{name}.nseq"][t])
            if idx < 0 or g[f"{name}.chosen_len"][t] == 0:
                assert np.array_equal(pre[t], after[t])
                continue
            if not np.array_equal(out, after[t]):
                # a different pick is only acceptable inside the 1e-5 value tolerance
                X = orc.encode(after[t][None], pl)
                v_ref = orc.forward(w, X)[0]
                assert abs(v - v_ref) <= 1e-5 * abs(v_ref), (name, t)
                soft += 1
        assert soft <= max(2, len(pre) // 10), (name, soft)


# ------------------------------------------------------------------ live reference, fresh seeds

def test_fuzz_against_live_reference(orc, ref):
    from bgx.synth import make_queries
    q, _ = make_queries(4000, seed=987654)
    for r in q:
        s = r[:28].astype(np.int32)
        a = orc.turn_sequences(s, r[28], r[29], r[30])
        b = ref.turn_sequences(s, r[28], r[29], r[30])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


# ------------------------------------------------------------------ TD(lambda) per-game parity fixture

# max|dw - dw_ref| / max|dw_ref| per game of the fp32 C restatement (oracle/bgx_oracle.c: the reference's arithmetic step for
# step, ascending-index fp32 sums) against the reference's own apply_td_updates (torch) on the 1,024 fixture games:
# (p50, p99, max) as measured in the build container.  This is how far apart two faithful fp32 implementations of
# train.py:124-172 are - the noise floor under every "1e-5" statement about TD updates - and the yardstick the GPU kernel's
# distance from torch is printed beside (tests/test_gpu_parity.py, bench.py td_round.parity).
RESTATEMENT_VS_TORCH = {
    "rand": {"W1": (3.31e-6, 1.07e-5, 1.61e-5), "b1": (2.79e-6, 1.21e-5, 1.76e-5), "w2": (1.00e-6, 3.77e-6, 5.50e-6), "b2": (7.68e-7, 3.38e-6, 5.40e-6)},
    "trained": {"W1": (1.14e-5, 4.72e-5, 8.39e-5), "b1": (7.43e-6, 4.21e-5, 9.73e-5), "w2": (3.57e-6, 1.57e-5, 2.54e-5), "b2": (6.19e-7, 2.62e-5, 1.81e-4)},
}


@pytest.mark.parametrize("tag", ["rand", "trained"])
def test_td_parity_fixture_vs_c_restatement(orc, golden, tag):
    """The C oracle replays all 1,024 fixture trajectories; its distance from the reference's torch results is the
    documented fp32 noise floor (measured + 20 %), and the per-step TD errors agree at the value tolerance."""
    from concurrent.futures import ThreadPoolExecutor
    from td_fixture import TdFixture, flat, quantiles
    fx = TdFixture(golden, tag)

    def one(g):
        r = fx.trajectory(g)
        X = np.concatenate([orc.encode(r[t:t + 1, :28].astype(np.int32), int(r[t, 28])) for t in range(len(r))])
        new, sq = orc.td_replay(fx.w0, X, fx.p1_won[g], fx.lr, fx.lam)
        return fx.rel_error(g, flat(new)), float(np.max(np.abs(np.sqrt(sq) - np.sqrt(fx.ref_losses(g))))) if len(sq) else 0.0
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        res = list(ex.map(one, range(fx.n)))
    err = np.array([r[0] for r in res])
    assert max(r[1] for r in res) <= 1e-5
    q = quantiles(err)
    for name, want in RESTATEMENT_VS_TORCH[tag].items():
        for got, w, what in zip(q[name], want, ("p50", "p99", "max")):
            assert got <= 1.2 * w, (tag, name, what, got, w)
    # the headline: a faithful fp32 restatement does NOT stay within 1e-5 of torch on every game
    assert q["W1"][2] > 1e-5


@pytest.mark.parametrize("tag", ["rand", "trained"])
def test_td_parity_fixture_vs_torch_restatement(golden, tag):
    """tests/ref_td.py (the torch checker the GPU tests run live) against the fixture made by the reference's own
    apply_td_updates: same torch ops, so the same numbers up to the CPU's sgemv kernel choice."""
    from ref_td import replay_many
    from td_fixture import TdFixture
    fx = TdFixture(golden, tag)
    games = list(range(0, fx.n, 64))
    enc = _reference_encoder()
    jobs = [(fx.w0, enc(fx.trajectory(g)), int(fx.p1_won[g]), fx.lr, fx.lam) for g in games]
    for g, (new, sq) in zip(games, replay_many(jobs, procs=1)):
        e = fx.rel_error(g, new)
        assert np.all(e <= np.array([RESTATEMENT_VS_TORCH[tag][k][2] for k in ("W1", "b1", "w2", "b2")])), (tag, g, e)
        assert np.allclose(sq, fx.ref_losses(g), rtol=1e-3, atol=1e-12)


def _reference_encoder():
    """_encode_states_np (model.py:111-144) through the oracle's bit-exact restatement: records int8[T,32] -> X float32[T,198]"""
    from oracle.oracle import Oracle
    o = Oracle()
    return lambda r: np.concatenate([o.encode(r[t:t + 1, :28].astype(np.int32), int(r[t, 28])) for t in range(len(r))])


def test_cpu_baseline_driver_restatement_plays_the_reference_models_moves():
    """oracle/ref_play.py times the reference's own model.py where /root/reference is visible and a restatement of it on the
    GPU box: on the same Philox dice both choose the same sequence at every ply of whole games (reference engine underneath).
    Runs in a process of its own: `backgammon_env` there must be the reference's build, not the compat module."""
    import subprocess
    import sys
    from oracle import ref_play
    if not (ref_play.available() and ref_play.reference_model_available()):
        pytest.skip("needs oracle/_ref and /root/reference")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = """
import sys
import numpy as np
import torch
sys.path.insert(0, %r)
from oracle import ref_play
torch.set_num_threads(1)
bg = ref_play._import_reference_module()
with np.load(%r) as z:
    for tag in ("rand", "trained"):
        w = tuple(z[tag + "_" + k] for k in ("W1", "b1", "w2", "b2"))
        a, b = [], []
        ref_play.play(bg, w, [11, 12], 120.0, ref_play.PhiloxDice(), True, record=a)
        ref_play.play(bg, w, [11, 12], 120.0, ref_play.PhiloxDice(), False, record=b)
        assert len(a) > 60 and a == b, tag
print("SAME MOVES")
""" % (root, os.path.join(root, "tests", "golden", "model.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SAME MOVES" in r.stdout, r.stderr[-2000:]
