"""libbgx section 1 (single-position host functions behind the compat Game object) against
the reference's golden vectors, the tests.cpp known answers and the oracle.  These functions
share bgx_core.h's mask algebra with the CUDA kernels, so this suite also exercises the
device rules on CPU."""
import numpy as np
import pytest

from bgx import host

START = [2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2]


def state(board, jail=(0, 0), off=(0, 0)):
    return np.array(list(board) + list(jail) + list(off), np.int32)


def test_known_answers_tests_cpp():
    s = state(START)
    assert host.legal_moves(s, 0, 1) == [(1, 2), (17, 18), (19, 20)]         # tests.cpp:287-298
    assert host.legal_moves(s, 1, 1) == [(6, 5), (8, 7), (24, 23)]           # tests.cpp:300-311
    assert host.legal_moves(s, 0, 5) == [(12, 17), (17, 22)]                 # tests.cpp:313-323
    assert host.legal_moves(state(START, jail=(1, 0)), 0, 6) == []           # tests.cpp:325-333
    assert host.legal_moves(state(START, jail=(1, 0)), 0, 5) == [(0, 5)]     # tests.cpp:335-344
    ok, err, _ = host.try_move(s, 0, 5, 1, 6)                                 # tests.cpp:99-111
    assert not ok and err == "Invalid destination."
    ok, err, _ = host.try_move(state(START, jail=(1, 0)), 0, 5, 12, 7)        # tests.cpp:113-127
    assert not ok and err == "Invalid origin"
    seqs, _ = host.sequences_as_lists(s, 0, 1, 1)                             # tests.cpp:367-388
    assert len(seqs) == 245 and all(len(q) == 4 for q in seqs) and [(19, 20)] * 4 in seqs
    lone = state([-1] + [0] * 23)
    assert host.sequences_as_lists(lone, 1, 3, 2)[0] == [[(1, 0)], [(1, 0)]]  # tests.cpp:462-488
    assert host.game_over(state(START, off=(0, 15))) == 1                     # tests.cpp:189-202
    assert host.game_over(state(START, off=(15, 15))) == 0
    assert host.game_over(s) == -1


def test_enumeration_full_golden(golden):
    g = golden("enum_full.npz")
    offs = g["offsets"]
    for i, r in enumerate(g["queries"]):
        mv, ln, st = host.turn_sequences(r[:28].astype(np.int32), r[28], r[29], r[30])
        a, b = offs[i], offs[i + 1]
        assert len(ln) == b - a, i
        assert np.array_equal(mv.reshape(-1, 8), g["moves"][a:b]), i
        assert np.array_equal(ln, g["lens"][a:b]), i
        assert np.array_equal(st.astype(np.int8), g["states"][a:b]), i


def test_moves_golden(golden):
    g = golden("moves.npz")
    for i, r in enumerate(g["queries"]):
        s = r[:28].astype(np.int32)
        for die in range(1, 7):
            n = g["legal_n"][i, die - 1]
            exp = [tuple(int(x) for x in p) for p in g["legal"][i, die - 1, :n]]
            assert host.legal_moves(s, r[28], die) == exp, (i, die)
        pl, dice, o, d = (int(x) for x in g["try_in"][i])
        ok, err, out = host.try_move(s, pl, dice, o, d)
        assert ok == bool(g["try_ok"][i]) and err == str(g["try_err"][i]), (i, err)
        assert np.array_equal(out.astype(np.int8), g["try_out"][i]), i


def test_fuzz_against_oracle_odd_arguments(orc):
    """Arguments outside the kernels' domain take the scalar path: odd dice, odd players, big stacks."""
    from bgx.synth import make_queries
    rng = np.random.default_rng(5)
    q, _ = make_queries(1500, seed=99)
    for r in q:
        s = r[:28].astype(np.int32)
        die = int(rng.integers(-3, 12))
        assert host.legal_moves(s, int(r[28]), die) == orc.legal_moves(s, int(r[28]), die), (s, die)
        if rng.random() < 0.3:
            s = s.copy()
            s[int(rng.integers(0, 24))] = int(rng.integers(-20, 21))
        pl, dice = int(rng.integers(0, 2)), int(rng.integers(-1, 9))
        o, d = int(rng.integers(-3, 29)), int(rng.integers(-3, 29))
        a, b = host.try_move(s, pl, dice, o, d), orc.try_move(s, pl, dice, o, d)
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
        assert host.legal_moves(s, pl, int(r[29])) == orc.legal_moves(s, pl, int(r[29]))


def test_capacity_and_errors():
    import ctypes as C
    from bgx import lib as L
    lib = L.load()
    s = state(START)
    n = C.c_int64()
    rc = lib.bgx_turn_sequences(s.ctypes.data, 0, 3, 3, 10, None, None, None, C.byref(n))
    assert rc == L.E_CAPACITY and n.value == 536 and b"536" in lib.bgx_last_error()
    assert lib.bgx_turn_sequences(None, 0, 3, 3, 10, None, None, None, C.byref(n)) == L.E_INVALID


def test_library_exports_every_declared_symbol():
    """include/bgx.h <-> libbgx.so <-> bgx/lib.py: every declared entry point is exported and bound."""
    import os
    import re
    from bgx import lib as L
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "bgx.h")).read()
    declared = set(re.findall(r"\b(bgx_[a-z0-9_]+)\s*\(", hdr))
    bound = set(L.exported_symbols())
    assert declared == bound, declared ^ bound
    lib = L.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bgx import BgxError
    from bgx.engine import BatchEngine
    with pytest.raises(BgxError) as ei:
        BatchEngine(0)
    assert ei.value.code == -2 and "no CPU path" in str(ei.value)


def test_advance_host_matches_the_dice_rule_and_game_over(orc):
    """bgx_advance_host = the rest of a host-driven ply (train.py:113-121): winner as Game::over reports it
    (PLAYER1 first, game.cpp:388-407), mover flipped unless the game ended, Philox dice of (seed, ply, game id)."""
    from bgx import host
    rng = np.random.default_rng(3)
    n = 500
    c = np.zeros((n, 32), np.int8)
    c[:, :24] = rng.integers(-3, 4, (n, 24))
    c[:, 26] = rng.choice([0, 3, 15], n)
    c[:, 27] = rng.choice([0, 7, 15], n)
    c[:, 28] = rng.integers(0, 2, n)
    ply = rng.integers(0, 400, n).astype(np.int32)
    gid = rng.integers(0, 2 ** 40, n).astype(np.int64)
    win = np.zeros(n, np.int8)
    seed = 0x0123456789ABCDEF
    nxt = host.advance(c, np.zeros_like(c), seed, ply, gid, win)
    for i in range(n):
        w = orc.game_over(c[i, :28].astype(np.int32))
        assert win[i] == w and nxt[i, 31] == w + 1
        assert np.array_equal(nxt[i, :28], c[i, :28])
        assert nxt[i, 28] == (c[i, 28] if w >= 0 else c[i, 28] ^ 1)
        x = orc.philox(seed, int(ply[i]), int(gid[i]) & 0xFFFFFFFF, int(gid[i]) >> 32)
        assert (nxt[i, 29], nxt[i, 30]) == (orc.die(x[0]), orc.die(x[1]))
    # in place, defaults (ply 0, game id = index)
    d = c.copy()
    host.advance(d, d, seed)
    x = orc.philox(seed, 0, 7, 0)
    assert (d[7, 29], d[7, 30]) == (orc.die(x[0]), orc.die(x[1]))
