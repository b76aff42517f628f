"""Single-position host functions of libbgx (section 1 of include/bgx.h)."""
import ctypes as C

import numpy as np

from . import lib as L


def _pos(s):
    a = np.ascontiguousarray(s, dtype=np.int32).reshape(-1)
    if a.size != 28:
        raise ValueError("a position is 28 ints")
    return a


def legal_moves(position, player, die):
    """Game.legalMoves (cppsrc/game.cpp:80-105) -> list[(origin, dest)]"""
    out = np.zeros(52, np.int8)
    n = C.c_int()
    L.check(L.load().bgx_legal_moves(_pos(position).ctypes.data, int(player), int(die), out.ctypes.data, 26, C.byref(n)))
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n.value)]


def try_move(position, player, dice, origin, dest):
    """Game.tryMove (cppsrc/game.cpp:573-663) -> (ok, err, new position); input untouched."""
    st = _pos(position).copy()
    code = C.c_int()
    lib = L.load()
    L.check(lib.bgx_try_move(st.ctypes.data, int(player), int(dice), int(origin), int(dest), C.byref(code)))
    return code.value == 0, lib.bgx_move_error_string(code.value).decode(), st


def game_over(position):
    """Game.over (cppsrc/game.cpp:388-407) -> -1 / 0 / 1"""
    w = C.c_int()
    L.check(L.load().bgx_game_over(_pos(position).ctypes.data, C.byref(w)))
    return w.value


def turn_sequences(position, player, d1, d2):
    """Game.evaluateTurnSequences (cppsrc/game.cpp:193-222)
    -> moves int8[N,4,2], lens int8[N], states int32[N,28] in reference order."""
    s = _pos(position)
    lib = L.load()
    cap = 512
    while True:
        mv = np.zeros((cap, 4, 2), np.int8)
        ln = np.zeros(cap, np.int8)
        st = np.zeros((cap, 28), np.int32)
        n = C.c_int64()
        rc = lib.bgx_turn_sequences(s.ctypes.data, int(player), int(d1), int(d2), cap, mv.ctypes.data,
                                    ln.ctypes.data, st.ctypes.data, C.byref(n))
        if rc == L.E_CAPACITY:
            cap = int(n.value)
            continue
        L.check(rc)
        return mv[: n.value], ln[: n.value], st[: n.value]


def sequences_as_lists(position, player, d1, d2):
    mv, ln, st = turn_sequences(position, player, d1, d2)
    return [[(int(mv[i, j, 0]), int(mv[i, j, 1])) for j in range(ln[i])] for i in range(len(ln))], st


def advance(chosen, nxt, seed, ply=None, game_id=None, winner=None):
    """bgx_advance_host: the rest of a host-driven ply for n games (train.py:113-121) on numpy buffers.
    chosen int8[n,32] afterstate records -> nxt int8[n,32] (may be `chosen`): byte 31 = 1/2 when the game ended,
    else mover flipped; bytes 29,30 = Philox dice of (seed, ply[i], game_id[i]).  winner int8[n] optional."""
    n = chosen.shape[0]
    for a in (chosen, nxt):
        assert a.dtype == np.int8 and a.flags["C_CONTIGUOUS"] and a.shape == (n, 32)
    L.check(L.load().bgx_advance_host(chosen.ctypes.data, nxt.ctypes.data, n, int(seed),
                                      None if ply is None else ply.ctypes.data,
                                      None if game_id is None else game_id.ctypes.data,
                                      None if winner is None else winner.ctypes.data))
    return nxt
