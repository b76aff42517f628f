"""ctypes front-end to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``   — oracle/libbgx_oracle.so, the C restatement (bgx_oracle.c); always buildable.
* ``RefHarness`` — oracle/_ref/libref_harness.so, the UNMODIFIED reference engine
  (cppsrc/game.cpp …) behind a C shim; present only where oracle/Makefile could
  see /root/reference at build time (the prebuilt .so travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under backgammon-engine_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")

def td_replay_f64(weights, X, player1_won, lr, lam):
    """apply_td_updates (train.py:124-172) in float64, with the closed-form gradients of the 2-layer sigmoid net
    (model.py:63-67): the exact arithmetic that every fp32 implementation - torch, the C oracle,
    the GPU kernel - approximates.  A yardstick, not a parity target: a TD error is the difference of two fp32 values
    and fp32 implementations scatter around this result by ~1e-5 of max|dw| (more with large trained weights).
    -> (W1, b1, w2, b2) float64 after the replay."""
    W1, b1, w2, b2 = (np.asarray(a, np.float64).copy() for a in weights)
    W1 = W1.reshape(128, 198); w2 = w2.reshape(-1); b2 = b2.reshape(-1)
    X = np.asarray(X, np.float64).reshape(-1, 198)
    eW1 = np.zeros_like(W1); eb1 = np.zeros_like(b1); ew2 = np.zeros_like(w2); eb2 = 0.0

    def sig(z):
        return 1.0 / (1.0 + np.exp(-z))
    T = len(X)
    for t in range(T):
        x = X[t]
        h = sig(W1 @ x + b1)
        v = sig(w2 @ h + b2[0])
        if t < T - 1:                                   # train.py:153-160: v' with the CURRENT weights
            d = sig(w2 @ sig(W1 @ X[t + 1] + b1) + b2[0]) - v
        else:                                           # train.py:165-168
            d = (1.0 if player1_won else 0.0) - v
        gv = v * (1.0 - v)
        gh = gv * w2 * h * (1.0 - h)
        eW1 = lam * eW1 + np.outer(gh, x); eb1 = lam * eb1 + gh      # train.py:141-147
        ew2 = lam * ew2 + gv * h; eb2 = lam * eb2 + gv
        c = lr * d
        W1 += c * eW1; b1 += c * eb1; w2 += c * ew2; b2[0] += c * eb2
    return W1, b1, w2.reshape(1, -1), b2


ERR_STRINGS = {
    0: "",
    1: "Invalid origin",
    2: "Origin out of range",
    3: "Destination out of range",
    4: "Cannot move in that direction.",
    5: "Move does not match dice.",
    6: "Invalid destination.",
    7: "Cannot bear off from jail",
}


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile. Building the checker is not using it."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _state(s):
    a = np.ascontiguousarray(s, dtype=np.int32).reshape(-1)
    assert a.size == 28, "a position is 28 ints"
    return a


class _SeqAPI:
    """Shared wrapper for the two libraries' turn-sequence entry point."""

    _turn = None

    def turn_sequences(self, s, player, d1, d2):
        """-> (moves int8[N,4,2], lens int8[N], states int32[N,28]) in reference order."""
        s = _state(s)
        cap = 1024
        while True:
            mv = np.zeros((cap, 4, 2), np.int8)
            ln = np.zeros(cap, np.int8)
            st = np.zeros((cap, 28), np.int32)
            n = self._turn(s, int(player), int(d1), int(d2), cap, mv.reshape(-1), ln, st.reshape(-1))
            if n >= 0:
                return mv[:n], ln[:n], st[:n]
            cap = -n

    def sequences_as_lists(self, s, player, d1, d2):
        """Same value the reference's Python API returns: list[list[tuple[int,int]]], ndarray[N,28]."""
        mv, ln, st = self.turn_sequences(s, player, d1, d2)
        seqs = [[(int(mv[i, j, 0]), int(mv[i, j, 1])) for j in range(ln[i])] for i in range(len(ln))]
        return seqs, st


class Oracle(_SeqAPI):
    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "libbgx_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_legal_moves.argtypes = [_i32p, C.c_int, C.c_int, _i8p]
        L.orc_legal_moves.restype = C.c_int
        L.orc_try_move.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_try_move.restype = C.c_int
        L.orc_game_over.argtypes = [_i32p]
        L.orc_game_over.restype = C.c_int
        L.orc_turn_sequences.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_long, _i8p, _i8p, _i32p]
        L.orc_turn_sequences.restype = C.c_long
        L.orc_turn_summary.argtypes = [_i32p, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
        L.orc_turn_summary.restype = None
        L.orc_turn_summary_batch.argtypes = [_i8p, C.c_long, C.c_int,
                                             np.ctypeslib.ndpointer(np.int64, flags="C"),
                                             np.ctypeslib.ndpointer(np.int64, flags="C"),
                                             np.ctypeslib.ndpointer(np.uint64, flags="C")]
        L.orc_turn_summary_batch.restype = None
        L.orc_greedy_batch.argtypes = [_f32p, _f32p, _f32p, _f32p, _i8p, C.c_long, C.c_int, _i8p, _f32p,
                                       np.ctypeslib.ndpointer(np.int64, flags="C"), _i8p, _i8p]
        L.orc_greedy_batch.restype = None
        L.orc_encode.argtypes = [_i32p, C.c_long, C.c_int, _f32p]
        L.orc_encode.restype = None
        L.orc_forward.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_long, _f32p, C.c_void_p]
        L.orc_forward.restype = None
        L.orc_td_replay.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_long, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p]
        L.orc_td_replay.restype = None
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_philox4x32.restype = None
        L.orc_die.argtypes = [C.c_uint32]
        L.orc_die.restype = C.c_int
        L.orc_greedy_ply.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p, C.c_int, C.c_int, C.c_int,
                                     _i32p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
        L.orc_greedy_ply.restype = C.c_long
        self._turn = L.orc_turn_sequences

    # -- rules
    def legal_moves(self, s, player, die):
        out = np.zeros(52, np.int8)
        n = self.lib.orc_legal_moves(_state(s), int(player), int(die), out)
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]

    def try_move(self, s, player, dice, origin, dest):
        """-> (ok, err string, new state).  The input is not modified."""
        st = _state(s).copy()
        code = self.lib.orc_try_move(st, int(player), int(dice), int(origin), int(dest))
        return code == 0, ERR_STRINGS[code], st

    def game_over(self, s):
        return self.lib.orc_game_over(_state(s))

    def turn_summary(self, s, player, d1, d2):
        n, u, d = C.c_int64(), C.c_int64(), C.c_uint64()
        self.lib.orc_turn_summary(_state(s), int(player), int(d1), int(d2), C.byref(n), C.byref(u), C.byref(d))
        return n.value, u.value, d.value

    def turn_summary_batch(self, records, threads=None):
        """records int8[n,32] (state, mover, d1, d2, pad) -> (N int64[n], U int64[n], digest uint64[n])."""
        r = np.ascontiguousarray(records, dtype=np.int8).reshape(-1, 32)
        n = r.shape[0]
        ns, nu, dg = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.uint64)
        if n:
            self.lib.orc_turn_summary_batch(r.reshape(-1), n, int(threads or os.cpu_count() or 1), ns, nu, dg)
        return ns, nu, dg

    # -- model
    def encode(self, states, turn):
        st = np.ascontiguousarray(states, dtype=np.int32).reshape(-1, 28)
        X = np.zeros((st.shape[0], 198), np.float32)
        self.lib.orc_encode(st.reshape(-1), st.shape[0], int(turn), X.reshape(-1))
        return X

    @staticmethod
    def _w(weights):
        W1, b1, w2, b2 = weights
        return (np.ascontiguousarray(W1, np.float32).reshape(-1), np.ascontiguousarray(b1, np.float32).reshape(-1),
                np.ascontiguousarray(w2, np.float32).reshape(-1), np.ascontiguousarray(b2, np.float32).reshape(-1))

    def forward(self, weights, X):
        X = np.ascontiguousarray(X, np.float32).reshape(-1, 198)
        V = np.zeros(X.shape[0], np.float32)
        self.lib.orc_forward(*self._w(weights), X.reshape(-1), X.shape[0], V, None)
        return V

    def td_replay(self, weights, X, player1_won, lr, lam):
        """-> (new weights tuple, squared td errors of the non-terminal steps)."""
        W1, b1, w2, b2 = (a.copy() for a in self._w(weights))
        X = np.ascontiguousarray(X, np.float32).reshape(-1, 198)
        T = X.shape[0]
        sq = np.zeros(max(T - 1, 1), np.float64)
        self.lib.orc_td_replay(W1, b1, w2, b2, X.reshape(-1), T, int(bool(player1_won)),
                               float(lr), float(lam), sq.ctypes.data_as(C.c_void_p))
        return (W1.reshape(128, 198), b1, w2.reshape(1, 128), b2), sq[: max(T - 1, 0)]

    # -- dice
    def philox(self, seed, c0, c1, c2=0, c3=0):
        out = (C.c_uint32 * 4)()
        self.lib.orc_philox4x32(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, c0 & 0xFFFFFFFF,
                                c1 & 0xFFFFFFFF, c2 & 0xFFFFFFFF, c3 & 0xFFFFFFFF, out)
        return [int(x) for x in out]

    def die(self, x):
        return self.lib.orc_die(x & 0xFFFFFFFF)

    def greedy_ply(self, weights, s, player, d1, d2):
        """-> (index or -1, afterstate int32[28], value, N)"""
        out = np.zeros(28, np.int32)
        v = C.c_float()
        n = C.c_int64()
        idx = self.lib.orc_greedy_ply(*self._w(weights), _state(s), int(player), int(d1), int(d2),
                                      out, C.byref(v), C.byref(n))
        return idx, out, v.value, n.value


    def greedy_batch(self, weights, records, threads=None):
        """-> dict(after int8[n,28], value f32[n], n_seq int64[n], moves int8[n,4,2], moves_len int8[n])"""
        r = np.ascontiguousarray(records, dtype=np.int8).reshape(-1, 32)
        n = r.shape[0]
        out = {"after": np.zeros((n, 28), np.int8), "value": np.zeros(n, np.float32), "n_seq": np.zeros(n, np.int64),
               "moves": np.zeros((n, 4, 2), np.int8), "moves_len": np.zeros(n, np.int8)}
        if n:
            self.lib.orc_greedy_batch(*self._w(weights), r.reshape(-1), n, int(threads or os.cpu_count() or 1),
                                      out["after"].reshape(-1), out["value"], out["n_seq"], out["moves"].reshape(-1), out["moves_len"])
        return out


class RefHarness(_SeqAPI):
    """The real reference engine. ``RefHarness.available()`` says whether the prebuilt .so is there."""

    PATH = os.path.join(_HERE, "_ref", "libref_harness.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        L = self.lib = C.CDLL(self.PATH)
        L.ref_legal_moves.argtypes = [_i32p, C.c_int, C.c_int, _i8p]
        L.ref_legal_moves.restype = C.c_int
        L.ref_try_move.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_try_move.restype = C.c_int
        L.ref_game_over.argtypes = [_i32p]
        L.ref_game_over.restype = C.c_int
        L.ref_turn_sequences.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_long, _i8p, _i8p, _i32p]
        L.ref_turn_sequences.restype = C.c_long
        L.ref_bench_enumerate.argtypes = [_i32p, _i8p, _i8p, C.c_long, C.POINTER(C.c_int64)]
        L.ref_bench_enumerate.restype = C.c_double
        L.ref_turn_summary_batch.argtypes = [_i8p, C.c_long, C.c_int, np.ctypeslib.ndpointer(np.int64, flags="C"),
                                             np.ctypeslib.ndpointer(np.uint64, flags="C")]
        L.ref_turn_summary_batch.restype = None
        self._turn = L.ref_turn_sequences

    def turn_summary_batch(self, records, threads=None):
        """The reference's own evaluateTurnSequences over records int8[n,32] -> (N int64[n], digest uint64[n]); the digest
        of the ordered (sequence, state) list is computed in the harness from the reference's output."""
        r = np.ascontiguousarray(records, dtype=np.int8).reshape(-1, 32)
        n = r.shape[0]
        ns, dg = np.zeros(n, np.int64), np.zeros(n, np.uint64)
        if n:
            self.lib.ref_turn_summary_batch(r.reshape(-1), n, int(threads or os.cpu_count() or 1), ns, dg)
        return ns, dg

    def legal_moves(self, s, player, die):
        out = np.zeros(64, np.int8)
        n = self.lib.ref_legal_moves(_state(s), int(player), int(die), out)
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]

    def try_move(self, s, player, dice, origin, dest):
        st = _state(s).copy()
        buf = C.create_string_buffer(64)
        ok = self.lib.ref_try_move(st, int(player), int(dice), int(origin), int(dest), buf, 64)
        return bool(ok), buf.value.decode(), st

    def game_over(self, s):
        return self.lib.ref_game_over(_state(s))

    def bench_enumerate(self, states, player, dice):
        """Time Game::evaluateTurnSequences over a batch -> (seconds, total sequences)."""
        st = np.ascontiguousarray(states, np.int32).reshape(-1, 28)
        pl = np.ascontiguousarray(player, np.int8)
        dc = np.ascontiguousarray(dice, np.int8).reshape(-1)
        tot = C.c_int64()
        sec = self.lib.ref_bench_enumerate(st.reshape(-1), pl, dc, st.shape[0], C.byref(tot))
        return sec, tot.value
