"""CPU baseline driver: greedy 1-ply self-play on the REFERENCE engine.  TEST/BENCH INFRASTRUCTURE ONLY.

The engine is the unmodified reference module oracle/_ref/backgammon_env*.so (built from
/root/reference/cppsrc by oracle/Makefile, -O2; oracle/_ref/debug/ holds the -O0 -g build the
reference's own makefile ships, makefile:15); it travels to the GPU box, the Python sources under
/root/reference/pysrc do not.  Where /root/reference is visible the model side IS the reference's:
`TDLGammonModel.make_move` imported from its model.py, unmodified.  Elsewhere (the GPU box) it is
the restatement below of model.py as it stands - NumPy `_encode_states_np` (model.py:111-144), a
torch CPU 198-128-1 sigmoid MLP (model.py:36-37, 63-67), `make_move` (model.py:180-222) - which
tests/test_oracle_golden.py checks move for move against the imported reference.  The loop is
play_game's (train.py:103-121); dice are Philox4x32-10 keyed by (seed, game id), counter = ply -
the rolls the GPU arm plays - fed through Game.setDice (backgammon_bindings.cpp:86).
One process per core, torch pinned to one thread, which is how the reference parallelises
self-play (train.py:240-259, 324-325).
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


REF_MODEL_DIR = os.path.join(os.environ.get("BGX_REFERENCE", "/root/reference"), "pysrc", "TD(λ) model")
SEED = 0x5EED2026


def _import_reference_module(debug=False):
    sys.path.insert(0, os.path.join(_HERE, "_ref", "debug") if debug else os.path.join(_HERE, "_ref"))
    import backgammon_env as bg
    if "_ref" not in bg.__file__ or (debug and "debug" not in bg.__file__):
        raise ImportError("backgammon_env resolved to %s, not the reference build asked for" % bg.__file__)
    return bg


def available(debug=False):
    d = os.path.join(_HERE, "_ref", "debug") if debug else os.path.join(_HERE, "_ref")
    return os.path.isdir(d) and any(f.startswith("backgammon_env") for f in os.listdir(d))


def reference_model_available():
    return os.path.exists(os.path.join(REF_MODEL_DIR, "model.py"))


class PhiloxDice:
    """The dice of the GPU arm: Philox4x32-10, key = seed, counter = (ply, game id lo, hi, 0) (oracle/bgx_oracle.c)."""

    def __init__(self, seed=SEED):
        import ctypes as C
        self.C, self.seed = C, seed
        L = self.lib = C.CDLL(os.path.join(_HERE, "libbgx_oracle.so"))
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_philox4x32.restype = None
        L.orc_die.argtypes = [C.c_uint32]
        L.orc_die.restype = C.c_int
        self.out = (C.c_uint32 * 4)()

    def roll(self, game, ply):
        self.lib.orc_philox4x32(self.seed & 0xFFFFFFFF, self.seed >> 32, ply & 0xFFFFFFFF, game & 0xFFFFFFFF, (game >> 32) & 0xFFFFFFFF, 0, self.out)
        return self.lib.orc_die(self.out[0]), self.lib.orc_die(self.out[1])


def encode_states_np(states, turn):
    """model.py:111-144"""
    states = np.asarray(states)
    N = states.shape[0]
    board = states[:, :24]
    X = np.zeros((N, 198), dtype=np.float32)
    rows = np.arange(N)
    for i in range(24):
        col = board[:, i]
        n = np.abs(col)
        base = 8 * i + np.where(col > 0, 0, 4)
        for k in range(3):
            m = n >= k + 1
            X[rows[m], base[m] + k] = 1.0
        m = n >= 4
        X[rows[m], base[m] + 3] = (n[m] - 3) / 2
    X[:, 192] = 1.0 if turn == 0 else 0.0
    X[:, 193] = 0.0 if turn == 0 else 1.0
    X[:, 194] = states[:, 24] / 2
    X[:, 195] = states[:, 25] / 2
    X[:, 196] = states[:, 26] / 15.0
    X[:, 197] = states[:, 27] / 15.0
    return X


def play(bg, weights, games, seconds, dice, use_reference_model, record=None):
    """Greedy self-play of the games in `games` (ids) until `seconds` have passed -> (plies, sequences).  `record`, if a
    list, receives (game id, ply, chosen sequence) of every ply."""
    import torch
    W1, b1, w2, b2 = (torch.from_numpy(np.asarray(a, np.float32)) for a in weights)
    if use_reference_model:                                          # the reference's own model.py, as it is
        if REF_MODEL_DIR not in sys.path:
            sys.path.insert(0, REF_MODEL_DIR)
        import model as ref_model
        assert ref_model.bg is bg, "the reference model must run on the reference engine build chosen here"
        net = ref_model.TDLGammonModel()
        net.load_state_dict({"fc1.weight": W1.reshape(128, 198), "fc1.bias": b1.reshape(128), "fc2.weight": w2.reshape(1, 128), "fc2.bias": b2.reshape(1)})
    else:
        fc1 = torch.nn.Linear(198, 128)
        fc2 = torch.nn.Linear(128, 1)
        with torch.no_grad():
            fc1.weight.copy_(W1.reshape(128, 198)); fc1.bias.copy_(b1.reshape(128))
            fc2.weight.copy_(w2.reshape(1, 128)); fc2.bias.copy_(b2.reshape(1))
    plies = sequences = 0
    t0 = time.perf_counter()
    for gid in games:
        if time.perf_counter() - t0 >= seconds:
            break
        game = bg.Game(gid % 2)                                    # first mover g % 2 (benchmark.py:74)
        p1 = bg.Player("White", bg.PlayerType.PLAYER1)
        p2 = bg.Player("Black", bg.PlayerType.PLAYER2)
        game.setPlayers(p1, p2)
        players = {0: p1, 1: p2}
        ply = 0
        while time.perf_counter() - t0 < seconds:
            d1, d2 = dice.roll(gid, ply)
            game.setDice(d1, d2)                                     # backgammon_bindings.cpp:86
            turn = game.getTurn()
            if use_reference_model:
                n = len(game.legalTurnSequences(turn, d1, d2)) if record is not None else 0
                seq = net.make_move(game, gid, epsilon=0.0)          # model.py:180-222
                sequences += n
            else:
                actions, states = game.evaluateTurnSequences(turn, d1, d2)       # model.py:201
                sequences += len(actions)
                seq = []
                if actions:
                    X = torch.from_numpy(encode_states_np(states, turn))
                    with torch.inference_mode():
                        values = torch.sigmoid(fc2(torch.sigmoid(fc1(X)))).squeeze(1)
                    idx = int(torch.argmax(values) if turn == 0 else torch.argmin(values))
                    seq = actions[idx]
                    for o, dst in seq:
                        game.tryMove(players[turn], abs(o - dst), o, dst)
            plies += 1
            if record is not None:
                record.append((gid, ply, [tuple(m) for m in seq]))
            over, _ = game.is_game_over()
            if over:
                break
            game.setTurn(1 - turn)
            ply += 1
    return plies, sequences


def worker(args):
    """Play greedy games for `seconds`; returns (plies, sequences, seconds)."""
    widx, weights, seed, seconds, debug, n_workers = args
    import torch
    torch.set_num_threads(1)
    bg = _import_reference_module(debug)
    t0 = time.perf_counter()
    games = iter(range(widx, 1 << 40, n_workers))                    # worker w plays games w, w + W, ...
    plies, sequences = play(bg, weights, games, seconds, PhiloxDice(seed), reference_model_available())
    return plies, sequences, time.perf_counter() - t0


def run(weights, seconds, processes, seed=SEED, debug=False):
    """-> dict(plies, sequences, seconds (max over workers), processes, model: which model.py ran)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(processes) as pool:
        res = pool.map(worker, [(i, weights, seed, seconds, debug, processes) for i in range(processes)])
    return {"plies": sum(r[0] for r in res), "sequences": sum(r[1] for r in res),
            "seconds": max(r[2] for r in res), "processes": processes,
            "model": "reference model.py (imported)" if reference_model_available() else "restatement of model.py (oracle/ref_play.py)"}
