"""CPU baseline driver: greedy 1-ply self-play on the REFERENCE engine.  TEST/BENCH INFRASTRUCTURE ONLY.

The engine is the unmodified reference module oracle/_ref/backgammon_env*.so (built from
/root/reference/cppsrc by oracle/Makefile; it travels to the GPU box, the Python sources under
/root/reference/pysrc do not).  The model side restates the reference's model.py as it stands —
NumPy `_encode_states_np` (model.py:111-144), a torch CPU 198-128-1 sigmoid MLP (model.py:36-37,
63-67) and `make_move` (model.py:180-222) — and the loop is play_game's (train.py:103-121) with
dice injected through Game.setDice.  One process per core, torch pinned to one thread, which is
how the reference parallelises self-play (train.py:240-259, 324-325).
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference_module():
    sys.path.insert(0, os.path.join(_HERE, "_ref"))
    import backgammon_env as bg
    if "_ref" not in bg.__file__:
        raise ImportError("backgammon_env resolved to %s, not the reference build" % bg.__file__)
    return bg


def available():
    d = os.path.join(_HERE, "_ref")
    return os.path.isdir(d) and any(f.startswith("backgammon_env") for f in os.listdir(d))


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


def worker(args):
    """Play greedy games for `seconds`; returns (plies, sequences, seconds)."""
    widx, weights, seed, seconds = args
    import torch
    torch.set_num_threads(1)
    bg = _import_reference_module()
    W1, b1, w2, b2 = (torch.from_numpy(np.asarray(a, np.float32)) for a in weights)
    fc1 = torch.nn.Linear(198, 128)
    fc2 = torch.nn.Linear(128, 1)
    with torch.no_grad():
        fc1.weight.copy_(W1.reshape(128, 198)); fc1.bias.copy_(b1.reshape(128))
        fc2.weight.copy_(w2.reshape(1, 128)); fc2.bias.copy_(b2.reshape(1))
    rng = np.random.default_rng(seed + widx)
    plies = sequences = 0
    t0 = time.perf_counter()
    gid = widx
    while time.perf_counter() - t0 < seconds:
        game = bg.Game(gid % 2)                                    # first mover g % 2 (benchmark.py:74)
        p1 = bg.Player("White", bg.PlayerType.PLAYER1)
        p2 = bg.Player("Black", bg.PlayerType.PLAYER2)
        game.setPlayers(p1, p2)
        players = {0: p1, 1: p2}
        while time.perf_counter() - t0 < seconds:
            d1, d2 = (int(x) for x in rng.integers(1, 7, 2))
            game.setDice(d1, d2)
            turn = game.getTurn()
            actions, states = game.evaluateTurnSequences(turn, d1, d2)       # model.py:201
            plies += 1
            sequences += len(actions)
            if actions:
                X = torch.from_numpy(encode_states_np(states, turn))
                with torch.inference_mode():
                    values = torch.sigmoid(fc2(torch.sigmoid(fc1(X)))).squeeze(1)
                idx = int(torch.argmax(values) if turn == 0 else torch.argmin(values))
                for o, dst in actions[idx]:
                    game.tryMove(players[turn], abs(o - dst), o, dst)
            over, _ = game.is_game_over()
            if over:
                break
            game.setTurn(1 - turn)
        gid += 1000
    return plies, sequences, time.perf_counter() - t0


def run(weights, seconds, processes, seed=1):
    """-> dict(plies, sequences, seconds (max over workers), processes)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(processes) as pool:
        res = pool.map(worker, [(i, weights, seed, seconds) for i in range(processes)])
    return {"plies": sum(r[0] for r in res), "sequences": sum(r[1] for r in res),
            "seconds": max(r[2] for r in res), "processes": processes}
