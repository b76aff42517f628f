"""Torch CPU restatement of the reference's play_game / apply_td_updates (train.py:64-121, 124-172).

TEST INFRASTRUCTURE ONLY: the fp32 torch reference the CUDA TD(lambda) kernel is compared with at test time
(same torch ops as the reference: nn.Linear, sigmoid, autograd backward, scalar-tensor arithmetic).  It is pinned
to the reference's UNMODIFIED apply_td_updates by tests/golden/td_parity.npz (tests/test_oracle_golden.py).
Nothing under backgammon-engine_b200/ imports this file.
"""
import numpy as np
import torch


def play_game(model, game_idx, epsilon=0.0):
    """One self-play game on a compat Game. Returns (winner, states, total_moves) like train.py:64-121."""
    model.eval()
    import backgammon_env as bg
    game = bg.Game(0)
    white = bg.Player("White", bg.PlayerType.PLAYER1)
    black = bg.Player("Black", bg.PlayerType.PLAYER2)
    game.setPlayers(white, black)
    while True:                                       # opening roll-off by dice sums (train.py:89-97)
        a, b = sum(game.roll_dice()), sum(game.roll_dice())
        if a != b:
            break
    game.setTurn(bg.PlayerType.PLAYER1 if a > b else bg.PlayerType.PLAYER2)
    states, total_moves = [], 0
    while True:
        states.append(model.encode_state_np(game))    # pre-move encoding with the mover's flag
        game.roll_dice()
        model.make_move(game, game_idx, epsilon=epsilon)
        over, winner = game.is_game_over()
        if over:
            return winner, states, total_moves
        game.setTurn(bg.PlayerType.PLAYER2 if game.getTurn() == bg.PlayerType.PLAYER1 else bg.PlayerType.PLAYER1)
        total_moves += 1


def apply_td_updates(model, optimizer, states, player1_won):
    """Online TD(lambda) over one recorded game (train.py:124-172): traces reset by the caller,
    weights and model.eligibility_traces updated in place, returns the squared TD errors."""
    device = next(model.parameters()).device
    sq_errors = []

    def step(value, td_error):
        optimizer.zero_grad()
        value.backward()
        with torch.no_grad():
            for name, p in model.named_parameters():
                if p.requires_grad and p.grad is not None:
                    model.eligibility_traces[name] = model.lambda_decay * model.eligibility_traces[name] + p.grad.data
                    p.add_(model.learning_rate * td_error * model.eligibility_traces[name])

    tensors = [torch.from_numpy(s).to(device).unsqueeze(0) for s in states]
    for t in range(len(tensors) - 1):
        model.eval()
        with torch.no_grad():
            v_next = model(tensors[t + 1])
        model.train()
        v = model(tensors[t])
        delta = (v_next - v).item()
        step(v, delta)
        sq_errors.append(delta ** 2)
    if tensors:
        model.train()
        v = model(tensors[-1])
        step(v, (1.0 if player1_won else 0.0) - v.item())
    return sq_errors


def replay_worker(job):
    """(weights tuple, X float32[T,198], p1_won, lr, lam) -> flat new weights float32[25601], squared TD errors.
    Module-level so that a spawn-context process pool can run it."""
    w, X, p1_won, lr, lam = job
    torch.set_num_threads(1)
    import torch.nn as nn

    class Net(nn.Module):                                  # model.py:31-67, nothing but the parameters and forward
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(198, 128)
            self.fc2 = nn.Linear(128, 1)

        def forward(self, x):
            return torch.sigmoid(self.fc2(torch.sigmoid(self.fc1(x))))
    m = Net()
    with torch.no_grad():
        for p, a in zip((m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias), w):
            p.copy_(torch.from_numpy(np.asarray(a, np.float32).reshape(p.shape)))
    m.learning_rate, m.lambda_decay = lr, lam
    m.eligibility_traces = {name: torch.zeros_like(p.data) for name, p in m.named_parameters()}
    sq = apply_td_updates(m, torch.optim.SGD(m.parameters(), lr=0.1), [x for x in X], bool(p1_won))
    new = np.concatenate([p.detach().numpy().reshape(-1) for p in (m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias)])
    return new.astype(np.float32), np.asarray(sq, np.float64)


def replay_many(jobs, procs=None):
    """replay_worker over a spawn-context process pool (torch CPU, one thread per process)."""
    import multiprocessing as mp
    import os
    procs = procs or os.cpu_count() or 1
    if procs <= 1 or len(jobs) <= 2:
        return [replay_worker(j) for j in jobs]
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return pool.map(replay_worker, jobs, chunksize=max(1, len(jobs) // (8 * procs)))
