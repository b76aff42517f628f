"""Seeded synthetic positions for the enumeration sweep (BASELINE.json configs[1]).

Constructive classes follow SURVEY.md §8(d) row 2: contact mid-game, mover on the
bar (1-4 checkers, incl. fully blocked entries), all-home bear-off with 0..14 borne
off, bear-off with opponent checkers inside the mover's home (the PLAYER2 overrun
quirk, game.cpp:542-552), PLAYER1 overrun with gaps (game.cpp:526-536), pure races
and single-checker endings.  Every position has 15 checkers per side and no point
shared by both sides.  Dice are uniform over the 36 ordered pairs, the mover is
uniform over both players.

A query is one 32-byte record (the layout every batched C-ABI call takes):
  bytes 0..23 board, 24/25 jailed P1/P2, 26/27 borne-off P1/P2, 28 player, 29 d1, 30 d2, 31 zero.
"""
import numpy as np

START_BOARD = np.array([2, 0, 0, 0, 0, -5, 0, -3, 0, 0, 0, 5, -5, 0, 0, 0, 3, 0, 5, 0, 0, 0, 0, -2], np.int8)

CLASSES = ("contact", "bar", "bearoff", "bearoff_mixed", "race", "endgame")


def start_record(player=0, d1=1, d2=1):
    r = np.zeros(32, np.int8)
    r[:24] = START_BOARD
    r[28], r[29], r[30] = player, d1, d2
    return r


def _scatter(rng, owned, n_checkers):
    """Distribute n_checkers[b] checkers uniformly over the points where owned[b] is True."""
    B = owned.shape[0]
    n_owned = owned.sum(1)
    counts = np.zeros((B, 24), np.int64)
    rank = np.cumsum(owned, 1) - 1                      # rank of each owned point
    u = rng.random((B, 15))
    pick = np.minimum((u * n_owned[:, None]).astype(np.int64), np.maximum(n_owned - 1, 0)[:, None])
    live = (np.arange(15)[None, :] < n_checkers[:, None]) & (n_owned[:, None] > 0)
    for c in range(15):
        hit = owned & (rank == pick[:, c, None]) & live[:, c, None]
        counts += hit
    return counts


def make_queries(n, seed=20260101, classes=CLASSES):
    """-> int8[n, 32] query records, class ids int8[n]."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 32), np.int8)
    cls = rng.integers(0, len(classes), n).astype(np.int8)
    player = rng.integers(0, 2, n)
    pts = np.arange(24)[None, :]
    for ci, name in enumerate(classes):
        idx = np.nonzero(cls == ci)[0]
        B = idx.size
        if B == 0:
            continue
        mover = player[idx]
        # zones in the MOVER's frame, mapped to absolute points afterwards:
        # P1 home = points 19..24 (index 18..23), P2 home = 1..6 (index 0..5)
        home_m = np.where(mover[:, None] == 0, pts >= 18, pts <= 5)
        home_o = np.where(mover[:, None] == 0, pts <= 5, pts >= 18)
        bar_m = np.zeros(B, np.int64)
        bar_o = np.zeros(B, np.int64)
        off_m = np.zeros(B, np.int64)
        off_o = np.zeros(B, np.int64)
        r = rng.random((B, 24))
        if name == "contact":
            own_m = r < 0.30
            own_o = (r >= 0.30) & (r < 0.60)
            off_m = rng.integers(0, 4, B) * (rng.random(B) < 0.2)
            bar_o = rng.integers(0, 3, B) * (rng.random(B) < 0.3)
        elif name == "bar":
            own_m = r < 0.25
            own_o = (r >= 0.25) & (r < 0.65)
            bar_m = rng.integers(1, 5, B)
            bar_o = rng.integers(0, 3, B) * (rng.random(B) < 0.3)
            # a third of them: opponent owns most of its home board (blocked entries)
            blk = rng.random(B) < 0.33
            own_o = np.where(blk[:, None] & home_o, rng.random((B, 24)) < 0.85, own_o)
            own_m = own_m & ~own_o
        elif name == "bearoff":
            own_m = home_m & (r < 0.6)
            own_o = ~home_m & (r < 0.3)
            off_m = rng.integers(0, 15, B)
        elif name == "bearoff_mixed":
            own_m = home_m & (r < 0.5)
            # opponent checkers inside / next to the mover's home, mover stragglers outside
            near = np.where(mover[:, None] == 0, pts >= 16, pts <= 7)
            own_o = ~own_m & near & (rng.random((B, 24)) < 0.35)
            strag = ~near & (rng.random((B, 24)) < 0.04) & (rng.random(B) < 0.3)[:, None]
            own_m = own_m | strag
            own_o = own_o | (~near & ~strag & (rng.random((B, 24)) < 0.15))
            off_m = rng.integers(0, 15, B)
            bar_o = rng.integers(0, 3, B) * (rng.random(B) < 0.2)
        elif name == "race":
            split = rng.integers(6, 18, B)[:, None]
            side_m = np.where(mover[:, None] == 0, pts >= split, pts < split)
            own_m = side_m & (r < 0.45)
            own_o = ~side_m & (r < 0.45)
            off_m = rng.integers(0, 8, B) * (rng.random(B) < 0.5)
            off_o = rng.integers(0, 8, B) * (rng.random(B) < 0.5)
        else:  # endgame: 1-3 checkers left for the mover
            own_m = home_m & (r < 0.35)
            own_o = ~own_m & (rng.random((B, 24)) < 0.2)
            off_m = rng.integers(12, 15, B)
            off_o = rng.integers(0, 15, B)
        own_m = own_m & ~own_o
        n_m = 15 - bar_m - off_m
        n_o = 15 - bar_o - off_o
        cm = _scatter(rng, own_m, n_m)
        co = _scatter(rng, own_o, n_o)
        # whatever could not be placed (no owned point) goes off the board
        off_m = 15 - bar_m - cm.sum(1)
        off_o = 15 - bar_o - co.sum(1)
        sign = np.where(mover == 0, 1, -1)[:, None]
        board = sign * cm - sign * co
        rec = np.zeros((B, 32), np.int8)
        rec[:, :24] = board
        p1 = mover == 0
        rec[:, 24] = np.where(p1, bar_m, bar_o)
        rec[:, 25] = np.where(p1, bar_o, bar_m)
        rec[:, 26] = np.where(p1, off_m, off_o)
        rec[:, 27] = np.where(p1, off_o, off_m)
        rec[:, 28] = mover
        out[idx] = rec
    out[:, 29] = rng.integers(1, 7, n)
    out[:, 30] = rng.integers(1, 7, n)
    return out, cls


SWEEP_SEED = 20260101
PLAYOUT_SAMPLES_PER_GAME = 8


def make_sweep_queries(eng, n, seed=SWEEP_SEED):
    """The enumeration sweep of BASELINE.json configs[1] as SURVEY.md 8(d) config 2 defines it: half of the n positions are
    constructive (make_queries), half are sampled from random-vs-random playouts (config 1: benchmark.py:54-96, epsilon = 1
    self-play on `eng`, Philox dice under `seed`, first mover id % 2) at uniformly drawn plies, with the dice the playout
    rolled there.  `eng` is a BatchEngine with weights set (the random policy never looks at them); its self-play
    population is replaced.  -> int8[n, 32] queries, class ids int8[n] (len(CLASSES) = playout)."""
    from .lib import FIRST_PARITY
    n_con = n - n // 2
    q, cls = make_queries(n_con, seed=seed)
    n_play = n - n_con
    if n_play:
        per = PLAYOUT_SAMPLES_PER_GAME
        games = (n_play + per - 1) // per
        eng.selfplay_init(games, first_id=0, id_stride=games, seed=seed, first_mover=FIRST_PARITY, traj_cap=1024)
        st = eng.selfplay_round(epsilon=1.0)
        if st["truncated"]:
            raise RuntimeError(f"{st['truncated']} random playouts exceeded 1024 plies")
        rec = eng.selfplay_sample_host(per, seed=seed)[:n_play]
        q = np.concatenate([q, rec])
        cls = np.concatenate([cls, np.full(n_play, len(CLASSES), np.int8)])
    return q, cls
