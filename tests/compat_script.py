"""Scripted session against a `backgammon_env` module; run once with the reference build and once
with ours (tests/test_compat_module.py) and the two transcripts must be identical.
usage: python compat_script.py <dir containing backgammon_env*.so> <out.json>   (board pictures go to stdout)"""
import json
import random
import zlib
import sys

sys.path.insert(0, sys.argv[1])
import backgammon_env as bg  # noqa: E402

T = []
rec = T.append
rec(["module_has", sorted(n for n in ("Game", "Player", "Pieces", "PlayerType") if hasattr(bg, n))])
rec(["enum", int(bg.PlayerType.PLAYER1), int(bg.PlayerType.PLAYER2), bg.PlayerType.PLAYER1 == 0, 1 == bg.PlayerType.PLAYER2,
     {bg.PlayerType.PLAYER1: "a", bg.PlayerType.PLAYER2: "b"}.get(1)])
for bad in (lambda: bg.Player("W", 0), lambda: bg.Pieces()):
    try:
        bad()
        rec(["no error"])
    except Exception as e:
        rec([type(e).__name__])
for arg in (0, 1, 2, 3, -1, 7):
    rec(["ctor turn", arg, bg.Game(arg).getTurn()])
p1, p2 = bg.Player("White", bg.PlayerType.PLAYER1), bg.Player("Black", bg.PlayerType.PLAYER2)
rec(["player", p1.getName(), p1.getNum(), p2.getName(), p2.getNum()])
g = bg.Game(0)
g.setPlayers(p1, p2)
rec(["getPlayers", g.getPlayers(0).getName(), g.getPlayers(1).getName(), g.getPlayers(5).getName(),
     g.getPlayers(bg.PlayerType.PLAYER2).getNum()])
g.setTurn(5)
rec(["setTurn verbatim", g.getTurn()])
g.setTurn(0)
rec(["dice default", list(g.get_last_dice())])
d = g.roll_dice()
rec(["roll range", len(d), all(1 <= x <= 6 for x in d), list(g.get_last_dice()) == list(d)])
g.setDice(3, 4)
c = g.clone()
rec(["clone dice", list(c.get_last_dice()), list(g.get_last_dice()), c.getGameBoard() == g.getGameBoard()])
c.tryMove(p1, 1, 1, 2)
rec(["clone independent", c.getGameBoard() != g.getGameBoard()])
rec(["legalMoves enum arg", g.legalMoves(bg.PlayerType.PLAYER1, 3), g.getJailedCount(bg.PlayerType.PLAYER1)])
seqs, states = g.evaluateTurnSequences(0, 3, 3)
rec(["eval type", type(seqs).__name__, type(seqs[0]).__name__, type(seqs[0][0]).__name__, str(states.dtype), list(states.shape),
     bool(states.flags["C_CONTIGUOUS"]), bool(states.flags["WRITEABLE"])])
rec(["tryMove errors",
     g.tryMove(p2, 1, 1, 2), g.tryMove(p1, 5, 1, 6), g.tryMove(p1, 2, 1, 2), g.tryMove(p1, 1, 1, 0), g.tryMove(p1, 1, 12, 11),
     g.tryMove(p1, 3, 1, 30), g.tryMove(p1, 3, 30, 1), g.tryMove(p1, 3, -1, 2)])
pc = g.getPieces()
rec(["pieces", pc.numJailed(0), pc.numFreed(1), pc.numJailed(bg.PlayerType.PLAYER2)])
g.setBorneOffPieces(0, 15)
g.setBorneOffPieces(1, 15)
rec(["both off", list(g.is_game_over())])
g.reset()
rec(["reset keeps counters", g.getBornOffCount(0), g.getBornOffCount(1)])
g.setBorneOffPieces(0, 0)
g.setBorneOffPieces(1, 0)
rec(["not over", list(g.is_game_over())])
g.printGameBoard()

# seeded random playouts: every API result along the way
rng = random.Random(2026)
for game_no in range(12):
    g = bg.Game(game_no)
    g.setPlayers(p1, p2)
    for ply in range(400):
        d1, d2 = rng.randint(1, 6), rng.randint(1, 6)
        g.setDice(d1, d2)
        turn = g.getTurn()
        mover = g.getPlayers(turn)
        lm = [g.legalMoves(turn, d1), g.legalMoves(turn, d2)]
        seqs = g.legalTurnSequences(turn, d1, d2)
        seqs2, states = g.evaluateTurnSequences(turn, d1, d2)
        assert seqs == seqs2
        rec([game_no, ply, turn, d1, d2, lm, len(seqs), zlib.crc32(str(seqs).encode()), states.tolist()[-1] if len(seqs) else None])
        if seqs:
            pick = seqs[rng.randrange(len(seqs))]
            for o, dst in pick:
                ok, err = g.tryMove(mover, abs(o - dst), o, dst)
                assert ok, err
        rec([g.getGameBoard(), g.getJailedCount(0), g.getJailedCount(1), g.getBornOffCount(0), g.getBornOffCount(1)])
        over, winner = g.is_game_over()
        if over:
            rec(["winner", winner])
            break
        g.setTurn(1 - turn)
    if game_no == 3:
        g.printGameBoard()
json.dump(T, open(sys.argv[2], "w"))
