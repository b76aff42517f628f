"""World-size-2 gloo test of the N>1 host logic (sharding + weight-delta all-reduce) on CPU.
The per-rank deltas come from the oracle's TD replay of the reference's golden games, so the
reduced update is checked against a single-process sum."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG, golden_weights, load_golden


def _delta_of(orc, g, gm, name):
    w0 = golden_weights(gm, "rand" if name.startswith("rand") else "trained")
    new, _ = orc.td_replay(w0, g[f"{name}.enc"], int(g[f"{name}.winner"]) == 0, 0.1, 0.9)
    d = np.zeros(25604, np.float32)
    d[:25601] = np.concatenate([np.asarray(a, np.float32).reshape(-1) - np.asarray(b, np.float32).reshape(-1) for a, b in zip(new, w0)])
    return d


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bgx.parallel import allreduce_delta, shard
    from oracle.oracle import Oracle
    g, gm = load_golden("games.npz"), load_golden("model.npz")
    first, n, stride = shard(6, rank, world)
    assert (first, n, stride) == (3 * rank, 3, 6)
    names = ["trained5", "trained6", "trained7"]
    mine = names[rank::world] if rank < len(names) else []
    orc = Oracle()
    delta = torch.zeros(25604)
    for name in mine:
        delta += torch.from_numpy(_delta_of(orc, g, gm, name))
    allreduce_delta(delta, dist)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), delta.numpy())
    dist.destroy_process_group()


def test_two_rank_delta_allreduce(tmp_path, orc):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)                       # every rank ends with the same update
    g, gm = load_golden("games.npz"), load_golden("model.npz")
    want = sum(_delta_of(orc, g, gm, n).astype(np.float64) for n in ("trained5", "trained6", "trained7"))
    assert np.max(np.abs(a - want)) <= 2e-7 * np.max(np.abs(want)) + 1e-12
    assert not a[25601:].any()


def test_shard_covers_population_once():
    from bgx.parallel import shard
    for world in (1, 2, 4, 8):
        ids = []
        for r in range(world):
            first, n, stride = shard(1 << 20, r, world)
            assert stride == 1 << 20
            ids.append((first, first + n))
        assert ids[0][0] == 0 and ids[-1][1] == 1 << 20
        assert all(ids[i][1] == ids[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard(10, 0, 4)
