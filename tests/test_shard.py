"""Host-side multi-GPU logic on CPU: LPT partition of the halo list and the
gather of per-halo rows into catalog order over torch.distributed (gloo,
world_size 2)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from shellfish_b200 import shard


def test_partition_balances_and_keeps_catalog_order():
    rng = np.random.default_rng(0)
    w = rng.uniform(1, 100, 1000)
    parts = shard.partition_halos(w, 8)
    allidx = np.concatenate(parts)
    assert sorted(allidx) == list(range(1000))
    loads = np.array([w[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.02
    for p in parts:
        assert np.all(np.diff(p) > 0)


def test_partition_edge_cases():
    assert [len(p) for p in shard.partition_halos([5.0], 4)] == [1, 0, 0, 0]
    parts = shard.partition_halos([], 2)
    assert all(len(p) == 0 for p in parts)
    w = shard.estimate_weights(np.array([[0, 0, 0, 1.0], [0, 0, 0, -1.0], [0, 0, 0, 2.0]]))
    assert w[1] == 0 and abs(w[2] / w[0] - 8) < 1e-12


def test_gather_single_process():
    full = shard.gather_rows([2, 0], np.array([[1.0, 2.0], [3.0, 4.0]]), 3)
    assert np.array_equal(full, [[3, 4], [0, 0], [1, 2]])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = np.arange(1, n + 1, dtype=float)
    mine = shard.partition_halos(w, world)[rank]
    rows = np.stack([mine * 10.0 + j for j in range(k)], axis=1) if len(mine) else np.zeros((0, k))
    full = shard.gather_rows(mine, rows, n, dist=dist)
    ret[rank] = full
    dist.destroy_process_group()


def _worker_table(rank, world, port, n, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    w = rng.uniform(1, 5, n)
    groups = rng.integers(0, 4, n)
    parts = shard.partition_halos(w, world, groups)
    mine = parts[rank]
    rows = np.stack([mine * 10.0 + j for j in range(k)], axis=1) if len(mine) else np.zeros((0, k))
    ret[rank] = shard.gather_table(parts, rank, rows, n, dist=dist)
    dist.destroy_process_group()


def test_gather_table_two_ranks_gloo():
    """The bench's exchange: block-affinity partition known to every rank, one all-gather."""
    n, k, world = 13, 4, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_table, args=(world, _free_port(), n, k, ret), nprocs=world, join=True)
    want = np.stack([np.arange(n) * 10.0 + j for j in range(k)], axis=1)
    for r in range(world):
        assert np.array_equal(ret[r], want)


def test_partition_with_block_affinity():
    rng = np.random.default_rng(1)
    w = rng.uniform(0.5, 2.0, 1000)
    groups = rng.integers(0, 512, 1000)
    for world in (2, 4, 8):
        parts = shard.partition_halos(w, world, groups)
        assert sorted(np.concatenate(parts)) == list(range(1000))
        owner = np.empty(1000, int)
        for r, p in enumerate(parts):
            owner[p] = r
            assert np.all(np.diff(p) > 0)
        for g in np.unique(groups):
            assert len(set(owner[groups == g])) == 1          # a block's halos stay together
        loads = np.array([w[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.01


def test_block_of_finds_the_block_of_each_centre():
    blocks = [dict(Origin=(0.0, 0.0, 0.0), Width=(5.0, 5.0, 5.0)), dict(Origin=(5.0, 0.0, 0.0), Width=(5.0, 5.0, 5.0))]
    halos = np.array([[1.0, 1, 1, 0.5], [5.0, 4.9, 0, 0.5], [11.0, 0, 0, 0.5]])
    assert list(shard.block_of(halos, blocks)) == [0, 1, -1]


def test_gather_two_ranks_gloo():
    n, k, world = 11, 3, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, k, ret), nprocs=world, join=True)
    want = np.stack([np.arange(n) * 10.0 + j for j in range(k)], axis=1)
    for r in range(world):
        assert np.array_equal(ret[r], want)
