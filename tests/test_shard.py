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


def test_gather_two_ranks_gloo():
    n, k, world = 11, 3, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, k, ret), nprocs=world, join=True)
    want = np.stack([np.arange(n) * 10.0 + j for j in range(k)], axis=1)
    for r in range(world):
        assert np.array_equal(ret[r], want)
