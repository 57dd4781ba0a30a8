"""The exchange step inside the library (sfk_comm_init / sfk_split_finish_comm, NCCL bound
with dlopen): a halo whose particles are spread over several GPUs, Halo.Split / Halo.Join of
los/halo.go:107-140.  One rank runs on any GPU box; the 2-rank case needs two GPUs (one
process per GPU, spawned here) and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from shellfish_b200 import api, synth

pytestmark = pytest.mark.gpu


def _case():
    halos, blocks, mass = synth.make_box(31, 12.0, 2, 40000, 6000, 3, centers=[[6.0, 6.0, 6.0], [2.0, 9.0, 3.0]],
                                         radii=[1.2, 0.8])
    return halos, blocks, mass, api.default_params(seed=1337, rings=10)


def test_one_rank_communicator_matches_the_plain_path():
    """nranks = 1: the reduce-scatter / all-gather are identities, rows are not padded."""
    halos, blocks, mass, params = _case()
    ref, c_ref, s_ref = api.run_shell(params, mass, halos, blocks)
    ctx = api.ShellContext(params)
    ctx.comm_init(api.comm_unique_id(), 0, 1)
    for _ in range(2):   # the communicator outlives a snapshot
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        for b in blocks:
            ctx.push_block(b)
        c, s = ctx.split_finish_comm()
        assert np.array_equal(c, c_ref) and np.array_equal(s, s_ref)
        for h in range(len(halos)):
            assert np.array_equal(ctx.rsp(h), ref.rsp(h), equal_nan=True)
            assert ctx.intersection_count(h) == ref.intersection_count(h)
    with pytest.raises(api.ShellfishError, match="already set"):
        ctx.comm_init(api.comm_unique_id(), 0, 1)


def test_split_finish_comm_needs_a_communicator():
    halos, blocks, mass, params = _case()
    ctx = api.ShellContext(params)
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    with pytest.raises(api.ShellfishError, match="sfk_comm_init first"):
        ctx.split_finish_comm()


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from shellfish_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    halos, blocks, mass, params = _case()
    params.device = rank
    ctx = api.ShellContext(params)
    shard.comm_init(ctx, dist, "cuda")
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    for b in blocks[rank::world]:
        ctx.push_block(b)
    c, s = shard.split_halo_finish(ctx, dist, in_library=True)
    ret[rank] = (c, s, [ctx.rsp(h) for h in range(len(halos))], ctx.stats()["n_intersections"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_are_bitwise_equal_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    halos, blocks, mass, params = _case()
    ref, c_ref, s_ref = api.run_shell(params, mass, halos, blocks)
    for r in range(2):
        c, st, rsp, ni = ret[r]
        assert np.array_equal(c, c_ref) and np.array_equal(st, s_ref)
        assert ni == ref.stats()["n_intersections"]
        for h in range(len(halos)):
            assert np.array_equal(rsp[h], ref.rsp(h), equal_nan=True)
