#!/usr/bin/env python
"""Strong scaling of ONE snapshot: the halo list of BASELINE configs[1]'s box is partitioned over
the GPUs by estimated particle count (longest-processing-time greedy, shard.partition_halos), every
rank scans all particle blocks for its own halos (sfk_set_owned keeps the others in Transform's
cumulative wrap), and the coefficient rows are gathered by halo index.  No data-path collective.
This is the LPT scheme of BASELINE config 4 on config 2's box.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29512 scripts/run_sharded.py [--halos 1000 --n-total 134217728]

Rank 0 also runs the whole list alone and checks that the gathered table is bitwise equal."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shellfish_b200 import api, shard, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--halos", type=int, default=1000)
ap.add_argument("--n-total", type=int, default=512 ** 3)
ap.add_argument("--box", type=float, default=250.0)
ap.add_argument("--cells", type=int, default=8)
ap.add_argument("--no-check", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
halos, blocks, mass = synth.config2(seed=2, n_halos=args.halos, n_total=args.n_total, box=args.box, cells=args.cells)
dev = [(torch.from_numpy(b["xs"]).cuda(), torch.from_numpy(b["ms"]).cuda()) for b in blocks]
parts = shard.partition_halos(shard.estimate_weights(halos), world)
params = api.default_params(seed=1337, device=local)


def run(owned_idx, together=True):
    """together: every rank is in this call (the single-GPU check on rank 0 is not)."""
    ctx = api.ShellContext(params)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1 and together:
        dist.barrier()
    e0.record(stream)
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    if owned_idx is not None:
        owned = np.zeros(len(halos), bool)
        owned[owned_idx] = True
        ctx.set_owned(owned)
    for b, (dx, dm) in zip(blocks, dev):
        ctx.push_block_device(b, dx, dm)
    coeffs, status = ctx.finish_snapshot()
    e1.record(stream)
    torch.cuda.synchronize()
    return coeffs, status, e0.elapsed_time(e1), ctx.stats()


run(parts[rank] if world > 1 else None)   # warm-up
coeffs, status, ms, st = run(parts[rank] if world > 1 else None)
t = torch.tensor([ms], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
table = shard.gather_rows(parts[rank], coeffs[parts[rank]], len(halos), dist if world > 1 else None, device="cuda") \
    if world > 1 else coeffs
if rank == 0:
    out = {"config": "configs[1] box, %d halos, halo list LPT-partitioned over the GPUs (strong scaling)" % args.halos,
           "n_gpus": world, "ms": float(t.item()), "halos_per_s": args.halos / (float(t.item()) * 1e-3),
           "rank0_halos": int(len(parts[0])), "rank0_ms_cull": st["ms_cull"], "rank0_ms_bin": st["ms_bin"]}
    if world > 1 and not args.no_check:
        c1, s1, ms1, _ = run(None, together=False)
        out["single_gpu_ms"] = ms1
        out["bitwise_equal_to_single_gpu"] = bool(np.array_equal(table, c1))
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
