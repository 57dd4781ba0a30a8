#!/usr/bin/env python
"""BASELINE config 3: one cluster-mass halo whose particles are split across
the GPUs of one box, bin grids all-reduced over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/run_split.py [--particles 10000000] [--rings 24]

Every rank builds the same synthetic snapshot and pushes its contiguous 1/N share of
every block's particle list (BASELINE config 3: "particle list split contiguous-equal");
the result is compared (bitwise) with rank 0's single-GPU run of all particles."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shellfish_b200 import api, shard, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=10_000_000)
ap.add_argument("--rings", type=int, default=24)
ap.add_argument("--cells", type=int, default=4)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
box = 60.0
halos, blocks, mass = synth.make_box(3, box, 1, args.particles, args.particles // 10, args.cells,
                                     centers=[[box / 2] * 3], radii=[2.0])
params = api.default_params(seed=1337, rings=args.rings, device=local)


def run(my_blocks, split):
    ctx = api.ShellContext(params)
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    for b in my_blocks:
        ctx.push_block(b)
    torch.cuda.synchronize()
    if split:
        dist.barrier()   # every rank has pushed its blocks: time only the exchange + analysis
    t = time.time()
    out = shard.split_halo_finish(ctx, dist if split else None)
    torch.cuda.synchronize()
    return out, time.time() - t, ctx.stats()




def share(b):
    n = len(b["ms"])
    lo, hi = n * rank // world, n * (rank + 1) // world
    return dict(b, xs=np.ascontiguousarray(b["xs"][lo:hi]), ms=np.ascontiguousarray(b["ms"][lo:hi]))


mine = [share(b) for b in blocks] if world > 1 else blocks
run(mine, world > 1)   # warm-up
(coeffs, status), dt, st = run(mine, world > 1)
if rank == 0:
    (c1, s1), dt1, st1 = run(blocks, False)
    print(json.dumps({"config": "configs[2]: one halo, %d particles, %d rings x 256 LoS x 256 bins" %
                      (args.particles, args.rings), "n_gpus": world, "split_seconds": dt,
                      "single_gpu_seconds": dt1, "bitwise_equal_to_single_gpu": bool(np.array_equal(coeffs, c1)),
                      "status": int(status[0]), "intersections_rank0": st["n_intersections"],
                      "intersections_single": st1["n_intersections"], "allreduce_bytes": args.rings * 256 * 256 * 8}))
if world > 1:
    dist.destroy_process_group()
