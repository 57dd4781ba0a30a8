"""Times the `stats` M_sp pass (sfk_mass_*) on BASELINE configs[1]'s box:
every particle against every halo, blocks resident in HBM, CUDA events on the
library's stream; the oracle port runs a bounded sample on the host cores.
  PYTHONPATH=. python scripts/bench_mass.py [--halos 1000 --n-total 134217728]
"""
import argparse
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from oracle import pyoracle as po
from shellfish_b200 import api, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--halos", type=int, default=1000)
    ap.add_argument("--n-total", type=int, default=512 ** 3 // 1)
    ap.add_argument("--box", type=float, default=250.0)
    ap.add_argument("--cells", type=int, default=8)
    ap.add_argument("--cpu-pairs", type=float, default=2e9)
    a = ap.parse_args()
    halos, blocks, mass = synth.config2(seed=2, n_halos=a.halos, n_total=a.n_total, box=a.box, cells=a.cells)
    nh = len(halos)
    coeffs = np.zeros((nh, 18))
    rng = np.random.default_rng(7)
    coeffs[:, 0] = 1.3 * halos[:, 3]                      # R_sp ~ 1.3 R200m
    coeffs[:, 1:] = rng.normal(0, 0.03, (nh, 17)) * halos[:, 3:4]
    lo, hi = 0.8 * coeffs[:, 0], 1.25 * coeffs[:, 0]
    ctx = api.ShellContext(api.default_params(rings=3))
    dev = [(torch.from_numpy(b["xs"]).cuda(), torch.from_numpy(b["ms"]).cuda()) for b in blocks]
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    ctx.mass_contained(halos[:, :3], coeffs, lo, hi, blocks[:2], device_blocks=dev[:2])   # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    m = ctx.mass_contained(halos[:, :3], coeffs, lo, hi, blocks, device_blocks=dev)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n = sum(len(b["ms"]) for b in blocks)
    # CPU port on a bounded sample: first blocks x first halos, one thread per halo
    cores = os.cpu_count() or 1
    nb = 1
    hs = max(1, min(nh, int(a.cpu_pairs / max(1, len(blocks[0]["ms"])))))
    t0 = time.time()
    with ThreadPoolExecutor(cores) as ex:
        ref = list(ex.map(lambda h: po.mass_contained(blocks[0]["TotalWidth"], blocks[0]["xs"], blocks[0]["ms"], coeffs[h],
                                                      halos[h, :3].astype(np.float32), lo[h], hi[h]), range(hs)))
    dt = time.time() - t0
    pairs_cpu = hs * len(blocks[0]["ms"])
    print(json.dumps({"pass": "stats M_sp (massContained)", "halos": nh, "particles": n, "gpu_ms": ms,
                      "gpu_pairs_per_s": nh * n / (ms * 1e-3), "gpu_particle_bytes_per_s": 16.0 * n / (ms * 1e-3),
                      "cpu_port_pairs_per_s": pairs_cpu / dt, "cpu_cores": cores,
                      "cpu_sample": "%d halos x block 0 (%d particles)" % (hs, len(blocks[0]["ms"])),
                      "mass_sum": float(m.sum())}))


if __name__ == "__main__":
    main()
