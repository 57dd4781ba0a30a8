"""Times the shell property integrals of `stats` (sfk_shell_props: Volume, SurfaceArea,
Axes, RadialRange for every halo) for the 1000 halos of BASELINE configs[1] through the
C ABI with host buffers (H2D of the coefficients, kernel, D2H of the sums, host closing
arithmetic), and the oracle's literal Monte-Carlo restatement (4 x 50 000 samples per
halo, cmd/stats.go:274-289) on the host cores on a bounded sample of halos.
  PYTHONPATH=. python scripts/bench_shellprops.py [--halos 1000]
"""
import argparse
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from oracle import pyoracle as po
from shellfish_b200 import api


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--halos", type=int, default=1000)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--cpu-halos", type=int, default=32)
    a = ap.parse_args()
    rng = np.random.default_rng(7)
    n = 2 * a.order * a.order
    rows = rng.normal(0, 0.03, (a.halos, n))
    rows[:, 0] = rng.uniform(0.5, 2.0, a.halos)
    rows[:, 1:] *= rows[:, :1]
    ctx = api.ShellContext(api.default_params(rings=3))
    ctx.shell_props(rows[:8])   # warm-up (builds the Gauss-Legendre rule)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record(stream)
        props = ctx.shell_props(rows)
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    nodes = api.SHELL_RULE[0] * api.SHELL_RULE[1]
    cores = os.cpu_count() or 1
    hs = min(a.halos, a.cpu_halos)

    def cpu(h):
        cs = rows[h]
        return (po.shell_volume(cs, 50000, 1), po.shell_surface_area(cs, 50000, 2), po.shell_axes(cs, 50000, 3),
                po.shell_radial_range(cs, 50000, 4))
    t0 = time.time()
    with ThreadPoolExecutor(cores) as ex:
        ref = list(ex.map(cpu, range(hs)))
    dt = time.time() - t0
    dv = max(abs(props[h, 1] - ref[h][0]) / ref[h][0] for h in range(hs))
    print(json.dumps({"pass": "stats shell properties (Volume, SurfaceArea, Axes, RadialRange)", "halos": a.halos,
                      "order": a.order, "gpu_ms": ms, "gpu_halos_per_s": a.halos / (ms * 1e-3),
                      "surface_evaluations_per_halo": 5 * nodes,
                      "gpu_surface_evaluations_per_s": 5.0 * nodes * a.halos / (ms * 1e-3),
                      "cpu_port_halos_per_s": hs / dt, "cpu_cores": cores,
                      "cpu_sample": "%d halos, 4 x 50000 Monte-Carlo samples each" % hs,
                      "max_rel_volume_difference_vs_monte_carlo": dv}))


if __name__ == "__main__":
    main()
