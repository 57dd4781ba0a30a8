#!/usr/bin/env python
"""Benchmark of the `shellfish shell` line-of-sight hot path on B200.

One "step" = one full pass of loop() (cmd/shell.go:288-372 minus buf.Read) over
one synthetic snapshot: createHalos, every block pushed, Halo.Insert for every
surviving particle, haloAnalysis, Penna coefficients back on the host.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): BASELINE.json configs[1] -- 1000 halos (>= 5e4 particles each)
in a synthetic 512^3-particle LGadget-2-like box of 8^3 blocks, default
shell.config (100 rings x 256 LoS x 256 bins).  Under torchrun each rank runs
its own box (weak scaling; halos are independent, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from shellfish_b200 import synth  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--halos", type=int, default=1000)
    ap.add_argument("--n-total", type=int, default=512 ** 3)
    ap.add_argument("--box", type=float, default=250.0)
    ap.add_argument("--cells", type=int, default=8)
    ap.add_argument("--cpu-halos", type=int, default=6, help="halos in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--order", type=int, default=3, help="Penna order (BASELINE config 5 sweeps 2, 3, 4)")
    ap.add_argument("--subsample", type=int, default=1, help="SubsampleFactor (config 5: 1, 2)")
    ap.add_argument("--los-moments", action="store_true",
                    help="A/B switch: SFK_FLAG_LOS_MOMENTS (experimental sliding-moment Savitzky-Golay)")
    return ap.parse_args()


def workload(args, rank):
    t = time.time()
    halos, blocks, mass = synth.config2(seed=2 + 1000 * rank, n_halos=args.halos, n_total=args.n_total,
                                        box=args.box, cells=args.cells)
    desc = {"workload": "configs[1]: %d halos in a synthetic %d-particle LGadget-2-like box (%g Mpc/h, %d^3 blocks), "
                        "default shell.config 100 rings x 256 LoS x 256 bins" %
                        (args.halos, args.n_total, args.box, args.cells),
            "halos": args.halos, "n_particles": int(args.n_total), "blocks": len(blocks),
            "order": args.order, "subsample": args.subsample, "setup_s": round(time.time() - t, 1)}
    if args.los_moments:
        desc["los_moments"] = True   # experimental kernel variant: not the default product path
    return halos, blocks, mass, desc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(args, halos, blocks, mass, n_halos, threads):
    """Oracle (C restatement of the Go algorithm, strided workers + Split/Join
    like cmd/shell.go:423-500) on the first n_halos halos of the workload."""
    from oracle import pyoracle as po
    cfg = po.ShellConfig.default(threads=threads, seed=1337, order=args.order, subsampleFactor=args.subsample)
    sub = halos[:n_halos]
    t = time.time()
    out = po.shell_run(cfg, mass, sub, blocks, taps=("nIntersections",))
    dt = time.time() - t
    return dt, int(out["nIntersections"].sum()), int(out["ok"].sum())


def run_reference(args):
    """Reference arm: the reference's CPU algorithm on the host cores.  The Go
    binary cannot be built (no Go toolchain, un-vendored modules), so this is
    the oracle port with all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    halos, blocks, mass, desc = workload(args, 0)
    n = max(1, min(args.cpu_halos, len(halos)))
    for _ in range(args.warmup):
        cpu_sample(args, halos, blocks, mass, min(n, 2), cores)
    tot, inter = 0.0, 0
    for _ in range(args.steps):
        dt, ni, _ = cpu_sample(args, halos, blocks, mass, n, cores)
        tot += dt
        inter += ni
    value = n * args.steps / tot
    sample = "first %d halos of the workload per step, %d threads (GOMAXPROCS analogue)" % (n, cores)
    line = {"impl": "reference", "metric": "halos_per_s", "value": value, "unit": "halos/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": desc,
            "intersections_per_s": inter / tot,
            "cpu_baseline": {"value": value, "unit": "halos/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "halos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of one bin_kernel launch over its algorithmic bytes
NCU_TRAFFIC_RATIO = (2.206682 + 0.804353) / 1.704478


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from shellfish_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    halos, blocks, mass, desc = workload(args, rank)
    params = api.default_params(seed=1337, device=local, order=args.order, subsample_factor=args.subsample,
                                flags=api.SFK_FLAG_LOS_MOMENTS if args.los_moments else 0)
    ctx = api.ShellContext(params)
    R, n, B = int(params.rings), int(params.spokes), int(params.radial_bins)

    dev_blocks = [(torch.from_numpy(b["xs"]).cuda(), torch.from_numpy(b["ms"]).cuda()) for b in blocks]
    pin_blocks = []
    if not args.no_e2e:
        for b in blocks:
            pb = dict(b)
            pb["xs"] = torch.from_numpy(b["xs"]).pin_memory()
            pb["ms"] = torch.from_numpy(b["ms"]).pin_memory()
            pin_blocks.append(pb)
    torch.cuda.synchronize()

    def step(resident):
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        if resident:
            for b, (dx, dm) in zip(blocks, dev_blocks):
                ctx.push_block_device(b, dx, dm)
        else:
            for pb in pin_blocks:
                ctx.push_block(pb)
        coeffs, status = ctx.finish_snapshot()
        return coeffs, status, ctx.stats()

    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        agg = {}
        nok = 0
        for _ in range(steps):
            _, status, st = step(resident)
            nok += int((status == 0).sum())
            for k, v in st.items():
                agg[k] = agg.get(k, 0) + v
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, agg, nok

    for _ in range(args.warmup):
        step(True)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, agg, nok = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None
    e2e = None
    if not args.no_e2e:
        step(False)
        ms_e, agg_e, _ = timed(False, args.steps)
        h2d = sum(int(pb["xs"].numel() * 4 + pb["ms"].numel() * 4) for pb in pin_blocks)
        d2h = int(len(halos) * (ctx.P2 * 8 + 4 + 8 + 8 + 8 + R * 4))
        e2e = {"value": world * len(halos) * args.steps / (ms_e * 1e-3), "unit": "halos/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e / args.steps}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # algorithmic bytes of the binning kernel: 16*N_i + 2*8*R*n*B per halo
        alg_bytes = 16.0 * agg["n_candidates"] + 16.0 * R * n * B * agg["n_halos"]
        launches = max(1, agg["bin_launches"])
        bin_s = agg["ms_bin"] * 1e-3
        achieved = alg_bytes / bin_s / 1e9 if bin_s > 0 else 0.0
        line = {
            "metric": "halos_per_s", "value": world * len(halos) * args.steps / (ms * 1e-3),
            "unit": "halos/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(desc, l2="inputs larger than L2 (particle blocks %.1f GB, hit records > 100 MB per halo)"
                                    % (sum(b["xs"].nbytes + b["ms"].nbytes for b in blocks) / 1e9),
                           parallelism="halos sharded per GPU, no collective"),
            "halos_ok": nok // max(1, args.steps),
            "intersections_per_s": world * agg["n_intersections"] / (ms * 1e-3),
            "stage_ms_per_step": {k: agg[k] / args.steps for k in
                                  ("ms_cull", "ms_hits", "ms_bin", "ms_los", "ms_filter", "ms_fit")},
            "gpu_launches": int(agg["launches"]),
            "roofline": {"bound": "hbm", "kernel": "bin_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         # DRAM bytes of one launch = NCU_TRAFFIC_RATIO x its algorithmic bytes (ncu --set
                         # full capture of a 16-halo launch: dram read 2.207 GB + write 0.804 GB against
                         # 1.704 GB algorithmic, profiles/r01g_bin_kernel_v6_by_line.txt)
                         "traffic": NCU_TRAFFIC_RATIO * alg_bytes / launches / 1e9,
                         "traffic_unit": "GB per launch (ncu ratio x algorithmic bytes)",
                         "peak_source": peak_src,
                         "alg_bytes_per_launch": alg_bytes / launches,
                         "ms_per_launch": agg["ms_bin"] / launches,
                         "note": "bin_kernel is instruction-issue bound (FP64 chord geometry + shared-memory "
                                 "atomics), not HBM bound: ncu issue-active 73%, FP64 pipe 40%, DRAM 3% "
                                 "(profiles/r01g_bin_kernel_v6_by_line.txt, DESIGN.md)"},
            "clocks": clocks, "e2e": e2e,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            nh = max(1, min(args.cpu_halos, len(halos)))
            dt, ni, okc = cpu_sample(args, halos, blocks, mass, nh, cores)
            line["cpu_baseline"] = {"value": nh / dt, "unit": "halos/s", "cores": cores, "kind": "port",
                                    "sample": "first %d halos of the same workload, one pass, %d worker threads "
                                              "(C restatement of the Go algorithm; the Go binary cannot be built here)"
                                              % (nh, cores),
                                    "seconds": dt, "intersections_per_s": ni / dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
