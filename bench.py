#!/usr/bin/env python
"""Benchmark of the `shellfish shell` line-of-sight hot path on B200.

One "step" = one full pass of loop() (cmd/shell.go:288-372 minus buf.Read) over
one synthetic snapshot: createHalos, every needed block pushed, Halo.Insert for
every surviving particle, haloAnalysis, Penna coefficients back on the host (and,
with several GPUs, the catalog rows gathered in halo order).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: BASELINE.json configs[1] -- 1000 halos (>= 5e4 particles each) in a
synthetic 512^3-particle LGadget-2-like box of 8^3 blocks, default shell.config
(100 rings x 256 LoS x 256 bins).  Under torchrun every rank sees the SAME
snapshot; the halo list is partitioned across the ranks by estimated particle
count (north_star; shard.partition_spatial: equal-weight runs along a Z-curve), a rank pushes
only the blocks its halos touch, and there is no data-path collective: strong
scaling of one job.
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

from shellfish_b200 import shard, synth  # noqa: E402


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
    ap.add_argument("--cpu-halos", type=int, default=6, help="halos in the bounded CPU sample (N=1)")
    ap.add_argument("--parity-halos", type=int, default=2, help="halos checked against the oracle when N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sync-push", action="store_true",
                    help="e2e: sfk_push_block (returns once the block is copied) instead of sfk_push_block_async")
    ap.add_argument("--order", type=int, default=3, help="Penna order (BASELINE config 5 sweeps 2, 3, 4)")
    ap.add_argument("--subsample", type=int, default=1, help="SubsampleFactor (config 5: 1, 2)")
    ap.add_argument("--los-taploop", action="store_true",
                    help="A/B switch: SFK_FLAG_LOS_TAPLOOP (literal Savitzky-Golay tap loop instead of sliding moments)")
    return ap.parse_args()


def workload(args):
    """The same snapshot on every rank and in both arms; `config` is identical in both."""
    t = time.time()
    halos, blocks, mass = synth.config2(seed=2, n_halos=args.halos, n_total=args.n_total,
                                        box=args.box, cells=args.cells)
    config = {"workload": "configs[1]: %d halos in a synthetic %d-particle LGadget-2-like box (%g Mpc/h, %d^3 blocks), "
                          "default shell.config 100 rings x 256 LoS x 256 bins" %
                          (args.halos, args.n_total, args.box, args.cells),
              "halos": args.halos, "n_particles": int(args.n_total), "blocks": len(blocks),
              "order": args.order, "subsample": args.subsample,
              "l2": "inputs larger than L2: %.1f GB of particle blocks and > 100 MB of hit records per halo each step"
                    % (16.0 * args.n_total / 1e9)}
    if args.los_taploop:
        config["los_taploop"] = True   # A/B variant: not the default product path
    meta = {"setup_s": round(time.time() - t, 1)}
    return halos, blocks, mass, config, meta


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
    like cmd/shell.go:423-500) on the first n_halos halos of the workload.  A
    prefix of the halo list sees the same cumulative Transform wraps as in the
    full list, so its rows are comparable with the GPU's rows for those halos."""
    from oracle import pyoracle as po
    cfg = po.ShellConfig.default(threads=threads, seed=1337, order=args.order, subsampleFactor=args.subsample)
    sub = halos[:n_halos]
    t = time.time()
    out = po.shell_run(cfg, mass, sub, blocks, taps=("nIntersections",))
    dt = time.time() - t
    return dt, out


def parity_of(out, table):
    """GPU rows (coeffs..., status, n_intersections per halo) against the oracle's
    for the sampled halos: bit-exact intersection counts, coefficients within 1e-5
    of the halo's largest coefficient (north_star)."""
    nh = len(out["ok"])
    P2 = out["coeffs"].shape[1]
    rel, eq, ok_eq = 0.0, True, True
    for h in range(nh):
        eq &= int(out["nIntersections"][h]) == int(table[h, P2 + 1])
        ok_eq &= bool(out["ok"][h]) == (int(table[h, P2]) == 0)
        if out["ok"][h]:
            scale = float(np.max(np.abs(out["coeffs"][h])))
            rel = max(rel, float(np.max(np.abs(table[h, :P2] - out["coeffs"][h]))) / scale)
    return {"n_halos": nh, "max_rel_coeff": rel, "intersections_equal": bool(eq), "status_equal": bool(ok_eq),
            "tolerance": 1e-5, "pass": bool(eq and ok_eq and rel <= 1e-5)}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm on the host cores.  The Go
    binary cannot be built (no Go toolchain, un-vendored modules), so this is
    the oracle port with all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    halos, blocks, mass, config, meta = workload(args)
    # bounded sample: the first n halos per step, n chosen so that warm-up + timed steps take about
    # three minutes (one halo of this workload costs the port ~3.6 s on 16 cores); halos are
    # independent, so halos/s does not depend on n
    est = 3.6 * 16.0 / max(cores, 1)
    n = int(180.0 / (max(1, args.steps + args.warmup) * est))
    n = max(1, min(n, args.cpu_halos, len(halos)))
    for _ in range(args.warmup):
        cpu_sample(args, halos, blocks, mass, n, cores)
    tot, inter = 0.0, 0
    for _ in range(args.steps):
        dt, out = cpu_sample(args, halos, blocks, mass, n, cores)
        tot += dt
        inter += int(out["nIntersections"].sum())
    value = n * args.steps / tot
    sample = ("first %d halos of the workload per step (halos/s extrapolates linearly: halos are independent), "
              "%d threads = all host cores (GOMAXPROCS analogue)" % (n, cores))
    line = {"impl": "reference", "metric": "halos_per_s", "value": value, "unit": "halos/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "meta": meta,
            "intersections_per_s": inter / tot,
            "cpu_baseline": {"value": value, "unit": "halos/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "halos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ncu --set full capture of one bin_kernel launch (profiles/ncu_bin_constants.json, written by
# profiles/make_ncu_constants.py): dram__bytes_read.sum + dram__bytes_write.sum over the launch's
# algorithmic bytes, and warp instructions executed per ProfileRing.Insert
_NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_bin_constants.json")))
NCU_CAPTURE = _NCU["capture"]
NCU_TRAFFIC_RATIO = _NCU["traffic_ratio"]
NCU_WARP_INST_PER_INTERSECTION = _NCU["warp_inst_per_intersection"]


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

    halos, blocks, mass, config, meta = workload(args)
    nH = len(halos)
    params = api.default_params(seed=1337, device=local, order=args.order, subsample_factor=args.subsample,
                                flags=api.SFK_FLAG_LOS_TAPLOOP if args.los_taploop else 0)
    ctx = api.ShellContext(params)
    R, n, B = int(params.rings), int(params.spokes), int(params.radial_bins)
    P2 = ctx.P2

    # halo partition by estimated particle count (R200m^3) plus a per-halo constant for the
    # analysis stages, halos of one block kept together; every rank computes the same partition
    r3 = np.maximum(halos[:, 3], 0.0) ** 3
    weights = r3 / max(r3.mean(), 1e-300) + 0.25
    parts = shard.partition_spatial(weights, halos[:, :3], blocks[0]["TotalWidth"], world)
    mine = parts[rank]
    owned = np.zeros(nH, bool)
    owned[mine] = True
    ctx.begin_snapshot(blocks[0], mass)
    ctx.add_halos(halos)
    need = shard.blocks_needed(ctx, blocks, owned)
    my_blocks = [blocks[i] for i in need]

    # LGadget-2 snapshots carry one particle mass (io/lgadget2.go:228 fills ms with buf.MinMass()):
    # such blocks are pushed without a mass array, as the CLI does for LGadget-2 and gotetra files
    uniform = all(bool(np.all(b["ms"] == np.float32(mass))) for b in my_blocks)
    dev_blocks = [(torch.from_numpy(b["xs"]).cuda(), None if uniform else torch.from_numpy(b["ms"]).cuda())
                  for b in my_blocks]
    pin_blocks = []
    if not args.no_e2e:
        for b in my_blocks:
            pb = dict(b)
            pb["xs"] = torch.from_numpy(b["xs"]).pin_memory()
            pb["ms"] = None if uniform else torch.from_numpy(b["ms"]).pin_memory()
            pin_blocks.append(pb)
    torch.cuda.synchronize()

    n_max = max(len(p) for p in parts)
    cols = P2 + 2   # coefficients, status, intersections of the halo

    def step(resident):
        ctx.begin_snapshot(blocks[0], mass)
        ctx.add_halos(halos)
        if world > 1:
            ctx.set_owned(owned)
        if resident:
            for b, (dx, dm) in zip(my_blocks, dev_blocks):
                ctx.push_block_device(b, dx, dm)
        else:
            for pb in pin_blocks:
                ctx.push_block(pb, wait=args.sync_push)
        coeffs, status = ctx.finish_snapshot()
        rows = np.zeros((len(mine), cols))
        rows[:, :P2] = coeffs[mine]
        rows[:, P2] = status[mine]
        rows[:, P2 + 1] = [ctx.intersection_count(int(h)) for h in mine]
        # the catalog rows, gathered in halo order (the only exchange of the job)
        table = shard.gather_table(parts, rank, rows, nH, dist if world > 1 else None, device="cuda")
        return table, ctx.stats()

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
        table = None
        for _ in range(steps):
            table, st = step(resident)
            nok += int((table[:, P2] == 0).sum())
            for k, v in st.items():
                agg[k] = agg.get(k, 0) + v
        stream.wait_stream(torch.cuda.current_stream())   # the row gather ran on torch's stream
        e1.record(stream)
        barrier()
        ms_own = e0.elapsed_time(e1)
        ms = ms_own
        per_rank = [ms_own]
        if world > 1:
            t = torch.tensor([ms_own], device="cuda", dtype=torch.float64)
            allt = torch.empty(world, device="cuda", dtype=torch.float64)
            dist.all_gather_into_tensor(allt, t)
            per_rank = [float(v) for v in allt.cpu()]
            ms = max(per_rank)
        return ms, agg, nok, table, per_rank

    def sum_over_ranks(agg, keys):
        if world == 1:
            return {k: agg[k] for k in keys}
        t = torch.tensor([float(agg[k]) for k in keys], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {k: float(v) for k, v in zip(keys, t.cpu())}

    for _ in range(args.warmup):
        step(True)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, agg, nok, table, per_rank = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None
    work_keys = ("n_candidates", "n_intersections", "n_ring_hits", "launches", "bin_launches", "ms_bin", "n_los_literal")
    tot = sum_over_ranks(agg, work_keys)
    stage_keys = ("ms_cull", "ms_hits", "ms_bin", "ms_los", "ms_filter", "ms_fit")
    if world > 1:
        t = torch.tensor([agg[k] for k in stage_keys], device="cuda", dtype=torch.float64)
        allst = torch.empty((world, len(stage_keys)), device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allst, t)
        stage_by_rank = (allst.cpu().numpy() / args.steps).tolist()
    else:
        stage_by_rank = [[agg[k] / args.steps for k in stage_keys]]
    blocks_by_rank = [len(my_blocks)]
    if world > 1:
        t = torch.tensor([float(len(my_blocks))], device="cuda", dtype=torch.float64)
        allb = torch.empty(world, device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allb, t)
        blocks_by_rank = [int(v) for v in allb.cpu()]
    e2e = None
    if not args.no_e2e:
        step(False)
        ms_e, agg_e, _, _, _ = timed(False, args.steps)
        h2d_own = sum(int(pb["xs"].numel() * 4 + (pb["ms"].numel() * 4 if pb["ms"] is not None else 0)) for pb in pin_blocks)
        d2h_own = int(nH * (P2 * 8 + 4 + 8 + 8 + 8 + R * 4) + (0 if world == 1 else world * n_max * cols * 8))
        if world > 1:
            t = torch.tensor([h2d_own, d2h_own], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        else:
            h2d, d2h = h2d_own, d2h_own
        e2e = {"value": nH * args.steps / (ms_e * 1e-3), "unit": "halos/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / args.steps}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # algorithmic bytes of the binning kernel: 16*N_i + 2*8*R*n*B per halo;
        # achieved = bytes of all launches / their summed duration (per-GPU rate when N > 1)
        alg_bytes = 16.0 * tot["n_candidates"] + 16.0 * R * n * B * nH * args.steps
        launches = max(1, tot["bin_launches"])
        bin_s = tot["ms_bin"] * 1e-3
        achieved = alg_bytes / bin_s / 1e9 if bin_s > 0 else 0.0
        sm_clock = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        issue_peak = 148 * 4 * sm_clock * 1e6
        line = {
            "metric": "halos_per_s", "value": nH * args.steps / (ms * 1e-3),
            "unit": "halos/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "meta": dict(meta, parallelism="halo list partitioned over %d GPU(s) by estimated particle count "
                                          "(equal-weight runs along a Z-curve); no data-path collective; catalog rows gathered by "
                                          "halo index" % world,
                         masses="uniform: blocks pushed without a mass array" if uniform else "per-particle mass arrays",
                         e2e_push=("sfk_push_block (returns once the block is copied)" if args.sync_push else
                                   "sfk_push_block_async from page-locked buffers that hold the snapshot until "
                                   "sfk_finish_snapshot (copies queued back to back)")),
            "halos_ok": nok // max(1, args.steps),
            "intersections_per_s": tot["n_intersections"] / (ms * 1e-3),
            "stage_ms_per_step": dict(zip(stage_keys, stage_by_rank[int(np.argmax(per_rank))])),
            "gpu_launches": int(tot["launches"]),
            # lines of sight with long plateaus, smoothed by the literal tap loop (exact ties), per step
            "los_literal_per_step": tot["n_los_literal"] / args.steps,
            "roofline": {"bound": "hbm", "kernel": "bin_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_RATIO * alg_bytes / launches / 1e9,
                         "traffic_unit": "GB per launch = ncu DRAM bytes / algorithmic bytes of the captured launch "
                                         "(%s) x this run's algorithmic bytes per launch" % NCU_CAPTURE,
                         "peak_source": peak_src,
                         "alg_bytes_per_launch": alg_bytes / launches,
                         "ms_per_launch": tot["ms_bin"] / launches,
                         # the bound that does apply: warp instructions issued per second over
                         # SMs x 4 schedulers x SM clock (instructions per intersection from the same capture)
                         "issue_frac": NCU_WARP_INST_PER_INTERSECTION * tot["n_intersections"] / bin_s / issue_peak
                                       if bin_s > 0 else None,
                         "note": "bin_kernel is instruction-issue bound (FP64 chord geometry + shared-memory "
                                 "atomics), not HBM bound (DESIGN.md section 4); issue_frac is the relevant fraction"},
            "clocks": clocks, "e2e": e2e,
        }
        if world > 1:
            mean = float(np.mean(per_rank))
            stage_tot = [sum(s) for s in stage_by_rank]
            line["partition"] = {"ms_by_rank": [v / args.steps for v in per_rank],
                                 "imbalance": max(per_rank) / mean if mean > 0 else None,
                                 "halos_by_rank": [len(p) for p in parts],
                                 "blocks_by_rank": blocks_by_rank,
                                 "kernel_ms_by_rank": stage_tot,
                                 "cull_ms_by_rank": [s[0] for s in stage_by_rank],
                                 "residual": "max over ranks of (kernel time imbalance from the LPT estimate, "
                                             "cull_kernel over the rank's blocks, host orchestration)"}
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            nh = max(1, min(args.cpu_halos if world == 1 else args.parity_halos, nH))
            dt, out = cpu_sample(args, halos, blocks, mass, nh, cores)
            line["parity"] = parity_of(out, table)
            if world == 1:
                line["cpu_baseline"] = {"value": nh / dt, "unit": "halos/s", "cores": cores, "kind": "port",
                                        "sample": "first %d halos of the same workload, one pass, %d worker threads "
                                                  "(C restatement of the Go algorithm; the Go binary cannot be built here)"
                                                  % (nh, cores),
                                        "seconds": dt,
                                        "intersections_per_s": float(out["nIntersections"].sum()) / dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
