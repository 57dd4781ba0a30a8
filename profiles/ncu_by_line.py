#!/usr/bin/env python
"""Attributes the per-SASS-instruction counters of an ncu report to CUDA source
lines.  ncu's CSV source page lists SASS only, so the line table comes from
`nvdisasm -g` on the cubin embedded in libsfk.so (same instruction order).

  python profiles/ncu_by_line.py gpurun_out/prof.ncu-rep sfk_bin bin_kernelILb0 [top]

For a report holding several kernels set NCU_KERNEL (a name regex; the first matching launch
is used).  SFK_LIB / SFK_SRC point at the libsfk.so and csrc/ the report was captured from when
the tree has moved on since.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sh(cmd, **kw):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True, **kw).stdout


def main():
    rep, unit, func = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sel = ""
    if os.environ.get("NCU_KERNEL"):
        sel = " -k regex:%s -c 1" % os.environ["NCU_KERNEL"]
    raw = list(csv.reader(io.StringIO(sh("ncu -i %s%s --page raw --csv" % (rep, sel)))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    keys = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.per_cycle_active",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum"]
    keys += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            print("%s [%s] = %s" % (h, u, v))
    src = list(csv.reader(io.StringIO(sh("ncu -i %s%s --page source --csv" % (rep, sel)))))
    h2, data = src[1], src[2:]
    for i, r in enumerate(data):   # several launches in the report: keep the first
        if r and r[0] == "Kernel Name":
            data = data[:i]
            break
    iI, iS, iSm = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples")
    iT = h2.index("Thread Instructions Executed")
    with tempfile.TemporaryDirectory() as td:
        sh("cuobjdump -xelf all %s" % os.environ.get("SFK_LIB", os.path.join(ROOT, "shellfish_b200", "libsfk.so")), cwd=td)
        sass = sh("nvdisasm -g -c %s.sm_100a.cubin" % unit, cwd=td).split("\n")
    start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and func in l)
    cur, seq = None, []
    for l in sass[start + 1:]:
        if (l.startswith(".text.") or l.startswith(".section")) and seq:
            break
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith(unit + ".cu") else -1
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((cur, m.group(2)))
    assert len(seq) == len(data), (len(seq), len(data))
    per, smp, mix, thr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for (ln, _), r in zip(seq, data):
        per[ln] += int(r[iI])
        thr[ln] += int(r[iT])
        smp[ln] += int(r[iSm])
        t = r[iS].strip().split()
        mix[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[iI])
    tot = sum(per.values())
    print("\nSASS mix (warp instructions):")
    for k, v in mix.most_common(16):
        print("  %-10s %14d %5.1f%%" % (k, v, 100.0 * v / tot))
    text = open(os.path.join(os.environ.get("SFK_SRC", os.path.join(ROOT, "shellfish_b200", "csrc")), unit + ".cu")).read().split("\n")
    print("\nby source line (warp instructions, share, stall samples, mean active lanes):")
    for ln, v in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
        t = text[ln - 1].strip()[:88] if ln and ln > 0 else "<inlined from a header>"
        print("%5s %9.1fM %5.1f%% %7d %5.1f | %s" % (ln, v / 1e6, 100.0 * v / tot, smp[ln], thr[ln] / max(v, 1), t))


if __name__ == "__main__":
    main()
