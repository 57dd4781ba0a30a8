#!/usr/bin/env python
"""Per-kernel summary of an `ncu --page raw --csv` export (one row per captured launch).

  ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python profiles/ncu_summary.py X_raw.csv > profiles/X_summary.txt
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_inst_%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe_%"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__occupancy_limit_registers", "occ_lim_regs")]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    name = r[ki].replace("unnamed>::", "").replace("void ", "")
    print(name.split("(")[0])
    for col, label in COLS:
        if col in hdr:
            i = hdr.index(col)
            print("    %-16s %s %s" % (label, r[i], units[i]))
