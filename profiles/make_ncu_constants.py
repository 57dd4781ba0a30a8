#!/usr/bin/env python
"""Writes profiles/ncu_bin_constants.json from an `ncu --set full` capture of ONE bin_kernel launch
of the 64-halo bench (one launch per step) and that bench's own JSON line:

  python profiles/make_ncu_constants.py profiles/rXX_bin_kernel_by_line.txt gpurun_out/rXX_bench64.json

bench.py scales the capture's DRAM bytes and warp instructions to its own launches
(`roofline.traffic`, `roofline.issue_frac`) and cites the capture in the line."""
import json
import os
import re
import sys

by_line, bench = sys.argv[1], sys.argv[2]
txt = open(by_line).read()


def metric(name):
    m = re.search(re.escape(name) + r" \[(\w*)\] = ([0-9.]+)", txt)
    v = float(m.group(2))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "inst": 1.0, "": 1.0}[m.group(1)]


line = json.loads([l for l in open(bench) if l.startswith("{")][-1])
launches = line["roofline"]["alg_bytes_per_launch"]
inter_per_step = line["intersections_per_s"] * line["ms_per_step"] * 1e-3
out = {"capture": by_line,
       "dram_bytes": metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum"),
       "alg_bytes": launches,
       "warp_instructions": metric("smsp__inst_executed.sum"),
       "intersections": inter_per_step,
       "note": "one bin_kernel launch of `bench.py --halos 64 --n-total 16000000 --box 125 --cells 4` (one launch per step)"}
out["traffic_ratio"] = out["dram_bytes"] / out["alg_bytes"]
out["warp_inst_per_intersection"] = out["warp_instructions"] / out["intersections"]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_bin_constants.json"), "w"), indent=1)
print(json.dumps(out))
