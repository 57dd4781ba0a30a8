#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --clock-control none --csv`):

  python profiles/launch_summary.py profiles/rXX_launches_bench64.csv [bench64.json] > profiles/rXX_launch_summary.txt

With the bench line of the same command, the kernel shares are printed next to bench.py's own
stage shares (CUDA events): per-launch times under ncu are cold-cache and serialised, so the
SHARES must agree, not the absolute times."""
import csv
import json
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[kn])
    t = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[mu]]
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
allt = sum(v[1] for v in tot.values())
print("ncu launch list of `bench.py --halos 64 ...` (gpu__time_duration.sum, --clock-control none): per-kernel totals\n")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %5d launches %12.3f ms %5.1f%%" % (name[-40:], n, t, 100 * t / allt))
if len(sys.argv) > 2:
    line = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    st = line["stage_ms_per_step"]
    s = sum(st.values())
    print("\nbench.py's own stage times of the same command (CUDA events, ms per step):")
    for k, v in st.items():
        print("  %-10s %9.3f ms %5.1f%%" % (k, v, 100 * v / s))
