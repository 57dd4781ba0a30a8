#!/usr/bin/env python
"""Table of a scripts/sweep_config5.sh run: python scripts/summarize_sweep.py sweep.jsonl"""
import json
import sys

print("BASELINE configs[4] (Order x SubsampleFactor) on a 64-halo sample of configs[1]'s box, 1 x B200")
print("halos/s with blocks resident in HBM; stage times in ms per step; parity = GPU vs oracle on 2 halos")
print("order sub  halos/s  ok   cull  hits    bin    los filter   fit  intersections_equal  max_rel_coeff  pass")
for l in open(sys.argv[1]):
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    s, p = d["stage_ms_per_step"], d.get("parity", {})
    print("%-5d %-4d %7.1f  %-3d %5.2f %5.2f %6.2f %6.2f %6.2f %5.2f  %-19s  %-13.3g  %s" % (
        d["config"]["order"], d["config"]["subsample"], d["value"], d["halos_ok"], s["ms_cull"], s["ms_hits"],
        s["ms_bin"], s["ms_los"], s["ms_filter"], s["ms_fit"], p.get("intersections_equal"),
        p.get("max_rel_coeff", float("nan")), p.get("pass")))
