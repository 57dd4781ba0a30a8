"""Regenerates tests/golden/stats_small.json: the oracle's Monte-Carlo shell properties
(oracle/orc_shellprops.c, xorshift stand-in stream, fixed seeds) and the product rule's values
(host-compiled kernel source, tests/hostcheck) for the shell fitted to shell_small.json's halo
and for two synthetic coefficient rows.

Like shell_small.json this pins the ORACLE (and the quadrature) against drift and gives the
GPU tests numbers that need neither at run time; it is not a reference-generated vector --
the reference's stream is Go's unseeded global math/rand."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import hostcheck as hc  # noqa: E402
import shell_cases as shells  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SAMPLES = 50000
rng = np.random.default_rng(2024)
rows = {"shell_small": json.load(open(os.path.join(HERE, "shell_small.json")))["coeffs"],
        "order2": list(shells.random_shell(rng, 2, radius=1.5, amp=0.3)),
        "order4": list(shells.random_shell(rng, 4, radius=0.8, amp=0.3))}
tables = shells.axis_tables()
out = {"samples": SAMPLES, "cases": {}}
for name, cs in rows.items():
    a, b, c, vec = po.shell_axes(cs, SAMPLES, 3, tables)
    lo, hi = po.shell_radial_range(cs, SAMPLES, 4)
    out["cases"][name] = {
        "coeffs": [float(v) for v in cs],
        "monte_carlo": {"volume_seed1": po.shell_volume(cs, SAMPLES, 1), "area_seed2": po.shell_surface_area(cs, SAMPLES, 2),
                        "axes_seed3": [a, b, c], "avec_seed3": [float(v) for v in vec], "range_seed4": [lo, hi]},
        "rule_128x256": [float(v) for v in hc.shell_props(cs)],
    }
json.dump(out, open(os.path.join(HERE, "stats_small.json"), "w"), indent=1)
print("wrote stats_small.json")
