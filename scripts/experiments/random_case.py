"""Diagnostic for tests/test_gpu_parity.py::test_randomized_small_configurations[case]:
   python scripts/experiments/random_case.py CASE"""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from oracle import pyoracle as po  # noqa: E402
from shellfish_b200 import api, synth  # noqa: E402

case = int(sys.argv[1])
rng = np.random.default_rng(1000 + case)
n_halos = int(rng.integers(1, 5))
n_per = int(rng.choice([400, 1500, 6000, 20000]))
box = float(rng.choice([6.0, 12.0, 25.0]))
cells = int(rng.integers(1, 4))
halos, blocks, mass = synth.make_box(2000 + case, box, n_halos, n_per, int(rng.integers(0, 4000)), cells,
                                     r200m_range=(0.5, 0.9))
bins = int(rng.choice([256, 256, 128, 200]))
kw = dict(rings=int(rng.integers(3, 9)), spokes=int(rng.choice([256, 64, 100, 160])), radial_bins=bins,
          smoothing_window=int(rng.choice([w for w in (31, 61, 121) if w < bins - 8])),
          levels=int(rng.integers(1, 4)), order=int(rng.integers(1, 5)),
          subsample_factor=int(rng.choice([1, 1, 2])), eta=float(rng.choice([10.0, 6.0, 14.0])))
print("case", case, "halos", n_halos, "n_per", n_per, "box", box, "cells", cells, kw)
names = {"rings": "rings", "spokes": "spokes", "radial_bins": "radialBins", "order": "order", "levels": "levels",
         "smoothing_window": "smoothingWindow", "subsample_factor": "subsampleFactor", "eta": "eta"}
cfg = po.ShellConfig.default(seed=1337, threads=1, **{names[k]: v for k, v in kw.items()})
ref = po.shell_run(cfg, mass, halos, blocks, taps=("rsp", "profiles", "nInsert"))
for name, fl in (("default", api.SFK_FLAG_TAPS), ("taploop", api.SFK_FLAG_TAPS | api.SFK_FLAG_LOS_TAPLOOP)):
    ctx, c, s = api.run_shell(api.default_params(seed=1337, flags=fl, **kw), mass, halos, blocks)
    for h in range(n_halos):
        rsp, rr = ctx.rsp(h), ref["rsp"][h]
        nanmis = np.argwhere(np.isnan(rsp) != np.isnan(rr))
        m = ~np.isnan(rr) & ~np.isnan(rsp)
        bad = np.argwhere(m & (np.abs(rsp - rr) > 1e-5 * np.abs(rr)))
        ins = ctx.counts(h)[0]
        prof = ctx.profiles(h)
        rp = ref["profiles"][h]
        perr = np.max(np.abs(prof - rp) / np.maximum(np.abs(rp), rp.min()))
        print(name, "halo", h, "status", s[h], ref["status"][h], "ok-mask mismatches", len(nanmis), "R_sp mismatches", len(bad),
              "inserts equal", bool(np.array_equal(ins, ref["nInsert"][h])), "max profile rel err %.2e" % perr)
        for r_, l_ in bad[:5]:
            print("    ring", r_, "los", l_, "gpu", rsp[r_, l_], "oracle", rr[r_, l_],
                  "profile equal-run max", int(max(len(list(g)) for _, g in __import__("itertools").groupby(rp[r_, l_]))))
