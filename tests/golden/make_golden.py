"""Regenerates tests/golden/shell_small.json from the CPU oracle.

The reference (pure Go) cannot be run in this image (no Go toolchain, module
dependencies not vendored), so this fixture pins the ORACLE against drift and
gives the GPU tests a vector that does not need the oracle at run time.  It is
not a reference-generated vector."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as po  # noqa: E402
from shellfish_b200 import synth  # noqa: E402

SYNTH = dict(seed=1, n=20000, n_background=5000)
CONFIG = dict(rings=10, seed=1337, threads=1)

halos, blocks, mass = synth.single_halo(**SYNTH)
out = po.shell_run(po.ShellConfig.default(**CONFIG), mass, halos, blocks,
                   taps=("nIntersections", "nCandidates", "nFiltered", "rsp"))
rsp = out["rsp"][0]
json.dump({"synth": SYNTH, "config": CONFIG,
           "nIntersections": int(out["nIntersections"][0]), "nCandidates": int(out["nCandidates"][0]),
           "nFiltered": int(out["nFiltered"][0]), "coeffs": [float(c) for c in out["coeffs"][0]],
           "rsp_ring0": [None if r != r else float(r) for r in rsp[0]]},
          open(os.path.join(HERE, "shell_small.json"), "w"), indent=1)
print("wrote shell_small.json")
