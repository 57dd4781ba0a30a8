#!/usr/bin/env python
"""Oracle fixtures at the sizes the bench and BASELINE configs run at.

  python tests/golden/make_golden_fullsize.py        (about 3 minutes on 8 cores)

Writes tests/golden/shell_default_1e5.npz  (default shell.config, 100 rings x 256 LoS x
256 bins, one 1e5-particle halo + background: the per-halo shape of configs[1]) and
tests/golden/shell_config3_1e6.npz (24 rings x 256 x 256 on a 1e6-particle halo: the
angular resolution of configs[2], a tenth of its particles so that the oracle finishes in
a minute).  Inputs are regenerated from the seeds stored in the file (synth is seeded
numpy); `xs_sha256` guards against a drifting generator.  Outputs are the oracle's
(oracle/orc_shell.c, the C restatement of the Go loop): a drift guard for the oracle and,
on the B200, the expected values of tests/test_gpu_fullsize.py.  The Go binary cannot be
run here, so these are not reference-generated vectors.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po          # noqa: E402
from shellfish_b200 import synth           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "shell_default_1e5": dict(synth=dict(seed=101, n=100000, n_background=20000), rings=100),
    "shell_config3_1e6": dict(synth=dict(seed=103, n=1000000, n_background=50000), rings=24),
}


def inputs(case):
    return synth.single_halo(**CASES[case]["synth"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    for name, c in CASES.items():
        halos, blocks, mass = inputs(name)
        cfg = po.ShellConfig.default(seed=1337, threads=os.cpu_count() or 1, rings=c["rings"])
        out = po.shell_run(cfg, mass, halos, blocks,
                           taps=("rsp", "profiles", "nInsert", "startEdges", "endEdges", "nCandidates", "nIntersections"))
        st, en, prof = out["startEdges"][0], out["endEdges"][0], out["profiles"][0]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            rings=c["rings"], seed=1337, xs_sha256=sha(blocks[0]["xs"]), min_mass=mass,
            coeffs=out["coeffs"][0], status=out["status"][0],
            n_candidates=out["nCandidates"][0], n_intersections=out["nIntersections"][0],
            rsp=out["rsp"][0], n_insert=out["nInsert"][0],
            # bit-exact integer taps: full-array hashes plus marginals that localise a mismatch
            start_sha256=sha(st), end_sha256=sha(en),
            start_by_ring_bin=st.sum(axis=1, dtype=np.uint64), end_by_ring_bin=en.sum(axis=1, dtype=np.uint64),
            start_by_ring_los=st.sum(axis=2, dtype=np.uint64), end_by_ring_los=en.sum(axis=2, dtype=np.uint64),
            # Halo.GetRhos of every 8th LoS of every 4th ring
            profiles_sub=prof[::4, ::8, :])
        print(name, "ok=%d" % out["ok"][0], "intersections=%d" % out["nIntersections"][0],
              "c000=%.6g" % out["coeffs"][0, 0])


if __name__ == "__main__":
    main()
