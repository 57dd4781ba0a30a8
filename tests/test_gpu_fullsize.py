"""CUDA path against the committed oracle fixtures at bench / BASELINE sizes
(tests/golden/make_golden_fullsize.py): the default shell.config on a 1e5-particle halo
(the per-halo shape of configs[1]) and 24 x 256 x 256 on a 1e6-particle halo (configs[2]'s
angular resolution).  Integer taps bit-exact, profiles / R_sp / coefficients within 1e-5."""
import hashlib
import os

import numpy as np
import pytest

import importlib.util

from shellfish_b200 import api

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden_fullsize", os.path.join(GOLDEN, "make_golden_fullsize.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
RTOL = 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", sorted(gen.CASES))
def test_fixture_inputs_are_reproducible(case):
    """CPU: the seeded generator still produces the particles the fixture was made from."""
    fx = np.load(os.path.join(GOLDEN, case + ".npz"))
    halos, blocks, mass = gen.inputs(case)
    assert sha(blocks[0]["xs"]) == str(fx["xs_sha256"]) and mass == float(fx["min_mass"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(gen.CASES))
def test_gpu_matches_the_full_size_oracle_fixture(case):
    fx = np.load(os.path.join(GOLDEN, case + ".npz"))
    halos, blocks, mass = gen.inputs(case)
    assert sha(blocks[0]["xs"]) == str(fx["xs_sha256"])
    params = api.default_params(seed=int(fx["seed"]), rings=int(fx["rings"]), flags=api.SFK_FLAG_TAPS)
    ctx, coeffs, status = api.run_shell(params, mass, halos, blocks)
    assert status[0] == int(fx["status"]) == 0
    assert ctx.candidate_count(0) == int(fx["n_candidates"])
    assert ctx.intersection_count(0) == int(fx["n_intersections"]) == ctx.stats()["n_intersections"]
    ins, st, en = ctx.counts(0)
    assert np.array_equal(ins, fx["n_insert"]), "inserts per LoS"
    # marginals first (they localise a mismatch), then the full arrays through their hashes
    assert np.array_equal(st.sum(axis=1, dtype=np.uint64), fx["start_by_ring_bin"])
    assert np.array_equal(en.sum(axis=1, dtype=np.uint64), fx["end_by_ring_bin"])
    assert np.array_equal(st.sum(axis=2, dtype=np.uint64), fx["start_by_ring_los"])
    assert np.array_equal(en.sum(axis=2, dtype=np.uint64), fx["end_by_ring_los"])
    assert sha(st) == str(fx["start_sha256"]) and sha(en) == str(fx["end_sha256"])
    prof, rp = ctx.profiles(0)[::4, ::8, :], fx["profiles_sub"]
    assert np.max(np.abs(prof - rp) / np.maximum(np.abs(rp), rp.min())) <= RTOL
    rsp, rr = ctx.rsp(0), fx["rsp"]
    assert np.array_equal(np.isnan(rsp), np.isnan(rr)), "R_sp ok-mask"
    m = ~np.isnan(rr)
    assert np.all(np.abs(rsp[m] - rr[m]) <= RTOL * np.abs(rr[m]))
    assert np.all(np.abs(coeffs[0] - fx["coeffs"]) <= RTOL * np.abs(fx["coeffs"]).max())
    # the product instantiation (no taps) gives the same answer
    _, c2, s2 = api.run_shell(api.default_params(seed=int(fx["seed"]), rings=int(fx["rings"])), mass, halos, blocks)
    assert s2[0] == 0 and np.array_equal(c2, coeffs)
