"""Shell property integrals of `shellfish stats` on the GPU (sfk_shell_props,
shell_props_kernel) through the C ABI.  The kernel's per-node source is checked
against the oracle's literal Monte-Carlo restatement on the CPU tier
(tests/test_shellprops.py); here the kernel must return what that same source returns
when compiled for the host, and agree with the oracle statistically."""
import numpy as np
import pytest

import hostcheck as hc
import shell_cases as shells
from oracle import pyoracle as po
from shellfish_b200 import api

pytestmark = pytest.mark.gpu


def test_kernel_equals_its_host_compiled_source():
    rng = np.random.default_rng(77)
    ctx = api.ShellContext(api.default_params(rings=3))
    for order in (2, 3, 4):
        rows = np.array([shells.random_shell(rng, order, radius=rng.uniform(0.3, 4), amp=rng.uniform(0, 0.45))
                         for _ in range(23)])
        props = ctx.shell_props(rows)
        sums = ctx.shell_sums(len(rows))
        for h, cs in enumerate(rows):
            want = hc.shell_sums(cs, *api.SHELL_RULE)
            # sums: same terms, different summation tree and libm (CUDA vs glibc sincos / acos);
            # the first moments cancel to ~1e-2 of their terms, hence the absolute part
            np.testing.assert_allclose(sums[h, :11], want[:11], rtol=1e-12, atol=1e-13 * np.abs(want[:11]).max())
            np.testing.assert_allclose(sums[h, 11:], want[11:], rtol=1e-14)
            np.testing.assert_allclose(props[h], hc.shell_props(cs, *api.SHELL_RULE), rtol=1e-9, atol=1e-12)


def test_reference_quickstart_known_answer_through_the_c_abi():
    """doc/quickstart.md:99-106,150-155: the reference's printed `shell | stats` rows of one halo.
    sfk_shell_props must give the R_sp, Volume, Surface Area and axes the reference printed, within
    the noise of its 50 000-sample Monte-Carlo estimate (tests/test_shellprops.py has the details)."""
    ctx = api.ShellContext(api.default_params(rings=3))
    p = ctx.shell_props(np.array([shells.QUICKSTART_COEFFS]))[0]
    shells.assert_matches_quickstart(p)


def test_sphere_and_edge_cases():
    ctx = api.ShellContext(api.default_params(rings=3))
    R = 2.5
    p = ctx.shell_props(np.array([shells.sphere(R)]))[0]
    assert abs(p[1] - 4 * np.pi / 3 * R ** 3) < 1e-12 * R ** 3
    assert abs(p[2] - 4 * np.pi * R * R) < 1e-8 * R * R
    assert p[9] == R and p[10] == R
    assert ctx.shell_props(np.zeros((0, 18))).shape == (0, api.SHELL_PROPS)
    with pytest.raises(api.ShellfishError, match="Order above 4"):
        ctx.shell_props(np.ones((1, 50)))
    # deterministic: two calls, bitwise the same
    rows = np.array([shells.random_shell(np.random.default_rng(1), 3) for _ in range(5)])
    assert np.array_equal(ctx.shell_props(rows), ctx.shell_props(rows))


def test_agrees_with_monte_carlo_restatement_statistically():
    rng = np.random.default_rng(78)
    ctx = api.ShellContext(api.default_params(rings=3))
    cs = shells.random_shell(rng, 3, radius=1.3, amp=0.35)
    p = ctx.shell_props(np.array([cs]))[0]
    vols = np.array([po.shell_volume(cs, 50000, s) for s in range(1, 9)])
    sas = np.array([po.shell_surface_area(cs, 50000, 50 + s) for s in range(1, 9)])
    for got, mc in ((p[1], vols), (p[2], sas)):
        assert abs(got - mc.mean()) < 5 * mc.std(ddof=1) / np.sqrt(len(mc))
    abc = np.array([po.shell_axes(cs, 50000, s, shells.axis_tables())[:3] for s in range(1, 9)])
    assert np.all(np.abs(p[3:6] - abc.mean(0)) < 5 * abc.std(0, ddof=1) / np.sqrt(8) + 1e-4)


def test_radial_range_feeds_the_mass_pass():
    """stats.go:274-327: rangeSp's RMin / RMax are the r_low / r_high of massContained."""
    rng = np.random.default_rng(79)
    box, n, nh = 30.0, 200000, 12
    xs = (rng.random((n, 3)) * box).astype(np.float32)
    ms = rng.uniform(1.0, 3.0, n).astype(np.float32)
    blocks = [dict(xs=xs, ms=ms, TotalWidth=box, Origin=(0, 0, 0), Width=(box, box, box))]
    centers = 5 + rng.random((nh, 3)) * (box - 10)
    rows = np.array([shells.random_shell(rng, 3, radius=rng.uniform(1, 3), amp=0.3) for _ in range(nh)])
    ctx = api.ShellContext(api.default_params(rings=3))
    props = ctx.shell_props(rows)
    lo, hi = props[:, 9], props[:, 10]
    got = ctx.mass_contained(centers, rows, lo, hi, blocks)
    want = np.array([po.mass_contained(box, xs, ms, rows[h], centers[h].astype(np.float32), lo[h], hi[h])
                     for h in range(nh)])
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # and the range brackets the surface well enough that a 2 % wider one counts the same particles
    wide = ctx.mass_contained(centers, rows, 0.98 * lo, 1.02 * hi, blocks)
    np.testing.assert_allclose(wide, got, rtol=2e-4)


def test_angular_fraction_kernel():
    """sfk_angular_fraction = its host-compiled source (integer histogram: exactly), and the
    Monte-Carlo restatement within noise; NaN rows where no direction falls in range."""
    rng = np.random.default_rng(80)
    ctx = api.ShellContext(api.default_params(rings=3))
    for order in (2, 3, 4):
        rows = np.array([shells.random_shell(rng, order, radius=rng.uniform(0.5, 2), amp=0.4) for _ in range(9)])
        r200 = rows[:, 0] * rng.uniform(0.5, 1.2, len(rows))     # R_sp / R200m between 0.8 and 2
        r200[-1] = rows[-1, 0] / 8                               # a shell wholly beyond rMax: 0/0 rows
        rs, fs = ctx.angular_fraction(rows, 0.03 * r200, 3.0 * r200, 150)
        assert np.all(np.isnan(fs[-1])) and not np.isnan(fs[:-1]).any()
        for h, cs in enumerate(rows):
            want = hc.angular_fraction(cs, 0.03 * r200[h], 3.0 * r200[h], 150, *api.FRACTION_STRATA)
            # a node within an ulp of a bin edge may fall either side (CUDA vs glibc log / sincos)
            if h == len(rows) - 1:
                assert np.all(np.isnan(want))
                continue
            assert np.abs(fs[h] - want).max() < 1e-4
            assert np.sum(fs[h] != want) <= 4
            lr = np.log(0.03 * r200[h]) + (np.arange(150) + 0.5) * (np.log(3.0 * r200[h]) - np.log(0.03 * r200[h])) / 150
            np.testing.assert_allclose(rs[h], np.exp(lr), rtol=1e-14)
        fm = po.angular_fraction_profile(rows[0], 200000, 150, 0.03 * r200[0], 3.0 * r200[0], seed=9)[1]
        assert np.abs(fs[0] - fm).max() < 5 * np.sqrt(0.25 / 200000)
    _, fs = ctx.angular_fraction(np.array([shells.sphere(0.5)]), 1.0, 2.0, 10)
    assert np.all(np.isnan(fs))
    assert ctx.angular_fraction(np.zeros((0, 18)), 1.0, 2.0, 10)[1].shape == (0, 10)


def test_golden_stats_fixture_on_the_gpu():
    """The committed rule values of tests/golden/stats_small.json (no oracle, no host build needed)."""
    import json
    import os
    g = json.load(open(os.path.join(shells.ROOT, "tests", "golden", "stats_small.json")))
    ctx = api.ShellContext(api.default_params(rings=3))
    for name, case in g["cases"].items():
        got = ctx.shell_props(np.array([case["coeffs"]]))[0]
        np.testing.assert_allclose(got, case["rule_128x256"], rtol=1e-9, atol=1e-12)
