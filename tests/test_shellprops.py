"""Shell property integrals of `shellfish stats` on the CPU tier.

Three implementations meet here:
  * the oracle's literal Monte-Carlo restatement of los/analyze/shell.go
    (oracle/orc_shellprops.c; no sample-exact parity: the reference's sample stream is Go's
    unseeded global math/rand),
  * the product's per-node arithmetic, i.e. the source of shell_props_kernel compiled
    for the host (tests/hostcheck), on the product's quadrature rule,
  * closed forms / numpy / scipy.
The GPU tier (tests/test_gpu_shellprops.py) then only has to show that the kernel
returns what its host-compiled source returns."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
import hostcheck as hc
import shell_cases as shells


def test_gauss_legendre_rule_matches_numpy():
    for n in (2, 7, 64, 128):
        x, w = hc.gauss_legendre(n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.allclose(x, xr, rtol=0, atol=2e-15) and np.allclose(w, wr, rtol=1e-10, atol=0)
        assert abs(w.sum() - 2) < 1e-14
        k = 2 * n - 2       # the highest even degree the rule integrates exactly
        assert abs((w * x ** k).sum() * (k + 1) / 2 - 1) < 1e-12


@pytest.mark.parametrize("order", [2, 3, 4])
def test_penna_radius_and_cos_norm_equal_the_literal_restatement(order):
    """Hoisting the math.Pow calls out of the term loop changes nothing: the same
    products in the same association (penna.go:62-82)."""
    rng = np.random.default_rng(order)
    for _ in range(20):
        cs = shells.random_shell(rng, order, radius=rng.uniform(0.5, 3))
        for _ in range(20):
            phi, th = rng.uniform(-0.1, 2 * np.pi + 0.1), rng.uniform(0, np.pi)
            assert hc.penna_radius(cs, phi, th) == po.penna_func(cs, phi, th)
            assert hc.cos_norm(cs, phi, th) == po.cos_norm(cs, phi, th)


def test_sphere_closed_forms():
    R = 1.7
    p = hc.shell_props(shells.sphere(R))
    vol = 4 * np.pi / 3 * R ** 3
    assert abs(p[1] - vol) < 1e-13 * vol
    assert abs(p[0] - (vol / (4 * np.pi / 3)) ** 0.33333) < 1e-13 * R      # cmd/stats.go:279
    assert abs(p[2] - 4 * np.pi * R * R) < 1e-9 * R * R                     # finite-difference cosNorm
    assert p[9] == R and p[10] == R
    # Monte-Carlo restatement: every sample of a sphere gives the same value
    assert abs(po.shell_volume(shells.sphere(R), 1000) - vol) < 1e-12 * vol
    assert abs(po.shell_surface_area(shells.sphere(R), 1000) - 4 * np.pi * R * R) < 1e-8 * R * R
    assert po.shell_radial_range(shells.sphere(R), 100) == (R, R)


def test_offset_sphere_like_surface_closed_form():
    """R = 1 + e cos(theta): mean R^3 = 1 + e^2 exactly (odd powers of cos vanish)."""
    e = 0.3
    cs = shells.sphere(1.0)
    cs[9] = e
    s = hc.shell_sums(cs)
    assert abs(s[0] - (1 + e * e)) < 1e-14
    assert abs(s[11] - (1 - e)) < 2e-4 and abs(s[12] - (1 + e)) < 2e-4   # poles are not nodes
    assert s[11] >= 1 - e and s[12] <= 1 + e


@pytest.mark.parametrize("order", [2, 3, 4])
def test_rule_is_converged(order):
    rng = np.random.default_rng(10 + order)
    cs = shells.random_shell(rng, order, amp=0.4)
    a, b = hc.shell_sums(cs, 64, 128), hc.shell_sums(cs, 128, 256)
    assert np.allclose(a[:11], b[:11], rtol=1e-11, atol=1e-13)
    assert abs(a[11] - b[11]) < 1e-3 and abs(a[12] - b[12]) < 1e-3


@pytest.mark.parametrize("order", [2, 3, 4])
def test_quadrature_agrees_with_monte_carlo_statistically(order):
    """The deterministic rule must sit inside the scatter of the reference's estimator:
    |rule - MC mean| < 5 sigma_mean over independent 50 000-sample runs (the
    reference's MonteCarloSamples default, cmd/stats.go:111)."""
    rng = np.random.default_rng(20 + order)
    cs = shells.random_shell(rng, order, radius=1.3, amp=0.35)
    p = hc.shell_props(cs)
    seeds = range(1, 13)
    vols = np.array([po.shell_volume(cs, 50000, s) for s in seeds])
    sas = np.array([po.shell_surface_area(cs, 50000, 100 + s) for s in seeds])
    for got, mc in ((p[1], vols), (p[2], sas)):
        err = mc.std(ddof=1) / np.sqrt(len(mc))
        assert abs(got - mc.mean()) < 5 * err, (got, mc.mean(), err)
        assert err < 2e-3 * got
    lo, hi = po.shell_radial_range(cs, 50000, 7)
    # both estimates approach the true extrema from inside
    assert abs(p[9] - lo) < 2e-3 and abs(p[10] - hi) < 2e-3


def test_axes_agree_with_monte_carlo_statistically():
    rng = np.random.default_rng(31)
    tables = shells.axis_tables()
    for order in (2, 3):
        cs = shells.random_shell(rng, order, radius=1.0, amp=0.45)
        p = hc.shell_props(cs)
        runs = [po.shell_axes(cs, 50000, s, tables) for s in range(1, 9)]
        assert all(r is not None for r in runs)
        abc = np.array([r[:3] for r in runs])
        err = abc.std(0, ddof=1) / np.sqrt(len(runs))
        assert np.all(np.abs(p[3:6] - abc.mean(0)) < 5 * err + 1e-4), (p[3:6], abc.mean(0), err)
        assert p[3] >= p[4] >= p[5] > 0
        vec = np.array([r[3] for r in runs])
        cosang = np.abs(vec @ p[6:9])
        assert np.all(cosang > 0.98), cosang       # same major axis up to sign and noise
        assert abs(np.linalg.norm(p[6:9]) - 1) < 1e-14 and p[6:9][np.abs(p[6:9]).argmax()] > 0


def test_reference_quickstart_known_answer():
    """doc/quickstart.md: the coefficients `shellfish shell` printed for halo 80431577 and the row
    `shellfish stats` made of them.  The one vector the reference holds for Shell.Volume /
    SurfaceArea / Axes and the R_sp formula: the product's quadrature must reproduce it within
    the sampling noise of the reference's own 50 000-sample estimate, and so must the oracle's
    Monte-Carlo restatement (which pins it beyond the cross-checks above)."""
    cs = shells.QUICKSTART_COEFFS
    p = hc.shell_props(cs)
    shells.assert_matches_quickstart(p)
    # measured: R_sp -2.2e-4, volume -6.5e-4, area -9.2e-4, axes +2.8e-3 / -8e-4 / -2.8e-3 relative
    tables = shells.axis_tables()
    for seed in (1, 2, 3):
        ax = po.shell_axes(cs, 50000, seed, tables)
        vol = po.shell_volume(cs, 50000, 10 + seed)
        row = np.concatenate([[(vol / (4 * np.pi / 3)) ** 0.33333, vol, po.shell_surface_area(cs, 50000, 20 + seed)],
                              ax[:3], ax[3]])
        # two independent estimates: sqrt(2) x the one-run scatter
        shells.assert_matches_quickstart(row, n_sigma=4.5)
    # the estimators' scatter quoted in shell_cases.QUICKSTART_SIGMA
    vols = np.array([po.shell_volume(cs, 50000, s) for s in range(30, 38)])
    assert 0.3 * shells.QUICKSTART_SIGMA["volume"] < vols.std(ddof=1) / vols.mean() < 2.5 * shells.QUICKSTART_SIGMA["volume"]


def test_axes_closing_arithmetic_against_numpy():
    """sfk_axes_from_moments (libsfk.so host code) vs shell.go:140-198 written out with
    numpy.linalg.eigh and the oracle's BiCubic."""
    from shellfish_b200 import api
    ratios, acg, bcg, cg = shells.axis_tables()
    rng = np.random.default_rng(5)
    for _ in range(20):
        # moments of a thin homoeoid with semi-axes s in a random orientation, off-centre
        s = np.sort(rng.uniform(0.5, 1.5, 3))[::-1]
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        sec = q @ np.diag(s * s / 3) @ q.T
        mu = rng.normal(size=3) * 0.05
        sec = sec + np.outer(mu, mu)
        m = [sec[0, 0], sec[1, 1], sec[2, 2], sec[0, 1], sec[1, 2], sec[2, 0], mu[0], mu[1], mu[2]]
        got = api.axes_from_moments(m)
        cen = sec - np.outer(mu, mu)
        inertia = np.trace(cen) * np.eye(3) - cen
        w, v = np.linalg.eigh(inertia)
        ax2 = 3 * (w[1] + w[2] - w[0]) / 2
        c, b, a = np.sqrt(sorted([ax2, 3 * w[1] - ax2, 3 * w[2] - ax2], reverse=True))
        assert np.allclose([c, b, a], s, rtol=1e-12)         # the homoeoid relation recovers s
        ac, bc = a / c, b / c
        cr = po.bicubic_eval(ratios, ratios, cg, ac, bc)
        exp = [cr * c, po.bicubic_eval(ratios, ratios, bcg, ac, bc) * bc * cr * c,
               po.bicubic_eval(ratios, ratios, acg, ac, bc) * ac * cr * c]
        assert np.allclose(got[:3], exp, rtol=1e-12), (got, exp)
        assert abs(abs(got[3:] @ v[:, 0]) - 1) < 1e-10


def test_bicubic_against_scipy_natural_splines():
    from scipy.interpolate import CubicSpline
    ratios, acg, _, cg = shells.axis_tables()
    n = len(ratios)
    rng = np.random.default_rng(9)
    for grid in (acg, cg):
        g = grid.reshape(n, n)     # [yi][xi]
        for _ in range(10):
            x, y = rng.uniform(ratios[0], 1, 2)
            at_y = [CubicSpline(ratios, g[:, xi], bc_type="natural")(y) for xi in range(n)]
            ref = CubicSpline(ratios, at_y, bc_type="natural")(x)
            assert abs(po.bicubic_eval(ratios, ratios, grid, x, y) - ref) < 1e-13
    assert po.bicubic_eval(ratios, ratios, acg, 0.1, 0.5) is None      # Spline.Eval panics, spline.go:80-90


def test_elongated_shells_clamp_instead_of_panicking():
    """Axis ratios below the first table knot: the reference panics, the product clamps."""
    from shellfish_b200 import api
    m = [1.0 / 3, 0.01 / 3, 0.01 / 3, 0, 0, 0, 0, 0, 0]     # a homoeoid with ratios 0.1
    got = api.axes_from_moments(m)
    assert np.all(np.isfinite(got)) and got[0] > got[1] >= got[2] > 0
    assert abs(abs(got[3]) - 1) < 1e-12


def _ellipsoid_moments(a, b, c):
    """Area-weighted second and first moments of a true ellipsoid, by the generator's own
    quadrature (scripts/make_axis_tables.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(shells.ROOT, "scripts"))
    import make_axis_tables as mt
    u, gw = np.polynomial.legendre.leggauss(96)
    phi1 = 2 * np.pi * (np.arange(192) + 0.5) / 192
    th, phi = np.meshgrid(np.arccos(u), phi1, indexing="ij")
    w = (np.repeat(gw[:, None], 192, 1) * 0.5 / 192).ravel()
    th, phi = th.ravel(), phi.ravel()
    s = mt.ellipsoid(a, b, c)
    r = s(phi, th)
    area = w * r * r / mt.cos_norm(s, phi, th)
    x, y, z = mt.cartesian(phi, th, r)
    n = area.sum()
    return [(area * p * q).sum() / n for p, q in ((x, x), (y, y), (z, z), (x, y), (y, z), (z, x))] + \
           [(area * p).sum() / n for p in (x, y, z)]


def test_reference_test_case_ellipsoid_2_4_3():
    """los/analyze/shell_test.go:43-51 (TestEverything) runs Axes on ellipsoid(2, 4, 3) and only
    prints the result.  With the regenerated correction tables the axes come back within a
    few per cent -- as good as the reference's own tables allow -- and the major axis is y."""
    from shellfish_b200 import api
    got = api.axes_from_moments(_ellipsoid_moments(2, 4, 3))
    assert np.allclose(got[:3], [4, 3, 2], rtol=0.05), got
    assert abs(abs(got[4]) - 1) < 1e-9 and abs(got[3]) < 1e-9 and abs(got[5]) < 1e-9
    # a quirk kept on purpose: interplote.py flips the c grid top to bottom, so the entry a sphere
    # reads, CCorrectionGrid[23][23], is the correction of the most elongated body; the reference's
    # grid.go holds 1.1153 there, and so does the regenerated table
    sph = api.axes_from_moments(_ellipsoid_moments(1, 1, 1))
    assert abs(sph[0] - 1.1153) < 2e-4, sph


def test_reference_tables_are_imported_verbatim_and_match_the_recomputed_ones():
    """The product interpolates the reference's own data (ellipse_grid/grid.go, imported by
    scripts/import_axis_tables.py); the grids recomputed from first principles
    (scripts/make_axis_tables.py) agree with them within the Monte-Carlo noise of the reference's
    entries (1e6 samples each): median 5e-4, max 3e-3 relative."""
    ref, gen = shells.axis_tables(), shells.regenerated_axis_tables()
    assert np.array_equal(ref[0], gen[0]) and len(ref[0]) == 24
    for a, b in zip(ref[1:], gen[1:]):
        assert a.shape == b.shape == (576,)
        rel = np.abs(a - b) / np.abs(a)
        assert np.median(rel) < 8e-4 and rel.max() < 3.5e-3
    assert ref[3].reshape(24, 24)[23, 23] == 1.1153          # the entry a sphere reads (flipped c grid)
    src = "/root/reference/los/analyze/ellipse_grid/grid.go"  # only in the build container
    if os.path.exists(src):
        import re
        nums = [float(v) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", open(src).read().split("ACCorrectionGrid")[1].split("{")[1])]
        assert np.array_equal(np.array(nums[:576]), ref[1])


def test_committed_tables_are_what_the_generator_writes(tmp_path):
    import os
    import subprocess
    import sys
    out = tmp_path / "tables.inc"
    subprocess.run([sys.executable, os.path.join(shells.ROOT, "scripts", "make_axis_tables.py"), str(out)], check=True)
    committed = open(os.path.join(shells.ROOT, "shellfish_b200", "csrc", "sfk_axis_tables.inc")).read()
    assert out.read_text() == committed


def test_angular_fraction_profile():
    """Shell.AngularFractionProfile (shell.go:295-335): closed form for R = 1 + e cos(theta)
    -- the fraction of directions with R > r is (1 - (r - 1)/e)/2 -- and the Monte-Carlo
    restatement within its binomial noise."""
    e = 0.3
    cs = shells.sphere(1.0)
    cs[9] = e
    bins, r_min, r_max = 150, 0.3, 3.0
    f = hc.angular_fraction(cs, r_min, r_max, bins)
    edges = np.exp(np.log(r_min) + np.arange(bins) * (np.log(r_max) - np.log(r_min)) / bins)
    exact = np.clip((1 - (edges - 1) / e) / 2, 0, 1)     # fs[i] counts directions whose bin is >= i
    assert np.abs(f - exact).max() < 1.01 / 1024          # one stratum of cos(theta): the worst case, an axisymmetric shell
    assert f[0] == 1.0 and np.all(np.diff(f) <= 0) and f[-1] == 0.0
    rs, fm = po.angular_fraction_profile(cs, 200000, bins, r_min, r_max, seed=3)
    assert np.allclose(rs, np.exp(np.log(r_min) + (np.arange(bins) + 0.5) * (np.log(r_max) - np.log(r_min)) / bins))
    assert np.abs(f - fm).max() < 5 * np.sqrt(0.25 / 200000)
    rng = np.random.default_rng(41)
    for order in (2, 3, 4):
        cs = shells.random_shell(rng, order, radius=1.0, amp=0.4)
        f = hc.angular_fraction(cs, 0.5, 2.0, 60)
        fm = po.angular_fraction_profile(cs, 200000, 60, 0.5, 2.0, seed=order)[1]
        assert np.abs(f - fm).max() < 5 * np.sqrt(0.25 / 200000)


def test_angular_fraction_bin_truncation():
    """Go's int() truncates toward zero: a surface slightly inside rMin is still counted in
    bin 0 (shell.go:313-318); one far inside is skipped and the profile is 0/0 = NaN."""
    assert np.all(hc.angular_fraction(shells.sphere(0.999), 1.0, 2.0, 10)[:1] == 1.0)
    rs, fm = po.angular_fraction_profile(shells.sphere(0.999), 100, 10, 1.0, 2.0)
    assert fm[0] == 1.0 and fm[1] == 0.0
    assert np.all(np.isnan(hc.angular_fraction(shells.sphere(0.5), 1.0, 2.0, 10)))
    assert np.all(np.isnan(po.angular_fraction_profile(shells.sphere(0.5), 100, 10, 1.0, 2.0)[1]))


def test_golden_stats_fixture():
    """tests/golden/stats_small.json (tests/golden/make_golden_stats.py): the oracle's Monte-Carlo
    values at fixed seeds and the rule's values do not drift."""
    import json
    import os
    g = json.load(open(os.path.join(shells.ROOT, "tests", "golden", "stats_small.json")))
    tables = shells.axis_tables()
    for name, case in g["cases"].items():
        cs, mc, n = case["coeffs"], case["monte_carlo"], g["samples"]
        assert po.shell_volume(cs, n, 1) == pytest.approx(mc["volume_seed1"], rel=1e-13)
        assert po.shell_surface_area(cs, n, 2) == pytest.approx(mc["area_seed2"], rel=1e-13)
        a, b, c, vec = po.shell_axes(cs, n, 3, tables)
        assert [a, b, c] == pytest.approx(mc["axes_seed3"], rel=1e-11)
        assert list(po.shell_radial_range(cs, n, 4)) == pytest.approx(mc["range_seed4"], rel=1e-14)
        np.testing.assert_allclose(hc.shell_props(cs), case["rule_128x256"], rtol=1e-12, atol=1e-14)
        # the two estimates of the same integrals agree within the Monte-Carlo scatter (~1e-3)
        assert abs(mc["volume_seed1"] / case["rule_128x256"][1] - 1) < 3e-3
        assert abs(mc["area_seed2"] / case["rule_128x256"][2] - 1) < 3e-3


def test_degenerate_rows_do_not_crash():
    """An all-zero coefficient row (what `shell` prints for a halo whose fit failed) gives
    volume 0 and NaN shape columns; the reference's Axes would panic in mat64.Eigen there
    (los/analyze/shell.go:151-154)."""
    p = hc.shell_props(np.zeros(18))
    assert p[0] == 0 and p[1] == 0 and np.all(np.isnan(p[2:9])) and p[9] == 0 and p[10] == 0
    assert np.all(np.isnan(hc.shell_props(np.full(18, np.nan))[:9]))
