"""Cross-checks of the oracle functions the reference has NO golden vectors for
(Smooth, SplashbackRadius, splines, KDE filter, Penna fit) against independent
formulations in scipy / numpy.  These do not pin parity with the Go code (see
oracle/shellfish_oracle.h, "PARITY UNPINNED"); they pin the published
algorithms the Go code implements.  CPU only.
"""
import ctypes as C
import math

import numpy as np
import pytest
from scipy.interpolate import CubicSpline
from scipy.signal import savgol_coeffs

from oracle import pyoracle as po

L = po.lib()


@pytest.mark.parametrize("width,order", [(121, 4), (61, 4), (11, 2), (5, 2)])
def test_savgol_value_taps_match_scipy(width, order):
    got = po.savgol_kernel(order, width)
    want = savgol_coeffs(width, order, deriv=0)
    # the reference solves the (ill-conditioned, entries up to 60^8) moment
    # matrix by LU; scipy uses least squares: agreement to ~1e-10 is all there is
    assert np.allclose(got, want[::-1], rtol=0, atol=1e-9)


def test_savgol_derivative_taps_match_scipy():
    dx = 0.008994
    got = po.savgol_kernel(4, 121, ld=1, dx=dx)
    want = savgol_coeffs(121, 4, deriv=1, delta=dx, use="dot")
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    # a straight line of slope s has derivative s everywhere in the interior
    ys = 3.0 + 0.7 * dx * np.arange(400)
    out = np.zeros(400)
    L.orc_convolve_extension(po.dp(got), 121, po.dp(ys), 400, po.dp(out))
    assert np.allclose(out[60:-60], 0.7, atol=1e-9)


def test_convolve_extension_matches_numpy_edge_mode():
    rng = np.random.default_rng(0)
    xs = rng.normal(size=300)
    taps = po.savgol_kernel(4, 121)
    out = np.zeros(300)
    L.orc_convolve_extension(po.dp(taps), 121, po.dp(xs), 300, po.dp(out))
    pad = np.pad(xs, 60, mode="edge")
    want = np.array([np.dot(pad[i:i + 121], taps) for i in range(300)])
    assert np.allclose(out, want, atol=1e-12)


def test_spline_is_the_natural_cubic_spline():
    rng = np.random.default_rng(1)
    xs = np.sort(rng.uniform(0, 10, 40))
    ys = np.sin(xs) + 0.1 * rng.normal(size=40)
    sp = L.orc_spline_new(po.dp(xs), po.dp(ys), 40)
    ref = CubicSpline(xs, ys, bc_type="natural")
    y = C.c_double()
    for x in np.linspace(xs[0], xs[-1], 333):
        assert L.orc_spline_eval(sp, float(x), C.byref(y)) == 0
        assert abs(y.value - ref(x)) < 1e-10
    # Spline.Eval panics outside the table (spline.go:80-93)
    assert L.orc_spline_eval(sp, float(xs[0] - 1), C.byref(y)) == -1
    assert L.orc_spline_eval(sp, float(xs[-1] + 1), C.byref(y)) == -1
    L.orc_spline_free(sp)


def test_smooth_and_splashback_on_an_analytic_profile():
    """rho = r^-2 * (1 + (r/r0)^8)^-0.5: the log-slope dips from -2 to -6 with
    the steepest point of the *smoothed* profile near r0."""
    bins, window = 256, 121
    lr = np.log(0.3) + (np.log(3.0) - np.log(0.3)) / bins * (np.arange(bins) + 0.5)
    rs = np.exp(lr)
    r0 = 1.1
    rho = rs ** -2.0 / np.sqrt(1 + (rs / r0) ** 8) * (1 + 3 * (rs / 2.5) ** 6)
    dx = lr[1] - lr[0]
    k, kd = po.savgol_kernel(4, window), po.savgol_kernel(4, window, ld=1, dx=dx)
    ys, vals, ders = rho.copy(), np.zeros(bins), np.zeros(bins)
    assert L.orc_smooth(po.dp(k), po.dp(kd), window, po.dp(ys), bins, po.dp(vals), po.dp(ders)) == 1
    assert np.allclose(ys, rho, rtol=1e-14)           # ln/exp round trip restores the input
    # independent smoothing: scipy savgol on ln rho with edge extension
    from scipy.signal import savgol_filter
    want = np.exp(savgol_filter(np.log(rho), window, 4, mode="nearest"))
    wantd = savgol_filter(np.log(rho), window, 4, deriv=1, delta=dx, mode="nearest")
    assert np.allclose(vals, want, rtol=1e-7)     # taps agree to ~4e-11, see above
    assert np.allclose(ders, wantd, rtol=0, atol=1e-6)
    r = C.c_double()
    assert L.orc_splashback_radius(po.dp(rs), po.dp(vals), po.dp(ders), bins, 0.0, C.byref(r)) == 1
    # numpy restatement of the search
    rho_min, d_min, i_min = vals[0], 0.0, -1
    for i in range(1, bins - 1):
        if vals[i] < rho_min:
            rho_min = vals[i]
            if ders[i] < ders[i + 1] and ders[i] < ders[i - 1] and ders[i] < d_min:
                d_min, i_min = ders[i], i
    assert i_min > 0 and r.value == rs[i_min]
    assert r0 < r.value < 2.0      # the slope is steepest between the break and the upturn
    # too few bins -> not ok (smooth.go:51-53)
    assert L.orc_smooth(po.dp(k), po.dp(kd), window, po.dp(ys[:100].copy()), 100, po.dp(vals), po.dp(ders)) == 0


def _penna_design(x, y, z, P):
    r = np.sqrt(x * x + y * y + z * z)
    ct = z / r
    st = np.sqrt(1 - ct * ct)
    cp, sp_ = x / r / st, y / r / st
    cols = []
    for k in range(2):
        for j in range(P):
            for i in range(P):
                cols.append(st ** (i + j) * cp ** i * ct ** k * sp_ ** j)
    return np.stack(cols, axis=1), r


@pytest.mark.parametrize("P", [2, 3, 4])
def test_penna_fit_is_least_squares_and_well_enough_conditioned(P):
    """Conditioning study for the 1e-5 claim: the normal-equation solve of the
    oracle (gonum-style LU inverse) agrees with an orthogonal (SVD) solve of
    the same design matrix far inside 1e-5 of the coefficient scale."""
    rng = np.random.default_rng(5 + P)
    n = 6000
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    rad = 1.2 + 0.15 * v[:, 0] ** 2 - 0.1 * v[:, 2] + 0.03 * rng.normal(size=n)
    pts = v * rad[:, None]
    x, y, z = (np.ascontiguousarray(pts[:, i]) for i in range(3))
    cs = np.zeros(2 * P * P)
    assert L.orc_penna_coeffs(po.dp(x), po.dp(y), po.dp(z), n, P, P, 2, po.dp(cs)) == 0
    M, r = _penna_design(x, y, z, P)
    want = np.linalg.lstsq(M, r, rcond=None)[0]
    scale = np.max(np.abs(want))
    cond = np.linalg.cond(M.T @ M)
    assert cond * 2.3e-16 < 1e-7
    assert np.max(np.abs(cs - want)) <= 1e-8 * scale
    # PennaFunc reproduces the fitted radii
    phi = np.arctan2(y, x)
    th = np.arccos(z / np.sqrt(x * x + y * y + z * z))
    fit = np.array([L.orc_penna_func(po.dp(cs), P, P, 2, float(a), float(b)) for a, b in zip(phi[:50], th[:50])])
    assert np.allclose(fit, (M @ cs)[:50], atol=1e-10)


def test_filter_ring_keeps_the_shell_and_drops_outliers():
    n = 256
    phi = 2 * math.pi * np.arange(n) / n
    rng = np.random.default_rng(3)
    r = 1.2 + 0.1 * np.cos(2 * phi) + 0.01 * rng.normal(size=n)
    bad = rng.choice(n, 40, replace=False)
    r[bad] = rng.uniform(0.4, 2.6, 40)
    fr, fp, fi = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
    h = C.c_double()
    kept = L.orc_filter_ring(po.dp(r), po.dp(phi), n, 3, 0.3, po.dp(fr), po.dp(fp),
                             fi.ctypes.data_as(C.POINTER(C.c_int)), C.byref(h))
    assert kept > 150 and h.value == 0.3
    idx = fi[:kept]
    assert np.all(np.diff(idx) > 0)                       # order preserved
    truth = 1.2 + 0.1 * np.cos(2 * phi[idx])
    assert np.all(np.abs(fr[:kept] - truth) < 0.15 + 0.05)  # within h/2 of the curve
    good = np.setdiff1d(np.arange(n), bad)
    assert len(np.intersect1d(idx, good)) > 0.9 * len(good)
    # no input points -> the reference panics (kde.go:72-81)
    assert L.orc_filter_ring(po.dp(r), po.dp(phi), 0, 3, 0.3, po.dp(fr), po.dp(fp), None, None) == -2
