"""Oracle of the `stats` M_sp pass (oracle/orc_stats.c) against closed forms.
The reference holds no known answers for this pass (parity unpinned); these
tests pin the restatement to the formulas it implements."""
import numpy as np

from oracle import pyoracle as po


def sphere_coeffs(radius, order=3):
    cs = np.zeros(2 * order * order)
    cs[0] = radius   # i = j = k = 0: R(phi, theta) = c000
    return cs


def test_contains_sphere():   # shell.go:349-354 with PennaFunc == const
    cs = sphere_coeffs(2.0)
    assert po.shell_contains(cs, 1.0, 1.0, 1.0)          # r = 1.73
    assert not po.shell_contains(cs, 1.2, 1.2, 1.2)      # r = 2.08
    assert po.shell_contains(cs, 0.0, 0.0, -1.999)
    assert not po.shell_contains(cs, 0.0, -2.001, 0.0)


def test_contains_penna_terms():
    """R = c000 + c001 cos(theta) (k = 1 term, idx = order*order): an offset sphere-like surface."""
    order = 3
    cs = np.zeros(18)
    cs[0], cs[9] = 1.0, 0.5
    assert po.shell_contains(cs, 0, 0, 1.4) and not po.shell_contains(cs, 0, 0, 1.6)      # theta = 0: R = 1.5
    assert po.shell_contains(cs, 0, 0, -0.4) and not po.shell_contains(cs, 0, 0, -0.6)    # theta = pi: R = 0.5
    assert po.shell_contains(cs, 0.9, 0, 0) and not po.shell_contains(cs, 1.1, 0, 0)      # theta = pi/2: R = 1


def test_mass_contained_counts_a_sphere():
    rng = np.random.default_rng(3)
    n, box = 200000, 20.0
    xs = (rng.random((n, 3)) * box).astype(np.float32)
    ms = np.full(n, 2.5, np.float32)
    c = np.array([10.0, 10.0, 10.0], np.float32)
    d = xs - c
    r = np.sqrt((d.astype(np.float64) ** 2).sum(1))
    want = 2.5 * np.count_nonzero(r < 3.0)
    got = po.mass_contained(box, xs, ms, sphere_coeffs(3.0), c, 2.0, 4.0)
    assert abs(got - want) <= 2.5 * 2   # float32 r2 vs float64 r at the surface
    # rLow above the surface counts everything below it (cmd/stats.go:538: r2 < low2 short-circuits)
    got = po.mass_contained(box, xs, ms, sphere_coeffs(3.0), c, 5.0, 6.0)
    assert abs(got - 2.5 * np.count_nonzero(r < 5.0)) <= 2.5 * 2
    # workers only regroup the float64 sum
    a = po.mass_contained(box, xs, ms, sphere_coeffs(3.0), c, 2.0, 4.0, workers=1)
    b = po.mass_contained(box, xs, ms, sphere_coeffs(3.0), c, 2.0, 4.0, workers=7)
    assert abs(a - b) <= 1e-9 * a


def test_mass_wrap_moves_by_half_the_box():
    """cmd/stats.go:429-436 subtracts tw/2, not tw (SURVEY hazard 15): a particle
    half a box (+0.1) away along x is counted as if it were 0.1 from the centre,
    while the true periodic image one full box away is not."""
    box = 100.0
    c = np.array([10.0, 50.0, 50.0], np.float32)
    cs = sphere_coeffs(1.0)
    ghost = np.array([[60.1, 50.0, 50.0]], np.float32)       # dx = 50.1 -> 0.1
    image = np.array([[99.95, 50.0, 50.0]], np.float32)      # true periodic distance 10.05; dx = 89.95 -> 39.95
    near = np.array([[10.5, 50.0, 50.0]], np.float32)
    one = np.ones(1, np.float32)
    assert po.mass_contained(box, ghost, one, cs, c, 0.5, 2.0) == 1.0
    assert po.mass_contained(box, image, one, cs, c, 0.5, 2.0) == 0.0
    assert po.mass_contained(box, near, one, cs, c, 0.5, 2.0) == 1.0


def test_sphere_sheet_intersect():   # cmd/stats.go:685-692
    assert po.sphere_sheet_intersect([5, 5, 5], 1.0, [4, 4, 4], [2, 2, 2], 100.0)
    assert not po.sphere_sheet_intersect([5, 5, 5], 1.0, [7, 4, 4], [2, 2, 2], 100.0)
    assert po.sphere_sheet_intersect([0.5, 5, 5], 1.0, [98, 4, 4], [2, 2, 2], 100.0)   # periodic neighbour


def penna_numpy(cs, order, phi, th):
    """PennaFunc, penna.go:62-82, written independently with numpy broadcasting."""
    r = np.zeros_like(phi)
    idx = 0
    for k in range(2):
        for j in range(order):
            for i in range(order):
                r = r + cs[idx] * np.sin(th) ** (i + j) * np.cos(th) ** k * np.sin(phi) ** j * np.cos(phi) ** i
                idx += 1
    return r


def test_contains_general_shell_against_numpy():
    rng = np.random.default_rng(11)
    for order in (2, 3, 4):
        cs = np.zeros(2 * order * order)
        cs[0] = 1.0
        cs[1:] = rng.normal(0, 0.08, len(cs) - 1)
        pts = rng.normal(size=(4000, 3)) * 0.7
        r = np.linalg.norm(pts, axis=1)
        phi = np.arctan2(pts[:, 1], pts[:, 0])
        th = np.arccos(pts[:, 2] / r)
        surf = penna_numpy(cs, order, phi, th)
        want = surf > r
        got = np.array([po.shell_contains(cs, *p) for p in pts])
        sure = np.abs(surf - r) > 1e-9          # away from the surface both must agree
        assert np.array_equal(got[sure], want[sure]) and sure.mean() > 0.999
        assert 0.2 < want.mean() < 0.95         # the sample straddles the surface


def test_mass_contained_general_shell_against_numpy():
    rng = np.random.default_rng(12)
    order, box = 3, 50.0
    cs = np.zeros(18)
    cs[0] = 2.0
    cs[1:] = rng.normal(0, 0.15, 17)
    c = np.array([25.0, 24.0, 26.0], np.float32)
    xs = (c + rng.normal(size=(50000, 3)) * 1.5).astype(np.float32)
    ms = rng.uniform(1, 2, len(xs)).astype(np.float32)
    d = (xs - c).astype(np.float64)                      # no wrap needed: all within the box centre
    r = np.linalg.norm(d, axis=1)
    surf = penna_numpy(cs, order, np.arctan2(d[:, 1], d[:, 0]), np.arccos(d[:, 2] / r))
    lo, hi = 0.9 * surf.min(), 1.1 * surf.max()
    want = ms[(r < lo) | ((r < hi) & (surf > r))].astype(np.float64).sum()
    got = po.mass_contained(box, xs, ms, cs, c, lo, hi)
    assert abs(got - want) <= 1e-6 * want               # a particle within 1e-9 of the surface may flip


def test_shell_particles_closed_form_for_a_sphere():
    """appendShellParticlesChan (cmd/stats.go:545-597) on a spherical shell R = 1: the band keeps
    |r - 1| < ShellWidth * R200m, a negative ShellWidth keeps r < 1; the half-box wrap applies."""
    rng = np.random.default_rng(8)
    xs = (rng.random((20000, 3)) * 6 - 3 + 10).astype(np.float32)
    c = np.array([10.0, 10.0, 10.0], np.float32)
    cs = np.zeros(18)
    cs[0] = 1.0
    d = (xs - c).astype(np.float64)
    r = np.sqrt((d.astype(np.float32) ** 2).sum(1).astype(np.float32).astype(np.float64))
    band = po.shell_particles(100.0, xs, cs, c, 0.8, 0.9, 1.1, 0.25)     # delta = 0.8 * 0.25 = 0.2
    want = np.flatnonzero((r > 0.8) & (r < 1.2) & (r > 0.9 - 0.2) & (r < 1.1 + 0.2))
    assert np.array_equal(band, want) and len(band) > 100
    inside = po.shell_particles(100.0, xs, cs, c, 0.8, 0.9, 1.1, -1.0)
    assert np.array_equal(inside, np.flatnonzero(r < 1.0))
