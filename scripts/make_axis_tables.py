#!/usr/bin/env python
"""Generates shellfish_b200/csrc/sfk_axis_tables.inc: the axis-ratio correction
tables Shell.Axes interpolates in (los/analyze/shell.go:174-176, 358-374).

The reference ships these as three 24 x 24 grids that came out of a two-step
pipeline (los/analyze/ellipse_grid/scripts): make_grid.go ran the UNCORRECTED
Shell.Axes on true ellipsoids (1, a, b) for a <= b on a 30 x 30 grid in
[0.2, 1] with 10^6 Monte-Carlo samples each and stored true / measured ratios
with 4 significant digits; interplote.py interpolated the strictly-upper
triangle linearly onto the ratios k/29, k = 6..29, symmetrised with max(),
copied the left neighbour onto the diagonal and flipped the c grid top to
bottom.  This script recomputes every step from first principles -- the
moments by a Gauss-Legendre x uniform-phi quadrature instead of Monte Carlo, so
the tables carry no sampling noise -- and keeps the pipeline's quirks, because
they are part of what the reference's interpolators return.  The tables are
indexed by the reference with the MEASURED ratios although they were built on
the TRUE ones; that is reproduced as well.

Needs numpy + scipy (both in the image).  Run from the repo root:
    python scripts/make_axis_tables.py
"""
import os
import sys

import numpy as np
from scipy.interpolate import griddata

N_TRUE = 30
N_THETA, N_PHI = 96, 192


def ellipsoid(a, b, c):
    """make_grid.go:17-25"""
    def shell(phi, th):
        sp, cp = np.sin(phi), np.cos(phi)
        st, ct = np.sin(th), np.cos(th)
        return 1 / np.sqrt(cp * cp * st * st / (a * a) + sp * sp * st * st / (b * b) + ct * ct / (c * c))
    return shell


def cartesian(phi, th, r):
    return r * np.sin(th) * np.cos(phi), r * np.sin(th) * np.sin(phi), r * np.cos(th)


def cos_norm(s, phi, th):
    """shell.go:201-226"""
    dp = dt = 1e-3
    p00 = cartesian(phi - dp, th - dt, s(phi - dp, th - dt))
    p01 = cartesian(phi - dp, th + dt, s(phi - dp, th + dt))
    p10 = cartesian(phi + dp, th - dt, s(phi + dp, th - dt))
    p11 = cartesian(phi + dp, th + dt, s(phi + dp, th + dt))
    da = [p00[i] - p11[i] for i in range(3)]
    db = [p01[i] - p10[i] for i in range(3)]
    n = [da[1] * db[2] - da[2] * db[1], da[2] * db[0] - da[0] * db[2], da[0] * db[1] - da[1] * db[0]]
    norm = np.sqrt(n[0] ** 2 + n[1] ** 2 + n[2] ** 2)
    l = cartesian(phi, th, 1.0)
    return (l[0] * n[0] + l[1] * n[1] + l[2] * n[2]) / norm


def uncorrected_axes(s, phi, th, w):
    """Shell.Axes (shell.go:113-172) without the correction tables, the sampling
    loop replaced by the quadrature (phi, th, w)."""
    r = s(phi, th)
    area = w * r * r / cos_norm(s, phi, th)
    x, y, z = cartesian(phi, th, r)
    norm = area.sum()
    nxx, nyy, nzz = (area * x * x).sum() / norm, (area * y * y).sum() / norm, (area * z * z).sum() / norm
    nxy, nyz, nzx = (area * x * y).sum() / norm, (area * y * z).sum() / norm, (area * z * x).sum() / norm
    nx, ny, nz = (area * x).sum() / norm, (area * y).sum() / norm, (area * z).sum() / norm
    mat = np.array([
        [nyy + nzz - ny * ny - nz * nz, -nxy + nx * ny, -nzx + nz * nx],
        [-nxy + nx * ny, nxx + nzz - nx * nx - nz * nz, -nyz + ny * nz],
        [-nzx + nz * nx, -nyz + ny * nz, nxx + nyy - nx * nx - ny * ny]])
    ix, iy, iz = np.linalg.eigvalsh(mat)
    ax2 = 3 * (iy + iz - ix) / 2
    axes = np.sqrt(np.sort([ax2, 3 * iy - ax2, 3 * iz - ax2]))
    return axes[2], axes[1], axes[0]     # c (major), b, a


def g4(v):
    """make_grid.go prints with %8.4g; data.py holds those digits."""
    return float("%8.4g" % v)


def main():
    u, gw = np.polynomial.legendre.leggauss(N_THETA)
    phi1 = 2 * np.pi * (np.arange(N_PHI) + 0.5) / N_PHI
    th, phi = np.meshgrid(np.arccos(u), phi1, indexing="ij")
    w = np.repeat(gw[:, None], N_PHI, 1) * 0.5 / N_PHI
    th, phi, w = th.ravel(), phi.ravel(), w.ravel()

    # make_grid.go:40-79
    true = [0.2 + 0.8 * i / (N_TRUE - 1) for i in range(N_TRUE)]
    gas = np.zeros((N_TRUE, N_TRUE)); gbs = np.zeros_like(gas); cs = np.zeros_like(gas)
    for ia, a in enumerate(true):
        for ib, b in enumerate(true):
            if ia > ib:
                continue
            oc, ob, oa = uncorrected_axes(ellipsoid(1, a, b), phi, th, w)
            cs[ia, ib] = g4(1 / oc)
            gas[ia, ib] = g4(a / (oa / oc))
            gbs[ia, ib] = g4(b / (ob / oc))
    knots30 = np.array([g4(v) for v in true])

    # interplote.py:11-18
    bc, ac = np.meshgrid(knots30, knots30)
    bc, ac = bc.ravel(), ac.ravel()
    okay = bc > ac
    pts = np.stack([ac[okay], bc[okay]], 1)
    n = N_TRUE
    g_bc, g_ac = np.mgrid[n // 5:n, n // 5:n] / float(n - 1)

    def regrid(vals, fix00):
        grid = griddata(pts, vals.ravel()[okay], (g_ac, g_bc), method="linear")
        grid[np.isnan(grid)] = 0
        grid = np.maximum(grid, grid.T)
        for i in range(len(grid)):
            grid[i, i] = grid[i, i - 1]
        if fix00:
            grid[0, 0] = grid[0, 1]
        return grid

    ac_grid = regrid(gas, False)
    bc_grid = regrid(gbs, True)
    c_grid = regrid(cs, True)[::-1, :]
    ratios = g_ac[0]

    def arr(name, xs):
        out = ["static const double %s[] = {" % name]
        xs = ["%.5g" % v for v in xs]      # print_go_slice: 5 significant digits
        for i in range(0, len(xs), 8):
            out.append("    " + ", ".join(xs[i:i + 8]) + ",")
        out.append("};")
        return "\n".join(out)

    text = "\n".join([
        "// Generated by scripts/make_axis_tables.py -- do not edit.",
        "// Axis-ratio correction tables of Shell.Axes (los/analyze/shell.go:174-176):",
        "// knots (the same for both axes) and the three grids, vals[n * yi + xi].",
        "static const int kAxisN = %d;" % len(ratios),
        arr("kAxisRatios", ratios), arr("kACGrid", ac_grid.ravel()), arr("kBCGrid", bc_grid.ravel()),
        arr("kCGrid", c_grid.ravel()), ""])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "shellfish_b200", "csrc",
                       "sfk_axis_tables.inc")
    if len(sys.argv) > 1:
        dst = sys.argv[1]
    with open(dst, "w") as f:
        f.write(text)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
