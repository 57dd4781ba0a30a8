"""Synthetic NFW-plus-infall snapshots for parity tests and benchmarks.

Shapes follow BASELINE.json's configs: particles of a
periodic box cut into cubic blocks, each block carrying the io.Header fields
(io/io.go:71-83) the `shell` loop reads, plus a halo table (x, y, z, R200m).
Everything is seeded numpy; nothing here touches the GPU.
"""
import math

import numpy as np

# Planck-like cosmology used by every synthetic snapshot.
COSMO = dict(Z=0.0, OmegaM=0.27, OmegaL=0.73, H100=0.7)

# cosmo/const.go:3-8
_G, _MPC, _MSUN = 6.67259e-11, 3.08560e22, 1.98900e30


def rho_m0(cosmo=COSMO):
    """Mean matter density today in Msun h^2 / Mpc^3 (cosmo/param.go:36)."""
    h0 = (100.0 * 1000.0) / _MPC
    rho_c = 3.0 * h0 * h0 / (8.0 * math.pi * _G) * _MPC ** 3 / _MSUN
    return rho_c * cosmo["OmegaM"]


def uniform_mass(n_total, box, cosmo=COSMO):
    """io/gotetra.go:50-54 calcUniformMass."""
    return np.float32(box ** 3 * rho_m0(cosmo) / float(n_total))


def _nfw_radii(rng, n, c, x_max):
    """Radii (units of R200m) drawn from an NFW mass profile out to x_max."""
    def m(x):
        y = c * x
        return np.log1p(y) - y / (1.0 + y)
    grid = np.geomspace(1e-3, x_max, 4096)
    cdf = m(grid)
    cdf /= cdf[-1]
    return np.interp(rng.random(n), cdf, grid)


def halo_particles(rng, n, r200m, c=7.0, r_sp=1.25, x_max=3.4, infall_frac=0.35,
                   axes=(1.0, 0.85, 0.7)):
    """Offsets (n,3) float64 of a triaxial NFW halo truncated at r_sp*R200m
    with a rho ~ r^-1.5 infall envelope out to x_max*R200m."""
    n_in = int(round(n * (1.0 - infall_frac)))
    n_out = n - n_in
    x_in = _nfw_radii(rng, n_in, c, r_sp)
    # rho ~ r^-1.5  =>  M(<r) ~ r^1.5
    u = rng.random(n_out)
    x_out = (r_sp ** 1.5 + u * (x_max ** 1.5 - r_sp ** 1.5)) ** (1.0 / 1.5)
    x = np.concatenate([x_in, x_out])
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    a = np.asarray(axes, float)
    a = a / a.prod() ** (1.0 / 3.0)
    # random orientation of the ellipsoid
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    off = (v * x[:, None] * r200m * a[None, :]) @ q.T
    return off


def make_box(seed, box, n_halos, n_per_halo, n_background, cells, r200m_range=(0.4, 0.9),
             margin=0.0, cosmo=COSMO, centers=None, radii=None):
    """Builds one synthetic snapshot.

    Returns (halos, blocks, min_mass): halos is (H,4) float64 of x,y,z,R200m;
    blocks is a list of dicts (Z, OmegaM, OmegaL, H100, TotalWidth, Origin,
    Width, xs float32 (N,3), ms float32 (N,)) for the cells^3 cubic blocks that
    hold at least one particle; every particle carries the uniform LGadget-2 /
    gotetra mass (io/lgadget2.go:228).
    """
    rng = np.random.default_rng(seed)
    if centers is None:
        centers = margin + rng.random((n_halos, 3)) * (box - 2 * margin)
    centers = np.asarray(centers, float)
    if radii is None:
        lo, hi = r200m_range
        radii = np.exp(rng.uniform(np.log(lo), np.log(hi), n_halos))
    radii = np.asarray(radii, float)
    n_per = np.broadcast_to(np.asarray(n_per_halo), (n_halos,))
    parts = []
    for i in range(n_halos):
        off = halo_particles(rng, int(n_per[i]), radii[i])
        parts.append(centers[i][None, :] + off)
    if n_background > 0:
        parts.append(rng.random((n_background, 3)) * box)
    xs = np.concatenate(parts) if parts else np.zeros((0, 3))
    xs = np.mod(xs, box)
    xs32 = xs.astype(np.float32)
    xs32[xs32 >= np.float32(box)] = 0.0  # lgadget2.go:123-131 keeps x in [0, L)
    mass = uniform_mass(len(xs32), box, cosmo)
    w = box / cells
    idx = np.minimum((xs32 / np.float32(w)).astype(np.int64), cells - 1)
    flat = (idx[:, 0] * cells + idx[:, 1]) * cells + idx[:, 2]
    order = np.argsort(flat, kind="stable")
    flat_sorted = flat[order]
    xs_sorted = xs32[order]
    bounds = np.searchsorted(flat_sorted, np.arange(cells ** 3 + 1))
    blocks = []
    for b in range(cells ** 3):
        lo, hi = bounds[b], bounds[b + 1]
        if hi == lo:
            continue
        ix, iy, iz = b // (cells * cells), (b // cells) % cells, b % cells
        blocks.append(dict(cosmo, TotalWidth=float(box),
                           Origin=(np.float32(ix * w), np.float32(iy * w), np.float32(iz * w)),
                           Width=(np.float32(w),) * 3,
                           xs=np.ascontiguousarray(xs_sorted[lo:hi]),
                           ms=np.full(hi - lo, mass, np.float32)))
    halos = np.concatenate([centers, radii[:, None]], axis=1)
    return halos, blocks, float(mass)


def single_halo(seed=1, n=100000, r200m=1.0, box=62.5, n_background=20000, cells=1):
    """BASELINE config 1 analogue: one halo at the box centre."""
    return make_box(seed, box, 1, n, n_background, cells,
                    centers=[[box / 2] * 3], radii=[r200m])


def r200m_from_count(n200, m_p, cosmo=COSMO):
    """R200m (comoving Mpc/h) of a halo holding n200 particles of mass m_p."""
    return (3.0 * n200 * m_p / (4.0 * math.pi * 200.0 * rho_m0(cosmo))) ** (1.0 / 3.0)


def config2(seed=2, n_halos=1000, n_total=512 ** 3, box=250.0, cells=8,
            n_per_halo_range=(5.0e4, 2.0e5), cosmo=COSMO):
    """BASELINE config 2: `n_halos` halos with >= 5e4 particles each in a
    periodic `n_total`-particle LGadget-2-like box cut into cells^3 blocks.
    Per-halo particle counts are log-uniform in n_per_halo_range (scaled down
    if they would exceed 80% of n_total); R200m follows from the count so that
    densities in units of rho_m are self-consistent.  The rest of the budget is
    a uniform background."""
    rng = np.random.default_rng(seed)
    lo, hi = n_per_halo_range
    n_per = np.exp(rng.uniform(np.log(lo), np.log(hi), n_halos))
    if n_per.sum() > 0.8 * n_total:
        n_per *= 0.8 * n_total / n_per.sum()
    n_per = np.maximum(n_per.astype(np.int64), 1)
    m_p = float(uniform_mass(n_total, box, cosmo))
    radii = np.array([r200m_from_count(0.567 * n, m_p, cosmo) for n in n_per])
    centers = rng.random((n_halos, 3)) * box
    n_bg = int(n_total - n_per.sum())
    return make_box(seed + 1, box, n_halos, n_per, n_bg, cells, centers=centers, radii=radii, cosmo=cosmo)
