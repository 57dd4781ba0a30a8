"""`stats` M_sp pass on the GPU (sfk_mass_*) against the oracle: same particles
counted, float64 sums equal to 1e-12 (the reference adds its worker partial sums
in channel-arrival order, so the last bits are not defined there either)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from shellfish_b200 import api, synth

pytestmark = pytest.mark.gpu


def radial_range(cs, order, n=64):
    """Deterministic stand-in for Shell.RadialRange (Monte Carlo in the reference):
    min / max of the surface on a (phi, theta) grid, widened by 2 %."""
    phi = (np.arange(2 * n) + 0.5) * np.pi / n
    th = (np.arange(n) + 0.5) * np.pi / n
    P, T = np.meshgrid(phi, th, indexing="ij")
    R = np.zeros_like(P)
    idx = 0
    for k in range(2):
        for j in range(order):
            for i in range(order):
                R += cs[idx] * np.sin(T) ** (i + j) * np.cos(T) ** k * np.sin(P) ** j * np.cos(P) ** i
                idx += 1
    return 0.98 * R.min(), 1.02 * R.max()


def oracle_masses(halos, coeffs, r_low, r_high, blocks):
    out = np.zeros(len(halos))
    for h in range(len(halos)):
        for b in blocks:
            out[h] += po.mass_contained(b["TotalWidth"], b["xs"], b["ms"], coeffs[h], halos[h, :3].astype(np.float32),
                                        r_low[h], r_high[h], workers=1)
    return out


def test_mass_of_fitted_shells():
    """shell -> stats on one box: the Penna coefficients of the shell pass feed the mass pass."""
    halos, blocks, mass = synth.make_box(41, 62.5, 3, 30000, 20000, 2, margin=8.0)
    ctx, coeffs, status = api.run_shell(api.default_params(seed=1337, rings=12), mass, halos, blocks)
    assert np.all(status == 0)
    rng = np.random.default_rng(1)
    for b in blocks:   # unequal masses so that a wrong particle set cannot hide
        b["ms"] = (b["ms"] * rng.uniform(0.5, 2.0, len(b["ms"]))).astype(np.float32)
    lo, hi = np.array([radial_range(c, 3) for c in coeffs]).T
    got = ctx.mass_contained(halos[:, :3], coeffs, lo, hi, blocks)
    want = oracle_masses(halos, coeffs, lo, hi, blocks)
    assert np.all(want > 0)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # binSphereIntersections (cmd/stats.go:694-708) skips blocks that no R200m sphere touches --
    # even if a splashback shell, which is larger than R200m, reaches into them
    radii = halos[:, 3].astype(np.float32)
    kept = [b for b in blocks if any(po.sphere_sheet_intersect(halos[h, :3], radii[h], b["Origin"], b["Width"],
                                                               b["TotalWidth"]) for h in range(len(halos)))]
    got2 = ctx.mass_contained(halos[:, :3], coeffs, lo, hi, blocks, radii=radii)
    np.testing.assert_allclose(got2, oracle_masses(halos, coeffs, lo, hi, kept), rtol=1e-12)


def test_mass_half_box_wrap_quirk_and_many_halos():
    """Halos near the box edge: particles half a box away are counted (cmd/stats.go:429-436);
    300 halos exercise the shared-memory halo chunks; ragged particle count."""
    rng = np.random.default_rng(5)
    box, n = 40.0, 300001
    xs = (rng.random((n, 3)) * box).astype(np.float32)
    ms = rng.uniform(1.0, 3.0, n).astype(np.float32)
    blocks = [dict(xs=xs, ms=ms, TotalWidth=box, Origin=(0, 0, 0), Width=(box, box, box))]
    nh = 300
    halos = np.concatenate([rng.random((nh, 3)) * box, rng.uniform(0.5, 2.0, (nh, 1))], axis=1)
    order = 2
    coeffs = np.zeros((nh, 8))
    coeffs[:, 0] = halos[:, 3]
    coeffs[:, 1:] = rng.normal(0, 0.05, (nh, 7)) * halos[:, 3:4]
    lo, hi = np.array([radial_range(c, order) for c in coeffs]).T
    ctx = api.ShellContext(api.default_params(rings=3))
    got = ctx.mass_contained(halos[:, :3], coeffs, lo, hi, blocks)
    want = oracle_masses(halos, coeffs, lo, hi, blocks)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # the quirk is really exercised: counting with the true periodic distance gives another answer
    h = int(np.argmin(halos[:, 0]))   # a halo close to the x = 0 face
    d = xs.astype(np.float64) - halos[h, :3]
    d -= box * np.round(d / box)
    plain = ms[(d ** 2).sum(1) < lo[h] ** 2].sum()
    assert want[h] > 0 and abs(want[h] - plain) > 0.2 * plain


def test_mass_call_order():
    ctx = api.ShellContext(api.default_params(rings=3))
    with pytest.raises(api.ShellfishError, match="sfk_mass_begin first"):
        ctx._ck(api.lib().sfk_mass_finish(ctx._h, None), "sfk_mass_finish")


@pytest.mark.parametrize("width", [0.05, -1.0])
def test_shell_particle_selection_matches_the_oracle(width):
    """ShellParticleFile (cmd/stats.go:326-341, appendShellParticlesChan :545-597): the particles
    within ShellWidth * R200m of each fitted surface (or inside it for ShellWidth < 0), per block,
    in the reference's one-worker order.  Index lists must be identical."""
    halos, blocks, mass = synth.make_box(43, 30.0, 3, 25000, 20000, 2, margin=5.0)
    ctx, coeffs, status = api.run_shell(api.default_params(seed=1337, rings=10), mass, halos, blocks)
    assert np.all(status == 0)
    lo, hi = np.array([radial_range(c, 3) for c in coeffs]).T
    radii = halos[:, 3].astype(np.float32)
    masses, per_block = ctx.shell_particles(halos[:, :3], coeffs, lo, hi, radii, width, blocks)
    np.testing.assert_allclose(masses, oracle_masses(halos, coeffs, lo, hi, blocks), rtol=1e-12)
    total = 0
    for b, pairs in zip(blocks, per_block):
        want = []
        for h in range(len(halos)):
            idx = po.shell_particles(b["TotalWidth"], b["xs"], coeffs[h], halos[h, :3], radii[h], lo[h], hi[h], width)
            want += [(h << 32) | int(i) for i in idx]
        assert pairs.tolist() == want
        total += len(want)
    assert total > 1000
