"""Sliding-moment Savitzky-Golay smoothing (experimental los_kernel_fast_moments,
the default los_kernel_fast path): the host/device-shared arithmetic of
shellfish_b200/csrc/sfk_sgmath.cuh, compiled for the CPU (tests/hostcheck), against the
tap loop of Kernel.ConvolveAt (math/interpolate/filter.go:104-161) that the oracle and the
default kernel evaluate."""
import numpy as np
import pytest

import hostcheck as hc
from oracle import pyoracle as po


def direct(ys, k0, k1):
    w = len(k0)
    h = w // 2
    pad = np.concatenate([np.repeat(ys[:1], h), ys, np.repeat(ys[-1:], h)])      # Extension boundary
    win = np.lib.stride_tricks.sliding_window_view(pad, w)
    return win @ k0, win @ k1


@pytest.mark.parametrize("window", [7, 21, 121, 231])
def test_moments_equal_the_tap_loop(window):
    rng = np.random.default_rng(window)
    k0 = po.savgol_kernel(4, window, 0, 0.0)
    k1 = po.savgol_kernel(4, window, 1, 0.009)
    worst = 0.0
    for _ in range(40):
        ys = np.cumsum(rng.normal(0, 0.05, 256)) - np.linspace(0, 6, 256) + rng.uniform(-20, 20)
        v, d = direct(ys, k0, k1)
        s0, s1 = hc.sg_moments(ys, k0, k1)
        worst = max(worst, np.abs(s0 - v).max() / np.abs(v).max(), np.abs(s1 - d).max() / np.abs(d).max())
        p0, p1 = hc.sg_moments(ys, k0, k1, paired=True)     # taps at +-u folded (sg_accumulate_pair)
        worst = max(worst, np.abs(p0 - v).max() / np.abs(v).max(), np.abs(p1 - d).max() / np.abs(d).max())
    assert worst < 2e-11, worst     # relative to the curve; the rows carry offsets of up to 20


def test_moments_longer_runs_lose_accuracy_gracefully():
    """The shift is a unipotent recursion: without re-anchoring (per = 256) the error grows to
    ~1e-10, which is why the kernel restarts from a direct evaluation every 8 outputs."""
    rng = np.random.default_rng(1)
    k0, k1 = po.savgol_kernel(4, 121, 0, 0.0), po.savgol_kernel(4, 121, 1, 0.009)
    ys = np.cumsum(rng.normal(0, 0.05, 256)) - np.linspace(0, 6, 256)
    v, d = direct(ys, k0, k1)
    e8 = np.abs(hc.sg_moments(ys, k0, k1, per=8)[1] - d).max()
    e256 = np.abs(hc.sg_moments(ys, k0, k1, per=256)[1] - d).max()
    assert e8 < 1e-12 and e8 < e256 < 1e-7


def splashback_index(vals, ders, d_lim=0.0):
    """SplashbackRadius (splashback_radius.go:30-58), vectorised over lines of sight."""
    n, bins = vals.shape
    rmin = vals[:, 0].copy()
    dmin = np.full(n, d_lim)
    imin = np.full(n, -1)
    for i in range(1, bins - 1):
        new = vals[:, i] < rmin
        rmin = np.where(new, vals[:, i], rmin)
        cand = new & (ders[:, i] < ders[:, i - 1]) & (ders[:, i] < ders[:, i + 1]) & (ders[:, i] < dmin)
        dmin = np.where(cand, ders[:, i], dmin)
        imin = np.where(cand, i, imin)
    return imin


def test_splashback_decisions_do_not_change():
    """On the oracle's own profiles of a halo (12 800 lines of sight) the steepest-slope search
    picks the same bin from the moment-smoothed curves as from the tap-loop ones, and the
    tap-loop ones reproduce the oracle's R_sp."""
    from shellfish_b200 import synth
    halos, blocks, mass = synth.single_halo(seed=5, n=40000, n_background=15000)
    cfg = po.ShellConfig.default(rings=50, seed=1337, threads=4)
    ref = po.shell_run(cfg, mass, halos, blocks, taps=("profiles", "rsp"))
    prof = ref["profiles"][0].reshape(-1, 256)
    rsp = ref["rsp"][0].reshape(-1)
    r200 = halos[0, 3]
    lr0, lr1 = np.log(0.3 * r200), np.log(3.0 * r200)
    rs = np.exp(lr0 + (lr1 - lr0) / 256 * (np.arange(256) + 0.5))
    k0 = po.savgol_kernel(4, 121, 0, 0.0)
    k1 = po.savgol_kernel(4, 121, 1, np.log(rs[1]) - np.log(rs[0]))
    ys = np.log(prof)
    vd = np.zeros_like(ys); dd = np.zeros_like(ys); vm = np.zeros_like(ys); dm = np.zeros_like(ys)
    for i in range(len(ys)):
        vd[i], dd[i] = direct(ys[i], k0, k1)
        vm[i], dm[i] = hc.sg_moments(ys[i], k0, k1)
    i_d = splashback_index(np.exp(vd), dd)
    i_m = splashback_index(np.exp(vm), dm)
    ok = ~np.isnan(rsp)
    assert np.array_equal(ok, i_d >= 0)
    assert np.allclose(rs[i_d[ok]], rsp[ok], rtol=1e-12)
    assert np.array_equal(i_d, i_m)
    assert np.abs(dm - dd).max() < 1e-11 and np.abs(vm - vd).max() < 1e-11
    # exp is monotone and SplashbackRadius only compares the smoothed densities
    # (splashback_radius.go:36-41): the search can run on the smoothed logarithms
    assert np.array_equal(splashback_index(vd, dd), i_d)
