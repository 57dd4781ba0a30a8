"""Experiment behind SFK_FLAG_LOS_MOMENTS (CPU only, numpy): Savitzky-Golay smoothing by
sliding scaled moments (O(order^2) per output) vs the direct 121-tap FIR on the oracle's own
profiles -- do the SplashbackRadius decisions change?  Here WITHOUT re-anchoring (256 shifts in
a row, error ~2e-10); the kernel re-anchors every 8 outputs (error ~1e-13).
Result (5 halos, 128 000 lines of sight): 0 changed decisions.
    PYTHONPATH=. python scripts/experiments/sg_sliding_moments.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import pyoracle as po
from shellfish_b200 import synth

W, H = 121, 60
halos, blocks, mass = synth.single_halo(seed=5, n=60000, n_background=20000)
cfg = po.ShellConfig.default(rings=100, seed=1337, threads=8)
t0 = time.time()
ref = po.shell_run(cfg, mass, halos, blocks, taps=("profiles", "rsp"))
print("oracle", time.time() - t0, "s")
prof = ref["profiles"][0].reshape(-1, 256)      # GetRhos: [LoS][bin], clamped to rho_bg
rsp_ref = ref["rsp"][0].reshape(-1)
B = 256
r200 = halos[0, 3]
lrmin, lrmax = np.log(0.3 * r200), np.log(3.0 * r200)
dlr = (lrmax - lrmin) / B
rs = np.exp(lrmin + dlr * (np.arange(B) + 0.5))
dx = np.log(rs[1]) - np.log(rs[0])
k0 = po.savgol_kernel(4, W, 0, 0.0)
k1 = po.savgol_kernel(4, W, 1, dx)
ys = np.log(prof)
pad = np.concatenate([np.repeat(ys[:, :1], H, 1), ys, np.repeat(ys[:, -1:], H, 1)], 1)   # Extension boundary

def direct(pad):
    v = np.zeros((pad.shape[0], B)); d = np.zeros_like(v)
    for j in range(W):   # same tap order as ConvolveAt? (order only changes rounding)
        v += k0[j] * pad[:, j:j + B]
        d += k1[j] * pad[:, j:j + B]
    return np.exp(v), d

# taps as polynomials of the offset: k[j] = sum_k c_k ((j-H)/H)^k
jj = (np.arange(W) - H) / H
V = np.vander(jj, 5, increasing=True)
c0 = np.linalg.lstsq(V, k0, rcond=None)[0]; c1 = np.linalg.lstsq(V, k1, rcond=None)[0]
print("tap poly fit residual", np.abs(V @ c0 - k0).max() / np.abs(k0).max(), np.abs(V @ c1 - k1).max() / np.abs(k1).max())
# which convention does the FIR use: out[i] = sum_j k[j] pad[i+j]  (correlation) as coded above

def moments(pad):
    n = pad.shape[0]
    M = np.zeros((5, n))
    for k in range(5):
        M[k] = (pad[:, :W] * jj[None, :] ** k).sum(1)
    v = np.zeros((n, B)); d = np.zeros_like(v)
    from math import comb
    T = np.zeros((5, 5))
    for k in range(5):
        for l in range(k + 1):
            T[k, l] = comb(k, l) * (-1.0 / H) ** (k - l)
    lo_w = np.array([(-(H + 1) / H) ** k for k in range(5)])   # weight of the sample that leaves
    hi_w = np.array([1.0 for k in range(5)])                   # (h/h)^k of the sample that enters
    for i in range(B):
        v[:, i] = c0 @ M
        d[:, i] = c1 @ M
        if i + 1 < B:
            M = T @ M - lo_w[:, None] * pad[None, :, i] + hi_w[:, None] * pad[None, :, i + W]
    return np.exp(v), d

def splashback(rhos, derivs, dlim=0.0):
    """splashback_radius.go:30-58 vectorised over LoS."""
    n = rhos.shape[0]
    rmin = rhos[:, 0].copy(); dmin = np.full(n, dlim); imin = np.full(n, -1)
    for i in range(1, B - 1):
        new = rhos[:, i] < rmin
        rmin = np.where(new, rhos[:, i], rmin)
        cand = new & (derivs[:, i] < derivs[:, i - 1]) & (derivs[:, i] < derivs[:, i + 1]) & (derivs[:, i] < dmin)
        dmin = np.where(cand, derivs[:, i], dmin); imin = np.where(cand, i, imin)
    return imin

vd, dd = direct(pad)
vm, dm = moments(pad)
print("max rel diff vals", np.nanmax(np.abs(vm / vd - 1)), "derivs abs", np.nanmax(np.abs(dm - dd)), "scale", np.nanmax(np.abs(dd)))
i_d, i_m = splashback(vd, dd), splashback(vm, dm)
r_d = np.where(i_d >= 0, rs[np.clip(i_d, 0, B - 1)], np.nan)
ok = ~np.isnan(rsp_ref)
print("direct numpy vs oracle: ok-mask equal", np.array_equal(ok, i_d >= 0), "rsp mismatches", np.sum(np.abs(r_d[ok] / rsp_ref[ok] - 1) > 1e-12))
print("moments vs direct: differing LoS", np.sum(i_d != i_m), "of", len(i_d))

tot=0; flips=0
for seed in range(6, 10):
    halos, blocks, mass = synth.single_halo(seed=seed, n=40000, n_background=15000)
    ref = po.shell_run(cfg, mass, halos, blocks, taps=("profiles",))
    ys = np.log(ref["profiles"][0].reshape(-1, 256))
    pad = np.concatenate([np.repeat(ys[:, :1], H, 1), ys, np.repeat(ys[:, -1:], H, 1)], 1)
    vd, dd = direct(pad); vm, dm = moments(pad)
    a, b = splashback(vd, dd), splashback(vm, dm)
    tot += len(a); flips += np.sum(a != b)
    print(seed, "flips", np.sum(a != b), "maxerr", np.nanmax(np.abs(dm - dd)), np.nanmax(np.abs(vm / vd - 1)))
print("total", tot, "flips", flips)
