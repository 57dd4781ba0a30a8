"""ctypes loader for tests/hostcheck/libhostcheck.so (TEST INFRASTRUCTURE): the
host/device-shared arithmetic of shell_props_kernel compiled for the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libhostcheck.so"))
        L.hc_penna_radius.restype = C.c_double
        L.hc_penna_radius.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        L.hc_cos_norm.restype = C.c_double
        L.hc_cos_norm.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        L.hc_shell_sums.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        L.hc_shell_props.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        L.hc_angular_fraction.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _dp]
        L.hc_sg_moments.argtypes = [_dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
        L.hc_gauss_legendre.argtypes = [C.c_int, _dp, _dp]
        _LIB = L
    return _LIB


def _cs(cs):
    cs = np.ascontiguousarray(cs, np.float64)
    return cs, int(round((len(cs) / 2) ** 0.5))


def penna_radius(cs, phi, theta):
    cs, order = _cs(cs)
    return lib().hc_penna_radius(cs.ctypes.data_as(_dp), order, phi, theta)


def cos_norm(cs, phi, theta):
    cs, order = _cs(cs)
    return lib().hc_cos_norm(cs.ctypes.data_as(_dp), order, phi, theta)


def shell_sums(cs, n_theta=128, n_phi=256):
    cs, order = _cs(cs)
    out = np.zeros(13)
    lib().hc_shell_sums(cs.ctypes.data_as(_dp), order, n_theta, n_phi, out.ctypes.data_as(_dp))
    return out


def shell_props(cs, n_theta=128, n_phi=256):
    cs, order = _cs(cs)
    out = np.zeros(11)
    lib().hc_shell_props(cs.ctypes.data_as(_dp), order, n_theta, n_phi, out.ctypes.data_as(_dp))
    return out


def gauss_legendre(n):
    x, w = np.zeros(n), np.zeros(n)
    lib().hc_gauss_legendre(n, x.ctypes.data_as(_dp), w.ctypes.data_as(_dp))
    return x, w


def angular_fraction(cs, r_min, r_max, bins, n_theta=1024, n_phi=256):
    cs, order = _cs(cs)
    out = np.zeros(bins)
    lib().hc_angular_fraction(cs.ctypes.data_as(_dp), order, n_theta, n_phi, r_min, r_max, bins, out.ctypes.data_as(_dp))
    return out


def sg_moments(ys, taps0, taps1, per=8, paired=False):
    """Savitzky-Golay value / derivative of one ln(rho) row by the kernel's sliding moments
    (paired: the anchor folds the taps at +-u, sg_accumulate_pair)."""
    ys = np.ascontiguousarray(ys, np.float64)
    t0 = np.ascontiguousarray(taps0, np.float64)
    t1 = np.ascontiguousarray(taps1, np.float64)
    s0, s1 = np.zeros(len(ys)), np.zeros(len(ys))
    rc = lib().hc_sg_moments(ys.ctypes.data_as(_dp), len(ys), t0.ctypes.data_as(_dp), t1.ctypes.data_as(_dp), len(t0),
                             -per if paired else per, s0.ctypes.data_as(_dp), s1.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("savgol_poly failed")
    return s0, s1
