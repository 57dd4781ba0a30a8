"""ctypes bindings for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from shellfish_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


def build():
    """Compile oracle/liboracle.so with the committed Makefile."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_rho_average.restype = C.c_double
        L.orc_rho_average.argtypes = [C.c_double] * 4
        L.orc_go_pow.restype = C.c_double
        L.orc_go_pow.argtypes = [C.c_double] * 2
        L.orc_half_angular_width.restype = C.c_double
        L.orc_half_angular_width.argtypes = [C.c_double] * 2
        L.orc_one_val_intr_dist.restype = C.c_double
        L.orc_one_val_intr_dist.argtypes = [C.c_double] * 4
        L.orc_two_val_intr_dist.argtypes = [C.c_double] * 3 + [c_double_p] * 2
        L.orc_ring_new.restype = C.c_void_p
        L.orc_ring_new.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_ring_free.argtypes = [C.c_void_p]
        L.orc_ring_insert.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_ring_retrieve.argtypes = [C.c_void_p, C.c_int, c_double_p]
        L.orc_ring_derivs.restype = c_double_p
        L.orc_ring_derivs.argtypes = [C.c_void_p]
        L.orc_halo_new.restype = C.c_void_p
        L.orc_halo_new.argtypes = [c_float_p, C.c_int, c_double_p, C.c_double, C.c_double,
                                   C.c_int, C.c_int, C.c_double]
        L.orc_halo_free.argtypes = [C.c_void_p]
        L.orc_halo_enable_taps.argtypes = [C.c_void_p]
        L.orc_halo_transform.argtypes = [C.c_void_p, c_float_p, C.c_int64, C.c_double]
        L.orc_halo_intersect.argtypes = [C.c_void_p, c_float_p, C.c_int64, C.c_double,
                                         C.POINTER(C.c_uint8)]
        L.orc_halo_insert.argtypes = [C.c_void_p, c_float_p, C.c_double, C.c_double]
        L.orc_halo_insert_to_ring.argtypes = [C.c_void_p, c_float_p, C.c_double, C.c_double, C.c_int]
        L.orc_halo_sphere_intersect_ring.restype = C.c_int
        L.orc_halo_sphere_intersect_ring.argtypes = [C.c_void_p, c_float_p, C.c_double, C.c_int]
        L.orc_halo_idx_range.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_halo_get_rhos.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.orc_halo_get_rs.argtypes = [C.c_void_p, c_double_p]
        L.orc_halo_sheet_intersect.restype = C.c_int
        L.orc_halo_sheet_intersect.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_double]
        L.orc_halo_plane_to_volume.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, c_double_p]
        L.orc_halo_derivs.restype = c_double_p
        L.orc_halo_derivs.argtypes = [C.c_void_p, C.c_int]
        L.orc_halo_rot.restype = c_float_p
        L.orc_halo_rot.argtypes = [C.c_void_p, C.c_int]
        L.orc_halo_irot.restype = c_float_p
        L.orc_halo_irot.argtypes = [C.c_void_p, C.c_int]
        L.orc_spline_new.restype = C.c_void_p
        L.orc_spline_new.argtypes = [c_double_p, c_double_p, C.c_int]
        L.orc_spline_free.argtypes = [C.c_void_p]
        L.orc_spline_eval.restype = C.c_int
        L.orc_spline_eval.argtypes = [C.c_void_p, C.c_double, c_double_p]
        L.orc_savgol_kernel.restype = C.c_int
        L.orc_savgol_kernel.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, c_double_p]
        L.orc_convolve_extension.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, c_double_p]
        L.orc_lu_solve.restype = C.c_int
        L.orc_lu_solve.argtypes = [c_double_p, C.c_int, c_double_p]
        L.orc_lu_invert.restype = C.c_int
        L.orc_lu_invert.argtypes = [c_double_p, C.c_int, c_double_p]
        L.orc_smooth.restype = C.c_int
        L.orc_smooth.argtypes = [c_double_p, c_double_p, C.c_int, c_double_p, C.c_int,
                                 c_double_p, c_double_p]
        L.orc_splashback_radius.restype = C.c_int
        L.orc_splashback_radius.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int,
                                            C.c_double, c_double_p]
        L.orc_filter_ring.restype = C.c_int
        L.orc_filter_ring.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_double,
                                      c_double_p, c_double_p, C.POINTER(C.c_int), c_double_p]
        L.orc_penna_coeffs.restype = C.c_int
        L.orc_penna_coeffs.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int64,
                                       C.c_int, C.c_int, C.c_int, c_double_p]
        L.orc_penna_func.restype = C.c_double
        L.orc_penna_func.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_shell_run.restype = C.c_int
        L.orc_shell_contains.restype = C.c_int
        L.orc_shell_contains.argtypes = [c_double_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_mass_contained.restype = C.c_double
        L.orc_mass_contained.argtypes = [C.c_double, c_float_p, c_float_p, C.c_int64, c_double_p, C.c_int,
                                         c_float_p, C.c_double, C.c_double, C.c_int]
        L.orc_shell_particles.restype = C.c_int64
        L.orc_shell_particles.argtypes = [C.c_double, c_float_p, C.c_int64, c_double_p, C.c_int, c_float_p, C.c_float,
                                          C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int64)]
        L.orc_sphere_sheet_intersect.restype = C.c_int
        L.orc_sphere_sheet_intersect.argtypes = [c_float_p, C.c_float, c_float_p, c_float_p, C.c_double]
        L.orc_nth_largest.restype = C.c_double
        L.orc_nth_largest.argtypes = [c_double_p, C.c_int64, C.c_int64, c_double_p]
        L.orc_percentile.restype = C.c_double
        L.orc_percentile.argtypes = [c_double_p, C.c_int64, C.c_double, c_double_p]
        L.orc_calc_percentile.restype = None
        L.orc_calc_percentile.argtypes = [C.c_double, C.c_double, C.c_int, c_double_p, C.c_int64, C.c_double,
                                          c_double_p]
        _LIB = L
    return _LIB


def dp(a):
    return a.ctypes.data_as(c_double_p)


def fp(a):
    return a.ctypes.data_as(c_float_p)


class ShellConfig(C.Structure):
    """Mirrors orc_shell_config; defaults are cmd/shell.go:105-119."""
    _fields_ = [("radialBins", C.c_int64), ("spokes", C.c_int64), ("rings", C.c_int64),
                ("rMaxMult", C.c_double), ("rMinMult", C.c_double), ("rKernelMult", C.c_double),
                ("eta", C.c_double),
                ("order", C.c_int64), ("smoothingWindow", C.c_int64), ("levels", C.c_int64),
                ("subsampleFactor", C.c_int64),
                ("losSlopeCutoff", C.c_double), ("backgroundRhoMult", C.c_double),
                ("seed", C.c_uint64), ("threads", C.c_int64)]

    @classmethod
    def default(cls, **kw):
        c = cls(radialBins=256, spokes=256, rings=100, rMaxMult=3.0, rMinMult=0.3,
                rKernelMult=0.2, eta=10.0, order=3, smoothingWindow=121, levels=3,
                subsampleFactor=1, losSlopeCutoff=0.0, backgroundRhoMult=0.5,
                seed=1337, threads=1)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


class Block(C.Structure):
    _fields_ = [("Z", C.c_double), ("OmegaM", C.c_double), ("OmegaL", C.c_double),
                ("H100", C.c_double), ("N", C.c_int64), ("TotalWidth", C.c_double),
                ("Origin", C.c_float * 3), ("Width", C.c_float * 3),
                ("xs", c_float_p), ("ms", c_float_p)]


class Taps(C.Structure):
    _fields_ = [("rsp", c_double_p), ("profiles", c_double_p),
                ("nInsert", C.POINTER(C.c_uint32)), ("startEdges", C.POINTER(C.c_uint32)),
                ("endEdges", C.POINTER(C.c_uint32)), ("nFiltered", C.POINTER(C.c_int64)),
                ("nCandidates", C.POINTER(C.c_int64)), ("nIntersections", C.POINTER(C.c_int64))]


def set_exact_bins(on):
    """orc_set_exact_bins: float64 bins like the reference (False, default) or the exact-sum
    variant used to tell rounding-decided lines of sight apart (see orc_los.c)."""
    lib().orc_set_exact_bins(C.c_int(1 if on else 0))


def norm_vecs(seed, rings):
    out = np.zeros((rings, 3), np.float32)
    lib().orc_norm_vecs(C.c_uint64(seed), C.c_int(rings), fp(out))
    return out


def savgol_kernel(order, width, ld=0, dx=1.0):
    out = np.zeros(width, np.float64)
    rc = lib().orc_savgol_kernel(order, width, ld, dx, dp(out))
    if rc:
        raise ValueError("orc_savgol_kernel failed")
    return out


class Halo:
    """Thin handle on orc_halo (los.Halo)."""

    def __init__(self, norms, origin, rMin, rMax, bins, n, defaultRho=0.0, taps=False):
        norms = np.ascontiguousarray(norms, np.float32)
        origin = np.ascontiguousarray(origin, np.float64)
        self.rings, self.bins, self.n = len(norms), bins, n
        self._h = lib().orc_halo_new(fp(norms), len(norms), dp(origin), rMin, rMax, bins, n, defaultRho)
        if taps:
            lib().orc_halo_enable_taps(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_halo_free(self._h)
            self._h = None

    def insert(self, vec, radius, rho):
        v = np.ascontiguousarray(vec, np.float32)
        lib().orc_halo_insert(self._h, fp(v), radius, rho)

    def insert_to_ring(self, vec, radius, rho, ring):
        v = np.ascontiguousarray(vec, np.float32)
        lib().orc_halo_insert_to_ring(self._h, fp(v), radius, rho, ring)

    def sphere_intersect_ring(self, vec, radius, ring):
        v = np.ascontiguousarray(vec, np.float32)
        return bool(lib().orc_halo_sphere_intersect_ring(self._h, fp(v), radius, ring))

    def idx_range(self, lo, hi):
        out = (C.c_int * 4)()
        lib().orc_halo_idx_range(self._h, lo, hi, out)
        return tuple(out)

    def get_rhos(self, ring, los):
        buf = np.zeros(self.bins)
        lib().orc_halo_get_rhos(self._h, ring, los, dp(buf))
        return buf

    def get_rs(self):
        buf = np.zeros(self.bins)
        lib().orc_halo_get_rs(self._h, dp(buf))
        return buf

    def derivs(self, ring):
        p = lib().orc_halo_derivs(self._h, ring)
        return np.ctypeslib.as_array(p, shape=(self.n, self.bins)).copy()

    def transform(self, xs, totalWidth):
        assert xs.dtype == np.float32 and xs.flags.c_contiguous
        lib().orc_halo_transform(self._h, fp(xs), len(xs), totalWidth)

    def intersect(self, xs, rad):
        out = np.zeros(len(xs), np.uint8)
        lib().orc_halo_intersect(self._h, fp(xs), len(xs), rad, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def sheet_intersect(self, origin, width, tw):
        o = np.ascontiguousarray(origin, np.float32)
        w = np.ascontiguousarray(width, np.float32)
        return bool(lib().orc_halo_sheet_intersect(self._h, fp(o), fp(w), tw))

    def rot(self, ring):
        return np.ctypeslib.as_array(lib().orc_halo_rot(self._h, ring), shape=(9,)).copy()

    def irot(self, ring):
        return np.ctypeslib.as_array(lib().orc_halo_irot(self._h, ring), shape=(9,)).copy()


def shell_run(cfg, minMass, halos, blocks, taps=(), reset_cache=True):
    """Runs orc_shell_run.

    halos: (H,4) float64 array of x,y,z,r200m.  blocks: list of dicts with
    keys Z, OmegaM, OmegaL, H100, TotalWidth, Origin, Width, xs (N,3) f32,
    ms (N,) f32.  taps: names from {"rsp","profiles","nInsert","startEdges",
    "endEdges","nFiltered","nCandidates","nIntersections"}.
    Returns dict(coeffs, ok, status, <taps...>).
    """
    L = lib()
    if reset_cache:
        L.orc_reset_kernel_cache()
    halos = np.ascontiguousarray(halos, np.float64)
    H = len(halos)
    hx, hy, hz, hr = (np.ascontiguousarray(halos[:, i]) for i in range(4))
    barr = (Block * len(blocks))()
    keep = []
    for i, b in enumerate(blocks):
        xs = np.ascontiguousarray(b["xs"], np.float32)
        ms = np.ascontiguousarray(b["ms"], np.float32)
        keep += [xs, ms]
        barr[i].Z, barr[i].OmegaM, barr[i].OmegaL, barr[i].H100 = b["Z"], b["OmegaM"], b["OmegaL"], b["H100"]
        barr[i].N = len(xs)
        barr[i].TotalWidth = b["TotalWidth"]
        for k in range(3):
            barr[i].Origin[k] = b["Origin"][k]
            barr[i].Width[k] = b["Width"][k]
        barr[i].xs, barr[i].ms = fp(xs), fp(ms)
    R, n, B = int(cfg.rings), int(cfg.spokes), int(cfg.radialBins)
    P2 = int(2 * cfg.order * cfg.order)
    out = {"coeffs": np.zeros((H, P2)), "ok": np.zeros(H, np.uint8), "status": np.zeros(H, np.int32)}
    t = Taps()
    shapes = {"rsp": ((H, R, n), np.float64), "profiles": ((H, R, n, B), np.float64),
              "nInsert": ((H, R, n), np.uint32), "startEdges": ((H, R, n, B), np.uint32),
              "endEdges": ((H, R, n, B), np.uint32), "nFiltered": ((H,), np.int64),
              "nCandidates": ((H,), np.int64), "nIntersections": ((H,), np.int64)}
    for name in taps:
        shp, dt = shapes[name]
        a = np.zeros(shp, dt)
        out[name] = a
        setattr(t, name, a.ctypes.data_as(dict(Taps._fields_)[name]))
    rc = L.orc_shell_run(C.byref(cfg), C.c_float(minMass), C.c_int64(H), dp(hx), dp(hy), dp(hz), dp(hr),
                         C.c_int64(len(blocks)), barr, dp(out["coeffs"]),
                         out["ok"].ctypes.data_as(C.POINTER(C.c_uint8)),
                         out["status"].ctypes.data_as(C.POINTER(C.c_int32)), C.byref(t))
    if rc:
        raise RuntimeError("orc_shell_run failed: %d" % rc)
    return out


def nth_largest(xs, n):
    """math/sort/median.go:40-65 (1-indexed)."""
    xs = np.ascontiguousarray(xs, np.float64)
    buf = np.empty_like(xs)
    return lib().orc_nth_largest(dp(xs), len(xs), n, dp(buf))


def percentile(xs, p):
    """math/sort/median.go:8-22, p in [0, 1]."""
    xs = np.ascontiguousarray(xs, np.float64)
    buf = np.empty_like(xs)
    return lib().orc_percentile(dp(xs), len(xs), p, dp(buf))


def calc_percentile(r200m, profiles, percentile_value, rMinMult=0.3, rMaxMult=3.0):
    """calcPercentile (cmd/shell.go:629-660) of one halo from its GetRhos tap
    [rings][spokes][bins]: returns [2*bins] = radii, percentile densities."""
    prof = np.ascontiguousarray(profiles, np.float64)
    bins = prof.shape[-1]
    prof = prof.reshape(-1, bins)
    out = np.zeros(2 * bins)
    lib().orc_calc_percentile(r200m * rMinMult, r200m * rMaxMult, bins, dp(prof), len(prof), percentile_value, dp(out))
    return out


def shell_contains(cs, x, y, z):
    """Shell.Contains, los/analyze/shell.go:349-354 (order from the coefficient count)."""
    cs = np.ascontiguousarray(cs, np.float64)
    order = int(round((len(cs) / 2) ** 0.5))
    return bool(lib().orc_shell_contains(dp(cs), order, x, y, z))


def mass_contained(total_width, xs, ms, cs, center, r_low, r_high, workers=1):
    """massContained, cmd/stats.go:451-543, for one halo over one block."""
    xs = np.ascontiguousarray(xs, np.float32)
    ms = np.ascontiguousarray(ms, np.float32)
    cs = np.ascontiguousarray(cs, np.float64)
    c = np.ascontiguousarray(center, np.float32)
    order = int(round((len(cs) / 2) ** 0.5))
    return lib().orc_mass_contained(total_width, fp(xs), fp(ms), len(ms), dp(cs), order, fp(c), r_low, r_high,
                                    workers)


def shell_particles(total_width, xs, cs, center, sphere_r, r_low, r_high, shell_width):
    """appendShellParticles (cmd/stats.go:482-597) for one halo over one block, one worker:
    indices of the selected particles in block order."""
    xs = np.ascontiguousarray(xs, np.float32)
    cs = np.ascontiguousarray(cs, np.float64)
    c = np.ascontiguousarray(center, np.float32)
    order = int(round((len(cs) / 2) ** 0.5))
    out = np.zeros(len(xs), np.int64)
    m = lib().orc_shell_particles(total_width, fp(xs), len(xs), dp(cs), order, fp(c), float(np.float32(sphere_r)),
                                  r_low, r_high, shell_width, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out[:m]


def sphere_sheet_intersect(center, r, origin, width, total_width):
    c = np.ascontiguousarray(center, np.float32)
    o = np.ascontiguousarray(origin, np.float32)
    w = np.ascontiguousarray(width, np.float32)
    return bool(lib().orc_sphere_sheet_intersect(fp(c), float(r), fp(o), fp(w), total_width))


# ---- `stats` shell property integrals (orc_shellprops.c) ------------------------------

class _Stream(C.Structure):
    _fields_ = [("w", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32)]


class _AxisTables(C.Structure):
    _fields_ = [("n", C.c_int), ("acs", c_double_p), ("bcs", c_double_p),
                ("acGrid", c_double_p), ("bcGrid", c_double_p), ("cGrid", c_double_p)]


def _stream(seed):
    L = lib()
    st = _Stream()
    L.orc_stream_init(C.byref(st), C.c_uint64(seed))
    return st, C.cast(L.orc_stream_next, C.c_void_p)


def _cs_order(cs):
    cs = np.ascontiguousarray(cs, np.float64)
    return cs, int(round((len(cs) / 2) ** 0.5))


def shell_volume(cs, samples, seed=1):
    """Shell.Volume, los/analyze/shell.go:59-68, on the xorshift stand-in stream."""
    L = lib()
    cs, order = _cs_order(cs)
    st, fn = _stream(seed)
    L.orc_shell_volume.restype = C.c_double
    L.orc_shell_volume.argtypes = [c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L.orc_shell_volume(dp(cs), order, samples, fn, C.byref(st))


def shell_surface_area(cs, samples, seed=1):
    """Shell.SurfaceArea, los/analyze/shell.go:229-237."""
    L = lib()
    cs, order = _cs_order(cs)
    st, fn = _stream(seed)
    L.orc_shell_surface_area.restype = C.c_double
    L.orc_shell_surface_area.argtypes = [c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return L.orc_shell_surface_area(dp(cs), order, samples, fn, C.byref(st))


def shell_radial_range(cs, samples, seed=1):
    """Shell.RadialRange, los/analyze/shell.go:268-283."""
    L = lib()
    cs, order = _cs_order(cs)
    st, fn = _stream(seed)
    lo, hi = C.c_double(), C.c_double()
    L.orc_shell_radial_range.argtypes = [c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_double_p, c_double_p]
    L.orc_shell_radial_range(dp(cs), order, samples, fn, C.byref(st), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def cos_norm(cs, phi, theta):
    """cosNorm, los/analyze/shell.go:201-226."""
    L = lib()
    cs, order = _cs_order(cs)
    L.orc_cos_norm.restype = C.c_double
    L.orc_cos_norm.argtypes = [c_double_p, C.c_int, C.c_double, C.c_double]
    return L.orc_cos_norm(dp(cs), order, phi, theta)


def penna_func(cs, phi, theta):
    """PennaFunc(cs, order, order, 2)(phi, theta), los/analyze/penna.go:62-82."""
    L = lib()
    cs, order = _cs_order(cs)
    L.orc_penna_func.restype = C.c_double
    L.orc_penna_func.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    return L.orc_penna_func(dp(cs), order, order, 2, phi, theta)


def bicubic_eval(xs, ys, vals, x, y):
    """BiCubic.Eval, math/interpolate/cubic_interpolators.go:92-103; None where Go panics."""
    L = lib()
    xs = np.ascontiguousarray(xs, np.float64)
    ys = np.ascontiguousarray(ys, np.float64)
    vals = np.ascontiguousarray(vals, np.float64)
    out = C.c_double()
    L.orc_bicubic_eval.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_double, C.c_double,
                                   c_double_p]
    rc = L.orc_bicubic_eval(dp(xs), len(xs), dp(ys), len(ys), dp(vals), x, y, C.byref(out))
    return None if rc else out.value


def shell_axes(cs, samples, seed=1, tables=None):
    """Shell.Axes, los/analyze/shell.go:113-199 -> (a, b, c, aVec) or None where Go
    panics.  tables = (ratios, acGrid, bcGrid, cGrid) or None for the uncorrected axes."""
    L = lib()
    cs, order = _cs_order(cs)
    st, fn = _stream(seed)
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    vec = np.zeros(3)
    tp = None
    if tables is not None:
        keep = [np.ascontiguousarray(t, np.float64) for t in tables]
        tp = _AxisTables(len(keep[0]), dp(keep[0]), dp(keep[0]), dp(keep[1]), dp(keep[2]), dp(keep[3]))
    L.orc_shell_axes.argtypes = [c_double_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 c_double_p, c_double_p, c_double_p, c_double_p]
    rc = L.orc_shell_axes(dp(cs), order, samples, fn, C.byref(st), C.byref(tp) if tp is not None else None,
                          C.byref(a), C.byref(b), C.byref(c), dp(vec))
    return None if rc else (a.value, b.value, c.value, vec)


def angular_fraction_profile(cs, samples, bins, r_min, r_max, seed=1):
    """Shell.AngularFractionProfile, los/analyze/shell.go:295-335 -> (rs, fs)."""
    L = lib()
    cs, order = _cs_order(cs)
    st, fn = _stream(seed)
    rs, fs = np.zeros(bins), np.zeros(bins)
    L.orc_angular_fraction_profile.argtypes = [c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                               C.c_void_p, C.c_void_p, c_double_p, c_double_p]
    L.orc_angular_fraction_profile.restype = None
    L.orc_angular_fraction_profile(dp(cs), order, samples, bins, r_min, r_max, fn, C.byref(st), dp(rs), dp(fs))
    return rs, fs
