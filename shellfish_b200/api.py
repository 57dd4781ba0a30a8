"""ctypes binding of the C ABI in include/shellfish_b200.h.

This is the Python stand-in for the cgo shim a Go maintainer would add to
cmd/shell.go: the same calls in the same order as
ShellConfig.Run / loop / sphereLoop / haloAnalysis.  There is no CPU path: if
libsfk.so is missing or no sm_100 GPU is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfk.so")

SFK_FLAG_TAPS = 1
SFK_FLAG_LOS_TAPLOOP = 2   # A/B: literal Savitzky-Golay tap loop instead of sliding moments
SFK_FLAG_NO_PRUNE = 4      # tests only: literal insertToRing / idxRange records
SFK_FLAG_KEEP_GRID = 8     # tests only: keep the int64 bin grids for ShellContext.grid
SPLIT_HIT_COUNTS, SPLIT_RHO_MAX, SPLIT_GRID = 0, 1, 2
HALO_STATUS = {0: "ok", 1: "skipped", 2: "no_points", 3: "no_maximum", 4: "spline", 5: "singular",
               6: "not_owned"}


class Params(C.Structure):
    """sfk_params == ShellConfig (cmd/shell.go:24-35) + seed/device/flags."""
    _fields_ = [("radial_bins", C.c_int64), ("spokes", C.c_int64), ("rings", C.c_int64),
                ("r_max_mult", C.c_double), ("r_min_mult", C.c_double), ("r_kernel_mult", C.c_double),
                ("eta", C.c_double),
                ("order", C.c_int64), ("smoothing_window", C.c_int64), ("levels", C.c_int64),
                ("subsample_factor", C.c_int64),
                ("los_slope_cutoff", C.c_double), ("background_rho_mult", C.c_double),
                ("percentile_profile", C.c_int32), ("percentile", C.c_double),
                ("seed", C.c_uint64), ("device", C.c_int32), ("flags", C.c_uint32)]


class Cosmo(C.Structure):
    _fields_ = [("z", C.c_double), ("omega_m", C.c_double), ("omega_l", C.c_double), ("h100", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_halos", C.c_int64), ("n_candidates", C.c_int64), ("n_ring_hits", C.c_int64),
                ("n_intersections", C.c_int64), ("n_points", C.c_int64),
                ("ms_cull", C.c_double), ("ms_hits", C.c_double), ("ms_bin", C.c_double),
                ("ms_los", C.c_double), ("ms_filter", C.c_double), ("ms_fit", C.c_double),
                ("bin_launches", C.c_int64), ("launches", C.c_int64), ("n_los_literal", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/shellfish_b200.h declares
EXPORTS = ["sfk_device_count", "sfk_default_params", "sfk_create", "sfk_destroy", "sfk_last_error", "sfk_begin_snapshot",
           "sfk_add_halos", "sfk_set_owned", "sfk_block_halo_mask", "sfk_push_block", "sfk_push_block_async", "sfk_push_sync", "sfk_push_block_device",
           "sfk_finish_snapshot", "sfk_row_length", "sfk_split_count", "sfk_split_bin", "sfk_split_buffer", "sfk_split_finish",
           "sfk_comm_unique_id", "sfk_comm_init", "sfk_split_finish_comm",
           "sfk_mass_begin", "sfk_mass_push_block", "sfk_mass_push_block_device", "sfk_mass_finish", "sfk_mass_set_band", "sfk_mass_take_band",
           "sfk_sphere_block_hit", "sfk_shell_props", "sfk_shell_sums", "sfk_angular_fraction", "sfk_axes_from_moments", "sfk_get_stats", "sfk_stream", "sfk_get_rsp", "sfk_get_candidate_count", "sfk_get_intersection_count",
           "sfk_get_profiles", "sfk_get_counts", "sfk_get_grid", "sfk_get_ring_tables"]

SHELL_PROPS = 11   # SFK_SHELL_PROPS
SHELL_SUMS = 13    # SFK_SHELL_SUMS
SHELL_RULE = (128, 256)   # SFK_SHELL_RULE_THETA, SFK_SHELL_RULE_PHI
FRACTION_STRATA = (1024, 256)   # SFK_FRACTION_STRATA_THETA, SFK_FRACTION_STRATA_PHI

_lib = None


def lib():
    """Loads libsfk.so (built by __graft_entry__.build()); never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, dp, fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.sfk_default_params.argtypes = [C.POINTER(Params)]
        L.sfk_default_params.restype = None
        L.sfk_create.argtypes = [C.POINTER(vp), C.POINTER(Params)]
        L.sfk_destroy.argtypes = [vp]
        L.sfk_destroy.restype = None
        L.sfk_last_error.argtypes = [vp]
        L.sfk_last_error.restype = C.c_char_p
        L.sfk_begin_snapshot.argtypes = [vp, C.POINTER(Cosmo), C.c_float]
        L.sfk_add_halos.argtypes = [vp, C.c_int64, dp, dp, dp, dp]
        L.sfk_set_owned.argtypes = [vp, C.c_int64, C.POINTER(C.c_uint8)]
        L.sfk_block_halo_mask.argtypes = [vp, fp, fp, C.c_double, C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
        L.sfk_push_block.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(Cosmo), C.c_double, fp, fp]
        L.sfk_push_block_async.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(Cosmo), C.c_double, fp, fp]
        L.sfk_push_sync.argtypes = [vp]
        L.sfk_push_block_device.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(Cosmo), C.c_double, fp, fp]
        L.sfk_finish_snapshot.argtypes = [vp, dp, C.POINTER(C.c_int32)]
        L.sfk_split_count.argtypes = [vp]
        L.sfk_split_bin.argtypes = [vp]
        L.sfk_split_buffer.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64)]
        L.sfk_split_finish.argtypes = [vp, dp, C.POINTER(C.c_int32)]
        L.sfk_comm_unique_id.argtypes = [vp]
        L.sfk_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
        L.sfk_split_finish_comm.argtypes = [vp, dp, C.POINTER(C.c_int32)]
        L.sfk_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.sfk_stream.argtypes = [vp]
        L.sfk_stream.restype = C.c_void_p
        L.sfk_get_rsp.argtypes = [vp, C.c_int64, dp]
        L.sfk_get_candidate_count.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64)]
        L.sfk_get_intersection_count.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64)]
        L.sfk_get_profiles.argtypes = [vp, C.c_int64, dp]
        u32p = C.POINTER(C.c_uint32)
        L.sfk_get_counts.argtypes = [vp, C.c_int64, u32p, u32p, u32p]
        L.sfk_get_ring_tables.argtypes = [vp, fp, fp, fp]
        L.sfk_get_grid.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64), dp]
        for name in EXPORTS:
            if getattr(L, name).restype is not None and name not in ("sfk_last_error", "sfk_stream"):
                getattr(L, name).restype = C.c_int
        L.sfk_row_length.restype = C.c_int64
        _lib = L
    return _lib


def default_params(**kw):
    p = Params()
    lib().sfk_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class ShellfishError(RuntimeError):
    pass


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class ShellContext:
    """One GPU's shell finder; mirrors the call sequence of cmd/shell.go."""

    def __init__(self, params):
        self.params = params
        self._h = C.c_void_p()
        rc = lib().sfk_create(C.byref(self._h), C.byref(params))
        if rc:
            msg = lib().sfk_last_error(self._h).decode() if self._h else "sfk_create failed"
            if self._h:
                lib().sfk_destroy(self._h)
                self._h = C.c_void_p()
            raise ShellfishError("sfk_create: %s (code %d)" % (msg, rc))
        self.n_halos = 0

    def close(self):
        if self._h:
            lib().sfk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    def _ck(self, rc, what):
        if rc:
            raise ShellfishError("%s: %s (code %d)" % (what, lib().sfk_last_error(self._h).decode(), rc))

    @property
    def P2(self):
        """Row length of the result table (sfk_row_length)."""
        return int(lib().sfk_row_length(self._h))

    def begin_snapshot(self, cosmo, min_mass):
        c = Cosmo(cosmo["Z"], cosmo["OmegaM"], cosmo["OmegaL"], cosmo["H100"])
        self._ck(lib().sfk_begin_snapshot(self._h, C.byref(c), C.c_float(min_mass)), "sfk_begin_snapshot")

    def add_halos(self, halos):
        halos = np.ascontiguousarray(halos, np.float64)
        cols = [np.ascontiguousarray(halos[:, i]) for i in range(4)]
        ptrs = [c.ctypes.data_as(C.POINTER(C.c_double)) for c in cols]
        self._ck(lib().sfk_add_halos(self._h, len(halos), *ptrs), "sfk_add_halos")
        self.n_halos = len(halos)

    def set_owned(self, owned):
        """owned: bool per halo; the rest is left to other ranks (shard.py)."""
        o = np.ascontiguousarray(owned, np.uint8)
        self._ck(lib().sfk_set_owned(self._h, len(o), o.ctypes.data_as(C.POINTER(C.c_uint8))),
                 "sfk_set_owned")

    def block_halo_mask(self, origin, width, total_width):
        hit = np.zeros(self.n_halos, np.uint8)
        n = C.c_int64()
        self._ck(lib().sfk_block_halo_mask(self._h, _f3(origin), _f3(width), float(total_width),
                                           hit.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n)),
                 "sfk_block_halo_mask")
        return hit.astype(bool)

    def push_block(self, block, wait=True):
        """block: dict with Z, OmegaM, OmegaL, H100, TotalWidth, Origin, Width,
        xs (N,3) float32 and ms (N,) float32 host arrays (numpy, or pinned
        torch CPU tensors).  ms may be None: every particle then has the min_mass
        given to begin_snapshot (uniform-mass snapshots; a quarter less to transfer).
        wait=False (sfk_push_block_async): the copy is only queued; xs / ms must stay
        unchanged until push_sync() or finish_snapshot() returns."""
        xs, ms = block["xs"], block.get("ms")
        c = Cosmo(block["Z"], block["OmegaM"], block["OmegaL"], block["H100"])
        if hasattr(xs, "data_ptr"):
            n, px, pm = xs.shape[0], xs.data_ptr(), (ms.data_ptr() if ms is not None else None)
        else:
            assert xs.dtype == np.float32 and xs.flags.c_contiguous and (ms is None or ms.dtype == np.float32)
            n, px, pm = len(xs), xs.ctypes.data, (ms.ctypes.data if ms is not None else None)
        fn = lib().sfk_push_block if wait else lib().sfk_push_block_async
        self._ck(fn(self._h, px, pm, n, C.byref(c), float(block["TotalWidth"]),
                    _f3(block["Origin"]), _f3(block["Width"])), "sfk_push_block" if wait else "sfk_push_block_async")

    def push_sync(self):
        self._ck(lib().sfk_push_sync(self._h), "sfk_push_sync")

    def push_block_device(self, block, d_xs, d_ms):
        """Same with positions/masses already on this context's GPU (torch CUDA tensors; d_ms may
        be None for uniform masses)."""
        c = Cosmo(block["Z"], block["OmegaM"], block["OmegaL"], block["H100"])
        self._ck(lib().sfk_push_block_device(self._h, d_xs.data_ptr(), d_ms.data_ptr() if d_ms is not None else None,
                                             d_xs.shape[0],
                                             C.byref(c), float(block["TotalWidth"]),
                                             _f3(block["Origin"]), _f3(block["Width"])),
                 "sfk_push_block_device")

    def finish_snapshot(self):
        coeffs = np.zeros((self.n_halos, self.P2))
        status = np.zeros(self.n_halos, np.int32)
        self._ck(lib().sfk_finish_snapshot(self._h, coeffs.ctypes.data_as(C.POINTER(C.c_double)),
                                           status.ctypes.data_as(C.POINTER(C.c_int32))),
                 "sfk_finish_snapshot")
        return coeffs, status

    # ---- split-halo mode (one halo's particles spread over several GPUs)
    def split_count(self):
        self._ck(lib().sfk_split_count(self._h), "sfk_split_count")

    def split_bin(self):
        self._ck(lib().sfk_split_bin(self._h), "sfk_split_bin")

    def split_buffer(self, which):
        """Device view (torch tensor sharing memory) of one exchange buffer:
        SPLIT_HIT_COUNTS (int32), SPLIT_RHO_MAX (int64), SPLIT_GRID (int64)."""
        import torch
        ptr, n = C.c_void_p(), C.c_int64()
        self._ck(lib().sfk_split_buffer(self._h, which, C.byref(ptr), C.byref(n)), "sfk_split_buffer")
        typestr = "<i4" if which == SPLIT_HIT_COUNTS else "<i8"

        class _View:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": typestr,
                                        "data": (int(ptr.value or 0), False), "version": 2}
        if n.value == 0:
            return torch.zeros(0, dtype=torch.int32 if which == SPLIT_HIT_COUNTS else torch.int64,
                               device="cuda:%d" % self.params.device)
        return torch.as_tensor(_View(), device="cuda:%d" % self.params.device)

    def split_finish(self):
        coeffs = np.zeros((self.n_halos, self.P2))
        status = np.zeros(self.n_halos, np.int32)
        self._ck(lib().sfk_split_finish(self._h, coeffs.ctypes.data_as(C.POINTER(C.c_double)),
                                        status.ctypes.data_as(C.POINTER(C.c_int32))), "sfk_split_finish")
        return coeffs, status

    # ---- split-halo mode with the exchange inside the library (NCCL)
    def comm_init(self, unique_id, rank, nranks):
        """Collective: joins this context to the NCCL communicator named by `unique_id`
        (comm_unique_id() of one rank, handed to the others by the host)."""
        _preload_nccl()
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._ck(lib().sfk_comm_init(self._h, buf, rank, nranks), "sfk_comm_init")

    def split_finish_comm(self):
        """Collective: hit-count / rho_max all-reduce, private grids, reduce-scatter by ring,
        ring-partitioned analysis, all-gather, fit (sfk_split_finish_comm)."""
        coeffs = np.zeros((self.n_halos, self.P2))
        status = np.zeros(self.n_halos, np.int32)
        self._ck(lib().sfk_split_finish_comm(self._h, coeffs.ctypes.data_as(C.POINTER(C.c_double)),
                                             status.ctypes.data_as(C.POINTER(C.c_int32))), "sfk_split_finish_comm")
        return coeffs, status

    # ---- stats mode: M_sp ----------------------------------------------------------
    def shell_particles(self, centers, coeffs, r_low, r_high, radii, shell_width, blocks):
        """The ShellParticleFile selection (sfk_mass_set_band / sfk_mass_take_band): per block, an
        int64 array of halo << 32 | particle index, sorted.  Returns (masses, [pairs per block])."""
        centers = np.ascontiguousarray(centers, np.float32)
        coeffs = np.ascontiguousarray(coeffs, np.float64)
        r_low = np.ascontiguousarray(r_low, np.float64)
        r_high = np.ascontiguousarray(r_high, np.float64)
        radii = np.ascontiguousarray(radii, np.float32)
        n = len(centers)
        order = int(round((coeffs.shape[1] / 2) ** 0.5))
        dp_, fp_ = C.POINTER(C.c_double), C.POINTER(C.c_float)
        L = lib()
        L.sfk_mass_set_band.argtypes = [C.c_void_p, C.c_double, fp_, dp_, dp_]
        L.sfk_mass_take_band.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_int64))]
        self._ck(L.sfk_mass_begin(self._h, C.c_int64(n), C.c_int32(order), centers.ctypes.data_as(fp_),
                                  coeffs.ctypes.data_as(dp_), r_low.ctypes.data_as(dp_), r_high.ctypes.data_as(dp_)),
                 "sfk_mass_begin")
        self._ck(L.sfk_mass_set_band(self._h, float(shell_width), radii.ctypes.data_as(fp_), r_low.ctypes.data_as(dp_),
                                     r_high.ctypes.data_as(dp_)), "sfk_mass_set_band")
        per_block = []
        for b in blocks:
            xs = np.ascontiguousarray(b["xs"], np.float32)
            ms = np.ascontiguousarray(b["ms"], np.float32)
            self._ck(L.sfk_mass_push_block(self._h, xs.ctypes.data_as(fp_), ms.ctypes.data_as(fp_), C.c_int64(len(ms)),
                                           C.c_double(b["TotalWidth"])), "sfk_mass_push_block")
            cnt, ptr = C.c_int64(), C.POINTER(C.c_int64)()
            self._ck(L.sfk_mass_take_band(self._h, C.byref(cnt), C.byref(ptr)), "sfk_mass_take_band")
            per_block.append(np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy() if cnt.value else np.zeros(0, np.int64))
        out = np.zeros(n)
        self._ck(L.sfk_mass_finish(self._h, out.ctypes.data_as(dp_)), "sfk_mass_finish")
        return out, per_block

    def mass_contained(self, centers, coeffs, r_low, r_high, blocks, radii=None, device_blocks=None):
        """massContained (cmd/stats.go:451-543) for every halo over `blocks`
        (dicts with xs, ms, TotalWidth, Origin, Width).  With `radii` given, blocks
        no bounding sphere touches are skipped like cmd/stats.go:319-320 does."""
        centers = np.ascontiguousarray(centers, np.float32)
        coeffs = np.ascontiguousarray(coeffs, np.float64)
        r_low = np.ascontiguousarray(r_low, np.float64)
        r_high = np.ascontiguousarray(r_high, np.float64)
        n = len(centers)
        order = int(round((coeffs.shape[1] / 2) ** 0.5))
        dp_, fp_ = C.POINTER(C.c_double), C.POINTER(C.c_float)
        self._ck(lib().sfk_mass_begin(self._h, C.c_int64(n), C.c_int32(order), centers.ctypes.data_as(fp_),
                                      coeffs.ctypes.data_as(dp_), r_low.ctypes.data_as(dp_),
                                      r_high.ctypes.data_as(dp_)), "sfk_mass_begin")
        for bi, b in enumerate(blocks):
            if radii is not None:
                o = np.ascontiguousarray(b["Origin"], np.float32)
                w = np.ascontiguousarray(b["Width"], np.float32)
                if not any(lib().sfk_sphere_block_hit(centers[i].ctypes.data_as(fp_), C.c_float(radii[i]),
                                                      o.ctypes.data_as(fp_), w.ctypes.data_as(fp_),
                                                      C.c_double(b["TotalWidth"])) for i in range(n)):
                    continue
            if device_blocks is not None:
                dx, dm = device_blocks[bi]
                self._ck(lib().sfk_mass_push_block_device(self._h, C.c_void_p(dx.data_ptr()), C.c_void_p(dm.data_ptr()),
                                                          C.c_int64(dm.numel()), C.c_double(b["TotalWidth"])),
                         "sfk_mass_push_block_device")
                continue
            xs = np.ascontiguousarray(b["xs"], np.float32)
            ms = np.ascontiguousarray(b["ms"], np.float32)
            self._ck(lib().sfk_mass_push_block(self._h, xs.ctypes.data_as(fp_), ms.ctypes.data_as(fp_),
                                               C.c_int64(len(ms)), C.c_double(b["TotalWidth"])), "sfk_mass_push_block")
        out = np.zeros(n)
        self._ck(lib().sfk_mass_finish(self._h, out.ctypes.data_as(dp_)), "sfk_mass_finish")
        return out

    # ---- stats mode: shell property integrals ----------------------------------------
    def shell_props(self, coeffs):
        """Shell.Volume / SurfaceArea / Axes / RadialRange (los/analyze/shell.go:58-283)
        for every coefficient row: [n][SHELL_PROPS] = R_sp, Volume, SurfaceArea, a, b, c,
        Ax, Ay, Az, RMin, RMax (the float columns of cmd/stats.go:343-349 after M_sp)."""
        coeffs = np.ascontiguousarray(coeffs, np.float64)
        n = len(coeffs)
        order = int(round((coeffs.shape[1] / 2) ** 0.5))
        out = np.zeros((n, SHELL_PROPS))
        dp_ = C.POINTER(C.c_double)
        self._ck(lib().sfk_shell_props(self._h, C.c_int64(n), C.c_int32(order), coeffs.ctypes.data_as(dp_),
                                       out.ctypes.data_as(dp_)), "sfk_shell_props")
        return out

    def angular_fraction(self, coeffs, r_min, r_max, bins):
        """Shell.AngularFractionProfile (los/analyze/shell.go:295-335) for every row:
        (rs, fs), each [n][bins]."""
        coeffs = np.ascontiguousarray(coeffs, np.float64)
        n = len(coeffs)
        order = int(round((coeffs.shape[1] / 2) ** 0.5))
        r_min = np.ascontiguousarray(np.broadcast_to(r_min, (n,)), np.float64)
        r_max = np.ascontiguousarray(np.broadcast_to(r_max, (n,)), np.float64)
        rs, fs = np.zeros((n, bins)), np.zeros((n, bins))
        dp_ = C.POINTER(C.c_double)
        self._ck(lib().sfk_angular_fraction(self._h, C.c_int64(n), C.c_int32(order), coeffs.ctypes.data_as(dp_),
                                            r_min.ctypes.data_as(dp_), r_max.ctypes.data_as(dp_), C.c_int32(bins),
                                            rs.ctypes.data_as(dp_), fs.ctypes.data_as(dp_)), "sfk_angular_fraction")
        return rs, fs

    def shell_sums(self, n):
        """Parity tap: raw sums of the last shell_props call, [n][SHELL_SUMS]."""
        out = np.zeros((n, SHELL_SUMS))
        self._ck(lib().sfk_shell_sums(self._h, C.c_int64(n), out.ctypes.data_as(C.POINTER(C.c_double))),
                 "sfk_shell_sums")
        return out

    def stream_ptr(self):
        """cudaStream_t of this context as an integer (torch.cuda.ExternalStream)."""
        return lib().sfk_stream(self._h)

    def stats(self):
        s = Stats()
        self._ck(lib().sfk_get_stats(self._h, C.byref(s)), "sfk_get_stats")
        return s.as_dict()

    # ---- parity taps
    def _shape(self):
        p = self.params
        return int(p.rings), int(p.spokes), int(p.radial_bins)

    def rsp(self, halo):
        R, n, _ = self._shape()
        out = np.zeros((R, n))
        self._ck(lib().sfk_get_rsp(self._h, halo, out.ctypes.data_as(C.POINTER(C.c_double))), "sfk_get_rsp")
        return out

    def candidate_count(self, halo):
        n = C.c_int64()
        self._ck(lib().sfk_get_candidate_count(self._h, halo, C.byref(n)), "sfk_get_candidate_count")
        return n.value

    def intersection_count(self, halo):
        n = C.c_int64()
        self._ck(lib().sfk_get_intersection_count(self._h, halo, C.byref(n)), "sfk_get_intersection_count")
        return n.value

    def profiles(self, halo):
        R, n, B = self._shape()
        out = np.zeros((R, n, B))
        self._ck(lib().sfk_get_profiles(self._h, halo, out.ctypes.data_as(C.POINTER(C.c_double))),
                 "sfk_get_profiles")
        return out

    def counts(self, halo):
        R, n, B = self._shape()
        ins = np.zeros((R, n), np.uint32)
        st = np.zeros((R, n, B), np.uint32)
        en = np.zeros((R, n, B), np.uint32)
        u32p = C.POINTER(C.c_uint32)
        self._ck(lib().sfk_get_counts(self._h, halo, ins.ctypes.data_as(u32p), st.ctypes.data_as(u32p),
                                      en.ctypes.data_as(u32p)), "sfk_get_counts")
        return ins, st, en

    def grid(self, halo):
        """(int64 [rings][spokes][bins], quantum): ProfileRing.derivs in fixed point (SFK_FLAG_KEEP_GRID)."""
        R, n, B = self._shape()
        g = np.zeros((R, n, B), np.int64)
        q = C.c_double()
        self._ck(lib().sfk_get_grid(self._h, halo, g.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(q)), "sfk_get_grid")
        return g, q.value

    def ring_tables(self):
        R = int(self.params.rings)
        norms, rot, irot = (np.zeros((R, 3), np.float32), np.zeros((R, 9), np.float32),
                            np.zeros((R, 9), np.float32))
        fp = C.POINTER(C.c_float)
        self._ck(lib().sfk_get_ring_tables(self._h, norms.ctypes.data_as(fp), rot.ctypes.data_as(fp),
                                           irot.ctypes.data_as(fp)), "sfk_get_ring_tables")
        return norms, rot, irot


COMM_ID_BYTES = 128   # SFK_COMM_ID_BYTES


def _preload_nccl():
    """libsfk binds NCCL with dlopen, preferring a libnccl.so.2 already in the process.  In a
    Python process that will also import torch, that copy must be the one torch was built
    against (its bundled nvidia-nccl wheel): the loader shares libraries by SONAME, so an older
    system libnccl loaded first would be handed to torch as well.  Load the wheel's copy first."""
    if os.environ.get("SFK_NCCL_LIB"):
        return
    try:
        import glob
        import nvidia.nccl
        for d in list(getattr(nvidia.nccl, "__path__", [])):
            for f in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so.*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
                return
    except (ImportError, OSError):
        pass


def comm_unique_id():
    """ncclGetUniqueId through the library: bytes to hand to every rank's comm_init."""
    _preload_nccl()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if lib().sfk_comm_unique_id(buf) != 0:
        raise ShellfishError("sfk_comm_unique_id: NCCL is not available (set SFK_NCCL_LIB)")
    return buf.raw


def axes_from_moments(moments):
    """Shell.Axes from los/analyze/shell.go:140 on: area-weighted means {xx, yy, zz, xy,
    yz, zx, x, y, z} -> (a, b, c, Ax, Ay, Az).  Host arithmetic of libsfk.so."""
    m = np.ascontiguousarray(moments, np.float64)
    out = np.zeros(6)
    dp_ = C.POINTER(C.c_double)
    if lib().sfk_axes_from_moments(m.ctypes.data_as(dp_), out.ctypes.data_as(dp_)) != 0:
        raise ShellfishError("sfk_axes_from_moments: bad arguments")
    return out


def run_shell(params, min_mass, halos, blocks):
    """loop() of cmd/shell.go:288-372 for one snapshot of in-memory blocks.
    Returns (ctx, coeffs, status); the context keeps the parity taps."""
    ctx = ShellContext(params)
    ctx.begin_snapshot(blocks[0], min_mass)   # hds[0] of the first snapshot, cmd/shell.go:305
    ctx.add_halos(halos)
    for b in blocks:
        ctx.push_block(b)
    coeffs, status = ctx.finish_snapshot()
    return ctx, coeffs, status
