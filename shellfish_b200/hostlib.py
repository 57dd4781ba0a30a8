"""ctypes view of libsfhost.so: the C++ host side of `shellfish shell`
(config parser, catalog text, snapshot readers).  Used by the CPU test tier
and by tools; the CLI binary `shellfish_cli` links the same objects."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsfhost.so")
        if not os.path.exists(path):
            raise RuntimeError("libsfhost.so is missing: run `make -C host` (or __graft_entry__.build())")
        _lib = C.CDLL(path)
        for name in ("sfh_go_g6", "sfh_comment_string", "sfh_format_cols", "sfh_last_text", "sfh_read_test_config",
                     "sfh_read_shell_config", "sfh_read_stats_config", "sfh_read_prof_config", "sfh_read_global_config", "sfh_catalog_name"):
            getattr(_lib, name).restype = C.c_char_p
        _lib.sfh_go_g6.argtypes = [C.c_double]
        _lib.sfh_read_block.restype = C.c_int64
        _lib.sfh_catalog_name.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                          C.c_int64, C.c_int, C.c_int, C.c_void_p]
    return _lib


def cli_path():
    return os.path.join(_HERE, "shellfish_cli")


def partition_halos(weights, groups, n_ranks):
    """PartitionHalos of host/sf_shell.cpp: owner rank of every halo."""
    w = np.ascontiguousarray(weights, np.float64)
    g = np.ascontiguousarray(groups, np.int64)
    own = np.zeros(len(w), np.int32)
    f = lib().sfh_partition_halos
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f.restype = None
    f(w.ctypes.data, g.ctypes.data, len(w), n_ranks, own.ctypes.data)
    return own


def partition_spatial(weights, xyz, total_width, n_ranks):
    """PartitionHalosSpatial of host/sf_shell.cpp (what the CLI uses with SHELLFISH_GPUS)."""
    w = np.ascontiguousarray(weights, np.float64)
    x = np.ascontiguousarray(xyz, np.float64)
    own = np.zeros(len(w), np.int32)
    f = lib().sfh_partition_spatial
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    f.restype = None
    f(w.ctypes.data, x.ctypes.data, len(w), float(total_width), n_ranks, own.ctypes.data)
    return own


def go_g6(v):
    return lib().sfh_go_g6(float(v)).decode()


def _iarr(xs):
    return (C.c_int * len(xs))(*xs)


def comment_string(int_names, float_names, order, sizes):
    return lib().sfh_comment_string("".join(n + "\n" for n in int_names).encode(),
                                    "".join(n + "\n" for n in float_names).encode(),
                                    _iarr(order), len(order), _iarr(sizes), len(sizes)).decode()


def format_cols(int_cols, float_cols, order):
    if not len(int_cols) and not len(float_cols):
        return []
    ic = np.ascontiguousarray(int_cols, np.int64).reshape(len(int_cols), -1)
    fc = np.ascontiguousarray(float_cols, np.float64).reshape(len(float_cols), -1)
    rows = ic.shape[1] if len(int_cols) else fc.shape[1]
    out = lib().sfh_format_cols(ic.ctypes.data_as(C.c_void_p), len(int_cols), fc.ctypes.data_as(C.c_void_p),
                                len(float_cols), rows, _iarr(order), len(order)).decode()
    return out.split("\n")[:-1]


def parse_catalog(text, icols, fcols, cap=4096):
    iout = np.zeros((max(len(icols), 1), cap), np.int64)
    fout = np.zeros((max(len(fcols), 1), cap), np.float64)
    n = lib().sfh_parse_catalog(text.encode(), _iarr(icols), len(icols), _iarr(fcols), len(fcols),
                                iout.ctypes.data_as(C.c_void_p), fout.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        raise ValueError(lib().sfh_last_text().decode())
    return iout[:len(icols), :n], fout[:len(fcols), :n]


def read_test_config(text, flags=(), fname="test.config"):
    num, flt, okay = C.c_int64(), C.c_double(), C.c_int()
    bufs = [C.create_string_buffer(512) for _ in range(5)]
    err = lib().sfh_read_test_config(None if text is None else text.encode(), fname.encode(),
                                     "\n".join(flags).encode(), C.byref(num), C.byref(flt), C.byref(okay),
                                     *bufs, 512).decode()
    word, nums, floats, okays, words = [b.value.decode() for b in bufs]
    sp = lambda s: [x for x in s.split(",")] if s else []
    return err, dict(num=num.value, float=flt.value, okay=bool(okay.value), word=word,
                     nums=[int(x) for x in sp(nums)], floats=[float(x) for x in sp(floats)],
                     okays=[x == "true" for x in sp(okays)], words=sp(words))


def read_shell_config(fname=None, flags=()):
    ints, floats, pp = (C.c_int64 * 7)(), (C.c_double * 7)(), C.c_int()
    err = lib().sfh_read_shell_config(None if fname is None else fname.encode(), "\n".join(flags).encode(),
                                      ints, floats, C.byref(pp)).decode()
    keys_i = ["SubsampleFactor", "RadialBins", "Spokes", "Rings", "Order", "Levels", "SmoothingWindow"]
    keys_f = ["RMaxMult", "RMinMult", "RKernelMult", "Eta", "LOSSlopeCutoff", "BackgroundRhoMult", "Percentile"]
    out = dict(zip(keys_i, list(ints)))
    out.update(zip(keys_f, list(floats)))
    out["PercentileProfile"] = bool(pp.value)
    return err, out


def read_stats_config(fname=None, flags=()):
    """StatsConfig.ReadConfig + validate, cmd/stats.go:109-172."""
    ints, width, bools = (C.c_int64 * 2)(), C.c_double(), (C.c_int * 2)()
    strat = C.create_string_buffer(64)
    err = lib().sfh_read_stats_config(None if fname is None else fname.encode(), "\n".join(flags).encode(),
                                      ints, C.byref(width), bools, strat, 64).decode()
    return err, dict(MonteCarloSamples=ints[0], Order=ints[1], ShellWidth=width.value, SkipMass=bool(bools[0]),
                     shellFilter=bool(bools[1]), ExclusionStrategy=strat.value.decode())


def read_prof_config(fname=None, flags=()):
    """ProfConfig.ReadConfig + validate, cmd/prof.go:103-169."""
    ints, floats = (C.c_int64 * 4)(), (C.c_double * 3)()
    ptype = C.create_string_buffer(64)
    err = lib().sfh_read_prof_config(None if fname is None else fname.encode(), "\n".join(flags).encode(),
                                     ints, floats, ptype, 64).decode()
    return err, dict(Bins=ints[0], Order=ints[1], Samples=ints[2], MedianPixelLevel=ints[3], RMaxMult=floats[0],
                     RMinMult=floats[1], Percentile=floats[2], ProfileType=ptype.value.decode())


def read_global_config(fname):
    st = C.create_string_buffer(64)
    blocks = C.c_int64(-1)
    err = lib().sfh_read_global_config(fname.encode(), st, 64, C.byref(blocks)).decode()
    return err, st.value.decode(), blocks.value


def catalog_name(fmt, meanings, block_mins, block_maxes, snap_min, snap_max, snap, block):
    nb = C.c_int()
    mins = (C.c_int64 * max(len(block_mins), 1))(*block_mins)
    maxes = (C.c_int64 * max(len(block_maxes), 1))(*block_maxes)
    out = lib().sfh_catalog_name(fmt.encode(), ",".join(meanings).encode(), mins, maxes, len(block_mins),
                                 snap_min, snap_max, snap, block, C.byref(nb)).decode()
    if nb.value < 0:
        raise ValueError(out)
    return out, nb.value


def read_block(kind, fname, endianness="SystemOrder", npart_num=2, pos_units=1.0, mass_units=1.0, cap=1 << 22,
               dm_types=(), single_mass=()):
    """Returns (xyz, mass, header dict, min_mass, total_particles).  dm_types / single_mass:
    GadgetDMTypeIndices / GadgetSingleMassIndices (empty = the reference's defaults, [1] and [1])."""
    xs = np.zeros((cap, 3), np.float32)
    ms = np.zeros(cap, np.float32)
    hd = np.zeros(72, np.uint8)
    mm, tot = C.c_float(), C.c_int64()
    n = lib().sfh_read_block(kind.encode(), fname.encode(), endianness.encode(), C.c_int64(npart_num),
                             C.c_double(pos_units), C.c_double(mass_units), xs.ctypes.data_as(C.c_void_p),
                             ms.ctypes.data_as(C.c_void_p), C.c_int64(cap), hd.ctypes.data_as(C.c_void_p),
                             C.byref(mm), C.byref(tot), C.c_int(sum(1 << i for i in dm_types)),
                             C.c_int(sum(1 << i for i in single_mass)))
    if n < 0:
        raise ValueError(lib().sfh_last_text().decode())
    h = dict(zip(["z", "omega_m", "omega_l", "h100"], hd[:32].view(np.float64)))
    h["n"] = int(hd[32:40].view(np.int64)[0])
    h["total_width"] = float(hd[40:48].view(np.float64)[0])
    h["origin"] = hd[48:60].view(np.float32).copy()
    h["width"] = hd[60:72].view(np.float32).copy()
    return xs[:n], ms[:n], h, mm.value, tot.value


def read_ids(kind, fname, endianness="SystemOrder", npart_num=2, cap=1 << 22, dm_types=()):
    """Particle IDs of a block file (the reader's ReadIDs), aligned with read_block's positions."""
    ids = np.zeros(cap, np.int64)
    f = lib().sfh_read_ids
    f.restype = C.c_int64
    n = f(kind.encode(), fname.encode(), endianness.encode(), C.c_int64(npart_num), ids.ctypes.data_as(C.c_void_p),
          C.c_int64(cap), C.c_int(sum(1 << i for i in dm_types)))
    if n < 0:
        raise ValueError(lib().sfh_last_text().decode())
    return ids[:n]


def write_filter(path, endianness, snaps, ids, particles):
    """io.WriteFilter (io/filter_data.go:36-63): the ShellParticleFile."""
    snaps = np.ascontiguousarray(snaps, np.int64)
    ids = np.ascontiguousarray(ids, np.int64)
    lens = np.array([len(p) for p in particles], np.int64)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(p, np.int64) for p in particles]) if len(particles) else [], np.int64)
    rc = lib().sfh_write_filter(path.encode(), endianness.encode(), C.c_int(len(snaps)), snaps.ctypes.data_as(C.c_void_p),
                                ids.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p),
                                flat.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError(lib().sfh_last_text().decode())


def bounding_box(xyz, total_width):
    xyz = np.ascontiguousarray(xyz, np.float32)
    o, w = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib().sfh_bounding_box(xyz.ctypes.data_as(C.c_void_p), C.c_int64(len(xyz)), C.c_double(total_width),
                           o.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    return o, w
