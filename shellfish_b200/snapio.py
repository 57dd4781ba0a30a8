"""Writers for the three particle-snapshot formats BASELINE's configs name, so
that tests and the CLI can be driven from synthetic boxes (there is no network
for real simulation data).  Layouts follow the readers they feed:

  LGadget-2  io/lgadget2.go:10-23 (256-byte header), :84-118 (Fortran records)
  Gadget-2   io/gadget2.go:11-25, :64-144
  gotetra    io/gotetra.go:98-120 (152-byte raw header), :230-256

Only positions, ids and (for Gadget-2) masses carry information; velocities are
zero.  Little endian unless stated.
"""
import struct

import numpy as np


def _record(payload: bytes, order="<") -> bytes:
    n = struct.pack(order + "i", len(payload) & 0x7FFFFFFF)
    return n + payload + n


def write_lgadget2(path, xyz, box, cosmo, n_total=None, order="<", npart_num=2):
    """xyz: (N,3) float32 positions in Mpc/h; cosmo: dict z, omega_m, omega_l, h100.
    n_total: particles in the whole simulation (NPartTotal; default N)."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    n = len(xyz)
    n_total = n if n_total is None else n_total
    npart = [0] * 6
    nparttot = [0] * 6
    if npart_num == 2:
        npart[1], npart[0] = n & 0xFFFFFFFF, n >> 32
        nparttot[1], nparttot[0] = n_total & 0xFFFFFFFF, n_total >> 32
    else:
        npart[0], nparttot[0] = n, n_total
    a = 1.0 / (1.0 + cosmo["z"])
    # binary.Read(f, binary.LittleEndian, gh): the header is always little endian (lgadget2.go:81)
    hd = struct.pack("<6I6d2d2i6I2i4d2i88s", *npart, *([0.0] * 6), a, cosmo["z"], 0, 0, *nparttot, 0, 1,
                     float(box), cosmo["omega_m"], cosmo["omega_l"], cosmo["h100"], 0, 0, b"")
    assert len(hd) == 256
    dt = np.dtype(np.float32).newbyteorder(order)
    with open(path, "wb") as f:
        f.write(_record(hd, order))
        f.write(_record(xyz.astype(dt).tobytes(), order))
        f.write(_record(np.zeros((n, 3), dt).tobytes(), order))
        f.write(_record(np.arange(n, dtype=np.dtype(np.int64).newbyteorder(order)).tobytes(), order))


def write_gadget2(path, xyz_by_type, box, cosmo, table_mass, masses_by_type=None, order="<", id_bytes=8):
    """xyz_by_type: list of 6 (N_i,3) arrays (None = empty) in file position units;
    table_mass: 6 floats (header Mass, used for single-mass types);
    masses_by_type: per-type mass arrays for the types that are NOT single-mass
    (None entries = type is single-mass or empty)."""
    xs = [np.zeros((0, 3), np.float32) if x is None else np.ascontiguousarray(x, np.float32) for x in xyz_by_type]
    npart = [len(x) for x in xs]
    a = 1.0 / (1.0 + cosmo["z"])
    hd = struct.pack(order + "6I6d2d2i6I2i4d2i6Ii60s", *npart, *[float(m) for m in table_mass], a, cosmo["z"], 0, 0,
                     *npart, 0, 1, float(box), cosmo["omega_m"], cosmo["omega_l"], cosmo["h100"], 0, 0,
                     *([0] * 6), 0, b"")
    assert len(hd) == 256
    allx = np.concatenate(xs) if sum(npart) else np.zeros((0, 3), np.float32)
    n = len(allx)
    fdt = np.dtype(np.float32).newbyteorder(order)
    idt = np.dtype(np.int64 if id_bytes == 8 else np.int32).newbyteorder(order)
    multi = []
    if masses_by_type is not None:
        for m in masses_by_type:
            if m is not None:
                multi.append(np.asarray(m, np.float32))
    multi = np.concatenate(multi) if multi else np.zeros(0, np.float32)
    with open(path, "wb") as f:
        f.write(_record(hd, order))
        f.write(_record(allx.astype(fdt).tobytes(), order))
        f.write(_record(np.zeros((n, 3), fdt).tobytes(), order))
        f.write(_record(np.arange(n).astype(idt).tobytes(), order))
        f.write(_record(multi.astype(fdt).tobytes(), order))


def write_gotetra(path, grid_xyz, idx, cells, segment_width, count_width, box, cosmo, origin, width, order="<"):
    """grid_xyz: (grid_width^3, 3) float32 sheet positions, grid_width = segment_width + 1."""
    gw = segment_width + 1
    grid_xyz = np.ascontiguousarray(grid_xyz, np.float32)
    assert grid_xyz.shape == (gw ** 3, 3)
    # endianness flag, gotetra.go:140-148: 0 -> little endian, -1 -> big endian
    flag = 0 if order == "<" else -1
    raw = struct.pack(order + "4d7q2d3f3f3f3f", cosmo["z"], cosmo["omega_m"], cosmo["omega_l"], cosmo["h100"],
                      count_width ** 3, count_width, segment_width, gw, gw ** 3, idx, cells, 0.0, float(box),
                      *[float(v) for v in origin], *[float(v) for v in width], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert len(raw) == 152
    with open(path, "wb") as f:
        f.write(struct.pack("<i", flag))
        f.write(struct.pack(order + "i", len(raw)))
        f.write(raw)
        f.write(grid_xyz.astype(np.dtype(np.float32).newbyteorder(order)).tobytes())


def write_box_lgadget2(dirname, fmt, blocks, snap=100, n_total=None):
    """Writes synth.make_box blocks (dicts with xs, Z, OmegaM, OmegaL, H100,
    TotalWidth) as one LGadget-2 file per block; fmt takes (snap, block).
    Returns the file names."""
    import os
    n_total = sum(len(b["xs"]) for b in blocks) if n_total is None else n_total
    names = []
    for i, b in enumerate(blocks):
        name = os.path.join(dirname, fmt % (snap, i))
        cosmo = dict(z=b["Z"], omega_m=b["OmegaM"], omega_l=b["OmegaL"], h100=b["H100"])
        write_lgadget2(name, b["xs"], b["TotalWidth"], cosmo, n_total=n_total)
        names.append(name)
    return names


def write_box_gadget2(dirname, fmt, blocks, snap=100, pos_units=1e-3, mass_units=1e10):
    """Writes synth.make_box blocks as one Gadget-2 file per block: all particles are
    type 1 (dark matter) with the mass in the header table; positions are stored in
    1/pos_units (kpc/h for 1e-3), masses in 1/mass_units (1e10 Msun/h)."""
    import os
    names = []
    for i, b in enumerate(blocks):
        name = os.path.join(dirname, fmt % (snap, i))
        cosmo = dict(z=b["Z"], omega_m=b["OmegaM"], omega_l=b["OmegaL"], h100=b["H100"])
        m = float(b["ms"][0]) / mass_units
        write_gadget2(name, [None, b["xs"] / np.float32(pos_units), None, None, None, None],
                      b["TotalWidth"] / pos_units, cosmo, [0.0, m, 0.0, 0.0, 0.0, 0.0])
        names.append(name)
    return names


def write_box_gotetra(dirname, blocks, cells, segment_width, fmt="sheet%d%d%d.dat"):
    """Writes the particles of synth.make_box blocks as cells^3 gotetra sheet files of
    segment_width^3 particles each (the total must be exactly (cells*segment_width)^3).
    Particles are dealt to files in block order; each header carries the bounding box of
    its own particles, as gotetra's do for a Lagrangian segment.  Returns the file names
    in the reference's block order (Block0 fastest, cmd/env/env.go:187-234)."""
    import os
    sw, gw = segment_width, segment_width + 1
    xs = np.concatenate([b["xs"] for b in blocks]).astype(np.float32)
    assert len(xs) == (cells * sw) ** 3, (len(xs), (cells * sw) ** 3)
    b0 = blocks[0]
    cosmo = dict(z=b0["Z"], omega_m=b0["OmegaM"], omega_l=b0["OmegaL"], h100=b0["H100"])
    names = []
    per = sw ** 3
    k = 0
    for iz in range(cells):
        for iy in range(cells):
            for ix in range(cells):
                part = xs[k * per:(k + 1) * per]
                grid = np.zeros((gw, gw, gw, 3), np.float32)
                grid[:sw, :sw, :sw] = part.reshape(sw, sw, sw, 3)   # [z][y][x] like gotetra.go:73-85
                lo, hi = part.min(0), part.max(0)
                name = os.path.join(dirname, fmt % (ix, iy, iz))
                write_gotetra(name, grid.reshape(-1, 3), k, cells, sw, cells * sw, b0["TotalWidth"], cosmo,
                              origin=lo, width=hi - lo)
                names.append(name)
                k += 1
    return names
