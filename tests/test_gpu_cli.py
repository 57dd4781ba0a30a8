"""End to end through the `shellfish shell` binary (host/shellfish_main.cpp):
LGadget-2 files on disk -> stdin halo table -> catalog text on stdout, compared
with (a) the same run driven through the Python C-ABI binding and (b) the CPU
oracle, at the precision the catalog prints (%.6g, cmd/catalog/catalog.go:118).
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from shellfish_b200 import api, hostlib as H, snapio, synth

pytestmark = pytest.mark.gpu

GLOBAL = """[config]
Version = 1.0.4
SnapshotFormat = %s/snap_%%03d.%%d
SnapshotFormatMeanings = Snapshot, Block
SnapshotType = LGadget-2
BlockMins = 0
BlockMaxes = %d
SnapMin = 100
SnapMax = 100
MemoDir = %s/memo
"""


def _setup(tmp_path, seed, n_halos, n_per, n_bg, cells, box=62.5):
    halos, blocks, mass = synth.make_box(seed, box, n_halos, n_per, n_bg, cells, margin=8.0)
    names = snapio.write_box_lgadget2(str(tmp_path), "snap_%03d.%d", blocks)
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL % (tmp_path, len(blocks) - 1, tmp_path))
    # what the reader hands to the loop: tight bounding boxes, uniform mass from NPartTotal
    rb = []
    for f in names:
        xs, ms, hd, min_mass, _ = H.read_block("LGadget-2", f)
        rb.append(dict(Z=hd["z"], OmegaM=hd["omega_m"], OmegaL=hd["omega_l"], H100=hd["h100"],
                       TotalWidth=hd["total_width"], Origin=tuple(hd["origin"]), Width=tuple(hd["width"]),
                       xs=xs.copy(), ms=ms.copy()))
    return halos, rb, min_mass, str(cfg)


def _run_cli(cfg, halos, flags, seed=1337, snap=100, gpus=None):
    stdin = "# ID Snap X Y Z R200m\n" + "".join(
        "%d %d %.17g %.17g %.17g %.17g\n" % (i, snap, *h[:4]) for i, h in enumerate(halos))
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg, SHELLFISH_SEED=str(seed))
    if gpus is not None:
        env["SHELLFISH_GPUS"] = gpus
    r = subprocess.run([H.cli_path(), "shell"] + flags, input=stdin, capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout.split("\n")[:-1]


def test_cli_matches_binding_and_oracle(tmp_path):
    halos, blocks, min_mass, cfg = _setup(tmp_path, 11, 3, 30000, 20000, 2)
    lines = _run_cli(cfg, halos, ["--Rings", "12"])
    assert lines[0] == ("# Column contents: ID(0) Snapshot(1) X [cMpc/h](2) Y [cMpc/h](3) Z [cMpc/h](4) "
                        "R200m [cMpc/h](5) P_ijk(6-23)")
    assert len(lines) == 1 + len(halos)

    # (a) same run through the ctypes binding, printed by the same formatter: text-identical
    params = api.default_params(seed=1337, rings=12)
    _, coeffs, status = api.run_shell(params, min_mass, halos, blocks)
    assert np.all(status == 0)
    ids = np.arange(len(halos))
    fcols = [halos[:, k] for k in range(4)] + [coeffs[:, k] for k in range(coeffs.shape[1])]
    want = H.format_cols([ids, np.full(len(halos), 100)], fcols, list(range(2 + len(fcols))))
    assert lines[1:] == want

    # (b) the oracle on the same blocks: every printed coefficient within 1e-5 of the halo's scale
    ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=1, rings=12), min_mass, halos, blocks)
    ic, fc = H.parse_catalog("\n".join(lines) + "\n", [0, 1], list(range(2, 24)))
    assert ic[0].tolist() == ids.tolist() and np.all(ic[1] == 100)
    got = fc[4:].T
    for h in range(len(halos)):
        scale = np.max(np.abs(ref["coeffs"][h]))
        assert np.all(np.abs(got[h] - ref["coeffs"][h]) <= (1e-5 + 5e-6) * scale)   # + the %.6g rounding


def test_cli_several_contexts_give_the_same_catalog(tmp_path):
    """SHELLFISH_GPUS: one context per listed device, halos partitioned by estimated particle
    count with block affinity, every block pushed to the contexts whose halos touch it, rows
    merged in input order.  Two and three contexts on GPU 0 (and every GPU of the box when there
    are several) print the catalog of the single-context run, byte for byte."""
    import torch
    halos, blocks, min_mass, cfg = _setup(tmp_path, 15, 7, 12000, 20000, 3, box=30.0)
    one = _run_cli(cfg, halos, ["--Rings", "8"], gpus="0")
    assert len(one) == 1 + len(halos)
    assert _run_cli(cfg, halos, ["--Rings", "8"], gpus="0,0") == one
    assert _run_cli(cfg, halos, ["--Rings", "8"], gpus="0,0,0") == one
    if torch.cuda.device_count() > 1:
        assert _run_cli(cfg, halos, ["--Rings", "8"], gpus="all") == one
        assert _run_cli(cfg, halos, ["--Rings", "8"], gpus="1,0") == one


def test_cli_memo_and_skipped_halo(tmp_path):
    """Second run reuses MemoDir/hd_snap100.dat; a halo with R200m <= 0 and a
    Snap == -1 row keep all-zero coefficient rows (cmd/shell.go:325,535-538)."""
    halos, blocks, min_mass, cfg = _setup(tmp_path, 12, 2, 20000, 10000, 2)
    first = _run_cli(cfg, halos, ["--Rings", "6"])
    memo = os.path.join(os.path.dirname(cfg), "memo", "hd_snap100.dat")
    assert os.path.getsize(memo) == 72 * len(blocks)
    assert _run_cli(cfg, halos, ["--Rings", "6"]) == first
    bad = halos.copy()
    bad[1, 3] = -1.0
    lines = _run_cli(cfg, bad, ["--Rings", "6"])
    assert lines[1].split() == first[1].split()
    assert all(float(v) == 0.0 for v in lines[2].split()[6:])
    stdin_extra = _run_cli(cfg, halos, ["--Rings", "6"], snap=100)
    assert stdin_extra == first


def test_cli_changed_config_is_refused(tmp_path):   # checkMemoDir, shellfish.go:491-521
    halos, blocks, min_mass, cfg = _setup(tmp_path, 13, 1, 5000, 2000, 1)
    _run_cli(cfg, halos, ["--Rings", "4"])
    text = open(cfg).read().replace("Endianness", "Endianness")
    open(cfg, "w").write(text + "Endianness = LittleEndian\n")
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg)
    r = subprocess.run([H.cli_path(), "shell"], input="0 100 1 1 1 1\n", capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "invlalidate the files Shellfish cached" in r.stderr
    assert r.stdout == "Shellfish terminating.\n"


def test_cli_percentile_profile(tmp_path):   # cmd/shell.go:629-660 through shell.config
    halos, blocks, min_mass, cfg = _setup(tmp_path, 14, 2, 15000, 5000, 1)
    sc = tmp_path / "shell.config"
    sc.write_text("[shell.config]\nRings = 6\nPercentileProfile = true\nPercentile = 75\n")
    stdin = "".join("%d 100 %.17g %.17g %.17g %.17g\n" % (i, *h[:4]) for i, h in enumerate(halos))
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg, SHELLFISH_SEED="1337")
    r = subprocess.run([H.cli_path(), "shell", str(sc)], input=stdin, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")[:-1]
    assert lines[0].endswith("R200m [cMpc/h](5) P_ijk(6-517)")
    ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=1, rings=6), min_mass, halos, blocks,
                       taps=("profiles",))
    _, fc = H.parse_catalog("\n".join(lines) + "\n", [0, 1], list(range(6, 518)))
    for h in range(len(halos)):
        want = po.calc_percentile(halos[h, 3], ref["profiles"][h], 75.0)
        np.testing.assert_allclose(fc[:, h], want, rtol=2e-5)


GLOBAL_G2 = """[config]
Version = 1.0.4
SnapshotFormat = %s/g2_%%03d.%%d
SnapshotFormatMeanings = Snapshot, Block
SnapshotType = Gadget-2
BlockMins = 0
BlockMaxes = %d
SnapMin = 100
SnapMax = 100
MemoDir = %s/memo
GadgetPositionUnits = 0.001
GadgetMassUnits = 1e10
"""

GLOBAL_GT = """[config]
Version = 1.0.4
SnapshotFormat = %s/sheet%%d%%d%%d.dat
SnapshotFormatMeanings = Block0, Block1, Block2
SnapshotType = gotetra
BlockMins = 0, 0, 0
BlockMaxes = 1, 1, 1
SnapMin = 100
SnapMax = 100
MemoDir = %s/memo
"""


def _check_against_binding(lines, halos, names, kind, tmp_path, rings, **reader_kw):
    rb = []
    for f in names:
        xs, ms, hd, min_mass, _ = H.read_block(kind, f, **reader_kw)
        rb.append(dict(Z=hd["z"], OmegaM=hd["omega_m"], OmegaL=hd["omega_l"], H100=hd["h100"],
                       TotalWidth=hd["total_width"], Origin=tuple(hd["origin"]), Width=tuple(hd["width"]),
                       xs=xs.copy(), ms=ms.copy()))
    _, coeffs, status = api.run_shell(api.default_params(seed=1337, rings=rings), min_mass, halos, rb)
    ids = np.arange(len(halos))
    fcols = [halos[:, k] for k in range(4)] + [coeffs[:, k] for k in range(coeffs.shape[1])]
    want = H.format_cols([ids, np.full(len(halos), 100)], fcols, list(range(2 + len(fcols))))
    assert lines[1:] == want
    return rb, min_mass, coeffs, status


def test_cli_gadget2_snapshot(tmp_path):
    """BASELINE config 1's format: one halo, Gadget-2 files in kpc/h and 1e10 Msun/h.  The
    Gadget-2 reader never sets its mass, so MinMass() == 0 and the background density is 0
    (io/gadget2.go:327,402; SURVEY hazard 6): empty bins become ln 0 = -Inf downstream."""
    halos, blocks, mass = synth.make_box(21, 62.5, 1, 60000, 10000, 2, margin=15.0)
    names = snapio.write_box_gadget2(str(tmp_path), "g2_%03d.%d", blocks)
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL_G2 % (tmp_path, len(blocks) - 1, tmp_path))
    lines = _run_cli(str(cfg), halos, ["--Rings", "8"])
    rb, min_mass, coeffs, status = _check_against_binding(lines, halos, names, "Gadget-2", tmp_path, 8,
                                                          pos_units=1e-3, mass_units=1e10)
    assert min_mass == 0.0
    np.testing.assert_allclose(rb[0]["ms"][0], mass, rtol=1e-6)
    np.testing.assert_allclose(np.concatenate([b["xs"] for b in rb]), np.concatenate([b["xs"] for b in blocks]),
                               rtol=2e-7, atol=1e-5)


def test_cli_gotetra_snapshot(tmp_path):
    """gotetra sheet files (BASELINE config 4's format): 2^3 files of 32^3 particles."""
    sw, cells = 32, 2
    n_total = (sw * cells) ** 3
    halos, blocks, mass = synth.make_box(22, 62.5, 2, 40000, n_total - 80000, 2, margin=12.0)
    names = snapio.write_box_gotetra(str(tmp_path), blocks, cells, sw)
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL_GT % (tmp_path, tmp_path))
    lines = _run_cli(str(cfg), halos, ["--Rings", "8"])
    rb, min_mass, coeffs, status = _check_against_binding(lines, halos, names, "gotetra", tmp_path, 8)
    assert np.all(status == 0)
    assert min_mass == pytest.approx(mass, rel=1e-6)   # calcUniformMass(CountWidth^3), gotetra.go:43-54
    ref = po.shell_run(po.ShellConfig.default(seed=1337, threads=1, rings=8), min_mass, halos, rb)
    for h in range(len(halos)):
        scale = np.max(np.abs(ref["coeffs"][h]))
        assert np.all(np.abs(coeffs[h] - ref["coeffs"][h]) <= 1e-5 * scale)


def test_cli_two_snapshots_in_one_table(tmp_path):
    """Rows of several snapshots in one halo table (binBySnap, cmd/shell.go:299-303,322-344):
    snapshots are processed in ascending order, each with its own header memo, and the rows come
    back in input order."""
    halos, blocks, mass = synth.make_box(15, 62.5, 2, 20000, 8000, 2, margin=10.0)
    for snap in (100, 101):
        snapio.write_box_lgadget2(str(tmp_path), "snap_%03d.%d", blocks, snap=snap)
    cfg = tmp_path / "g.config"
    cfg.write_text((GLOBAL % (tmp_path, len(blocks) - 1, tmp_path)).replace("SnapMax = 100", "SnapMax = 101"))
    rows = [(0, 101, halos[0]), (1, 100, halos[1]), (2, 100, halos[0]), (3, 101, halos[1]), (4, -1, halos[0])]
    stdin = "".join("%d %d %.17g %.17g %.17g %.17g\n" % (i, s, *h[:4]) for i, s, h in rows)
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=str(cfg), SHELLFISH_SEED="1337")
    r = subprocess.run([H.cli_path(), "shell", "--Rings", "6"], input=stdin, capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    out = [l.split() for l in r.stdout.split("\n")[1:-1]]
    assert [int(l[0]) for l in out] == [0, 1, 2, 3, 4] and [int(l[1]) for l in out] == [101, 100, 100, 101, -1]
    assert out[0][6:] == out[2][6:] and out[1][6:] == out[3][6:]      # same particles in both snapshots
    assert out[0][6:] != out[1][6:]
    assert all(float(v) == 0.0 for v in out[4][6:])                   # Snap == -1 rows are skipped, cmd/shell.go:325
    for snap in (100, 101):
        assert os.path.getsize(tmp_path / "memo" / ("hd_snap%d.dat" % snap)) == 72 * len(blocks)


def test_shell_stats_pipeline(tmp_path):
    """`shellfish shell | shellfish stats` (cmd/stats.go:174-360): the stats catalog printed by
    the CLI is text-identical to the same two device passes driven through the ctypes binding,
    M_sp agrees with the oracle's massContained on the blocks binSphereIntersections keeps, and
    volume / area sit inside the scatter of the oracle's Monte-Carlo restatement."""
    halos, blocks, min_mass, cfg = _setup(tmp_path, 15, 3, 30000, 20000, 2)
    shell_lines = _run_cli(cfg, halos, ["--Rings", "12"])
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg)
    r = subprocess.run([H.cli_path(), "stats"], input="\n".join(shell_lines) + "\n", capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")[:-1]
    assert lines[0] == ("# Column contents: ID(0) Snapshot(1) M_sp [M_sun/h](2) R_sp [cMpc/h](3) "
                        "Volume [cMpc^3/h^3](4) Surface Area [cMpc^2/h^2](5) Major Axis [cMpc/h](6) "
                        "Intermediate Axis [cMpc/h](7) Minor Axis [cMpc/h](8) Ax(9) Ay(10) Az(11) "
                        "RMin [cMpc/h](12) RMax [cMpc/h](13)")
    assert len(lines) == 1 + len(halos)

    # the catalog text is the interface between the two modes: stats sees the %.6g coefficients
    ic, fc = H.parse_catalog("\n".join(shell_lines) + "\n", [0, 1], list(range(2, 24)))
    centers, r200, coeffs = fc[:3].T.copy(), fc[3].copy(), fc[4:].T.copy()
    ctx = api.ShellContext(api.default_params(rings=3))
    props = ctx.shell_props(coeffs)
    masses = ctx.mass_contained(centers, coeffs, props[:, 9], props[:, 10], blocks, radii=r200.astype(np.float32))
    want = H.format_cols([ic[0], ic[1]], [masses] + [props[:, k] for k in range(11)], list(range(14)))
    assert lines[1:] == want

    kept = [b for b in blocks if any(po.sphere_sheet_intersect(centers[h], np.float32(r200[h]), b["Origin"], b["Width"],
                                                               b["TotalWidth"]) for h in range(len(halos)))]
    for h in range(len(halos)):
        m = sum(po.mass_contained(b["TotalWidth"], b["xs"], b["ms"], coeffs[h], centers[h].astype(np.float32),
                                  props[h, 9], props[h, 10]) for b in kept)
        assert m > 0 and abs(masses[h] - m) <= 1e-12 * m
        vols = np.array([po.shell_volume(coeffs[h], 50000, s) for s in range(1, 7)])
        assert abs(props[h, 1] - vols.mean()) < 5 * vols.std(ddof=1) / np.sqrt(len(vols))
        assert props[h, 9] < halos[h, 3] * 3 and props[h, 10] > props[h, 9] > 0

    # SkipMass leaves the mass column at zero and reads no particle file (cmd/stats.go:316)
    r2 = subprocess.run([H.cli_path(), "stats", "--SkipMass", "true"], input="\n".join(shell_lines) + "\n",
                        capture_output=True, text=True, env=env, timeout=600)
    assert r2.returncode == 0, r2.stderr
    rows = [l.split() for l in r2.stdout.split("\n")[1:-1]]
    assert all(float(row[2]) == 0.0 for row in rows)
    assert [row[3:] for row in rows] == [l.split()[3:] for l in lines[1:]]
    # ShellParticleFile (cmd/stats.go:326-346): IDs of the particles within ShellWidth * R200m of the
    # surfaces, written as io.WriteFilter does; LGadget-2 IDs of the synthetic files are 0 .. N-1 per file
    sp = os.path.join(os.path.dirname(cfg), "shell-particles.dat")
    r4 = subprocess.run([H.cli_path(), "stats", "--ShellParticleFile", sp, "--ShellWidth", "0.05"],
                        input="\n".join(shell_lines) + "\n", capture_output=True, text=True, env=env, timeout=600)
    assert r4.returncode == 0, r4.stderr
    assert r4.stdout.split("\n")[:-1] == lines            # the catalog itself is unchanged
    raw = open(sp, "rb").read()
    flag, nh = np.frombuffer(raw[:8], np.int32)
    assert flag == 0 and nh == len(halos)
    info = np.frombuffer(raw[8:8 + 32 * nh], np.int64).reshape(nh, 4)
    assert info[:, 0].tolist() == ic[0].tolist() and np.all(info[:, 1] == 100) and np.all(info[:, 2] == 8 + 32 * nh)
    got_ids = np.frombuffer(raw[8 + 32 * nh:], np.int64)
    assert len(got_ids) == info[:, 3].sum()
    want_ids = []
    for h in range(len(halos)):
        for b in kept:
            want_ids += po.shell_particles(b["TotalWidth"], b["xs"], coeffs[h], centers[h], np.float32(r200[h]),
                                           props[h, 9], props[h, 10], 0.05).tolist()
    assert got_ids.tolist() == want_ids and len(want_ids) > 100
    # the pipe sentinel of a failed upstream stage is passed on (shellfish.go:366-373)
    r3 = subprocess.run([H.cli_path(), "stats"], input="Shellfish terminating.\n", capture_output=True, text=True, env=env)
    assert r3.returncode == 1 and r3.stdout == "Shellfish terminating.\n"


def test_prof_angular_fraction_cli(tmp_path):
    """`shellfish shell | shellfish prof --ProfileType angular-fraction` (cmd/prof.go:628-661)."""
    halos, blocks, min_mass, cfg = _setup(tmp_path, 16, 2, 20000, 10000, 1)
    shell_lines = _run_cli(cfg, halos, ["--Rings", "8"])
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg)
    r = subprocess.run([H.cli_path(), "prof", "--ProfileType", "angular-fraction", "--Bins", "40"],
                       input="\n".join(shell_lines) + "\n", capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")[:-1]
    assert lines[0] == "# Column contents: ID(0) Snapshot(1) R [cMpc/h](2-41) Volume Fraction Contained(42-81)"
    ic, fc = H.parse_catalog("\n".join(shell_lines) + "\n", [0, 1], list(range(2, 24)))
    r200, coeffs = fc[3].copy(), fc[4:].T.copy()
    ctx = api.ShellContext(api.default_params(rings=3))
    rs, fs = ctx.angular_fraction(coeffs, 0.03 * r200, 3.0 * r200, 40)
    want = H.format_cols([ic[0], ic[1]], [rs[:, j] for j in range(40)] + [fs[:, j] for j in range(40)], list(range(82)))
    assert lines[1:] == want
    assert np.all(fs[:, 0] == 1.0) and np.all(fs[:, -1] < 1.0)


def test_cli_performance_logging(tmp_path):
    """Logging = performance (cmd/shell.go:352-367): stage times on stderr, the file-read time
    separately; the catalog on stdout is unchanged."""
    halos, blocks, min_mass, cfg = _setup(tmp_path, 17, 1, 10000, 4000, 1)
    quiet = _run_cli(cfg, halos, ["--Rings", "4"])
    open(cfg, "a").write("Logging = performance\n")
    env = dict(os.environ, SHELLFISH_GLOBAL_CONFIG=cfg, SHELLFISH_SEED="1337")
    stdin = "".join("%d 100 %.17g %.17g %.17g %.17g\n" % (i, *h[:4]) for i, h in enumerate(halos))
    os.remove(os.path.join(os.path.dirname(cfg), "memo", "memo.config"))   # the config changed on purpose
    r = subprocess.run([H.cli_path(), "shell", "--Rings", "4"], input=stdin, capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split("\n")[:-1] == quiet
    for needle in ("RNG Seed is 1337", "Snap 100, sphereLoop ended", "Read: ", "Snap 100, haloAnalysis ended", "Device: cull"):
        assert needle in r.stderr, r.stderr
