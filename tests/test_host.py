"""CPU tier for the C++ host side (host/*.cpp through libsfhost.so): the config
parser, catalog text, file-name expansion and the three snapshot readers.
Vectors are the reference's own (parse/config_test.go,
cmd/catalog/comment_string_test.go, cmd/env/interleave_test.go), restated here
because /root/reference does not travel."""
import os
import struct

import numpy as np
import pytest

from shellfish_b200 import hostlib as H
from shellfish_b200 import snapio, synth

COSMO = dict(z=0.0, omega_m=0.27, omega_l=0.73, h100=0.7)

SUCCESS = """# Header comment

[config]

float = -1.2e4
 FLOATS = 2.5, 2.5, 2.5

# Body comment

num = 3 # In-line comment
NuMs =1, 1,2, 3,5

    okay = true

okAys = true, false, true

words = dorothy, maddy, sahil
woRd=meow
"""


def test_valid_config():   # parse/config_test.go:323-363
    err, c = H.read_test_config(SUCCESS)
    assert err == ""
    assert c == dict(num=3, float=-1.2e4, okay=True, word="meow", nums=[1, 1, 2, 3, 5], floats=[2.5, 2.5, 2.5],
                     okays=[True, False, True], words=["dorothy", "maddy", "sahil"])


HDR = "I expected the config file test.config to have the header [config] at the top, but didn't find it."


@pytest.mark.parametrize("text,want", [
    ("\n", HDR),                                                              # empty.config
    ("[wrong_header]\n", HDR),                                                # wrong_header.config
    ("[config]\n\nnum = 3\nnums = 1, 2, 3\n\nmeow\n",                        # non_assignment.config
     "I could not parse line 6 of the config file test.config because it did not take the form of a variable "
     "assignment."),
    ("[config]\n\nword=cat\n =dog\nwords=1,2,3\n",                            # no_variable.config
     "I could not parse line 4 of the config file test.config because it did not take the form of a variable "
     "assignment."),
    ("[config]\n\nword = cat\n\nword = dog\n",                                # dupicates.config
     "Lines 3 and 5 of the config file test.config both assign a value to the variable 'word'."),
    ("[config]\n\ncat=3\n",                                                   # invalid_var.config
     "Line 3 of the config file test.config assigns a value to the variable 'cat', but config files of type "
     "config don't have that variable."),
])
def test_invalid_config(text, want):   # parse/config_test.go:365-389, messages of parse/config.go:300-333
    err, _ = H.read_test_config(text)
    assert err == want


def test_invalid_type_config():   # invalid_type.config; message of parse/config.go:349-355
    err, _ = H.read_test_config("[config]\n\nnum=cat\n")
    assert err.startswith("I could not parse line 3 of the config file test.config because 'num' expects values "
                          "of type int and '")
    assert err.endswith("' cannnot be converted to an int.")


def test_valid_flags():   # parse/config_test.go:391-424 (flag names are case-insensitive, any number of dashes)
    flags = ["--Num", "16", "--Nums", "1, 2, 3, 4, 5", "--Float", "16", "--Floats", "1", "2", "3", "4", "5",
             "--Okay", "true", "---Okays", "true, true", "false"]
    err, c = H.read_test_config(None, flags)
    assert err == ""
    assert c["num"] == 16 and c["nums"] == [1, 2, 3, 4, 5] and c["float"] == 16
    assert c["floats"] == [1, 2, 3, 4, 5] and c["okay"] and c["okays"] == [True, True, False]


def test_flags_override_file():
    err, c = H.read_test_config(SUCCESS, ["--num", "7"])
    assert err == "" and c["num"] == 7 and c["word"] == "meow"


def test_invalid_flag():
    err, _ = H.read_test_config(None, ["--Meow", "3"])
    assert err != ""
    err, _ = H.read_test_config(None, ["--Num", "cat"])
    assert err != ""


COMMENT = [   # cmd/catalog/comment_string_test.go:14-27
    (["A"], [], [0], [1], "# Column contents: A(0)"),
    ([], ["A"], [0], [1], "# Column contents: A(0)"),
    (["A"], [], [0], [11], "# Column contents: A(0-10)"),
    (["A"], ["B"], [0, 1], [1, 1], "# Column contents: A(0) B(1)"),
    (["A"], ["B"], [1, 0], [1, 1], "# Column contents: B(0) A(1)"),
    (["A"], ["B"], [0, 1], [1, 2], "# Column contents: A(0) B(1-2)"),
    (["A"], ["B"], [1, 0], [1, 2], "# Column contents: B(0-1) A(2)"),
]


@pytest.mark.parametrize("ints,floats,order,sizes,want", COMMENT)
def test_comment_string(ints, floats, order, sizes, want):
    assert H.comment_string(ints, floats, order, sizes) == want


def test_comment_string_reference_case8():
    # The 8th vector of comment_string_test.go:28-29 expects "A(0) B(1-2) C(3)" for
    # order [0, 1], which catalog.go:26-36 cannot produce (only len(order) tokens are
    # emitted and token 1 is the second INT name).  The code is the behaviour a user
    # sees, so it is what we mirror.
    assert H.comment_string(["A", "C"], ["B"], [0, 1], [1, 2]) == "# Column contents: A(0) C(1-2)"


def test_shell_header_line():   # cmd/shell.go:255-262 with the default Order = 3
    got = H.comment_string(["ID", "Snapshot"], ["X [cMpc/h]", "Y [cMpc/h]", "Z [cMpc/h]", "R200m [cMpc/h]", "P_ijk"],
                           [0, 1, 2, 3, 4, 5, 6], [1, 1, 1, 1, 1, 1, 18])
    assert got == ("# Column contents: ID(0) Snapshot(1) X [cMpc/h](2) Y [cMpc/h](3) Z [cMpc/h](4) "
                   "R200m [cMpc/h](5) P_ijk(6-23)")


def test_go_g6():   # fmt "%.6g" as strconv formats it (catalog.go:118)
    for v, want in [(0.0, "0"), (1.0, "1"), (1.5, "1.5"), (123456.0, "123456"), (1234567.0, "1.23457e+06"),
                    (1e-4, "0.0001"), (1e-5, "1e-05"), (-2.5e-10, "-2.5e-10"), (0.1 + 0.2, "0.3"),
                    (float("nan"), "NaN"), (float("inf"), "+Inf"), (float("-inf"), "-Inf"), (1e21, "1e+21"),
                    (99999.95, "99999.9"), (999999.5, "1e+06"), (62.5, "62.5")]:
        assert H.go_g6(v) == want, (v, H.go_g6(v))


def test_format_cols():   # catalog.go:52-141: right-aligned to the widest entry, single spaces
    lines = H.format_cols([[1, 20], [100, 100]], [[1.5, 2.25], [1e-7, 3.0]], [0, 1, 2, 3])
    assert lines == [" 1 100  1.5 1e-07", "20 100 2.25     3"]
    assert H.format_cols([[1, 2]], [[0.5, 0.25]], [1, 0]) == [" 0.5 1", "0.25 2"]
    assert H.format_cols([], [], []) == []


def test_format_cols_reference_documents():
    """Rows the reference's documentation prints (its only catalog text besides the unit tests):
    doc/quickstart.md:105 (one `shell` row: no padding, "%.6g") and doc/walkthrough.md:86-89 (three
    `stats` rows: every column right-aligned to its widest entry -- "0.3129  0.429")."""
    import shell_cases as shells
    row = [13.6225, 86.3578, 53.1017, 0.815028] + list(shells.QUICKSTART_COEFFS)
    lines = H.format_cols([[80431577], [100]], [[v] for v in row], list(range(2 + len(row))))
    assert lines == ["80431577 100 13.6225 86.3578 53.1017 0.815028 0.979897 -0.0616729 0.35461 0.122959 0.281588 "
                     "-0.0869048 -0.0560629 0.0831245 0.244373 -0.0405277 -0.0488017 -0.302187 -0.61362 0.357011 "
                     "2.46993 0.0243595 0.525989 2.74402"]
    ints, flts = H.parse_catalog(lines[0] + "\n", [0, 1], list(range(2, 24)))
    assert ints[:, 0].tolist() == [80431577, 100] and flts[:, 0].tolist() == row
    lines = H.format_cols([[169665239, 168208646, 168863226], [100, 100, 100]],
                          [[1.373e12, 1.244e12, 1.284e12], [0.4084, 0.3683, 0.3576], [0.3121, 0.3129, 0.2613],
                           [0.5297, 0.429, 0.4081]], [0, 1, 2, 3, 4, 5])
    assert lines == ["169665239 100 1.373e+12 0.4084 0.3121 0.5297",
                     "168208646 100 1.244e+12 0.3683 0.3129  0.429",
                     "168863226 100 1.284e+12 0.3576 0.2613 0.4081"]


def test_parse_catalog():   # catalog.go:143-231
    text = "# comment\n1 100 1.5 2.5 3.5 0.5\n\n  2 100 4 5 6 0.25 # trailing\n"
    ic, fc = H.parse_catalog(text, [0, 1], [2, 3, 4, 5])
    assert ic.tolist() == [[1, 2], [100, 100]]
    assert fc.tolist() == [[1.5, 4.0], [2.5, 5.0], [3.5, 6.0], [0.5, 0.25]]
    with pytest.raises(ValueError):
        H.parse_catalog("1 100 1.5\n", [0, 1], [2, 3, 4, 5])       # too few columns
    with pytest.raises(ValueError):
        H.parse_catalog("1 x 1 2 3 4\n", [0, 1], [2, 3, 4, 5])     # not an int


def test_format_parse_round_trip():
    rng = np.random.default_rng(0)
    ids = np.arange(50)
    vals = rng.normal(size=(3, 50)) * 10.0 ** rng.integers(-8, 8, size=(3, 50))
    lines = H.format_cols([ids], vals, [0, 1, 2, 3])
    ic, fc = H.parse_catalog("\n".join(lines) + "\n", [0], [1, 2, 3])
    assert ic[0].tolist() == ids.tolist()
    np.testing.assert_allclose(fc, vals, rtol=5e-6)


# ---- cmd/env/interleave_test.go through the file-name expansion -----------------

def _names(fmt, meanings, mins, maxes, smin, smax):
    _, nb = H.catalog_name(fmt, meanings, mins, maxes, smin, smax, smin, 0)
    return [[H.catalog_name(fmt, meanings, mins, maxes, smin, smax, s, b)[0] for b in range(nb)]
            for s in range(smin, smax + 1)]


def test_interleave_singleton():   # TestSingletonInterleave
    assert _names("%d", ["Snapshot"], [], [], 1, 3) == [["1"], ["2"], ["3"]]


def test_interleave_pair():   # TestPairInterleave cases 2-4
    assert _names("%d,%d", ["Snapshot", "Block"], [4], [4], 1, 3) == [["1,4"], ["2,4"], ["3,4"]]
    assert _names("%d,%d", ["Snapshot", "Block"], [4], [5], 1, 3) == [["1,4", "1,5"], ["2,4", "2,5"], ["3,4", "3,5"]]
    assert _names("%d,%d", ["Block", "Snapshot"], [4], [5], 1, 3) == [["4,1", "5,1"], ["4,2", "5,2"], ["4,3", "5,3"]]


def test_interleave_multi():   # TestMultiInterleave
    assert _names("%d,%d,%d", ["Snapshot", "Block0", "Block1"], [3, 5], [4, 6], 1, 2) == [
        ["1,3,5", "1,4,5", "1,3,6", "1,4,6"], ["2,3,5", "2,4,5", "2,3,6", "2,4,6"]]
    got = _names("%d,%d,%d,%d", ["Snapshot", "Block0", "Block1", "Block2"], [3, 5, 6], [4, 5, 8], 1, 2)
    assert got[0] == ["1,3,5,6", "1,4,5,6", "1,3,5,7", "1,4,5,7", "1,3,5,8", "1,4,5,8"]
    assert got[1] == ["2,3,5,6", "2,4,5,6", "2,3,5,7", "2,4,5,7", "2,3,5,8", "2,4,5,8"]


def test_snapshot_format_verbs():
    name, nb = H.catalog_name("/data/snapdir_%03d/snapshot_%03d.%d", ["Snapshot", "Snapshot", "Block"], [0], [511],
                              0, 100, 100, 37)
    assert nb == 512 and name == "/data/snapdir_100/snapshot_100.37"


# ---- snapshot readers -------------------------------------------------------------

def _box(seed=3, n=4000, box=62.5):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3)) * box).astype(np.float32)


def test_lgadget2_reader(tmp_path):   # io/lgadget2.go
    box, n_total = 62.5, 1 << 20
    xyz = _box()
    xyz[0] = [-0.25, box + 0.5, 1.0]   # wrapped by the reader, lgadget2.go:123-131
    f = str(tmp_path / "snap.0")
    snapio.write_lgadget2(f, xyz, box, COSMO, n_total=n_total)
    xs, ms, hd, min_mass, tot = H.read_block("LGadget-2", f)
    want = xyz.copy()
    want[0] = [np.float32(-0.25) + np.float32(box), np.float32(box + 0.5) - np.float32(box), 1.0]
    assert np.array_equal(xs, want)
    assert tot == n_total and hd["n"] == len(xyz) and hd["total_width"] == box
    assert (hd["z"], hd["omega_m"], hd["omega_l"], hd["h100"]) == (0.0, 0.27, 0.73, 0.7)
    m = synth.uniform_mass(n_total, box, dict(OmegaM=0.27, OmegaL=0.73, h100=0.7))
    assert min_mass == pytest.approx(float(m), rel=1e-6)
    assert np.all(ms == np.float32(min_mass))


def test_lgadget2_big_endian_and_npart1(tmp_path):
    xyz = _box(n=100)
    f = str(tmp_path / "be.0")
    snapio.write_lgadget2(f, xyz, 62.5, COSMO, order=">", npart_num=1)
    xs, _, hd, _, tot = H.read_block("LGadget-2", f, endianness="BigEndian", npart_num=1)
    assert np.array_equal(xs, xyz) and hd["n"] == 100 and tot == 100


def test_lgadget2_corruption(tmp_path):   # lgadget2.go:133-141
    xyz = _box(n=10)
    xyz[3, 1] = np.nan
    f = str(tmp_path / "bad.0")
    snapio.write_lgadget2(f, xyz, 62.5, COSMO)
    with pytest.raises(ValueError, match="Corruption detected in the file"):
        H.read_block("LGadget-2", f)


def test_gadget2_reader(tmp_path):   # io/gadget2.go: DM = type 1, single mass from the header table
    box = 62500.0   # kpc/h in the file, GadgetPositionUnits = 1e-3
    gas, dm = _box(1, 50, box), _box(2, 300, box)
    f = str(tmp_path / "g2.0")
    snapio.write_gadget2(f, [gas, dm, None, None, None, None], box, COSMO, [0, 0.0123, 0, 0, 0, 0],
                         masses_by_type=[np.full(50, 0.5, np.float32), None, None, None, None, None])
    xs, ms, hd, min_mass, tot = H.read_block("Gadget-2", f, pos_units=1e-3, mass_units=1e10)
    assert len(xs) == 300 and hd["n"] == 300
    np.testing.assert_allclose(xs, dm * np.float32(1e-3), rtol=1e-6)
    np.testing.assert_allclose(ms, np.float32(0.0123) * np.float32(1e10), rtol=1e-6)
    assert hd["total_width"] == pytest.approx(62.5)
    assert min_mass == 0.0   # the Gadget-2 buffer never sets its mass (io/gadget2.go:403)


def test_gotetra_reader(tmp_path):   # io/gotetra.go:69-93: segment = grid minus the duplicated upper faces
    sw, cw, cells, box = 4, 8, 2, 62.5
    gw = sw + 1
    rng = np.random.default_rng(5)
    grid = (rng.random((gw ** 3, 3)) * box).astype(np.float32)
    f = str(tmp_path / "sheet000.dat")
    snapio.write_gotetra(f, grid, 0, cells, sw, cw, box, COSMO, origin=[1, 2, 3], width=[30, 31, 32])
    xs, ms, hd, min_mass, tot = H.read_block("gotetra", f)
    g = grid.reshape(gw, gw, gw, 3)[:sw, :sw, :sw].reshape(-1, 3)   # z, y, x order
    assert np.array_equal(xs, g)
    assert hd["n"] == sw ** 3 and tot == -1
    assert hd["origin"].tolist() == [1, 2, 3] and hd["width"].tolist() == [30, 31, 32]
    m = synth.uniform_mass(cw ** 3, box, dict(OmegaM=0.27, OmegaL=0.73, h100=0.7))
    assert min_mass == pytest.approx(float(m), rel=1e-6) and np.all(ms == np.float32(min_mass))


def test_bounding_box():   # io/io.go:263-304
    o, w = H.bounding_box(np.array([[1, 2, 3], [4, 6, 5]], np.float32), 100.0)
    assert o.tolist() == [1, 2, 3] and w.tolist() == [3, 4, 2]
    # a block that straddles the periodic boundary: the short way round
    o, w = H.bounding_box(np.array([[99, 50, 50], [1, 51, 52]], np.float32), 100.0)
    assert o[0] == 99 and w[0] == pytest.approx(2) and w[1] == pytest.approx(1)


# ---- GlobalConfig / ShellConfig ------------------------------------------------------

GLOBAL = """[config]
Version = 1.0.4
SnapshotFormat = %s/snap_%%03d.%%d
SnapshotFormatMeanings = Snapshot, Block
SnapshotType = LGadget-2
BlockMins = 0
BlockMaxes = 7
SnapMin = 100
SnapMax = 100
MemoDir = %s
"""


def test_global_config(tmp_path):   # cmd/cmd.go:98-356
    cfg = tmp_path / "g.config"
    memo = tmp_path / "memo"
    cfg.write_text(GLOBAL % (tmp_path, memo))
    err, snap_type, blocks = H.read_global_config(str(cfg))
    assert err == "" and snap_type == "LGadget-2" and blocks == 8
    assert memo.is_dir()   # validateDir creates it, cmd/cmd.go:372-381


@pytest.mark.parametrize("edit,needle", [
    (("Version = 1.0.4", "Version = 1.1.0"), "from the future"),
    (("Version = 1.0.4", "Version = meow"), "couldn't parse the 'Version'"),
    (("SnapshotType = LGadget-2", "SnapshotType = HDF5"), "which I don't recognize"),
    (("SnapshotType = LGadget-2\n", ""), "'SnapshotType variable isn't set"),
    (("BlockMins = 0", "BlockMins = 0\nEndianness = Middle"), "'Endianness'"),
    (("BlockMins = 0", "BlockMins = 0\nLGadgetNpartNum = 3"), "only valid values are 1 and 2"),
    (("BlockMins = 0", "BlockMins = 0\nHaloType = Text"), "'HaloPositionUnits' variable isn't set"),
])
def test_global_config_errors(tmp_path, edit, needle):
    cfg = tmp_path / "g.config"
    text = GLOBAL % (tmp_path, tmp_path / "memo")
    assert edit[0] in text
    cfg.write_text(text.replace(*edit))
    err, _, _ = H.read_global_config(str(cfg))
    assert needle in err, err


def test_shell_config_defaults_and_flags(tmp_path):   # cmd/shell.go:102-187
    err, c = H.read_shell_config()
    assert err == ""
    assert c == dict(SubsampleFactor=1, RadialBins=256, Spokes=256, Rings=100, RMaxMult=3.0, RMinMult=0.3,
                     RKernelMult=0.2, Eta=10.0, Order=3, Levels=3, SmoothingWindow=121, LOSSlopeCutoff=0.0,
                     BackgroundRhoMult=0.5, PercentileProfile=False, Percentile=50.0)
    f = tmp_path / "shell.config"
    f.write_text("[shell.config]\nRings = 24\nEta = 5\n")
    err, c = H.read_shell_config(str(f), ["--Spokes", "128"])
    assert err == "" and c["Rings"] == 24 and c["Eta"] == 5.0 and c["Spokes"] == 128
    err, _ = H.read_shell_config(None, ["--Rings", "0"])
    assert err == "The variable 'Rings' was set to 0."
    err, _ = H.read_shell_config(None, ["--RMinMult", "4"])
    assert err == "The variable 'RMinMult' was set to 4, but the variable 'RMaxMult' was set to 3."


# ---- CLI protocol without a GPU --------------------------------------------------------

def _cli(args, stdin="", env=None):
    import subprocess
    e = dict(os.environ)
    e.pop("SHELLFISH_GLOBAL_CONFIG", None)
    e.update(env or {})
    return subprocess.run([H.cli_path()] + args, input=stdin, capture_output=True, text=True, env=e, timeout=120)


def test_cli_basics():   # shellfish.go:305-349
    r = _cli([])
    assert r.returncode == 1 and "I was not supplied with a mode." in r.stderr
    assert _cli(["version"]).stdout == "Shellfish version 1.0.4\n"
    assert "Installation was successful" in _cli(["hello"]).stdout
    assert "RadialBins" in _cli(["help", "shell.config"]).stdout
    assert _cli(["shell"], stdin="").returncode == 0   # empty stdin: nothing to do


def test_cli_upstream_termination():   # shellfish.go:364-369: a failed stage upstream ends the pipe
    r = _cli(["shell"], stdin="Shellfish terminating.\n")
    assert r.returncode == 1 and r.stdout == "Shellfish terminating.\n"


def test_cli_errors(tmp_path):   # shellfish.go:436-444
    r = _cli(["shell"], stdin="1 100 1 1 1 1\n")
    assert r.returncode == 1 and "$SHELLFISH_GLOBAL_CONFIG has not been set." in r.stderr
    assert r.stderr.startswith("Error running mode shell:\n") and r.stdout == "Shellfish terminating.\n"
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL % (tmp_path, tmp_path / "memo"))
    r = _cli(["shell", "--Rings", "-3"], stdin="1 100 1 1 1 1\n", env={"SHELLFISH_GLOBAL_CONFIG": str(cfg)})
    assert r.returncode == 1 and "The variable 'Rings' was set to -3." in r.stderr


def test_cli_no_gpu_fails_loudly(tmp_path):
    """Without a GPU the CLI must stop with the library's error: no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    halos, blocks, mass = synth.single_halo(seed=4, n=2000, n_background=500)
    n_total = sum(len(b["xs"]) for b in blocks)
    snapio.write_box_lgadget2(str(tmp_path), "snap_%03d.%d", blocks)
    cfg = tmp_path / "g.config"
    cfg.write_text((GLOBAL % (tmp_path, tmp_path / "memo")).replace("BlockMaxes = 7", "BlockMaxes = 0"))
    line = "0 100 %.6g %.6g %.6g %.6g\n" % tuple(halos[0][:4])
    r = _cli(["shell", "--Rings", "6"], stdin=line, env={"SHELLFISH_GLOBAL_CONFIG": str(cfg)})
    assert r.returncode == 1 and r.stdout == "Shellfish terminating.\n"
    assert "Error running mode shell:" in r.stderr
    # the header memo was still written by the host side before the device was needed
    hd = (tmp_path / "memo" / "hd_snap100.dat").read_bytes()
    assert len(hd) == 72
    z, om, ol, h, n, tw = struct.unpack("<4dqd", hd[:48])
    assert (om, ol, h, n, tw) == (blocks[0]["OmegaM"], blocks[0]["OmegaL"], blocks[0]["H100"], n_total, 62.5)
    assert (tmp_path / "memo" / "memo.config").exists()


def test_cli_snapshot_column_checks(tmp_path):
    """Snap == -1 rows are skipped (cmd/shell.go:325); the reference indexes its file table with -1
    before it gets there and panics, the C++ host reports instead."""
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL % (tmp_path, tmp_path / "memo"))
    env = {"SHELLFISH_GLOBAL_CONFIG": str(cfg)}
    r = _cli(["shell"], stdin="1 -1 1 1 1 1\n2 -1 2 2 2 1\n", env=env)
    assert r.returncode == 1 and "Every row of the halo table has Snapshot = -1" in r.stderr
    assert r.stdout == "Shellfish terminating.\n"
    r = _cli(["shell"], stdin="1 7 1 1 1 1\n", env=env)
    assert r.returncode == 1 and "Snapshot 7 is outside [SnapMin, SnapMax] = [100, 100]." in r.stderr


def test_gadget2_multi_mass_species_and_32bit_ids(tmp_path):
    """Gas with its own masses, DM type 1 with the table mass, DM type 2 with its own masses:
    unpackMass / packVec (io/gadget2.go:202-295) keep types 1 and 2; ids may be 32 bit (:110-120)."""
    box = 1000.0
    gas, dm1, dm2 = _box(1, 20, box), _box(2, 100, box), _box(3, 50, box)
    m_gas = np.linspace(1, 2, 20).astype(np.float32)
    m_dm2 = np.linspace(5, 9, 50).astype(np.float32)
    f = str(tmp_path / "multi.0")
    snapio.write_gadget2(f, [gas, dm1, dm2, None, None, None], box, COSMO, [0, 0.25, 0, 0, 0, 0],
                         masses_by_type=[m_gas, None, m_dm2, None, None, None], id_bytes=4)
    xs, ms, hd, _, tot = H.read_block("Gadget-2", f, dm_types=(1, 2), single_mass=(1,))
    assert hd["n"] == 150 and tot == 150
    assert np.array_equal(xs, np.concatenate([dm1, dm2]))
    assert np.array_equal(ms, np.concatenate([np.full(100, 0.25, np.float32), m_dm2]))


def test_gadget2_big_endian(tmp_path):
    """The particle read honours Endianness (io/gadget2.go:83); the stand-alone header read is
    hard-wired little endian (:60), so ReadHeader on a big-endian file sees a byte-swapped header --
    the reference's behaviour, reproduced: only Read() is checked here."""
    dm = _box(4, 64, 500.0)
    f = str(tmp_path / "be.0")
    snapio.write_gadget2(f, [None, dm, None, None, None, None], 500.0, COSMO, [0, 2.0, 0, 0, 0, 0], order=">")
    try:
        xs, ms, hd, _, _ = H.read_block("Gadget-2", f, endianness="BigEndian")
    except ValueError:
        return   # a swapped header may fail validation before the particles are reached: also the reference's fate
    assert np.array_equal(xs, dm) and np.all(ms == 2.0)


# ---- `shellfish stats` host side ---------------------------------------------------------

def test_stats_config_defaults_flags_and_validation(tmp_path):   # cmd/stats.go:109-172
    err, c = H.read_stats_config()
    assert err == ""
    assert c == dict(MonteCarloSamples=50000, Order=3, ShellWidth=0.0, SkipMass=False, shellFilter=False,
                     ExclusionStrategy="none")
    f = tmp_path / "stats.config"
    f.write_text("[stats.config]\nOrder = 4\nExclusionStrategy = overlap\nValues = m_sp, r_sp\n")
    err, c = H.read_stats_config(str(f), ["--SkipMass", "true"])
    assert err == "" and c["Order"] == 4 and c["ExclusionStrategy"] == "overlap" and c["SkipMass"]
    # shellFilter needs both a width and a file (cmd/stats.go:129-130,143-144)
    assert H.read_stats_config(None, ["--ShellWidth", "0.05"])[1]["shellFilter"] is False
    assert H.read_stats_config(None, ["--ShellWidth", "0.05", "--ShellParticleFile", "p.dat"])[1]["shellFilter"] is True
    err, _ = H.read_stats_config(None, ["--Values", "m_sp,volume"])
    assert err == "Item 1 of variable 'Values' is set to 'volume', which I don't recognize."
    err, _ = H.read_stats_config(None, ["--ExclusionStrategy", "all"])
    assert err == "variable 'ExclusionStrategy' set to 'all', which I don't recognize."
    err, _ = H.read_stats_config(None, ["--MonteCarloSamples", "0"])
    assert err == "The variable 'MonteCarloSamples' was set to %!g(int64=0)"   # Go's %g on an int64 (:167-168)
    f.write_text("[shell.config]\nOrder = 4\n")
    err, _ = H.read_stats_config(str(f))
    assert "stats.config" in err


def test_stats_cli_protocol_without_gpu(tmp_path):
    assert "MonteCarloSamples" in _cli(["help", "stats.config"]).stdout
    assert _cli(["stats"], stdin="").returncode == 0
    r = _cli(["stats"], stdin="Shellfish terminating.\n")
    assert r.returncode == 1 and r.stdout == "Shellfish terminating.\n"
    row = "0 100 1 1 1 1 " + " ".join(["0.5"] + ["0"] * 17) + "\n"
    r = _cli(["stats"], stdin=row)
    assert r.returncode == 1 and r.stderr.startswith("Error running mode stats:\n")
    assert "$SHELLFISH_GLOBAL_CONFIG has not been set." in r.stderr and r.stdout == "Shellfish terminating.\n"
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL % (tmp_path, tmp_path / "memo"))
    env = {"SHELLFISH_GLOBAL_CONFIG": str(cfg)}
    r = _cli(["stats", "--ExclusionStrategy", "most"], stdin=row, env=env)
    assert r.returncode == 1 and "which I don't recognize." in r.stderr
    # a shell catalog of another order has the wrong number of columns (catalog.Parse, cmd/stats.go:197-202)
    r = _cli(["stats", "--Order", "4"], stdin=row, env=env)
    assert r.returncode == 1 and r.stdout == "Shellfish terminating.\n"
    # ShellParticleFile is accepted (cmd/stats.go:114-141); the run then stops where any stats run does here
    r = _cli(["stats", "--ShellWidth", "0.1", "--ShellParticleFile", "x.dat"], stdin=row, env=env)
    assert r.returncode == 1 and "not supported" not in r.stderr
    import torch
    if not torch.cuda.is_available():   # no CPU path: the device library refuses
        r = _cli(["stats", "--SkipMass", "true"], stdin=row, env=env)
        assert r.returncode == 1 and "Error running mode stats:" in r.stderr and r.stdout == "Shellfish terminating.\n"


def test_prof_config(tmp_path):   # cmd/prof.go:103-169
    err, c = H.read_prof_config()
    assert err == "" and c["Bins"] == 150 and c["RMinMult"] == 0.03 and c["RMaxMult"] == 3.0 and c["Samples"] == 50000
    assert H.read_prof_config(None, ["--Bins", "10"])[0] == "The variable 'ProfileType' was not set."
    assert H.read_prof_config(None, ["--ProfileType", "mean"])[0] == "The varaiable 'ProfileType' was set to 'mean'."
    err, c = H.read_prof_config(None, ["--ProfileType", "angular-fraction", "--Bins", "40"])
    assert err == "" and c["ProfileType"] == "angular-fraction" and c["Bins"] == 40
    assert H.read_prof_config(None, ["--ProfileType", "density", "--Bins", "-2"])[0] == "The variable 'Bins' was set to -2."
    # the reference reports RMinMult (name and value) when RMaxMult is the bad one (cmd/prof.go:160-162)
    assert H.read_prof_config(None, ["--ProfileType", "density", "--RMaxMult", "-1"])[0] == \
        "The variable 'RMinMult' was set to 0.03."
    assert H.read_prof_config(None, ["--ProfileType", "density", "--MedianPixelLevel", "-1"])[0] == \
        "The variable 'MedianPixelLevel' was set to %!g(int64=-1)."
    r = _cli(["prof", "--ProfileType", "angular-fraction"], stdin="Shellfish terminating.\n")
    assert r.returncode == 1 and r.stdout == "Shellfish terminating.\n"
    cfg = tmp_path / "g.config"
    cfg.write_text(GLOBAL % (tmp_path, tmp_path / "memo"))
    row = "0 100 1 1 1 1 " + " ".join(["0.5"] + ["0"] * 17) + "\n"
    r = _cli(["prof", "--ProfileType", "median-density"], stdin=row, env={"SHELLFISH_GLOBAL_CONFIG": str(cfg)})
    assert r.returncode == 1 and "is not provided by this build" in r.stderr and r.stdout == "Shellfish terminating.\n"


def test_example_configs_parse(tmp_path):
    """cmd/config_file_test.go:9-39 (TestExampleFiles): every mode's ExampleConfig() is a file its
    own ReadConfig accepts."""
    for name, reader in (("shell.config", H.read_shell_config), ("stats.config", H.read_stats_config),
                         ("prof.config", H.read_prof_config)):
        text = _cli(["help", name]).stdout
        assert text.startswith("[%s]" % name)
        f = tmp_path / name
        f.write_text(text)
        err, _ = reader(str(f))
        assert err == "", (name, err)


def test_truncated_snapshots_are_reported(tmp_path):
    """A header that promises more particles than the file holds ends in the reader's error
    text (io.ReadFull's unexpected EOF in the reference), not in an allocation failure."""
    dm = _box(9, 1000, 100.0)
    f = str(tmp_path / "g2.0")
    snapio.write_gadget2(f, [None, dm, None, None, None, None], 100.0, COSMO, [0, 1.0, 0, 0, 0, 0])
    raw = open(f, "rb").read()
    open(f, "wb").write(raw[:4 + 256 + 4 + 4 + 1200])          # header + 100 of the 1000 positions
    with pytest.raises(ValueError, match="unexpected EOF"):
        H.read_block("Gadget-2", f)
    hdr = bytearray(raw)
    struct.pack_into("<I", hdr, 4 + 4, 0x7fffffff)             # NPart[1] := 2^31 - 1 in an otherwise intact file
    open(f, "wb").write(bytes(hdr))
    with pytest.raises(ValueError, match="unexpected EOF"):
        H.read_block("Gadget-2", f)


def test_cli_halo_partition_equals_the_python_one():
    """PartitionHalos (host/sf_shell.cpp, SHELLFISH_GPUS) == shard.partition_halos."""
    from shellfish_b200 import shard
    rng = np.random.default_rng(4)
    for n, ranks in ((1000, 8), (37, 3), (5, 8), (64, 1)):
        w = rng.uniform(0.3, 3.0, n)
        grp = rng.integers(-1, max(2, n // 2), n)
        own = H.partition_halos(w, grp, ranks)
        parts = shard.partition_halos(w, ranks, grp)
        want = np.empty(n, np.int32)
        for r, p in enumerate(parts):
            want[p] = r
        assert np.array_equal(own, want)


def test_cli_spatial_partition_equals_the_python_one():
    """PartitionHalosSpatial (what SHELLFISH_GPUS uses) == shard.partition_spatial (what bench.py uses)."""
    from shellfish_b200 import shard
    rng = np.random.default_rng(6)
    for n, ranks in ((1000, 8), (37, 3), (5, 8), (64, 1), (0, 4)):
        w = rng.uniform(0.3, 3.0, n)
        xyz = rng.random((n, 3)) * 250.0
        own = H.partition_spatial(w, xyz, 250.0, ranks)
        parts = shard.partition_spatial(w, xyz, 250.0, ranks)
        want = np.zeros(n, np.int32)
        for r, p in enumerate(parts):
            want[p] = r
        assert np.array_equal(own, want)
        if n >= 100:
            loads = np.array([w[p].sum() for p in parts])
            assert loads.max() / loads.mean() < 1.03


def test_filter_file_is_the_references_bytes(tmp_path):
    """io.WriteFilter, io/filter_data.go:36-63: flag, count, {ID, Snap, StartByte, Len} per halo,
    then the ID lists.  StartByte is the same for every halo (the reference never advances totLen)."""
    import struct
    parts = [[5, 7, 9], [], [11]]
    for endian, fmt in (("LittleEndian", "<"), ("BigEndian", ">"), ("SystemOrder", "<")):
        path = str(tmp_path / ("f_%s.dat" % endian))
        H.write_filter(path, endian, [100, 100, 99], [3, 4, 8], parts)
        raw = open(path, "rb").read()
        flag, n = struct.unpack(fmt + "ii", raw[:8])
        assert struct.unpack("<i", raw[:4])[0] == (-1 if fmt == ">" else 0) or fmt == ">"   # ReadFilter reads the flag little endian
        assert flag == (-1 if fmt == ">" else 0) and n == 3
        info = struct.unpack(fmt + "12q", raw[8:8 + 96])
        base = 4 + 4 + 3 * 32
        assert info == (3, 100, base, 3, 4, 100, base, 0, 8, 99, base, 1)
        assert struct.unpack(fmt + "4q", raw[8 + 96:]) == (5, 7, 9, 11) and len(raw) == 8 + 96 + 32


def test_readers_return_particle_ids(tmp_path):
    """ReadIDs: LGadget-2 int64 block (io/lgadget2.go:107-116), Gadget-2 4- or 8-byte IDs packed to
    the DM species (io/gadget2.go:109-133), gotetra zeros (io/gotetra.go:40,87)."""
    from shellfish_b200 import snapio
    rng = np.random.default_rng(0)
    xs = (rng.random((50, 3)) * 10).astype(np.float32)
    cosmo = dict(z=0.0, omega_m=0.27, omega_l=0.73, h100=0.7)
    p = str(tmp_path / "l.0")
    snapio.write_lgadget2(p, xs, 10.0, cosmo)
    assert np.array_equal(H.read_ids("LGadget-2", p), np.arange(50))
    g = str(tmp_path / "g.0")
    snapio.write_gadget2(g, [xs[:10], xs[10:40], None, None, None, None], 10.0, cosmo, [0.5, 1.0, 0, 0, 0, 0])
    assert np.array_equal(H.read_ids("Gadget-2", g), np.arange(10, 40))          # type 1 only, file order
    assert np.array_equal(H.read_ids("Gadget-2", g, dm_types=(0, 1)), np.arange(40))
