"""Synthetic Penna-Dines coefficient rows shared by the shell property tests."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sphere(radius, order=3):
    cs = np.zeros(2 * order * order)
    cs[0] = radius
    return cs


def random_shell(rng, order, radius=1.0, amp=0.25):
    """A star-shaped surface: R = radius + smooth perturbations of total size <= amp*radius."""
    n = 2 * order * order
    cs = rng.normal(size=n)
    cs[0] = 0
    cs *= amp * radius / np.abs(cs).sum()
    cs[0] = radius
    return cs


def _tables(fname, prefix):
    txt = open(os.path.join(ROOT, "shellfish_b200", "csrc", fname)).read()

    def arr(name):
        body = re.search(name + r"\[\] = \{([^}]*)\}", txt).group(1)
        return np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()])
    return arr(prefix + "AxisRatios"), arr(prefix + "ACGrid"), arr(prefix + "BCGrid"), arr(prefix + "CGrid")


def axis_tables():
    """The correction tables of the product: the reference's own data
    (los/analyze/ellipse_grid/grid.go through scripts/import_axis_tables.py)."""
    return _tables("sfk_axis_tables_ref.inc", "kRef")


def regenerated_axis_tables():
    """The same grids recomputed from first principles (scripts/make_axis_tables.py)."""
    return _tables("sfk_axis_tables.inc", "k")


# The reference's own known answer for `shellfish shell | shellfish stats`: doc/quickstart.md:99-106
# prints the 18 Penna coefficients (Order 3) of one halo of the author's simulation and
# doc/quickstart.md:150-155 what `stats` made of them -- R_sp, Volume, Surface Area, the three
# axes and the major-axis direction, each a 50 000-sample Monte-Carlo estimate printed with %.6g
# (M_sp needs the particles and cannot be checked).
QUICKSTART_COEFFS = np.array([
    0.979897, -0.0616729, 0.35461, 0.122959, 0.281588, -0.0869048, -0.0560629, 0.0831245, 0.244373,
    -0.0405277, -0.0488017, -0.302187, -0.61362, 0.357011, 2.46993, 0.0243595, 0.525989, 2.74402])
QUICKSTART_STATS = {"r_sp": 1.13591, "volume": 6.13931, "area": 18.3738,
                    "axes": np.array([1.50995, 1.11079, 1.01605]),
                    "a_vec": np.array([-0.928776, -0.35794, -0.096201])}
# one-run relative scatter of the reference's estimators at 50 000 samples for this shell (the
# oracle's restatement over independent sample streams): the printed numbers carry that noise
QUICKSTART_SIGMA = {"r_sp": 1.1e-3, "volume": 3.2e-3, "area": 2.1e-3, "axes": 2.8e-3}


def assert_matches_quickstart(p, n_sigma=3.0):
    """p: a row of shell properties (R_sp, Volume, Surface Area, a, b, c, Ax, Ay, Az, ...)."""
    ref, sig = QUICKSTART_STATS, QUICKSTART_SIGMA
    assert abs(p[0] / ref["r_sp"] - 1) < n_sigma * sig["r_sp"], (p[0], ref["r_sp"])
    assert abs(p[1] / ref["volume"] - 1) < n_sigma * sig["volume"], (p[1], ref["volume"])
    assert abs(p[2] / ref["area"] - 1) < n_sigma * sig["area"], (p[2], ref["area"])
    assert np.all(np.abs(np.asarray(p[3:6]) / ref["axes"] - 1) < n_sigma * sig["axes"]), (p[3:6], ref["axes"])
    # the same major axis up to sign (the reference takes whatever sign its eigen-solver returns)
    assert abs(np.dot(p[6:9], ref["a_vec"])) > 0.995, (p[6:9], ref["a_vec"])
