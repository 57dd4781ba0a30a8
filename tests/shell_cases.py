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
