"""The C-ABI library loads and exports every symbol include/*.h declares.
No compute calls are made here (no GPU in the CPU test tier)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shellfish_b200.h")
LIB = os.path.join(ROOT, "shellfish_b200", "libsfk.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "shellfish_b200", "csrc")], check=True)
    return C.CDLL(LIB)


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfk_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("sfk_create", "sfk_begin_snapshot", "sfk_add_halos", "sfk_push_block",
                 "sfk_finish_snapshot", "sfk_destroy", "sfk_last_error"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), "libsfk.so does not export %s" % name


def test_python_binding_lists_the_same_symbols():
    from shellfish_b200 import api
    assert sorted(api.EXPORTS) == declared_symbols()


def test_params_struct_matches_header_defaults(lib):
    from shellfish_b200 import api
    p = api.default_params()
    # cmd/shell.go:105-119
    assert (p.subsample_factor, p.radial_bins, p.spokes, p.rings) == (1, 256, 256, 100)
    assert (p.r_max_mult, p.r_min_mult, p.r_kernel_mult, p.eta) == (3.0, 0.3, 0.2, 10.0)
    assert (p.order, p.levels, p.smoothing_window) == (3, 3, 121)
    assert (p.los_slope_cutoff, p.background_rho_mult, p.percentile_profile, p.percentile) == (0.0, 0.5, 0, 50.0)


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device sfk_create must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shellfish_b200 import api
    with pytest.raises(api.ShellfishError, match="no CUDA device|CUDA"):
        api.ShellContext(api.default_params(rings=3))


def test_product_never_touches_the_oracle():
    """Nothing under shellfish_b200/ (python or C++) may import, include or
    link oracle/."""
    pkg = os.path.join(ROOT, "shellfish_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout if os.path.exists(LIB) else ""
    assert "liboracle" not in out
