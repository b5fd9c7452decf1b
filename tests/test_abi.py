"""The C-ABI shared library: loads without a GPU, exports every symbol include/help_b200.h
declares, validates arguments, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os.path as osp
import re

import numpy as np
import pytest

from pyhelp_b200 import _native as N
from pyhelp_b200 import build as B

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.lib()


def declared_functions():
    src = open(osp.join(ROOT, "include", "help_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(help_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/help_b200.h but not exported"


def test_header_cites_reference():
    src = open(osp.join(ROOT, "include", "help_b200.h")).read()
    for needle in ("pyhelp/processing.py:40-59", "HELP3O.FOR:17-39", "pyhelp/processing.py:64-147"):
        assert needle in src


def test_abi_version_and_enums(lib):
    assert lib.help_b200_abi_version() == 1
    src = open(osp.join(ROOT, "include", "help_b200.h")).read()
    # the python mirror of the enums must match the header order
    block = re.search(r"enum \{\s*HB_ULAT = 0,(.*?)HB_NCELL_F32", src, re.S).group(1)
    assert len(re.findall(r"HB_[A-Z0-9]+", block)) + 1 == N.HB_NCELL_F32
    assert N.HR_NROWS == 7 and N.HY_NROWS == 5 and N.RAW_ROWS == 14 and N.DAYS == 368


def test_struct_layout_matches_header(lib):
    # 6 int32 + 14 pointers; result: 5 pointers + 4 floats + 1 int32 (padded to 8)
    assert C.sizeof(N.Batch) == 6 * 4 + 14 * 8
    assert C.sizeof(N.Result) == 5 * 8 + 4 * 4 + 4 + 4


def test_invalid_arguments_return_einval(lib):
    h = C.c_void_p()
    assert lib.help_b200_engine_create(None, 1, 0, C.byref(h)) == -1
    assert b"NULL" in lib.help_b200_last_error()
    b = N.Batch()
    b.n_cells, b.n_years, b.max_layers = 4, 0, 3
    assert lib.help_b200_engine_create(C.byref(b), 1, 0, C.byref(h)) == -1
    b.n_years, b.max_layers = 3, 99
    assert lib.help_b200_engine_create(C.byref(b), 1, 0, C.byref(h)) == -1
    assert lib.help_b200_run(None, None) == -1
    assert lib.help_b200_engine_simulate(None, None) == -1


def test_algorithmic_bytes_formula(lib):
    """12*W*D + (56 + 36 L) per cell + 336 per cell-year."""
    n, ny, L = 10, 2, 3
    cell_i32 = np.zeros((N.HB_NCELL_I32, n), np.int32)
    cell_i32[N.HB_NLAY] = L
    cell_i32[N.HB_NODE_P] = cell_i32[N.HB_NODE_T] = cell_i32[N.HB_NODE_R] = np.arange(n) % 2
    years = np.array([2003, 2004], np.int32)
    b = N.Batch()
    b.n_cells, b.n_years, b.max_layers = n, ny, L
    b.n_nodes_p = b.n_nodes_t = b.n_nodes_r = 5
    b.years, b.cell_i32 = years.ctypes.data, cell_i32.ctypes.data
    got = lib.help_b200_algorithmic_bytes(C.byref(b))
    days = 365 + 366
    assert got == 12 * 2 * days + (56 + 36 * L) * n + 336 * n * ny


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry must fail loudly (ENODEV), never compute."""
    if lib.help_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from pyhelp_b200 import processing, grid as G
    from support import synth
    g = synth.grid_natural_3layer(4, seed=1)
    wx = synth.synth_weather(1, 2001, 1, seed=1)
    node = np.zeros(4, np.int32)
    batch = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)
    with pytest.raises(N.EngineError, match="no CUDA device"):
        processing.run_batch(batch)
    r = N.Result()
    assert lib.help_b200_run(C.byref(batch.as_struct()), C.byref(r)) == -2
    x = np.ones(4)
    assert lib.help_b200_math_eval(0, x.ctypes.data, x.ctypes.data, x.ctypes.data, 4) == -2
    with pytest.raises(N.EngineError):
        processing.run_help_allcells({"a": None})


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (or any CPU path)."""
    pkg = osp.join(ROOT, "pyhelp_b200")
    import os
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(osp.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "libhelp3o_oracle" not in txt, f
