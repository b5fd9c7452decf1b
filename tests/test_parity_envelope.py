"""Parity of the engine's arithmetic with the reference's -- stated, measured and held.

The reference's Fortran gets `**`, EXP, ALOG, SIN ... from the libm its compiler links; the reference itself
cannot be built here (no Fortran compiler), so "the reference's arithmetic" is the oracle port built against
the system libm (libm="glibc").  Three builds of the same oracle source:

  det          help_math_glibc.h -- the arithmetic the GPU engine executes (bit-equal to it: tests/test_gpu_parity.py)
  glibc        direct calls of libm.so.6
  glibc_noise  glibc with ONE ulp added to / taken from 0.1 % of the pow/log/powf/expf/logf results: what another
               correct libm (another glibc release, mingw's, Intel's) looks like from HELP's point of view

Two statements are tested on stand-ins of BASELINE configs[0]..[3], 256 cells x 10 yr each:

 1. det == glibc, BIT FOR BIT, every monthly value of every cell -- the `north_star` tolerance (<= 1e-3 relative or
    <= 0.1 mm/yr on yearly totals) is met with zero difference against the reference's arithmetic linked to this
    image's libm.
 2. HELP is chaotic in the last bit of libm: against glibc_noise only ~90-99 % of the cell-years stay inside that
    tolerance (per-cell multi-year means within ~1.5 % of the cell's precipitation, area means within 1e-3).  A
    per-cell-year tolerance can therefore only be promised against one given libm binary; across libm versions the
    reference does not reproduce itself any better than this envelope.
"""
import os
import platform

import numpy as np
import pytest

from oracle import oracle
from support import synth

KEYS = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")


def family(name, n):
    if name == "example":
        return synth.grid_example(n, seed=20260101), None, dict(sf_edepth=0.15, sf_cn=1.15), 11
    if name == "natural3":
        return synth.grid_natural_3layer(n, seed=1), None, {}, 10
    if name == "mixed":
        return synth.grid_mixed(n, seed=2), None, {}, 10
    g, ex = synth.grid_landfill(n, seed=3)
    return g, ex, {}, 10


def run3(name, n, workdir):
    g, ex, kw, ny = family(name, n)
    nn = 8
    wx = synth.synth_weather(nn, 2000, ny, seed=5)
    files = synth.write_weather_files(os.path.join(workdir, name), wx)
    cells = np.nonzero(np.asarray(g["run"]) == 1)[0][:n].tolist() if "run" in g else None
    node = synth.assign_nodes(len(g["cid"]), nn)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex, cells=cells, **kw)
    return cp, {m: oracle.run_help_allcells(cp, libm=m) for m in ("det", "glibc", "glibc_noise")}


def yearly(out, names):
    y = {k: np.stack([out[c][k].sum(axis=1) for c in names]).astype(np.float64) for k in ("precip", "runoff", "evapo", "rechg")}
    y["latdrain"] = np.stack([(out[c]["subrun1"] + out[c]["subrun2"]).sum(axis=1) for c in names]).astype(np.float64)
    return y


libm_is_cloned = platform.machine() == "x86_64" and platform.libc_ver() == ("glibc", "2.39")


@pytest.mark.parametrize("name", ["example", "natural3", "mixed", "landfill"])
def test_envelope(tmp_path, name):
    cp, out = run3(name, 256, str(tmp_path))
    names = list(cp.keys())
    assert len(names) >= 200
    # ---- 1. the engine's arithmetic IS the reference's arithmetic with this image's libm -----------------------------
    if libm_is_cloned:
        bad = [(c, k) for c in names for k in KEYS if not np.array_equal(out["det"][c][k], out["glibc"][c][k], equal_nan=True)]
        assert not bad, (len(bad), bad[:5])
    # ---- 2. the envelope across libm versions -------------------------------------------------------------------------
    a, b = yearly(out["glibc_noise"], names), yearly(out["glibc"], names)
    ok = ~(np.isnan(a["rechg"]).any(axis=1) | np.isnan(b["rechg"]).any(axis=1))      # cells HELP returns NaN for (inconsistent soils)
    assert ok.mean() > 0.9
    assert np.array_equal(a["precip"], b["precip"], equal_nan=True)                  # inputs only
    report = {}
    for k in ("runoff", "evapo", "latdrain", "rechg"):
        x, y = a[k][ok], b[k][ok]
        if y.mean() < 1.0:
            continue
        within = (np.abs(x - y) <= np.maximum(1e-3 * np.abs(y), 0.1)).mean()
        cell_mean = (np.abs(x.mean(axis=1) - y.mean(axis=1)) / b["precip"][ok].mean(axis=1)).max()
        area = abs(x.mean() - y.mean()) / y.mean()
        report[k] = (round(float(within), 3), round(float(cell_mean), 4), float(area))
        assert within > 0.85, (name, k, within)               # most cell-years agree to the north_star tolerance ...
        assert cell_mean < 0.03, (name, k, cell_mean)         # ... a cell's multi-year mean within 3 % of its precipitation ...
        assert area < 1e-3 or abs(x.mean() - y.mean()) < 0.1, (name, k, area)      # ... and area means within 1e-3 (or 0.1 mm/yr)
    assert any(v[0] < 1.0 for v in report.values()), report   # ... but one ulp in 0.1 % of the libm calls IS visible per cell-year
    print("envelope", name, report)
