"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  It wraps
``oracle/libhelp3o_oracle.so`` (built from ``oracle/help3o_oracle.cpp`` by
``oracle/Makefile``), the scalar C++ restatement of the reference's
``pyhelp/HELP3O.FOR:17-5440``.

The two public calls mirror the reference boundary they stand in for:

* :func:`run_simulation` ~ ``HELP3O.run_simulation`` (f2py, reference
  ``pyhelp/processing.py:33``) followed by ``read_monthly_help_output``
  (``pyhelp/processing.py:64-147``): same 12-tuple in, same dict out.
* :func:`run_help_allcells` ~ ``pyhelp/processing.py:40-59``: same
  ``cellparams`` dict in, same dict-of-dicts out, fanned out over a
  ``multiprocessing.Pool`` exactly as the reference does.
"""
from __future__ import annotations

import ctypes
import multiprocessing as mp
import os
import os.path as osp
import subprocess

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
_LIB_PATH = osp.join(_HERE, "libhelp3o_oracle.so")
_lib = None

ROWS_MONTHLY = ("pre", "run", "et", "drn1", "drn2", "drn3", "drn4", "drn5",
                "prc1", "prc2", "prc3", "prc4", "prc5", "prc6")
KEYS_PARSED = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")
DAILY_COLS = ("pre", "run", "ett", "swe", "hed1", "drn1", "prc1", "hed2", "drn2",
              "prc2", "hed3", "drn3", "prc3", "air_freeze", "soil_freeze", "sno")


def build(force: bool = False) -> str:
    """Compile the oracle shared library (gcc only; no CUDA, no reference code)."""
    src = osp.join(_HERE, "help3o_oracle.cpp")
    if force or not osp.exists(_LIB_PATH) or osp.getmtime(_LIB_PATH) < osp.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libhelp3o_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        f = lib.help_oracle_run
        f.restype = ctypes.c_int
        c = ctypes
        f.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p,
                      c.c_char_p, c.c_int, c.c_char_p, c.c_int, c.c_int,
                      c.c_int, c.c_float,
                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        _lib = lib
    return _lib


def _records(arr) -> tuple[bytes, int]:
    """S80 array of shape (n,) or (n, 1) (reference managers.py:410-411) -> bytes."""
    a = np.asarray(arr)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "ascii")
    a = np.char.ljust(a.reshape(-1), 80).astype("S80")
    return a.tobytes(), a.shape[0]


def run_simulation(fpath_precip, fpath_tasavg, fpath_solrad, d11_input, d10_input,
                   fpath_output=None, iod=0, iom=1, ioa=0, ios=0, lmyr=1, tfsoil=32.0,
                   *, daily: bool = False, full: bool = False):
    """One cell.  Returns the dict ``read_monthly_help_output`` would return
    (keys years/precip/runoff/evapo/subrun1/subrun2/perco/rechg); with
    ``full=True`` also the raw REAL*4 accumulators and diagnostics."""
    lib = _load()
    d11b, n11 = _records(d11_input)
    d10b, n10 = _records(d10_input)
    ny = int(lmyr)
    years = np.zeros(ny, dtype=np.int32)
    monthly = np.zeros((ny, 14, 12), dtype=np.float32)
    annual = np.zeros((ny, 14), dtype=np.float32)
    parsed = np.zeros((ny, 7, 12), dtype=np.float32)
    dly = np.zeros((ny * 366, 16), dtype=np.float64) if daily else None
    layer_sw = np.zeros((2, 20), dtype=np.float32)
    info = np.zeros(32, dtype=np.int32)
    bal = np.zeros(ny, dtype=np.float32)
    rc = lib.help_oracle_run(
        os.fsencode(fpath_precip), os.fsencode(fpath_tasavg), os.fsencode(fpath_solrad),
        d11b, n11, d10b, n10, 80, ny, float(tfsoil),
        years.ctypes.data, monthly.ctypes.data, annual.ctypes.data, parsed.ctypes.data,
        dly.ctypes.data if daily else None, layer_sw.ctypes.data, info.ctypes.data,
        bal.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"help_oracle_run failed with code {rc}")
    out = {"years": years.astype("uint16")}
    for i, k in enumerate(KEYS_PARSED):
        out[k] = parsed[:, i, :].copy()
    if full or daily:
        out["monthly_in"] = monthly
        out["annual_in"] = annual
        out["layer_sw"] = layer_sw
        out["balance_in"] = bal
        out["info"] = {
            "lay": int(info[0]), "nseg": int(info[1]), "idfs": int(info[2]),
            "nonconv": int(info[3]), "lp": info[4:10].tolist(), "nstep": info[10:16].tolist(),
            "nsegn": info[16:22].tolist(), "ld": info[22:27].tolist(),
            "lcase": info[27:32].tolist()}
        if daily:
            ndays = sum(366 if y % 4 == 0 else 365 for y in years.tolist())
            out["daily"] = dly[:ndays]
    return out


def _run_single(item):
    name, p = item
    return name, run_simulation(*p)


def run_help_allcells(cellparams: dict, ncore: int | None = None) -> dict:
    """Reference-shaped fan-out (pyhelp/processing.py:40-59) over the oracle."""
    ncore = max(mp.cpu_count() if ncore is None else ncore, 1)
    _load()
    if ncore == 1 or len(cellparams) < 2:
        return dict(_run_single(it) for it in cellparams.items())
    with mp.get_context("fork").Pool(ncore) as pool:
        chunk = max(1, len(cellparams) // (ncore * 8))
        return dict(pool.imap_unordered(_run_single, cellparams.items(), chunksize=chunk))
