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
# Two builds of the same source (oracle/Makefile): "det" uses the deterministic
# transcendentals of pyhelp_b200/csrc/help_math.h (the build the GPU engine must match
# bit for bit), "glibc" the host libm a gfortran build of the reference would link.
_LIB_PATHS = {"det": osp.join(_HERE, "libhelp3o_oracle.so"),
              "glibc": osp.join(_HERE, "libhelp3o_oracle_glibc.so"),
              "glibc_noise": osp.join(_HERE, "libhelp3o_oracle_glibc_noise.so")}
_LIB_PATH = _LIB_PATHS["det"]
_libs = {}

ROWS_MONTHLY = ("pre", "run", "et", "drn1", "drn2", "drn3", "drn4", "drn5",
                "prc1", "prc2", "prc3", "prc4", "prc5", "prc6")
KEYS_PARSED = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")
DAILY_COLS = ("pre", "run", "ett", "swe", "hed1", "drn1", "prc1", "hed2", "drn2",
              "prc2", "hed3", "drn3", "prc3", "air_freeze", "soil_freeze", "sno")


def build(force: bool = False) -> str:
    """Compile both oracle shared libraries (g++ only; no CUDA, no reference code)."""
    deps = [osp.join(_HERE, "help3o_oracle.cpp"),
            osp.join(_HERE, "..", "pyhelp_b200", "csrc", "help_math.h"),
            osp.join(_HERE, "..", "pyhelp_b200", "csrc", "help_math_tables.h")]
    newest = max(osp.getmtime(d) for d in deps)
    if force or any(not osp.exists(p) or osp.getmtime(p) < newest for p in _LIB_PATHS.values()):
        subprocess.check_call(["make", "-C", _HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load(libm: str = "det"):
    if libm not in _libs:
        build()
        lib = ctypes.CDLL(_LIB_PATHS[libm])
        f = lib.help_oracle_run
        f.restype = ctypes.c_int
        c = ctypes
        f.argtypes = [c.c_char_p, c.c_char_p, c.c_char_p,
                      c.c_char_p, c.c_int, c.c_char_p, c.c_int, c.c_int,
                      c.c_int, c.c_float,
                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        m = lib.help_oracle_math_eval
        m.restype = ctypes.c_int
        m.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_longlong]
        _libs[libm] = lib
    return _libs[libm]


MATH_FN = {"pow": 0, "exp": 1, "log": 2, "sin": 3, "cos": 4, "atan": 5, "r4_acos": 6, "r4_tan": 7,
           "r4_pow": 8, "r4_exp": 9, "r4_log": 10, "r4_sin": 11, "r4_cos": 12, "r4_log10": 13,
           "r4_sqrt": 14, "r4_pow8": 15, "r4_pow7": 16, "r4_pow4": 17, "pow2": 18,
           "pow_sb": 19, "fdiv": 20}


def math_eval(fn: str, x, y=None, libm: str = "det") -> np.ndarray:
    """Evaluate one of the oracle build's transcendental functions elementwise."""
    lib = _load(libm)
    x = np.ascontiguousarray(x, dtype=np.float64)
    yy = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    rc = lib.help_oracle_math_eval(MATH_FN[fn], x.ctypes.data, None if yy is None else yy.ctypes.data,
                                   out.ctypes.data, x.size)
    if rc != 0:
        raise ValueError(fn)
    return out


def _records(arr) -> tuple[bytes, int]:
    """S80 array of shape (n,) or (n, 1) (reference managers.py:410-411) -> bytes."""
    a = np.asarray(arr)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "ascii")
    a = np.char.ljust(a.reshape(-1), 80).astype("S80")
    return a.tobytes(), a.shape[0]


def run_simulation(fpath_precip, fpath_tasavg, fpath_solrad, d11_input, d10_input,
                   fpath_output=None, iod=0, iom=1, ioa=0, ios=0, lmyr=1, tfsoil=32.0,
                   *, daily: bool = False, full: bool = False, libm: str = "det"):
    """One cell.  Returns the dict ``read_monthly_help_output`` would return
    (keys years/precip/runoff/evapo/subrun1/subrun2/perco/rechg); with
    ``full=True`` also the raw REAL*4 accumulators and diagnostics."""
    lib = _load(libm)
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


def _run_single_glibc(item):
    name, p = item
    return name, run_simulation(*p, libm="glibc")


def _run_single_glibc_noise(item):
    name, p = item
    return name, run_simulation(*p, libm="glibc_noise")


def run_help_allcells(cellparams: dict, ncore: int | None = None, libm: str = "det") -> dict:
    """Reference-shaped fan-out (pyhelp/processing.py:40-59) over the oracle."""
    ncore = max(mp.cpu_count() if ncore is None else ncore, 1)
    _load(libm)
    fn = {"det": _run_single, "glibc": _run_single_glibc, "glibc_noise": _run_single_glibc_noise}[libm]
    if ncore == 1 or len(cellparams) < 2:
        return dict(fn(it) for it in cellparams.items())
    with mp.get_context("fork").Pool(ncore) as pool:
        chunk = max(1, len(cellparams) // (ncore * 8))
        return dict(pool.imap_unordered(fn, cellparams.items(), chunksize=chunk))
