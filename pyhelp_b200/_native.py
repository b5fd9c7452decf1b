"""ctypes binding of ``libhelp_b200.so`` (C ABI: ``include/help_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (or
``python -m pyhelp_b200.build``) with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
There is deliberately no fallback: if the library is missing, or no CUDA
device is visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import os.path as osp

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
LIB_PATH = os.environ.get("HELP_B200_LIB") or osp.join(_HERE, "libhelp_b200.so")   # (override: developer A/B builds)

# enums of include/help_b200.h
(HB_ULAT, HB_ULAI, HB_EDEPTH, HB_WIND, HB_HUM1, HB_HUM2, HB_HUM3, HB_HUM4, HB_OSNO, HB_FRUNOF,
 HB_CN2, HB_TFSOIL, HB_NCELL_F32) = range(13)
(HB_IU11, HB_IPL, HB_IHV, HB_IU10, HB_IPRE, HB_NLAY, HB_NODE_P, HB_NODE_T, HB_NODE_R,
 HB_NCELL_I32) = range(10)
(HL_THICK, HL_PORO, HL_FC, HL_WP, HL_SW, HL_XLENG, HL_SLOPE, HL_RECIR, HL_SUBIN, HL_PHOLE,
 HL_DEFEC, HL_TRANS, HL_NLAYER_F32) = range(13)
HL_TYPE, HL_ISOIL, HL_LAYR, HL_IPQ, HL_NLAYER_I32 = range(5)
HR_PRECIP, HR_RUNOFF, HR_EVAPO, HR_SUBRUN1, HR_SUBRUN2, HR_PERCO, HR_RECHG, HR_NROWS = range(8)
HY_PRECIP, HY_RUNOFF, HY_EVAPO, HY_LATDRAIN, HY_RECHG, HY_NROWS = range(6)
RAW_ROWS = 14
DAILY_COLS = 16
DAYS = 368
FLAG_NONCONV, FLAG_NAN, FLAG_OVERFLOW, FLAG_RECIRC, FLAG_BADCELL = 0x1, 0x2, 0x4, 0x8, 0x10

ERRORS = {0: "ok", -1: "invalid argument", -2: "no CUDA device", -3: "CUDA error", -4: "out of memory"}


class Batch(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32), ("n_years", C.c_int32), ("max_layers", C.c_int32),
        ("n_nodes_p", C.c_int32), ("n_nodes_t", C.c_int32), ("n_nodes_r", C.c_int32),
        ("years", C.c_void_p), ("pre", C.c_void_p), ("tmp", C.c_void_p), ("rad", C.c_void_p),
        ("iu4", C.c_void_p), ("iu7", C.c_void_p), ("iu13", C.c_void_p),
        ("tm_normals", C.c_void_p), ("idfs", C.c_void_p),
        ("cell_f32", C.c_void_p), ("cell_i32", C.c_void_p),
        ("layer_f32", C.c_void_p), ("layer_i32", C.c_void_p), ("layer_rc", C.c_void_p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("monthly", C.c_void_p), ("yearly", C.c_void_p), ("raw_monthly", C.c_void_p),
        ("flags", C.c_void_p), ("topo", C.c_void_p),
        ("ms_h2d", C.c_float), ("ms_setup", C.c_float), ("ms_simulate", C.c_float),
        ("ms_d2h", C.c_float), ("n_launches", C.c_int32),
    ]


class EngineError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA engine; raise loudly when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not osp.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). pyhelp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.help_b200_abi_version.restype = C.c_int
    L.help_b200_device_count.restype = C.c_int
    L.help_b200_last_error.restype = C.c_char_p
    L.help_b200_algorithmic_bytes.restype = C.c_double
    L.help_b200_algorithmic_bytes.argtypes = [C.POINTER(Batch)]
    L.help_b200_run.restype = C.c_int
    L.help_b200_run.argtypes = [C.POINTER(Batch), C.POINTER(Result)]
    L.help_b200_engine_create.restype = C.c_int
    L.help_b200_engine_create.argtypes = [C.POINTER(Batch), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.help_b200_engine_simulate.restype = C.c_int
    L.help_b200_engine_simulate.argtypes = [C.c_void_p, C.c_void_p]
    L.help_b200_engine_device_ptrs.restype = C.c_int
    L.help_b200_engine_device_ptrs.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
    L.help_b200_engine_fetch.restype = C.c_int
    L.help_b200_engine_fetch.argtypes = [C.c_void_p, C.POINTER(Result)]
    L.help_b200_engine_timings.restype = C.c_int
    L.help_b200_engine_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.POINTER(C.c_int32)]
    L.help_b200_engine_destroy.restype = None
    L.help_b200_engine_destroy.argtypes = [C.c_void_p]
    L.help_b200_release_memory.restype = C.c_int
    L.help_b200_set_device.restype = C.c_int
    L.help_b200_set_device.argtypes = [C.c_int]
    L.help_b200_math_eval.restype = C.c_int
    L.help_b200_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.help_b200_engine_select_daily.restype = C.c_int
    L.help_b200_engine_select_daily.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.help_b200_engine_fetch_daily.restype = C.c_int
    L.help_b200_engine_fetch_daily.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    L.help_b200_engine_post_process.restype = C.c_int
    L.help_b200_engine_post_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.help_b200_nearest_nodes.restype = C.c_int
    L.help_b200_nearest_nodes.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.help_b200_pack_weather.restype = C.c_int
    L.help_b200_pack_weather.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.help_b200_pack_weather_resident.restype = C.c_int
    L.help_b200_pack_weather_resident.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                                  C.POINTER(C.c_void_p)]
    L.help_b200_idfs_prescan_resident.restype = C.c_int
    L.help_b200_idfs_prescan_resident.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.help_b200_device_free.restype = None
    L.help_b200_device_free.argtypes = [C.c_void_p]
    L.help_b200_parse_fields.restype = C.c_int
    L.help_b200_parse_fields.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.help_b200_engine_fetch_monthly_by_cell.restype = C.c_int
    L.help_b200_engine_fetch_monthly_by_cell.argtypes = [C.c_void_p, C.c_void_p]
    if L.help_b200_abi_version() != 1:
        raise EngineError("libhelp_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().help_b200_last_error()
        raise EngineError(f"{what}: {ERRORS.get(rc, rc)}: {msg.decode(errors='replace') if msg else ''}")


def require_device():
    if lib().help_b200_device_count() < 1:
        raise EngineError("no CUDA device visible: the HELP engine runs on the GPU only (no CPU fallback)")


def set_device(index: int):
    check(lib().help_b200_set_device(int(index)), "set_device")


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


class DeviceTable:
    """A float32 table that lives in device memory (``help_b200_pack_weather_resident``): what ``CellBatch.pre/tmp/rad`` may
    hold instead of a numpy array.  It has a shape and a size but no host values; the buffer is released with the object."""

    def __init__(self, dev_ptr: int, shape):
        self.ptr = int(dev_ptr)
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(np.float32)

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape)) * 4

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", 0), 0
        if p and _lib is not None:
            _lib.help_b200_device_free(p)

