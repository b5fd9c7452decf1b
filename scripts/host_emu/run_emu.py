"""DEVELOPER DEBUGGING AID: run the kernels through the g++ shim (see cuda_shim.h) and
compare with the oracle.  Not a product path."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyhelp_b200 import _native as N

def emu_run(batch, raw=True):
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libemu.so"))
    n, ny = batch.n_cells, batch.n_years
    monthly = np.zeros((7, ny, 12, n), np.float32); yearly = np.zeros((5, ny, n), np.float32)
    rawm = np.zeros((14, ny, 12, n), np.float32); flags = np.zeros(n, np.uint32); topo = np.zeros((16, n), np.int32)
    s = batch.as_struct()
    ndays = int(sum(366 if int(y) % 4 == 0 else 365 for y in batch.years))
    daily = np.zeros((n, ndays, 16), np.float64)
    lib.emu_set_daily.argtypes = [C.c_void_p, C.c_longlong]
    lib.emu_set_daily(daily.ctypes.data, ndays)
    lib.emu_run.argtypes = [C.POINTER(N.Batch)] + [C.c_void_p] * 5
    lib.emu_run(C.byref(s), monthly.ctypes.data, yearly.ctypes.data, rawm.ctypes.data, flags.ctypes.data, topo.ctypes.data)
    return {"monthly": monthly, "yearly": yearly, "raw": rawm, "flags": flags, "topo": topo, "daily": daily}
