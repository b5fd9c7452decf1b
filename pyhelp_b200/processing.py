"""Drop-in for the reference's ``pyhelp/processing.py`` hot path, on the B200 engine.

Reference functions mirrored here (same names, arguments, return values):

* :func:`run_help_allcells` -- ``pyhelp/processing.py:40-59``.  Takes the
  ``cellparams`` dict ``HelpManager.calc_help_cells`` builds
  (``pyhelp/managers.py:393-428``) and returns ``{cellname: {'years', 'precip',
  'runoff', 'evapo', 'subrun1', 'subrun2', 'perco', 'rechg'}}`` with the dtypes
  and shapes of ``read_monthly_help_output`` (``pyhelp/processing.py:139-146``).
  The reference runs one Fortran call + one text file per cell in a
  ``multiprocessing.Pool``; here the whole dict becomes one GPU batch.
  Differences, both deliberate: results come back in ``cellparams`` order (the
  reference's ``imap_unordered`` order is nondeterministic) and
  ``ncore`` is accepted and ignored.
* :func:`run_help_singlecell` -- ``pyhelp/processing.py:30-37``.

:class:`HelpEngine` is the batch-level API underneath (upload once, simulate,
fetch), used by ``bench.py`` and by callers that want the arrays rather than
per-cell dicts.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _native as N
from .inputs import CellBatch, batch_from_cellparams

KEYS = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")
DAILY_COLS = ("pre", "run", "ett", "swe", "hed1", "drn1", "prc1", "hed2", "drn2", "prc2", "hed3", "drn3", "prc3",
              "air_freeze", "soil_freeze", "sno")


class HelpEngine:
    """Resident GPU engine for one :class:`CellBatch` (``help_b200_engine_*``)."""

    def __init__(self, batch: CellBatch, monthly: bool = True, raw: bool = False):
        N.require_device()
        self.batch = batch
        self._struct = batch.as_struct()
        self._h = C.c_void_p()
        self.want_monthly, self.want_raw = monthly, raw
        N.check(N.lib().help_b200_engine_create(C.byref(self._struct), int(monthly), int(raw),
                                                C.byref(self._h)), "engine_create")

    def simulate(self, stream: int | None = None):
        N.check(N.lib().help_b200_engine_simulate(self._h, C.c_void_p(stream) if stream else None),
                "engine_simulate")
        ms_setup, ms_sim, nl = C.c_float(), C.c_float(), C.c_int32()
        N.lib().help_b200_engine_timings(self._h, C.byref(ms_setup), C.byref(ms_sim), C.byref(nl))
        self.ms_setup, self.ms_simulate, self.n_launches = ms_setup.value, ms_sim.value, nl.value
        return self

    def device_ptrs(self) -> dict:
        p = [C.c_void_p() for _ in range(4)]
        N.check(N.lib().help_b200_engine_device_ptrs(self._h, *[C.byref(x) for x in p]), "device_ptrs")
        return dict(zip(("monthly", "yearly", "raw", "flags"), [x.value for x in p]))

    def fetch(self, monthly: bool | None = None, yearly: bool = True, raw: bool | None = None,
              topo: bool = False) -> dict:
        b = self.batch
        n, ny = b.n_cells, b.n_years
        monthly = self.want_monthly if monthly is None else monthly
        raw = self.want_raw if raw is None else raw
        out = {"flags": np.zeros(n, dtype=np.uint32)}
        if monthly:
            out["monthly"] = np.empty((N.HR_NROWS, ny, 12, n), dtype=np.float32)
        if yearly:
            out["yearly"] = np.empty((N.HY_NROWS, ny, n), dtype=np.float32)
        if raw:
            out["raw"] = np.empty((N.RAW_ROWS, ny, 12, n), dtype=np.float32)
        if topo:
            out["topo"] = np.zeros((16, n), dtype=np.int32)
        r = N.Result()
        r.monthly, r.yearly, r.raw_monthly = N.ptr(out.get("monthly")), N.ptr(out.get("yearly")), N.ptr(out.get("raw"))
        r.flags, r.topo = N.ptr(out["flags"]), N.ptr(out.get("topo"))
        N.check(N.lib().help_b200_engine_fetch(self._h, C.byref(r)), "engine_fetch")
        out["timing_ms"] = {"h2d": r.ms_h2d, "setup": r.ms_setup, "simulate": r.ms_simulate, "d2h": r.ms_d2h}
        out["n_launches"] = r.n_launches
        out["years"] = b.years.astype(np.uint16)
        return out

    def fetch_by_cell(self) -> np.ndarray:
        """The monthly result as the reference hands it out, one block per cell: float32 [n_cells][7][n_years][12]
        (rows ``KEYS``), transposed on the device."""
        b = self.batch
        out = np.empty((b.n_cells, N.HR_NROWS, b.n_years, 12), dtype=np.float32)
        N.check(N.lib().help_b200_engine_fetch_monthly_by_cell(self._h, N.ptr(out)), "fetch_monthly_by_cell")
        return out

    def select_daily(self, cells):
        """Ask the next :meth:`simulate` to record the reference's daily output columns (OUTDAY,
        ``HELP3O.FOR:5984-6029``; ``read_daily_help_output``, ``pyhelp/processing.py:150-220``)
        for these cell indices (a verification channel for a handful of cells)."""
        idx = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1)
        N.check(N.lib().help_b200_engine_select_daily(self._h, N.ptr(idx), idx.size), "select_daily")
        self._n_daily = idx.size
        return self

    def fetch_daily(self) -> np.ndarray:
        """float64 [n_selected, n_days, 16] in inches, unrounded; columns ``DAILY_COLS``."""
        nd = C.c_int64()
        N.check(N.lib().help_b200_engine_fetch_daily(self._h, None, C.byref(nd)), "fetch_daily")
        out = np.empty((self._n_daily, nd.value, N.DAILY_COLS), dtype=np.float64)
        N.check(N.lib().help_b200_engine_fetch_daily(self._h, N.ptr(out), C.byref(nd)), "fetch_daily")
        return out

    def post_process(self, context=None, year_from=None, year_to=None) -> dict:
        """PyHELP's result reductions on the device-resident monthly buffer (no 7 x (Np, Ny, 12)
        float64 arrays on the host): ``HelpManager._post_process_output``
        (reference ``pyhelp/managers.py:448-506``: NaN cells, context-2 recharge turned into
        subsurface runoff), ``HelpOutput.calc_area_monthly_avg`` (``pyhelp/output.py:127-152``)
        and ``HelpOutput.calc_cells_yearly_avg`` (``pyhelp/output.py:171-203``).
        Returns ``{'years', 'idx_nan', 'area_monthly_avg': {var: [Ny, 12]},
        'cells_yearly_avg': {var: [Np]}}`` keyed like ``pyhelp.output.VARNAMES``."""
        b = self.batch
        n, ny = b.n_cells, b.n_years
        ctx = None if context is None else np.ascontiguousarray(context, dtype=np.int32).reshape(-1)
        if ctx is not None and ctx.size != n:
            raise ValueError("context must have one entry per cell")
        area = np.empty((N.HR_NROWS, ny, 12), dtype=np.float64)
        cells = np.empty((N.HR_NROWS, n), dtype=np.float64)
        mask = np.zeros(n, dtype=np.uint8)
        yf = -(2 ** 31) if year_from is None or year_from == -np.inf else int(year_from)
        yt = 2 ** 31 - 1 if year_to is None or year_to == np.inf else int(year_to)
        N.check(N.lib().help_b200_engine_post_process(self._h, N.ptr(ctx), yf, yt, N.ptr(area), N.ptr(cells), N.ptr(mask)),
                "engine_post_process")
        return {"years": b.years.astype(np.uint16), "idx_nan": np.nonzero(mask)[0].tolist(),
                "area_monthly_avg": {k: area[i] for i, k in enumerate(KEYS)},
                "cells_yearly_avg": {k: cells[i] for i, k in enumerate(KEYS)}}

    def close(self):
        if self._h:
            N.lib().help_b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_batch(batch: CellBatch, monthly: bool = True, raw: bool = False, topo: bool = False) -> dict:
    """One-shot: upload, simulate, fetch (host buffers in and out)."""
    with HelpEngine(batch, monthly=monthly, raw=raw) as eng:
        eng.simulate()
        return eng.fetch(topo=topo)


def run_batch_devices(batch: CellBatch, n_devices: int, monthly: bool = True) -> dict:
    """The batch on ``n_devices`` GPUs of this process: contiguous shards of equal estimated cost
    (:mod:`pyhelp_b200.sharding`), one host thread per device (``help_b200_set_device`` is per thread,
    the C calls release the GIL), results joined in the caller's cell order.  Cells are independent,
    so nothing is exchanged between devices.  The torchrun/NCCL route of ``bench.py`` is the measured
    multi-GPU path; this one serves a caller that stays in one process (opt-in: ``HELP_B200_DEVICES``)."""
    import threading
    from . import sharding
    n_devices = max(1, min(int(n_devices), batch.n_cells))
    if n_devices == 1:
        return run_batch(batch, monthly=monthly)
    ranges = sharding.cost_balanced_ranges(sharding.estimate_cost(batch), n_devices)
    parts: list = [None] * n_devices
    errors: list = []

    def work(dev: int, lo: int, hi: int):
        try:
            N.set_device(dev)
            sub = sharding.compact_nodes(batch.take(np.arange(lo, hi)))
            parts[dev] = run_batch(sub, monthly=monthly)
        except BaseException as e:                     # re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d, lo, hi)) for d, (lo, hi) in enumerate(ranges) if hi > lo]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    done = [p for p in parts if p is not None]
    out = {"years": done[0]["years"], "flags": np.concatenate([p["flags"] for p in done])}
    for key in ("monthly", "yearly", "raw"):
        if key in done[0]:
            out[key] = np.concatenate([p[key] for p in done], axis=-1)
    out["timing_ms"] = {k: max(p["timing_ms"][k] for p in done) for k in done[0]["timing_ms"]}
    out["n_launches"] = sum(p["n_launches"] for p in done)
    return out


def split_by_cell(batch: CellBatch, res: dict) -> dict:
    """[row][year][month][cell] (or the already cell-major ``res["by_cell"]``) -> the reference's dict of per-cell dicts."""
    years = res["years"]
    per_cell = res["by_cell"] if "by_cell" in res else np.ascontiguousarray(np.transpose(res["monthly"], (3, 0, 1, 2)))   # (n, 7, ny, 12)
    # one dict per cell, like the reference returns: the arrays are views of one (n, 7, ny, 12) block and of one (n, ny)
    # block of years (a dict literal per cell: half the time of filling the dicts key by key, 0.2 s per 100 000 cells)
    assert KEYS == ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")
    ys = np.tile(np.asarray(years), (per_cell.shape[0], 1))
    return {name: {"years": y, "precip": r[0], "runoff": r[1], "evapo": r[2], "subrun1": r[3], "subrun2": r[4], "perco": r[5],
                   "rechg": r[6]} for name, r, y in zip(batch.names, per_cell, ys)}


LAST_TIMING: dict = {}      # wall-clock seconds of the stages of the last run_help_allcells call (developer information)


def run_help_allcells(cellparams: dict, ncore=None) -> dict:
    """Run HELP in batch for multiple cells (reference pyhelp/processing.py:40-59)."""
    tstart = time.perf_counter()
    if len(cellparams) == 0:
        return {}
    N.require_device()          # fail before any host work: this path has no CPU fallback
    batch = batch_from_cellparams(cellparams, parse="device")      # formatted reads of all records in a few kernels
    t_pack = time.perf_counter()
    want = os.environ.get("HELP_B200_DEVICES", "1")      # "all" or a count: fan out over several GPUs of this process
    ndev = N.lib().help_b200_device_count() if want.strip().lower() == "all" else int(want)
    if ndev > 1:
        res = run_batch_devices(batch, ndev, monthly=True)
        t_up = t_sim = t_fetch = time.perf_counter()
    else:
        with HelpEngine(batch, monthly=True) as eng:
            t_up = time.perf_counter()
            eng.simulate()
            t_sim = time.perf_counter()
            res = {"years": batch.years.astype(np.uint16), "by_cell": eng.fetch_by_cell()}
            t_fetch = time.perf_counter()
    output = split_by_cell(batch, res)
    t_end = time.perf_counter()
    LAST_TIMING.clear()
    LAST_TIMING.update({"records_and_weather_s": t_pack - tstart, "upload_s": t_up - t_pack, "setup_simulate_s": t_sim - t_up,
                        "transpose_download_s": t_fetch - t_sim, "per_cell_dicts_s": t_end - t_fetch, "total_s": t_end - tstart})
    print("\nTask completed in %0.2f sec" % (t_end - tstart))
    return output


def run_help_singlecell(item):
    """Run HELP for a single cell (reference pyhelp/processing.py:30-37)."""
    cellname, outparam = item
    out = run_help_allcells({cellname: outparam})
    return (cellname, out[cellname])


POST_KEYS = ("precip", "runoff", "evapo", "perco", "subrun1", "subrun2", "rechg")


def post_process_monthly(monthly: np.ndarray, years, cellnames, context, lat_dd=None, lon_dd=None) -> dict:
    """``HelpManager._post_process_output`` (reference ``pyhelp/managers.py:448-506``) on the engine's
    ``[row][year][month][cell]`` float32 buffer, vectorised: the ``HelpOutput.data`` layout, (Np, Ny, 12)
    float64 arrays, NaN cells listed in ``idx_nan``, recharge of ``context == 2`` cells added to
    ``subrun1`` (or ``subrun2`` when the cell has any) and zeroed."""
    data = {k: np.ascontiguousarray(np.transpose(monthly[i], (2, 0, 1))).astype(np.float64) for i, k in enumerate(KEYS)}
    data = {k: data[k] for k in POST_KEYS}
    rechg = data["rechg"]
    nan = np.isnan(rechg).any(axis=(1, 2))
    ctx = np.zeros(rechg.shape[0], dtype=np.int64) if context is None else np.asarray(context)
    stream = (ctx == 2) & ~nan
    deep = stream & ~(data["subrun2"] == 0).all(axis=(1, 2))
    shallow = stream & ~deep
    data["subrun1"][shallow] += rechg[shallow]
    data["subrun2"][deep] += rechg[deep]
    rechg[stream] = 0.0
    data["cid"] = list(cellnames)
    data["years"] = np.asarray(years)
    data["idx_nan"] = np.nonzero(nan)[0].tolist()
    if lat_dd is not None:
        data["lat_dd"] = np.asarray(lat_dd)
    if lon_dd is not None:
        data["lon_dd"] = np.asarray(lon_dd)
    return data


def post_process_output(output: dict, context=None, lat_dd=None, lon_dd=None) -> dict:
    """The same for the dict of per-cell dicts :func:`run_help_allcells` returns (the argument
    ``_post_process_output`` takes in the reference)."""
    cellnames = list(output.keys())
    monthly = np.stack([np.stack([output[c][k] for c in cellnames], axis=-1) for k in KEYS])
    return post_process_monthly(monthly, output[cellnames[0]]["years"], cellnames, context, lat_dd, lon_dd)
