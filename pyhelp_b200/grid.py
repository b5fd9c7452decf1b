"""Grid-level fast path: PyHELP's input grid + weather tables -> engine batch,
without going through one text record per cell.

This is the vectorised equivalent of what ``HelpManager.calc_help_cells`` does
per cell before the hot path (reference ``pyhelp/managers.py:234-340,393-428``):

* nearest weather node per cell (``_generate_d4d7d13_input_files``,
  managers.py:300-334, haversine ``pyhelp/utils.py:81-96``),
* D10/D11 formatting (``pyhelp/preprocessing.py:27-185``) and D4/D7/D13
  formatting (``pyhelp/weather_reader.py:22-104``).

The reference's formatters QUANTISE the inputs: precipitation to
0.1 mm, temperature to 0.1 C, radiation to 0.01 MJ/m2, CN and thickness to
integers, porosities to 3 decimals ...  ``quantize`` reproduces Python's
``format(x, '.Nf')`` (round-half-even on the exact binary value) and the
Fortran read that follows, so a batch built here is bit-identical to the batch
parsed back from the reference's text (tests/test_inputs.py proves it).
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .inputs import CellBatch, F32

MINEDEPTH, MAXEDEPTH, MINTHICK = 3, 80, 10        # preprocessing.py:20-22


def quantize(x, d: int) -> np.ndarray:
    """float(format(x, f'.{d}f')) element-wise, vectorised."""
    x = np.asarray(x, dtype=np.float64)
    s = 10.0 ** d
    p = x * s
    r = np.rint(p)
    frac = np.abs(p - np.trunc(p))
    near_tie = np.abs(frac - 0.5) < 1e-6
    out = r / s
    if near_tie.any():
        flat, idx = out.reshape(-1), np.nonzero(near_tie.reshape(-1))[0]
        xs = x.reshape(-1)
        for i in idx:
            flat[i] = float(format(xs[i], f".{d}f"))
        out = flat.reshape(x.shape)
    return out + 0.0          # -0.0 -> 0.0 is irrelevant to the Fortran read


def _year_blocks(years_of_day: np.ndarray):
    """(years, first row, number of rows) of every calendar year of a daily table; the row counts must
    be the ones weather_reader.py:87 pads to (calendar.isleap)."""
    years_of_day = np.asarray(years_of_day)
    years, day0, ndays = np.unique(years_of_day, return_index=True, return_counts=True)
    if not (np.diff(years_of_day) >= 0).all():
        raise ValueError("the daily table must be in chronological order")
    for y, k in zip(years.tolist(), ndays.tolist()):
        nd = 366 if (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)) else 365
        if k != nd:
            raise ValueError(f"year {y} has {k} days of data, expected {nd}")
    return years.astype(np.int32), day0.astype(np.int32), ndays.astype(np.int32)


def weather_to_nodes(years_of_day: np.ndarray, values: np.ndarray, decimals: int, where: str = "host"):
    """[day][node] table -> [node][year][368] float32 as the engine reads a D4/D7/D13
    file written by weather_reader.py (values quantised to `decimals`).  ``where="device"`` runs the
    format/read round trip and the transposition on the GPU (``help_b200_pack_weather``,
    ``csrc/help_weather.cuh``; same bits, tests/test_gpu_pack.py) -- the host version costs 0.4 s
    per variable per 1000 nodes x 30 yr and is what a caller without a device at packing time uses;
    ``where="resident"`` does the same and leaves the result in device memory (a ``DeviceTable``: the
    engine then copies it device to device instead of taking it through the host twice)."""
    years, day0, ndays = _year_blocks(years_of_day)
    values = np.asarray(values, dtype=np.float64)
    nn = values.shape[1]
    if where == "resident":            # quantised and laid out on the GPU, and left there (N.DeviceTable)
        import ctypes as C
        tab = np.ascontiguousarray(values)
        N.require_device()
        dev = C.c_void_p()
        N.check(N.lib().help_b200_pack_weather_resident(N.ptr(tab), tab.shape[0], nn, N.ptr(day0), N.ptr(ndays), years.size,
                                                        int(decimals), C.byref(dev)), "pack_weather_resident")
        return years, N.DeviceTable(dev.value, (nn, years.size, N.DAYS))
    out = np.zeros((nn, years.size, N.DAYS), dtype=F32)
    if where == "device":
        tab = np.ascontiguousarray(values)
        N.require_device()
        N.check(N.lib().help_b200_pack_weather(N.ptr(tab), tab.shape[0], nn, N.ptr(day0), N.ptr(ndays), years.size,
                                               int(decimals), N.ptr(out)), "pack_weather")
        return years, out
    if where != "host":
        raise ValueError("where must be 'host', 'device' or 'resident'")
    q = quantize(values, decimals).astype(F32)
    for k in range(years.size):
        out[:, k, :ndays[k]] = q[day0[k]:day0[k] + ndays[k]].T
    return years, out


def idfs_prescan_nodes(rad: np.ndarray, iu13: int = 2) -> np.ndarray:
    """:func:`pyhelp_b200.inputs.idfs_prescan` for every node at once: int32 [n_nodes][2] from the
    [node][year][368] radiation (float32 sequential accumulation over years x 31 days, F:1352-1373)."""
    rad = np.asarray(rad, dtype=F32)[:, :100]
    nn, nyrs = rad.shape[0], rad.shape[1]
    out = np.empty((nn, 2), dtype=np.int32)
    for h, (lo, hi) in enumerate(((334, 365), (151, 182))):
        seq = np.ascontiguousarray(rad[:, :, lo:hi]).reshape(nn, -1)
        radtot = np.cumsum(seq, axis=1, dtype=F32)[:, -1]
        radavg = (radtot / F32(nyrs)) / F32(31) if nyrs > 1 else radtot / F32(31)
        if iu13 != 1:
            radavg = radavg * F32(23.89)
        radavg = radavg.astype(F32)
        out[:, h] = np.maximum((F32(35.4) - (F32(0.154) * radavg).astype(F32)).astype(F32).astype(np.int32), 1)
    return out


def batch_from_grid(grid: dict, years_of_day, precip, airtemp, solrad, node_p, node_t, node_r,
                    solrad_lat=None, tfsoil_c: float = 0.0, sf_edepth: float = 1.0, sf_ulai: float = 1.0,
                    sf_cn: float = 1.0, extra_layers: dict | None = None, weather: str = "host") -> CellBatch:
    """Vectorised `format_d10d11_inputs` + weather formatting + Fortran read.

    ``grid``: mapping column -> array over cells with the columns of docs/data.rst
    (cid, lat_dd, wind, hum1-4, growth_start, growth_end, LAI, EZD, CN, nlayer,
    lay_type{i}, thick{i}, poro{i}, fc{i}, wp{i}, ksat{i}, dist_dr{i}, slope{i}).
    ``precip``/``airtemp``/``solrad``: [day][node] tables in mm, deg C, MJ/m2/day.
    ``extra_layers``: optional per-layer design fields PyHELP's formatter leaves blank
    (phole{i}, defec{i}, ipq{i}, trans{i}, recir{i}, layr{i}, subin{i}, sw{i}; per cell: ipre, osno) for landfill designs.
    ``weather``: ``"host"``, ``"device"`` or ``"resident"`` -- where the weather tables are quantised and laid
    out, and where they stay (:func:`weather_to_nodes`).
    Cells PyHELP would skip (nlayer == 0 or a -9999 layer field,
    preprocessing.py:32-35,150-153) must be removed by the caller (`runnable_cells`).
    """
    g = {k: np.asarray(v) for k, v in grid.items()}
    n = int(np.asarray(g["nlayer"]).shape[0])
    nlay = g["nlayer"].astype(np.int32)
    L = int(nlay.max()) if n else 1
    cell_f32 = np.zeros((N.HB_NCELL_F32, n), dtype=F32)
    cell_i32 = np.zeros((N.HB_NCELL_I32, n), dtype=np.int32)
    # D11 (preprocessing.py:63-74)
    cell_f32[N.HB_ULAT] = quantize(g["lat_dd"], 2)
    cell_f32[N.HB_ULAI] = quantize(g["LAI"].astype(float) * sf_ulai, 2)
    ezd = np.minimum(np.maximum(g["EZD"].astype(float) * sf_edepth, MINEDEPTH), MAXEDEPTH)
    cell_f32[N.HB_EDEPTH] = quantize(ezd, 1)
    cell_f32[N.HB_WIND] = quantize(g["wind"], 1)
    for q in range(4):
        cell_f32[N.HB_HUM1 + q] = quantize(g[f"hum{q + 1}"], 1)
    cell_i32[N.HB_IU11] = 2
    cell_i32[N.HB_IPL] = g["growth_start"].astype(np.int32)
    cell_i32[N.HB_IHV] = g["growth_end"].astype(np.int32)
    # D10 header (preprocessing.py:93-128): iu10=2, ipre=0, osno=0, frunof=100, irun=1
    cell_i32[N.HB_IU10] = 2
    cell_i32[N.HB_IPRE] = 0
    cell_f32[N.HB_OSNO] = 0.0
    cell_f32[N.HB_FRUNOF] = 100.0
    cell_f32[N.HB_CN2] = quantize(g["CN"].astype(float) * sf_cn, 0)
    cell_f32[N.HB_TFSOIL] = F32((tfsoil_c * 1.8) + 32)          # managers.py:387
    if extra_layers and "ipre" in extra_layers:            # user-specified initial moisture and snow (D10 record 2)
        cell_i32[N.HB_IPRE] = np.asarray(extra_layers["ipre"], dtype=np.int32)
    if extra_layers and "osno" in extra_layers:
        cell_f32[N.HB_OSNO] = np.asarray(extra_layers["osno"], dtype=float)
    cell_i32[N.HB_NLAY] = nlay
    cell_i32[N.HB_NODE_P] = np.asarray(node_p, dtype=np.int32)
    cell_i32[N.HB_NODE_T] = np.asarray(node_t, dtype=np.int32)
    cell_i32[N.HB_NODE_R] = np.asarray(node_r, dtype=np.int32)
    layer_f32 = np.zeros((N.HL_NLAYER_F32, L, n), dtype=F32)
    layer_i32 = np.zeros((N.HL_NLAYER_I32, L, n), dtype=np.int32)
    layer_rc = np.zeros((L, n), dtype=np.float64)
    ex = extra_layers or {}
    for j in range(L):
        s = str(j + 1)
        on = nlay > j
        if f"lay_type{s}" not in g:
            continue
        def col(name, default=0.0):
            a = g.get(name + s, ex.get(name + s))
            return np.where(on, np.asarray(a, dtype=float), 0.0) if a is not None else np.full(n, default) * on
        layer_i32[N.HL_TYPE, j] = np.where(on, g[f"lay_type{s}"].astype(np.int32), 0)
        layer_f32[N.HL_THICK, j] = np.where(on, quantize(np.maximum(col("thick"), MINTHICK), 0), 0)
        layer_f32[N.HL_PORO, j] = quantize(col("poro"), 3)
        layer_f32[N.HL_FC, j] = quantize(col("fc"), 3)
        layer_f32[N.HL_WP, j] = quantize(col("wp"), 3)
        layer_rc[j] = quantize(col("ksat"), 14)
        layer_f32[N.HL_XLENG, j] = quantize(col("dist_dr"), 0)
        layer_f32[N.HL_SLOPE, j] = quantize(col("slope"), 2)
        if ex:
            layer_f32[N.HL_PHOLE, j] = col("phole")
            layer_f32[N.HL_DEFEC, j] = col("defec")
            layer_f32[N.HL_TRANS, j] = col("trans")
            layer_f32[N.HL_RECIR, j] = col("recir")
            layer_f32[N.HL_SUBIN, j] = col("subin")
            layer_f32[N.HL_SW, j] = col("sw")
            layer_i32[N.HL_IPQ, j] = col("ipq").astype(np.int32)
            layer_i32[N.HL_LAYR, j] = col("layr").astype(np.int32)
            if f"thick_exact{s}" in ex:          # landfill liners are thinner than MINTHICK
                te = np.asarray(ex[f"thick_exact{s}"], dtype=float)
                layer_f32[N.HL_THICK, j] = np.where(on & (te > 0), te, layer_f32[N.HL_THICK, j])
    years, pre = weather_to_nodes(years_of_day, precip, 1, weather)
    _, tmp = weather_to_nodes(years_of_day, airtemp, 1, weather)
    _, rad = weather_to_nodes(years_of_day, solrad, 2, weather)
    npn, ntn, nrn = pre.shape[0], tmp.shape[0], rad.shape[0]
    if isinstance(rad, N.DeviceTable):
        idfs = np.zeros((nrn, 2), dtype=np.int32)
        N.check(N.lib().help_b200_idfs_prescan_resident(rad.ptr, nrn, int(years.size), 2, N.ptr(idfs)), "idfs_prescan_resident")
    else:
        idfs = idfs_prescan_nodes(rad, 2)
    names = [str(c) for c in g["cid"].tolist()] if "cid" in g else [str(i) for i in range(n)]
    return CellBatch(
        n, int(years.size), L, np.ascontiguousarray(years), pre, tmp, rad,
        np.full(npn, 2, dtype=np.int32), np.full(ntn, 2, dtype=np.int32), np.full(nrn, 2, dtype=np.int32),
        np.zeros((ntn, 12), dtype=F32), np.ascontiguousarray(idfs),
        cell_f32, cell_i32, np.ascontiguousarray(layer_f32), np.ascontiguousarray(layer_i32),
        np.ascontiguousarray(layer_rc), names)


def runnable_cells(grid: dict) -> np.ndarray:
    """Mask of cells PyHELP would run (preprocessing.py:32-35,150-153; managers.py `run`)."""
    g = {k: np.asarray(v) for k, v in grid.items()}
    nlay = g["nlayer"].astype(int)
    ok = nlay > 0
    if "run" in g:
        ok &= g["run"].astype(int) == 1
    for j in range(int(nlay.max()) if nlay.size else 0):
        s = str(j + 1)
        on = nlay > j
        # (thick is max()-ed with MINTHICK before the reference's -9999 test, so it never trips it)
        for name in ("poro", "fc", "wp", "ksat", "dist_dr", "slope"):
            if name + s in g:
                ok &= ~(on & (g[name + s].astype(float) == -9999))
    return ok
