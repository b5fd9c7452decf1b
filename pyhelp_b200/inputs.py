"""Host-side packing of the reference's inputs into the engine's SoA batch.

What the reference feeds ``HELP3O.run_simulation`` per cell (reference
``pyhelp/managers.py:393-428``) is text: three weather files (D4/D7/D13) and
the D11/D10 records as ``S80`` arrays.  The Fortran reads them with formatted
READs (``pyhelp/HELP3O.FOR:1100-1129`` weather years, ``:1215-1225`` headers
and D11, ``:1260-1300`` D10).  This module re-does exactly those reads -- fixed
columns, blank = 0, an explicit decimal point overrides ``Fw.d`` -- once per
distinct file and vectorised over cells, and nothing more: every derived
quantity (Brooks-Corey parameters, segments, sub-steps ...) is computed on the
GPU by the set-up kernel.

The one exception is ``IDFS`` (days to thaw): READIN pre-scans the *whole* D13
file for it (``HELP3O.FOR:1348-1375``), including years beyond the simulated
ones, so it is evaluated here per file (both hemispheres) in float32 with the
Fortran's accumulation order.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import _native as N

F32 = np.float32


# ----------------------------------------------------------------------------
# Fortran formatted-read semantics, vectorised
# ----------------------------------------------------------------------------
def _field(rec: np.ndarray, a: int, w: int) -> np.ndarray:
    """Columns [a, a+w) of an (n, reclen) uint8 record matrix as stripped byte strings."""
    sub = np.ascontiguousarray(rec[:, a:a + w])
    return np.char.strip(sub.view(f"S{w}").reshape(-1))


def _to_real(txt: np.ndarray, d: int, dtype) -> np.ndarray:
    out = np.zeros(txt.shape, dtype=np.float64)
    nz = txt != b""
    if nz.any():
        t = txt[nz]
        try:
            vals = t.astype(np.float64)                 # plain decimal fields: the usual case
        except ValueError:                              # embedded blanks (ignored, BN) or a D exponent
            t = np.char.replace(np.char.replace(np.char.replace(t, b" ", b""), b"D", b"E"), b"d", b"e")
            vals = t.astype(np.float64)
        if d > 0:
            nodot = np.char.find(t, b".") < 0          # Fw.d: no point -> the last d digits of the significand are decimals, exponent or not
            if nodot.any():
                vals[nodot] = vals[nodot] / (10.0 ** d)
        out[nz] = vals
    return out.astype(dtype)


def _to_int(txt: np.ndarray) -> np.ndarray:
    out = np.zeros(txt.shape, dtype=np.int32)
    nz = txt != b""
    if nz.any():
        t = txt[nz]
        try:
            out[nz] = t.astype(np.int64).astype(np.int32)
        except ValueError:
            out[nz] = np.char.replace(t, b" ", b"").astype(np.int64).astype(np.int32)
    return out


def _records(lines, reclen=80) -> np.ndarray:
    a = np.asarray(lines)
    if a.dtype.kind == "S" and a.dtype.itemsize == reclen:      # already fixed-length records (the cellparams case)
        rec = np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint8).reshape(-1, reclen).copy()
        rec[rec == 0] = 32                                        # numpy pads short byte strings with NULs
        return rec
    if a.dtype.kind == "U":
        a = np.char.encode(a, "ascii", "replace")
    a = np.char.ljust(a.reshape(-1).astype(f"S{reclen}"), reclen)
    return np.frombuffer(a.astype(f"S{reclen}").tobytes(), dtype=np.uint8).reshape(-1, reclen)


# ----------------------------------------------------------------------------
# Weather files (D4 / D7 / D13)
# ----------------------------------------------------------------------------
@dataclass
class WeatherFile:
    kind: str              # 'D4' | 'D7' | 'D13'
    units: int             # IU4 / IU7 / IU13: 1 = IP, 2 = SI
    years: np.ndarray      # int32 [nyears_in_file] (last record of each block, like READCD)
    data: np.ndarray       # float32 [nyears_in_file][366] raw values as read
    header12: np.ndarray   # float32 [12] header line 4 (monthly normals; D13: [0] = latitude)
    idfs: tuple = (1, 1)   # (ULAT >= 0, ULAT < 0), D13 only


def read_weather_file(path: str, kind: str) -> WeatherFile:
    """READIN header read (F:1215-1218) + READCD year blocks (F:1100-1129)."""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if len(lines) < 4:
        raise ValueError(f"{path}: truncated weather header")
    nblocks = (len(lines) - 4) // 37
    if nblocks < 1:
        raise ValueError(f"{path}: no complete year of data")
    # one byte matrix for the whole file (records cut or blank-padded to 80 columns, CR dropped)
    allrec = _records(np.array(lines[:4 + 37 * nblocks], dtype="S80"))
    allrec[allrec == 13] = 32
    allrec[allrec > 127] = 63
    hdr, rec = allrec[:4], allrec[4:]
    units = int(_to_int(_field(hdr[1:2], 0, 2))[0])
    h12 = np.ascontiguousarray(hdr[3, :72]).reshape(12, 6)
    header12 = _to_real(np.char.strip(h12.view("S6").reshape(-1)), 2, F32)
    if kind == "D4":      # FORMAT(I10,10F5.2)
        ycol, y_w, v0, v_w, d = 0, 10, 10, 5, 2
    else:                 # FORMAT(I5,10F6.1) / (I5,10F6.0)
        ycol, y_w, v0, v_w = 0, 5, 5, 6
        d = 1 if (kind == "D7" or units == 1) else 0
    yrs = _to_int(_field(rec, ycol, y_w)).reshape(nblocks, 37)[:, -1]
    vals = np.ascontiguousarray(rec[:, v0:v0 + 10 * v_w]).reshape(-1, v_w)            # [record x 10 fields][v_w]
    vals = _to_real(np.char.strip(vals.view(f"S{v_w}").reshape(-1)), d, F32)
    data = vals.reshape(nblocks, 370)[:, :366].astype(F32)
    wf = WeatherFile(kind, units, yrs.astype(np.int32), np.ascontiguousarray(data), header12)
    if kind == "D13":
        # READIN's pre-scan reads the same fields with F6.1 whatever the units (formats 5400/5410,
        # F:1364-1365) while READCD reads SI files with F6.0 (F:1128): fields written without a
        # decimal point are ten times smaller in the pre-scan
        pre = data if d == 1 else _to_real(np.char.strip(np.ascontiguousarray(rec[:, v0:v0 + 10 * v_w]).reshape(-1, v_w).view(f"S{v_w}").reshape(-1)), 1, F32).reshape(nblocks, 370)[:, :366]
        wf.idfs = idfs_prescan(pre, units)
    return wf


def idfs_prescan(rad: np.ndarray, iu13: int) -> tuple:
    """IDFS for both hemispheres (F:1348-1375): float32, sequential accumulation.

    ``rad``: [nyears_in_file][366] raw D13 values (at most the first 100 years count).
    """
    rad = np.asarray(rad, dtype=F32)[:100]
    nyrs = rad.shape[0]
    out = []
    for lo, hi in ((334, 365), (151, 182)):         # days 335..365 / 152..182 (1-based)
        seq = rad[:, lo:hi].reshape(-1)
        radtot = np.cumsum(seq, dtype=F32)[-1] if seq.size else F32(0)
        if nyrs > 1:
            radavg = F32(F32(radtot / F32(nyrs)) / F32(31))
        else:
            radavg = F32(radtot / F32(31))
        if iu13 != 1:
            radavg = F32(radavg * F32(23.89))
        idfs = int(F32(F32(35.4) - F32(F32(0.154) * radavg)))
        out.append(max(idfs, 1))
    return tuple(out)


# ----------------------------------------------------------------------------
# D11 / D10 records
# ----------------------------------------------------------------------------
def parse_d11(rec3: np.ndarray, cell_f32: np.ndarray, cell_i32: np.ndarray, idx: np.ndarray):
    """rec3: (m, 3, 80) uint8.  F:1221-1225: (I2) ; (F10.2,I4,I4,F7.0,F8.0,5F5.0)."""
    r1, r3 = rec3[:, 0], rec3[:, 2]
    cell_i32[N.HB_IU11, idx] = _to_int(_field(r1, 0, 2))
    cell_f32[N.HB_ULAT, idx] = _to_real(_field(r3, 0, 10), 2, F32)
    cell_i32[N.HB_IPL, idx] = _to_int(_field(r3, 10, 4))
    cell_i32[N.HB_IHV, idx] = _to_int(_field(r3, 14, 4))
    cell_f32[N.HB_ULAI, idx] = _to_real(_field(r3, 18, 7), 0, F32)
    cell_f32[N.HB_EDEPTH, idx] = _to_real(_field(r3, 25, 8), 0, F32)
    cell_f32[N.HB_WIND, idx] = _to_real(_field(r3, 33, 5), 0, F32)
    for q in range(4):
        cell_f32[N.HB_HUM1 + q, idx] = _to_real(_field(r3, 38 + 5 * q, 5), 0, F32)


def parse_d10(rec: np.ndarray, cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, idx: np.ndarray):
    """rec: (m, nrec, 80) uint8 with a common record count.  F:1260-1300."""
    m, nrec = rec.shape[0], rec.shape[1]
    r2 = rec[:, 1]                       # FORMAT(I2,I2,2F10.0,F6.0,I2)
    cell_i32[N.HB_IU10, idx] = _to_int(_field(r2, 0, 2))
    cell_i32[N.HB_IPRE, idx] = _to_int(_field(r2, 2, 2))
    cell_f32[N.HB_OSNO, idx] = _to_real(_field(r2, 4, 10), 0, F32)
    cell_f32[N.HB_FRUNOF, idx] = _to_real(_field(r2, 24, 6), 0, F32)
    irun = _to_int(_field(r2, 30, 2))
    r3 = rec[:, 2]
    cn = np.zeros(m, dtype=F32)
    nxt = np.full(m, 2, dtype=np.int64)  # record index of the first layer record
    for code, col in ((1, 0), (2, 21), (3, 20)):   # F:1268-1281: where CN2 sits for IRUN = 1 / 2 / 3
        sel = irun == code
        if sel.any():
            cn[sel] = _to_real(_field(r3[sel], col, 7), 0, F32)
            nxt[sel] = 3
    cell_f32[N.HB_CN2, idx] = cn
    # layers: two records each; a trailing unpaired first record still counts (F:1284-1297)
    nlay = np.minimum((nrec - nxt + 1) // 2, 20).astype(np.int32)
    nlay = np.maximum(nlay, 0)
    cell_i32[N.HB_NLAY, idx] = nlay
    for j in range(int(nlay.max()) if m else 0):
        has1 = nlay > j
        rows = np.nonzero(has1)[0]
        if rows.size == 0:
            break
        ra = rec[rows, nxt[rows] + 2 * j]           # FORMAT(I2,F7.0,I4,4F6.0,F16.0)
        tgt = idx[rows]
        layer_i32[N.HL_TYPE, j, tgt] = _to_int(_field(ra, 0, 2))
        layer_f32[N.HL_THICK, j, tgt] = _to_real(_field(ra, 2, 7), 0, F32)
        layer_i32[N.HL_ISOIL, j, tgt] = _to_int(_field(ra, 9, 4))
        layer_f32[N.HL_PORO, j, tgt] = _to_real(_field(ra, 13, 6), 0, F32)
        layer_f32[N.HL_FC, j, tgt] = _to_real(_field(ra, 19, 6), 0, F32)
        layer_f32[N.HL_WP, j, tgt] = _to_real(_field(ra, 25, 6), 0, F32)
        layer_f32[N.HL_SW, j, tgt] = _to_real(_field(ra, 31, 6), 0, F32)
        layer_rc[j, tgt] = _to_real(_field(ra, 37, 16), 0, np.float64)
        has2 = (nxt[rows] + 2 * j + 1) < nrec
        rows2 = rows[has2]
        if rows2.size:
            rb = rec[rows2, nxt[rows2] + 2 * j + 1]  # FORMAT(F7.0,2F6.0,I3,F13.0,2F7.0,I2,G14.6)
            tg2 = idx[rows2]
            layer_f32[N.HL_XLENG, j, tg2] = _to_real(_field(rb, 0, 7), 0, F32)
            layer_f32[N.HL_SLOPE, j, tg2] = _to_real(_field(rb, 7, 6), 0, F32)
            layer_f32[N.HL_RECIR, j, tg2] = _to_real(_field(rb, 13, 6), 0, F32)
            layer_i32[N.HL_LAYR, j, tg2] = _to_int(_field(rb, 19, 3))
            layer_f32[N.HL_SUBIN, j, tg2] = _to_real(_field(rb, 22, 13), 0, F32)
            layer_f32[N.HL_PHOLE, j, tg2] = _to_real(_field(rb, 35, 7), 0, F32)
            layer_f32[N.HL_DEFEC, j, tg2] = _to_real(_field(rb, 42, 7), 0, F32)
            layer_i32[N.HL_IPQ, j, tg2] = _to_int(_field(rb, 49, 2))
            layer_f32[N.HL_TRANS, j, tg2] = _to_real(_field(rb, 51, 14), 6, F32)


# ----------------------------------------------------------------------------
# The batch
# ----------------------------------------------------------------------------
@dataclass
class CellBatch:
    """Structure-of-arrays inputs of ``help_b200_run`` (see include/help_b200.h)."""
    n_cells: int
    n_years: int
    max_layers: int
    years: np.ndarray
    pre: np.ndarray
    tmp: np.ndarray
    rad: np.ndarray
    iu4: np.ndarray
    iu7: np.ndarray
    iu13: np.ndarray
    tm_normals: np.ndarray
    idfs: np.ndarray
    cell_f32: np.ndarray
    cell_i32: np.ndarray
    layer_f32: np.ndarray
    layer_i32: np.ndarray
    layer_rc: np.ndarray
    names: list = field(default_factory=list)

    def as_struct(self) -> N.Batch:
        b = N.Batch()
        b.n_cells, b.n_years, b.max_layers = self.n_cells, self.n_years, self.max_layers
        b.n_nodes_p, b.n_nodes_t, b.n_nodes_r = self.pre.shape[0], self.tmp.shape[0], self.rad.shape[0]
        for k in ("years", "pre", "tmp", "rad", "iu4", "iu7", "iu13", "tm_normals", "idfs",
                  "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc"):
            a = getattr(self, k)
            assert a.flags["C_CONTIGUOUS"], k
            setattr(b, k, a.ctypes.data)
        return b

    def nbytes(self) -> int:
        return sum(getattr(self, k).nbytes for k in (
            "years", "pre", "tmp", "rad", "iu4", "iu7", "iu13", "tm_normals", "idfs",
            "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc"))

    def take(self, sel: np.ndarray) -> "CellBatch":
        """Sub-batch of the given cells (weather nodes are kept as they are)."""
        sel = np.asarray(sel)
        return CellBatch(
            int(sel.size), self.n_years, self.max_layers, self.years, self.pre, self.tmp, self.rad,
            self.iu4, self.iu7, self.iu13, self.tm_normals, self.idfs,
            np.ascontiguousarray(self.cell_f32[:, sel]), np.ascontiguousarray(self.cell_i32[:, sel]),
            np.ascontiguousarray(self.layer_f32[:, :, sel]), np.ascontiguousarray(self.layer_i32[:, :, sel]),
            np.ascontiguousarray(self.layer_rc[:, sel]),
            [self.names[i] for i in sel.tolist()] if self.names else [])


def pack_weather(files_p, files_t, files_r, n_years: int):
    """Stack WeatherFile objects into the [node][year][368] arrays of the ABI."""
    def stack(files, what):
        out = np.zeros((len(files), n_years, N.DAYS), dtype=F32)
        for i, wf in enumerate(files):
            if wf.data.shape[0] < n_years:
                raise ValueError(f"{what} file {i} holds {wf.data.shape[0]} years, {n_years} requested "
                                 "(the reference aborts with a Fortran end-of-file error here)")
            out[i, :, :366] = wf.data[:n_years]
        return out
    years = files_p[0].years[:n_years].astype(np.int32)
    for wf in files_p:
        if not np.array_equal(wf.years[:n_years], years):
            raise ValueError("all D4 files of one batch must cover the same years")
    return (np.ascontiguousarray(years), stack(files_p, "D4"), stack(files_t, "D7"), stack(files_r, "D13"),
            np.array([w.units for w in files_p], dtype=np.int32),
            np.array([w.units for w in files_t], dtype=np.int32),
            np.array([w.units for w in files_r], dtype=np.int32),
            np.ascontiguousarray(np.stack([w.header12 for w in files_t]).astype(F32)),
            np.ascontiguousarray(np.array([w.idfs for w in files_r], dtype=np.int32).reshape(-1, 2)))


def batch_from_cellparams(cellparams: dict) -> CellBatch:
    """``cellparams`` exactly as ``HelpManager.calc_help_cells`` builds it
    (reference managers.py:422-428): name -> (d4, d7, d13, d11 S80[3], d10 S80[n],
    fpath_out, iod, iom, ioa, ios, simu_nyear, tfsoil_F)."""
    names = list(cellparams.keys())
    n = len(names)
    nyears = {int(p[10]) for p in cellparams.values()}
    if len(nyears) > 1:
        raise ValueError("all cells of one batch must simulate the same number of years")
    n_years = nyears.pop() if nyears else 1
    cache: dict = {}

    def node(path, kind, table, order):
        key = os.fspath(path)
        if isinstance(key, bytes):
            key = key.decode()
        key = key.strip()
        if key not in table:
            if key not in cache:
                if not os.path.exists(key):
                    raise FileNotFoundError(f"{kind} weather file not found: {key}")
                cache[key] = read_weather_file(key, kind)
            table[key] = len(order)
            order.append(cache[key])
        return table[key]

    tp, tt, tr, fp, ft, fr = {}, {}, {}, [], [], []
    cell_f32 = np.zeros((N.HB_NCELL_F32, n), dtype=F32)
    cell_i32 = np.zeros((N.HB_NCELL_I32, n), dtype=np.int32)
    # Records are gathered per cell and converted to byte matrices in a few large calls (one per
    # D10 record count): the per-cell work is list appends only.
    def as_s80(a):
        a = np.asarray(a).reshape(-1)
        if a.dtype.kind == "U":
            a = np.char.encode(a, "ascii", "replace")
        return a if (a.dtype.kind == "S" and a.dtype.itemsize == 80) else np.char.ljust(a.astype("S80"), 80).astype("S80")

    d10_lists: dict = {}
    d11_list = []
    node_ids = np.zeros((3, n), dtype=np.int32)
    tfsoil = np.zeros(n, dtype=np.float64)
    last = [None, None, None, 0, 0, 0]                       # consecutive cells usually share their weather files
    for i, name in enumerate(names):
        p = cellparams[name]
        for k, (kind, table, order) in enumerate((("D4", tp, fp), ("D7", tt, ft), ("D13", tr, fr))):
            if p[k] is not last[k]:
                last[k], last[3 + k] = p[k], node(p[k], kind, table, order)
            node_ids[k, i] = last[3 + k]
        r11 = as_s80(p[3])
        if r11.shape[0] < 3:
            raise ValueError(f"cell {name}: D11 needs 3 records")
        d11_list.append(r11[:3])
        r10 = as_s80(p[4])
        g10 = d10_lists.setdefault(r10.shape[0], ([], []))
        g10[0].append(i)
        g10[1].append(r10)
        tfsoil[i] = p[11]
    cell_i32[N.HB_NODE_P], cell_i32[N.HB_NODE_T], cell_i32[N.HB_NODE_R] = node_ids
    cell_f32[N.HB_TFSOIL] = tfsoil.astype(F32)
    d11_rec = _records(np.concatenate(d11_list)).reshape(n, 3, 80) if n else np.zeros((0, 3, 80), dtype=np.uint8)
    d10_groups = {nrec: (idx, _records(np.concatenate(recs)).reshape(len(idx), nrec, 80))
                  for nrec, (idx, recs) in d10_lists.items()}
    max_layers = 1
    for nrec in d10_groups:
        max_layers = max(max_layers, min(20, (nrec - 2) // 2 + 1))
    layer_f32 = np.zeros((N.HL_NLAYER_F32, max_layers, n), dtype=F32)
    layer_i32 = np.zeros((N.HL_NLAYER_I32, max_layers, n), dtype=np.int32)
    layer_rc = np.zeros((max_layers, n), dtype=np.float64)
    parse_d11(d11_rec, cell_f32, cell_i32, np.arange(n))
    for nrec, (idx, recs) in d10_groups.items():
        if nrec < 3:
            raise ValueError("a D10 input needs at least 3 records")
        parse_d10(recs, cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, np.asarray(idx))
    if n == 0:
        raise ValueError("cellparams is empty")
    used = max(1, int(cell_i32[N.HB_NLAY].max()))
    if used < max_layers:
        max_layers = used
        layer_f32 = np.ascontiguousarray(layer_f32[:, :used])
        layer_i32 = np.ascontiguousarray(layer_i32[:, :used])
        layer_rc = np.ascontiguousarray(layer_rc[:used])
    years, pre, tmp, rad, iu4, iu7, iu13, tm, idfs = pack_weather(fp, ft, fr, n_years)
    return CellBatch(n, n_years, max_layers, years, pre, tmp, rad, iu4, iu7, iu13, tm, idfs,
                     cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, names)
