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

import re

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
            try:
                vals = t.astype(np.float64)
            except ValueError:                          # "1.5+2": an exponent without its letter
                t = np.array([re.sub(rb"(?<=[0-9.])([+-])", rb"E\1", x, count=1) if b"E" not in x.upper() else x for x in t])
                vals = t.astype(np.float64)
        if d > 0:
            nodot = np.char.find(t, b".") < 0          # Fw.d: no point -> the last d digits of the significand are decimals, exponent or not
            if nodot.any():
                vals[nodot] = vals[nodot] / (10.0 ** d)  # an integer below 2^53 over a power of ten: one rounding
                # with an exponent the value read above is already rounded: convert significand x 10^(exponent - d)
                # in one step, as libgfortran does (and csrc/help_parse.cuh)
                for i in np.nonzero(nodot & (np.char.find(np.char.upper(t), b"E") >= 0))[0]:
                    mant, _, ex = t[i].decode().upper().partition("E")
                    vals[i] = float(f"{mant}E{int(ex) - d}")
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


def parse_matrix(rec: np.ndarray, specs, where: str = "host") -> np.ndarray:
    """Formatted READ of the fields ``specs`` = [(first column, width, implied decimals), ...] from every record of the
    (n, reclen) uint8 matrix ``rec``: float64 [n_fields][n].  ``where="device"`` does it in one kernel
    (``help_b200_parse_fields``, csrc/help_parse.cuh; same values: tests/test_gpu_pack.py); fields the kernel marks as
    needing more than one exact operation are converted here, fields that are not numbers raise ValueError like
    numpy's conversion does on the host path."""
    n = rec.shape[0]
    out = np.zeros((len(specs), n), dtype=np.float64)
    if n == 0:
        return out
    if where == "host":
        for k, (a, w, d) in enumerate(specs):
            out[k] = _to_real(_field(rec, a, w), d, np.float64)
        return out
    if where != "device":
        raise ValueError("where must be 'host' or 'device'")
    rec = np.ascontiguousarray(rec)
    col = np.array([s[0] for s in specs], dtype=np.int32)
    wid = np.array([s[1] for s in specs], dtype=np.int32)
    dec = np.array([s[2] for s in specs], dtype=np.int32)
    flags = np.zeros((len(specs), n), dtype=np.uint8)
    N.require_device()
    N.check(N.lib().help_b200_parse_fields(N.ptr(rec), n, rec.shape[1], N.ptr(col), N.ptr(wid), N.ptr(dec), len(specs),
                                           N.ptr(out), N.ptr(flags)), "parse_fields")
    if flags.any():
        for k, r in zip(*np.nonzero(flags)):
            a, w, d = specs[k]
            out[k, r] = _to_real(_field(rec[r:r + 1], a, w), d, np.float64)[0]      # raises ValueError for a non-number
    return out


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


def _load_weather_records(path: str):
    """(header records (4, 80), data records (37 * nblocks, 80), nblocks) of one D4/D7/D13 file as uint8 matrices
    (records cut or blank-padded to 80 columns, CR dropped)."""
    with open(path, "rb") as f:
        raw = f.read()
    # fast path: the year blocks of a file written by a program are records of ONE length (PyHELP's writers, HELP's own):
    # the data part is reshaped in one step instead of being split into a Python list of lines
    p4, k = -1, 0
    while k < 4:
        p4 = raw.find(b"\n", p4 + 1)
        if p4 < 0:
            break
        k += 1
    if k == 4:
        e1 = raw.find(b"\n", p4 + 1)
        L = e1 - p4 - 1
        nbytes = len(raw) - (p4 + 1)
        if e1 > 0 and 0 < L <= 80 and (nbytes % (L + 1) == 0 or (nbytes + 1) % (L + 1) == 0):
            nlines = (nbytes + 1) // (L + 1)
            body = np.frombuffer(raw, dtype=np.uint8, count=nlines * (L + 1) - (1 if nbytes % (L + 1) else 0), offset=p4 + 1)
            if body.size % (L + 1):
                body = np.concatenate([body, np.array([10], dtype=np.uint8)])
            body = body.reshape(nlines, L + 1)
            nblocks = nlines // 37
            if nblocks >= 1 and (body[:, L] == 10).all() and not (body[:, :L] == 10).any():
                data = np.full((37 * nblocks, 80), 32, dtype=np.uint8)
                data[:, :L] = body[:37 * nblocks, :L]
                hdr = _records(np.array(raw[:p4].split(b"\n"), dtype="S80"))
                for a in (hdr, data):
                    a[a == 13] = 32
                    a[a == 0] = 32
                    a[a > 127] = 63
                return hdr, data, nblocks
    lines = raw.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if len(lines) < 4:
        raise ValueError(f"{path}: truncated weather header")
    nblocks = (len(lines) - 4) // 37
    if nblocks < 1:
        raise ValueError(f"{path}: no complete year of data")
    allrec = _records(np.array(lines[:4 + 37 * nblocks], dtype="S80"))
    allrec[allrec == 13] = 32
    allrec[allrec > 127] = 63
    return allrec[:4], allrec[4:], nblocks


def read_weather_files(paths, kind: str, where: str = "host") -> list:
    """READIN header read (F:1215-1218) + READCD year blocks (F:1100-1129) of several files of one kind: the records
    of all files are converted in one :func:`parse_matrix` call."""
    loads = [_load_weather_records(os.fspath(p)) for p in paths]
    if not loads:
        return []
    hdr = np.stack([h for h, _, _ in loads])                               # (nf, 4, 80)
    units = _to_int(_field(np.ascontiguousarray(hdr[:, 1]), 0, 2))
    h12 = np.ascontiguousarray(hdr[:, 3, :72]).reshape(-1, 6)
    header12 = _to_real(np.char.strip(h12.view("S6").reshape(-1)), 2, F32).reshape(-1, 12)
    rec = np.concatenate([r for _, r, _ in loads]) if len(loads) > 1 else loads[0][1]
    if kind == "D4":      # FORMAT(I10,10F5.2)
        y_w, v0, v_w = 10, 10, 5
    else:                 # FORMAT(I5,10F6.1) / (I5,10F6.0)
        y_w, v0, v_w = 5, 5, 6
    nrec = np.array([r.shape[0] for _, r, _ in loads])
    file_of = np.repeat(np.arange(len(loads)), nrec)
    # the implied decimals depend on the file's units for D13 (F:1127-1128); READIN's pre-scan of D13 always reads F6.1
    # (formats 5400/5410, F:1364-1365): both variants are read, they differ only for fields without a decimal point
    d_main = {"D4": 2, "D7": 1}.get(kind)
    specs = [(0, y_w, 0)] + [(v0 + k * v_w, v_w, d_main if d_main is not None else 1) for k in range(10)]
    if kind == "D13":
        specs += [(v0 + k * v_w, v_w, 0) for k in range(10)]
    vals = parse_matrix(rec, specs, where)
    out, pos = [], 0
    for i, (_, r, nblocks) in enumerate(loads):
        sl = slice(pos, pos + r.shape[0])
        pos += r.shape[0]
        yrs = vals[0, sl].astype(np.int64).astype(np.int32).reshape(nblocks, 37)[:, -1]
        v1 = vals[1:11, sl].T.reshape(nblocks, 370)[:, :366].astype(F32)              # (.., d = 2 / 1)
        if kind == "D13" and int(units[i]) != 1:
            data = vals[11:21, sl].T.reshape(nblocks, 370)[:, :366].astype(F32)        # READCD: F6.0 for SI radiation
        else:
            data = v1
        wf = WeatherFile(kind, int(units[i]), yrs, np.ascontiguousarray(data), header12[i].astype(F32))
        if kind == "D13":
            wf.idfs = idfs_prescan(v1, int(units[i]))
        out.append(wf)
    return out


def read_weather_file(path: str, kind: str) -> WeatherFile:
    """One file (host conversion)."""
    return read_weather_files([path], kind, "host")[0]


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
D11_SPECS = [(0, 10, 2), (10, 4, 0), (14, 4, 0), (18, 7, 0), (25, 8, 0), (33, 5, 0), (38, 5, 0), (43, 5, 0), (48, 5, 0), (53, 5, 0)]
D10_HEAD_SPECS = [(0, 2, 0), (2, 2, 0), (4, 10, 0), (24, 6, 0), (30, 2, 0)]
D10_A_SPECS = [(0, 2, 0), (2, 7, 0), (9, 4, 0), (13, 6, 0), (19, 6, 0), (25, 6, 0), (31, 6, 0), (37, 16, 0)]
D10_B_SPECS = [(0, 7, 0), (7, 6, 0), (13, 6, 0), (19, 3, 0), (22, 13, 0), (35, 7, 0), (42, 7, 0), (49, 2, 0), (51, 14, 6)]


def _i32(v):
    return v.astype(np.int64).astype(np.int32)


def parse_d11(rec3: np.ndarray, cell_f32: np.ndarray, cell_i32: np.ndarray, idx: np.ndarray, where: str = "host"):
    """rec3: (m, 3, 80) uint8.  F:1221-1225: (I2) ; (F10.2,I4,I4,F7.0,F8.0,5F5.0)."""
    cell_i32[N.HB_IU11, idx] = _i32(parse_matrix(np.ascontiguousarray(rec3[:, 0]), [(0, 2, 0)], where)[0])
    v = parse_matrix(np.ascontiguousarray(rec3[:, 2]), D11_SPECS, where)
    cell_f32[N.HB_ULAT, idx] = v[0].astype(F32)
    cell_i32[N.HB_IPL, idx] = _i32(v[1])
    cell_i32[N.HB_IHV, idx] = _i32(v[2])
    cell_f32[N.HB_ULAI, idx] = v[3].astype(F32)
    cell_f32[N.HB_EDEPTH, idx] = v[4].astype(F32)
    cell_f32[N.HB_WIND, idx] = v[5].astype(F32)
    for q in range(4):
        cell_f32[N.HB_HUM1 + q, idx] = v[6 + q].astype(F32)


def parse_d10(rec: np.ndarray, cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, idx: np.ndarray, where: str = "host"):
    """rec: (m, nrec, 80) uint8 with a common record count.  F:1260-1300."""
    m, nrec = rec.shape[0], rec.shape[1]
    h = parse_matrix(np.ascontiguousarray(rec[:, 1]), D10_HEAD_SPECS, where)      # FORMAT(I2,I2,2F10.0,F6.0,I2)
    cell_i32[N.HB_IU10, idx] = _i32(h[0])
    cell_i32[N.HB_IPRE, idx] = _i32(h[1])
    cell_f32[N.HB_OSNO, idx] = h[2].astype(F32)
    cell_f32[N.HB_FRUNOF, idx] = h[3].astype(F32)
    irun = _i32(h[4])
    r3 = np.ascontiguousarray(rec[:, 2])
    cn = np.zeros(m, dtype=F32)
    nxt = np.full(m, 2, dtype=np.int64)  # record index of the first layer record
    for code, col in ((1, 0), (2, 21), (3, 20)):   # F:1268-1281: where CN2 sits for IRUN = 1 / 2 / 3
        sel = irun == code
        if sel.any():
            cn[sel] = parse_matrix(r3[sel], [(col, 7, 0)], where)[0].astype(F32)
            nxt[sel] = 3
    cell_f32[N.HB_CN2, idx] = cn
    # layers: two records each; a trailing unpaired first record still counts (F:1284-1297)
    nlay = np.minimum((nrec - nxt + 1) // 2, 20).astype(np.int32)
    nlay = np.maximum(nlay, 0)
    cell_i32[N.HB_NLAY, idx] = nlay
    L = int(nlay.max()) if m else 0
    if L == 0:
        return
    # every first ("A") and second ("B") layer record of the group in two matrices: two conversions for all layers
    rows_a, lay_a, rows_b, lay_b = [], [], [], []
    for j in range(L):
        rows = np.nonzero(nlay > j)[0]
        rows_a.append(rows); lay_a.append(np.full(rows.size, j))
        rows2 = rows[(nxt[rows] + 2 * j + 1) < nrec]
        rows_b.append(rows2); lay_b.append(np.full(rows2.size, j))
    rows_a, lay_a = np.concatenate(rows_a), np.concatenate(lay_a)
    rows_b, lay_b = np.concatenate(rows_b), np.concatenate(lay_b)
    va = parse_matrix(rec[rows_a, nxt[rows_a] + 2 * lay_a], D10_A_SPECS, where)        # FORMAT(I2,F7.0,I4,4F6.0,F16.0)
    tgt = idx[rows_a]
    layer_i32[N.HL_TYPE, lay_a, tgt] = _i32(va[0])
    layer_f32[N.HL_THICK, lay_a, tgt] = va[1].astype(F32)
    layer_i32[N.HL_ISOIL, lay_a, tgt] = _i32(va[2])
    layer_f32[N.HL_PORO, lay_a, tgt] = va[3].astype(F32)
    layer_f32[N.HL_FC, lay_a, tgt] = va[4].astype(F32)
    layer_f32[N.HL_WP, lay_a, tgt] = va[5].astype(F32)
    layer_f32[N.HL_SW, lay_a, tgt] = va[6].astype(F32)
    layer_rc[lay_a, tgt] = va[7]
    if rows_b.size:
        vb = parse_matrix(rec[rows_b, nxt[rows_b] + 2 * lay_b + 1], D10_B_SPECS, where)  # FORMAT(F7.0,2F6.0,I3,F13.0,2F7.0,I2,G14.6)
        tg2 = idx[rows_b]
        layer_f32[N.HL_XLENG, lay_b, tg2] = vb[0].astype(F32)
        layer_f32[N.HL_SLOPE, lay_b, tg2] = vb[1].astype(F32)
        layer_f32[N.HL_RECIR, lay_b, tg2] = vb[2].astype(F32)
        layer_i32[N.HL_LAYR, lay_b, tg2] = _i32(vb[3])
        layer_f32[N.HL_SUBIN, lay_b, tg2] = vb[4].astype(F32)
        layer_f32[N.HL_PHOLE, lay_b, tg2] = vb[5].astype(F32)
        layer_f32[N.HL_DEFEC, lay_b, tg2] = vb[6].astype(F32)
        layer_i32[N.HL_IPQ, lay_b, tg2] = _i32(vb[7])
        layer_f32[N.HL_TRANS, lay_b, tg2] = vb[8].astype(F32)


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
            if isinstance(a, N.DeviceTable):          # weather packed on the device and left there (grid.weather_to_nodes)
                assert k in ("pre", "tmp", "rad"), k
                setattr(b, k, a.ptr)
                continue
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


def batch_from_cellparams(cellparams: dict, parse: str = "host") -> CellBatch:
    """``cellparams`` exactly as ``HelpManager.calc_help_cells`` builds it
    (reference managers.py:422-428): name -> (d4, d7, d13, d11 S80[3], d10 S80[n],
    fpath_out, iod, iom, ioa, ios, simu_nyear, tfsoil_F).  ``parse``: where the formatted reads of the
    records and weather files run -- ``"host"`` (numpy) or ``"device"`` (:func:`parse_matrix`)."""
    names = list(cellparams.keys())
    n = len(names)
    nyears = {int(p[10]) for p in cellparams.values()}
    if len(nyears) > 1:
        raise ValueError("all cells of one batch must simulate the same number of years")
    n_years = nyears.pop() if nyears else 1
    tables = ({}, {}, {})                                    # path -> node index, per kind, in order of first use

    def node(path, kind, table):
        key = os.fspath(path)
        if isinstance(key, bytes):
            key = key.decode()
        key = key.strip()
        if key not in table:
            if not os.path.exists(key):
                raise FileNotFoundError(f"{kind} weather file not found: {key}")
            table[key] = len(table)
        return table[key]

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
    vals = list(cellparams.values())
    for k, kind in enumerate(("D4", "D7", "D13")):       # (consecutive cells usually share their weather files)
        ids, cur, cur_id, table = [0] * n, None, 0, tables[k]
        for i, p in enumerate(vals):
            q = p[k]
            if q is not cur:
                cur, cur_id = q, node(q, kind, table)
            ids[i] = cur_id
        node_ids[k] = ids
    tfsoil[:] = np.fromiter((p[11] for p in vals), dtype=np.float64, count=n)
    r11s, r10s = [p[3] for p in vals], [p[4] for p in vals]
    s80 = np.dtype("S80")

    def column_of_records(a):          # an S80 array of shape (k,) or (k, 1): what the reference's formatters build
        return type(a) is np.ndarray and a.dtype == s80 and (a.ndim == 1 or (a.ndim == 2 and a.shape[1] == 1))

    if all(column_of_records(a) for a in r11s) and all(column_of_records(a) for a in r10s) and \
       len({a.ndim for a in r11s}) == 1 and len({a.ndim for a in r10s}) == 1:
        # the usual case: no per-cell conversion, one concatenation per group of equal record count
        n11 = np.fromiter((a.shape[0] for a in r11s), dtype=np.int64, count=n)
        if (n11 < 3).any():
            raise ValueError(f"cell {names[int(np.argmax(n11 < 3))]}: D11 needs 3 records")
        d11_list = [np.concatenate(r11s if (n11 == 3).all() else [a[:3] for a in r11s]).reshape(-1)]
        n10 = np.fromiter((a.shape[0] for a in r10s), dtype=np.int64, count=n)
        for nrec in np.unique(n10).tolist():
            idx = np.nonzero(n10 == nrec)[0]
            d10_lists[nrec] = (idx.tolist(), [np.concatenate(r10s if idx.size == n else [r10s[i] for i in idx.tolist()]).reshape(-1)])
    else:
        for i, name in enumerate(names):
            r11 = as_s80(r11s[i])
            if r11.shape[0] < 3:
                raise ValueError(f"cell {name}: D11 needs 3 records")
            d11_list.append(r11[:3])
            r10 = as_s80(r10s[i])
            g10 = d10_lists.setdefault(r10.shape[0], ([], []))
            g10[0].append(i)
            g10[1].append(r10)
    if n == 0:
        raise ValueError("cellparams is empty")
    fp, ft, fr = (read_weather_files(list(t.keys()), kind, parse) for t, kind in zip(tables, ("D4", "D7", "D13")))
    cell_i32[N.HB_NODE_P], cell_i32[N.HB_NODE_T], cell_i32[N.HB_NODE_R] = node_ids
    cell_f32[N.HB_TFSOIL] = tfsoil.astype(F32)
    d11_rec = _records(np.concatenate(d11_list)).reshape(n, 3, 80)
    d10_groups = {nrec: (idx, _records(np.concatenate(recs)).reshape(len(idx), nrec, 80))
                  for nrec, (idx, recs) in d10_lists.items()}
    max_layers = 1
    for nrec in d10_groups:
        max_layers = max(max_layers, min(20, (nrec - 2) // 2 + 1))
    layer_f32 = np.zeros((N.HL_NLAYER_F32, max_layers, n), dtype=F32)
    layer_i32 = np.zeros((N.HL_NLAYER_I32, max_layers, n), dtype=np.int32)
    layer_rc = np.zeros((max_layers, n), dtype=np.float64)
    parse_d11(d11_rec, cell_f32, cell_i32, np.arange(n), parse)
    for nrec, (idx, recs) in d10_groups.items():
        if nrec < 3:
            raise ValueError("a D10 input needs at least 3 records")
        parse_d10(recs, cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, np.asarray(idx), parse)
    used = max(1, int(cell_i32[N.HB_NLAY].max()))
    if used < max_layers:
        max_layers = used
        layer_f32 = np.ascontiguousarray(layer_f32[:, :used])
        layer_i32 = np.ascontiguousarray(layer_i32[:, :used])
        layer_rc = np.ascontiguousarray(layer_rc[:used])
    years, pre, tmp, rad, iu4, iu7, iu13, tm, idfs = pack_weather(fp, ft, fr, n_years)
    return CellBatch(n, n_years, max_layers, years, pre, tmp, rad, iu4, iu7, iu13, tm, idfs,
                     cell_f32, cell_i32, layer_f32, layer_i32, layer_rc, names)
