"""The monthly part of a HELP ``.OUT`` file, written from the oracle's REAL*4 accumulators (TEST
INFRASTRUCTURE): what ``OUTMO`` / ``OUTMO2`` / ``OUTSW`` print with ``iom = 1`` and every other switch off
(reference ``pyhelp/HELP3O.FOR:6305-6406`` frame of asterisks, ``:6415-6682`` the block -- labels, layer
numbers, the x 25.4, ``12(1X,F7.3)`` --, ``:8543`` the 'FINAL WATER STORAGE' trailer the parser stops at).

Purpose: the reference's own parser ``read_monthly_help_output`` (``pyhelp/processing.py:64-147``) can be run
on this text, which pins the row -> key map, the ``split()[-12:]`` slicing, the F7.3 quantisation and the
units slip of the sixth subprofile (``HELP3O.FOR:6593-6594``: its Jul-Dec percolation is printed in inches)
with the reference's code instead of a reading of it (tests/test_outmo2_parser.py).

Not reproduced: the 'LATERAL DRAINAGE RECIRCULATED' rows (their accumulators are not exported; the parser
ignores those rows: 'LATERAL' does not contain 'LAT. DRAINAGE').
"""
from __future__ import annotations

import numpy as np

MDAY = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)


def f73(v) -> str:
    """Fortran F7.3 of a REAL*4 value: the exact binary value rounded to 3 decimals, '*******' when it does not fit."""
    s = "%.3f" % float(np.float32(v))
    if s.startswith("-0.") and float(s) == 0.0 and len(s) <= 7:
        pass                                   # gfortran prints -0.000 for a negative value that rounds to zero
    return s.rjust(7) if len(s) <= 7 else "*" * 7


def f83(v) -> str:
    s = "%.3f" % float(np.float32(v))
    return s.rjust(8) if len(s) <= 8 else "*" * 8


def monthly_block(year: int, rows: np.ndarray, lp, ld, subin_layers=()) -> list:
    """One year.  ``rows``: float32 [14][12] in inches (pre, run, et, drn1..5, prc1..6); ``lp``/``ld``: LP(1..6), LD(1..5);
    ``subin_layers``: (layer number, SUBIN in/day) of layers with subsurface inflow (row 6077)."""
    mm = np.float32(25.4)
    L = [" ", " ", " " + "*" * 126, " ", " " * 22 + "MONTHLY TOTALS (MM) FOR YEAR%5d" % year, " " + "-" * 126, " ",
         " " * 34 + "JAN" + "".join(" " * 5 + m for m in ("FEB", "MAR", "APR", "MAY", "JUN", "JUL", "AUG", "SEP", "OCT", "NOV", "DEC")),
         " " * 31 + " -------" * 12]
    row12 = lambda v: "".join(" " + f73(x) for x in v)
    L.append(" PRECIPITATION" + " " * 17 + row12(rows[0] * mm))
    L.append(" RUNOFF" + " " * 24 + row12(rows[1] * mm))
    L.append(" EVAPOTRANSPIRATION" + " " * 12 + row12(rows[2] * mm))
    L.append("")
    leap = (year % 4 == 0)
    first = 1
    for n in range(6):
        if lp[n] <= 0:
            continue
        for k, sub in subin_layers:
            if first <= k <= lp[n]:
                m = [np.float32(d) * np.float32(sub) for d in MDAY]
                if leap:
                    m[1] = np.float32(m[1] + np.float32(sub))
                L += [" ", " SUBSURFACE INFLOW INTO  " + " " * 7 + "".join(f83(x * mm) for x in m[:6]),
                      "   LAYER%3d         " % k + " " * 12 + "".join(f83(x * mm) for x in m[6:])]
        first = lp[n] + 1
        if n < 5 and ld[n] > 0:
            L.append(" LAT. DRAINAGE IN LAYER%3d" % ld[n] + " " * 5 + row12(rows[3 + n] * mm))
        prc = rows[8 + n]
        vals = list(prc * mm) if n < 5 else list(prc[:6] * mm) + list(prc[6:])        # sic: F:6593-6594
        L.append(" PERCOLATION THROUGH LAYER%3d" % lp[n] + "  " + row12(vals))
        L.append("")
    L += [" " + "*" * 126, "", ""]
    return L


def final_storage_trailer(last_year: int, n_layers: int) -> list:
    """Enough of OUTSW (F:8530-8566) for the parser: it stops at the first line with 'FINAL WATER STORAGE'."""
    L = [" \x0c", " " + "*" * 78, " ", " " * 20 + "FINAL WATER STORAGE AT END OF YEAR%5d" % last_year, " " * 5 + "-" * 70,
         " " * 21 + "LAYER" + " " * 10 + "(CM)" + " " * 9 + "(VOL/VOL)", " " * 21 + "-----" + " " * 9 + "------" + " " * 8 + "---------"]
    return L + [" ", " " + "*" * 78, " " + "*" * 78, "", "", "", ""]


def write_out(path: str, res: dict, d10_input=None):
    """``res`` = ``oracle.run_simulation(..., full=True)``.  Writes the file the reference would have written with
    iod = ioa = ios = 0, iom = 1 (monthly blocks + final-storage trailer)."""
    info = res["info"]
    subs = []
    if d10_input is not None:                   # layers with subsurface inflow (LSIN, F:2316-2387): D10 field SUBIN, mm/yr -> in/day
        recs = [r.decode() if isinstance(r, bytes) else str(r) for r in np.asarray(d10_input).reshape(-1)]
        iu10 = int(recs[1][0:2])
        irun = int(recs[1][30:32] or 0)
        first = 3 if irun in (1, 2, 3) else 2
        for j in range((len(recs) - first + 1) // 2):
            r2 = recs[first + 2 * j + 1] if first + 2 * j + 1 < len(recs) else ""
            txt = r2[22:35].strip()
            v = float(txt) if txt else 0.0
            if v > 0:
                sub = np.float32(np.float32(v) / np.float32(365.0))
                if iu10 != 1:
                    sub = np.float32(sub / np.float32(25.4))
                subs.append((j + 1, sub))
    lines = []
    for iy, year in enumerate(res["years"].tolist()):
        lines += monthly_block(int(year), res["monthly_in"][iy], info["lp"], info["ld"], subs)
    lines += final_storage_trailer(int(res["years"][-1]), info["lay"])
    with open(path, "w", newline="\n") as f:
        f.write("\n".join(lines) + "\n")
    return path
