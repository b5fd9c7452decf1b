"""Extract the reference's RCRA golden test case into committed fixtures.

Run once in the build container (``/root/reference`` present)::

    python tests/golden/make_rcra_golden.py

Source: ``/root/reference/pyhelp/tests/rca_original_testcase_1997`` -- the test
case distributed with HELP 3.07 (1997) that the reference uses to pin its
Fortran (``pyhelp/tests/test_help3o.py:33-118``).  This script copies the five
INPUT DATA files verbatim (they are data, not code) and parses ``RCRA.OUT``
(the expected output) into ``rcra_expected.json``:

* ``daily``      [1096][16]: day, air-freeze star, soil-freeze star, rain,
  runoff, ET, evap-zone water, then head/drain/leak for subprofiles 1-3
  (inches; RCRA.OUT:316-690, 819-1197, 1322-1700),
* ``monthly``    per year, per row label: 12 values (inches; RCRA.OUT:696-733 ...),
* ``annual``     per year: label -> inches (RCRA.OUT:759-800 ...),
* ``initial_sw`` per-layer vol/vol after the spin-up year (RCRA.OUT:54-221),
* ``final_sw``   per-layer vol/vol at the end of year 3 (RCRA.OUT:2040-2066),
* ``kat``        the yearly numbers hard-coded in test_help3o.py:87-118.

Nothing under tests/ reads /root/reference at run time; only this script does.
"""
import json
import os.path as osp
import re
import shutil

SRC = "/root/reference/pyhelp/tests/rca_original_testcase_1997"
DST = osp.join(osp.dirname(osp.abspath(__file__)), "rcra")


def fnum(tok: str) -> float:
    tok = tok.strip()
    if tok.startswith(".") or tok.startswith("-."):
        tok = tok.replace(".", "0.", 1)
    return float(tok)


def main():
    import os
    os.makedirs(DST, exist_ok=True)
    for ext in ("D4", "D7", "D13", "D10", "D11"):
        shutil.copyfile(osp.join(SRC, "RCRA." + ext), osp.join(DST, "RCRA." + ext))
    lines = open(osp.join(SRC, "RCRA.OUT"), errors="replace").read().splitlines()

    daily, monthly, annual, initial_sw, final_sw = [], {}, {}, [], []
    i = 0
    year = None
    while i < len(lines):
        ln = lines[i]
        m = re.search(r"DAILY OUTPUT FOR YEAR\s+(\d+)", ln)
        if m:
            year = int(m.group(1))
            i += 1
            continue
        # daily rows: "  ddd  A  S  rain runoff et ezone head drain leak x3"
        m = re.match(r"^\s{1,4}(\d{1,3})\s(.{6})\s*(-?\d*\.\d+)\s+(-?\d*\.\d+)\s+(-?\d*\.\d+)\s+(-?\d*\.\d+)\s+(.*)$", ln)
        if m and year is not None and "E. ZONE" not in ln and len(ln) > 100:
            day = int(m.group(1))
            stars = m.group(2)
            air = 1 if stars[1:3].strip() == "*" or stars[:3].count("*") else 0
            # columns: day(I5) 2X A1 2X A1 ... ; use positions of the two star columns
            air = 1 if ln[7:8] == "*" else 0
            soil = 1 if ln[10:11] == "*" else 0
            rest = [fnum(t) for t in m.group(7).split()]
            row = [year, day, air, soil, fnum(m.group(3)), fnum(m.group(4)), fnum(m.group(5)),
                   fnum(m.group(6))] + rest
            assert len(row) == 17, (ln, row)
            daily.append(row)
            i += 1
            continue
        m = re.search(r"MONTHLY TOTALS \(IN INCHES\) FOR YEAR\s+(\d+)", ln)
        if m:
            y = int(m.group(1))
            rows = {}
            i += 1
            while "MONTHLY SUMMARIES" not in lines[i]:
                l1 = lines[i]
                if re.match(r"^ (PRECIPITATION|RUNOFF|EVAPOTRANSPIRATION|LATERAL DRAINAGE COLLECTED|PERCOLATION/LEAKAGE THROUGH)", l1):
                    l2 = lines[i + 1]
                    label = l1[:32].strip()
                    lm = re.search(r"LAYER\s+(\d+)", l2[:32])
                    if lm:
                        label += " LAYER %d" % int(lm.group(1))
                    vals = [fnum(t) for t in l1[32:].split()] + [fnum(t) for t in l2[32:].split()]
                    assert len(vals) == 12, (l1, l2)
                    rows[label] = vals
                    i += 2
                    continue
                i += 1
            monthly[str(y)] = rows
            continue
        m = re.search(r"ANNUAL TOTALS FOR YEAR\s+(\d+)", ln)
        if m:
            y = int(m.group(1))
            rows = {}
            i += 1
            while "ANNUAL WATER BUDGET BALANCE" not in lines[i]:
                mm = re.match(r"^\s{3}([A-Z][A-Z./ ]+?(?:LAYER\s+\d+)?)\s{2,}(-?\d*\.\d+)", lines[i])
                if mm:
                    rows[re.sub(r"\s+", " ", mm.group(1).strip())] = fnum(mm.group(2))
                i += 1
            annual[str(y)] = rows
            continue
        m = re.search(r"INITIAL SOIL WATER CONTENT\s+=\s+(-?\d*\.\d+)", ln)
        if m:
            initial_sw.append(fnum(m.group(1)))
        if "FINAL WATER STORAGE AT END OF YEAR" in ln:
            i += 1
            while i < len(lines) and "SNOW WATER" not in lines[i]:
                mm = re.match(r"^\s+(\d+)\s+(-?\d*\.\d+)\s+(-?\d*\.\d+)\s*$", lines[i])
                if mm:
                    final_sw.append(fnum(mm.group(3)))
                i += 1
            continue
        i += 1

    assert len(daily) == 365 + 365 + 365 + (1 if False else 0) or len(daily) in (1095, 1096), len(daily)
    kat = {  # pyhelp/tests/test_help3o.py:87-118 (inches per year, tolerance 0.1 in)
        "precip": [48.53, 58.32, 56.71],
        "runoff": [0.596, 4.130, 2.397],
        "rechg": [0.000004, 0.000006, 0.000007],
        "evapo": [32.178, 33.420, 36.540],
        "subrun1": [15.4971, 22.8698, 19.1360],
        "subrun2": [0.0543 + 0.0833, 0.1481 + 0.1658, 0.1835 + 0.1935],
        "tolerance_in": 0.1,
        "mm_to_in": 0.0393701,
    }
    out = {"source": "pyhelp/tests/rca_original_testcase_1997/RCRA.OUT (HELP 3.07, 1997)",
           "daily_columns": ["year", "day", "air_freeze", "soil_freeze", "rain", "runoff", "et",
                             "ezone_water", "head1", "drain1", "leak1", "head2", "drain2",
                             "leak2", "head3", "drain3", "leak3"],
           "daily": daily, "monthly": monthly, "annual": annual,
           "initial_sw": initial_sw, "final_sw": final_sw, "kat": kat}
    with open(osp.join(DST, "rcra_expected.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("daily rows", len(daily), "monthly years", list(monthly), "annual", list(annual),
          "init", initial_sw, "final", final_sw)
    print(annual["1"])
    print(monthly["1"].keys())


if __name__ == "__main__":
    main()
