"""Packs the reference's example weather tables (BASELINE configs[0]:
/root/reference/example/{precip,airtemp,solrad}_input_data.csv, 24/24/1 nodes, 2000-2010) into
tests/golden/example_weather.npz so that the config-1 parity test can run where /root/reference
does not exist.  Values are stored at the precision the reference writes them to the HELP input
files (precip and air temperature 1 decimal, solar radiation 2 decimals; weather_reader.py:22-58),
which is all the simulation ever sees.  Also stores the first rows of each table verbatim
(float64) for the csv-reader test.

    python tests/golden/make_example_weather.py        (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pyhelp_b200 import managers  # noqa: E402

REF = "/root/reference/example/"
out = {}
for var, fname, dec in (("precip", "precip_input_data.csv", 1), ("airtemp", "airtemp_input_data.csv", 1),
                        ("solrad", "solrad_input_data.csv", 2)):
    t = managers.load_weather_from_csv(REF + fname)
    out[var + "_lat"] = t.columns.get_level_values("lat_dd").values.astype(np.float64)
    out[var + "_lon"] = t.columns.get_level_values("lon_dd").values.astype(np.float64)
    # the value the reference's '{:>6.1f}' / '{:>6.2f}' prints, kept as float32 (exactly re-printable)
    out[var] = np.array([[float(("{:." + str(dec) + "f}").format(v)) for v in row] for row in t.values], dtype=np.float32)
    out[var + "_head_exact"] = t.values[:40].astype(np.float64)
    if var == "precip":
        idx = t.index
        out["year"] = idx.year.values.astype(np.int16)
        out["month"] = idx.month.values.astype(np.int8)
        out["day"] = idx.day.values.astype(np.int8)
np.savez_compressed(os.path.join(HERE, "example_weather.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
print(os.path.getsize(os.path.join(HERE, "example_weather.npz")), "bytes")
