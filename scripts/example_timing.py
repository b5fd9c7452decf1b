"""BASELINE configs[0] timing: the reference's published run is 17 178 cells x 11 yr (2000-2010) of the
Riviere du Nord example in 388.95 s on an i7-7700HQ (docs/example.rst:92-93 = 486 cell-years/s).
Same call here -- HelpManager.calc_help_cells(cellnames, tfsoil=-3, sf_edepth=0.15, sf_cn=1.15) with the
example weather tables and a schema-exact stand-in grid (synth.grid_example; the example's grid blob is
missing from the checkout) -- wall time of the whole call: nearest nodes, packing, upload, simulation,
download, post-processing.      python scripts/example_timing.py"""
import os, sys, time
import numpy as np
import pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import managers
from support import synth

z = np.load(os.path.join(ROOT, "tests", "golden", "example_weather.npz"))
idx = pd.DatetimeIndex(pd.to_datetime({"year": z["year"].astype(int), "month": z["month"].astype(int), "day": z["day"].astype(int)}), name="date")
hm = managers.HelpManager()
for var in ("precip", "airtemp", "solrad"):
    cols = pd.MultiIndex.from_tuples(list(zip(z[var + "_lat"], z[var + "_lon"])), names=["lat_dd", "lon_dd"])
    setattr(hm, var + "_data", pd.DataFrame(z[var].astype(np.float64), index=idx, columns=cols))
g = synth.grid_example(19_200)
hm.grid = pd.DataFrame(g).set_index("cid", drop=False)
cells = hm.get_run_cellnames()[:17_178]
print("runnable cells", len(cells), flush=True)
for it in range(3):
    t = time.perf_counter()
    out = hm.calc_help_cells(cellnames=cells, tfsoil=-3, sf_edepth=0.15, sf_ulai=1, sf_cn=1.15)
    dt = time.perf_counter() - t
    n, ny = len(out.data["cid"]), len(out.data["years"])
    print("calc_help_cells: %d cells x %d yr in %.3f s -> %.0f cell-years/s (reference, published: 388.95 s, 486 cell-years/s); NaN cells %d"
          % (n, ny, dt, n * ny / dt, len(out.data["idx_nan"])), flush=True)
avg = out.calc_area_yearly_avg()
print("area yearly means (mm):", {k: round(float(v.mean()), 1) for k, v in avg.items()})
