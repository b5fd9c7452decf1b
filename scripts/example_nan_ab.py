"""Developer A/B: the example configuration with and without the cells HELP returns NaN for (scripts/example_timing.py)."""
import os, sys, time
import numpy as np
import pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import managers, processing
from support import synth
_sim = processing.HelpEngine.simulate
def simulate(self, *a, **k):
    r = _sim(self, *a, **k)
    print("   engine: %d cells  setup %.1f ms  simulate %.1f ms" % (self.batch.n_cells, self.ms_setup, self.ms_simulate), flush=True)
    return r
processing.HelpEngine.simulate = simulate
z = np.load(os.path.join(ROOT, "tests", "golden", "example_weather.npz"))
idx = pd.DatetimeIndex(pd.to_datetime({"year": z["year"].astype(int), "month": z["month"].astype(int), "day": z["day"].astype(int)}), name="date")
hm = managers.HelpManager()
for var in ("precip", "airtemp", "solrad"):
    cols = pd.MultiIndex.from_tuples(list(zip(z[var + "_lat"], z[var + "_lon"])), names=["lat_dd", "lon_dd"])
    setattr(hm, var + "_data", pd.DataFrame(z[var].astype(np.float64), index=idx, columns=cols))
g = synth.grid_example(19_200)
hm.grid = pd.DataFrame(g).set_index("cid", drop=False)
cells = hm.get_run_cellnames()[:17_178]
out = hm.calc_help_cells(cellnames=cells, tfsoil=-3, sf_edepth=0.15, sf_ulai=1, sf_cn=1.15)
bad = set(out.data["idx_nan"])
print("NaN cells", len(bad))
good = [c for k, c in enumerate(cells) if k not in bad]
for it in range(2):
    out2 = hm.calc_help_cells(cellnames=good, tfsoil=-3, sf_edepth=0.15, sf_ulai=1, sf_cn=1.15)
lt = hm.grid.loc[cells, [c for c in hm.grid.columns if c.startswith("lay_type")]].values
print("layer types present:", np.unique(lt, return_counts=True), "nlayer:", np.unique(hm.grid.loc[cells, "nlayer"].values, return_counts=True))
# the NaN cells on their own, the same number of good cells on their own, and one NaN cell among 31 good ones
badc = [c for k, c in enumerate(cells) if k in bad]
for label, sel in (("21 NaN cells", badc), ("21 good cells", good[:21]), ("1 NaN + 31 good", badc[:1] + good[:31]), ("32 good", good[:32])):
    print(label)
    hm.calc_help_cells(cellnames=sel, tfsoil=-3, sf_edepth=0.15, sf_ulai=1, sf_cn=1.15)
