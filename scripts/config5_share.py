"""One GPU's share of BASELINE config 5 (province scale: 5 M cells x 30 yr over 8 GPUs, mixed profile
classes, one weather node per 100 cells, yearly summaries only): 625 000 cells on one device.
python scripts/config5_share.py [n_cells] [n_years]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, processing, _native as N
from support import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 625000
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 30
nn = max(1, n // 100)
t = time.time()
g, ex = synth.grid_mixed(n, seed=4) if isinstance(synth.grid_mixed(8, seed=4), tuple) else (synth.grid_mixed(n, seed=4), None)
wx = synth.synth_weather(nn, 2001, ny, seed=5)
node = synth.assign_nodes(n, nn)
b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0, extra_layers=ex)
print("pack %.1f s, host batch %.2f GB" % (time.time() - t, b.nbytes() / 1e9), flush=True)
with processing.HelpEngine(b, monthly=False) as eng:
    eng.simulate()
    print("setup %.1f ms  simulate %.1f ms  -> %.0f cell-years/s" % (eng.ms_setup, eng.ms_simulate, n * ny / (eng.ms_simulate + eng.ms_setup) * 1e3), flush=True)
    r = eng.fetch(monthly=False, yearly=True)
    bad = int((r["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL | N.FLAG_OVERFLOW)).astype(bool).sum())
    y = r["yearly"]
    print("flagged cells", bad, " yearly mean mm (precip, runoff, et, latdrain, rechg):", np.round(y.mean(axis=(1, 2)), 2))
    # water balance of the area mean over 30 years: P = RO + ET + lateral drainage + recharge + storage change
    resid = y[0].sum() - y[1].sum() - y[2].sum() - y[3].sum() - y[4].sum()
    print("area water-balance residual %.3f mm per cell-year (storage change + print quantisation)" % (resid / (n * ny)))
