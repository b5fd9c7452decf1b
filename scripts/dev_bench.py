"""Developer timing: python scripts/dev_bench.py [n_cells] [n_years] [n_nodes] [kind]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, processing
from support import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 30
nn = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, n // 100)
kind = sys.argv[4] if len(sys.argv) > 4 else "natural3"
t = time.time()
res = {"natural3": synth.grid_natural_3layer, "mixed": synth.grid_mixed, "landfill": synth.grid_landfill}[kind](n, seed=1)
g, ex = res if isinstance(res, tuple) else (res, None)
wx = synth.synth_weather(nn, 2001, ny, seed=5)
node = synth.assign_nodes(n, nn)
b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                      tfsoil_c=-3.0, extra_layers=ex)
print("pack %.2fs" % (time.time() - t), flush=True)
with processing.HelpEngine(b, monthly=True) as eng:
    for it in range(3):
        eng.simulate()
        print("setup %.2f ms  simulate %.2f ms  -> %.0f cell-years/s" % (eng.ms_setup, eng.ms_simulate, n * ny / (eng.ms_simulate + eng.ms_setup) * 1e3), flush=True)
    r = eng.fetch(topo=True)
    print("flags", np.unique(r["flags"], return_counts=True), "timing", r["timing_ms"])
    print("yearly mean (mm):", r["yearly"].mean(axis=(1, 2)))
    print("nseg hist", np.unique(r["topo"][0], return_counts=True))
    print("nstep1 hist", np.unique(r["topo"][7], return_counts=True))
