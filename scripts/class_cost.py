"""Developer calibration: simulation time of each profile class of the BASELINE configs[2] grid on its own (one GPU, full
occupancy), next to the host-side cost estimate of pyhelp_b200.sharding -- the numbers COST_* in sharding.py come from.
usage: python scripts/class_cost.py [cells per class] [years]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, processing, sharding
from support import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 3
big = synth.grid_mixed(12 * n, seed=2)
nn = max(1, n // 100)
wx = synth.synth_weather(nn, 2001, ny, seed=5)
node = synth.assign_nodes(n, nn)
for c in range(4):
    idx = np.nonzero(np.asarray(big["class"]) == c)[0][:n]
    g = {k: np.asarray(v)[idx] for k, v in big.items()}
    m = idx.size
    cost, klass = sharding.grid_cost_and_class(g)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node[:m], node[:m], node[:m], tfsoil_c=-3.0)
    with processing.HelpEngine(b, monthly=True) as eng:
        eng.simulate(); eng.simulate()
        ms = eng.ms_simulate
        topo = eng.fetch(topo=True)["topo"]
    nseg = topo[0].astype(np.int64)
    work = sum(topo[1 + k].astype(np.int64) * topo[7 + k].astype(np.int64) for k in range(6)) if topo.shape[0] >= 13 else None
    print("class %d: %d cells x %d yr  %.1f ms  -> %.0f cell-years/s ; us per cell-year %.3f ; est cost/cell %.2f (classes %s) ; nseg mean %.2f ; sum nseg*nstep mean %s"
          % (c, m, ny, ms, m * ny / ms * 1e3, ms * 1e3 / (m * ny), cost.mean(), np.unique(klass).tolist(), nseg.mean(),
             ("%.1f" % work.mean()) if work is not None else "n/a"), flush=True)
