import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import inputs, processing
from support import synth
from oracle import oracle
np.set_printoptions(linewidth=200, precision=4, suppress=True)
n, ny = 16, 2
g = synth.grid_natural_3layer(n, seed=11)
wx = synth.synth_weather(2, 2001, ny, seed=5)
files = synth.write_weather_files("/tmp/help_b200_dev/dbg", wx)
node = synth.assign_nodes(n, 2)
cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0)
b = inputs.batch_from_cellparams(cp)
r = processing.run_batch(b, raw=True, topo=True)
names = list(cp.keys())
worst, wi = 0, 0
for i, nm in enumerate(names):
    o = oracle.run_simulation(*cp[nm], full=True)
    raw_g = r["raw"][:, :, :, i]            # [14][ny][12]
    raw_o = np.transpose(o["monthly_in"], (1, 0, 2))
    d = np.abs(raw_g - raw_o).max()
    t = r["topo"][:, i]
    ok_topo = (t[0] == o["info"]["nseg"] and t[13] == o["info"]["idfs"] and list(t[1:7]) == o["info"]["nsegn"] and list(t[7:13]) == o["info"]["nstep"])
    print(nm, "max raw diff (in)", d, "topo ok", ok_topo, t[0], o["info"]["nseg"], t[13], o["info"]["idfs"])
    if d > worst: worst, wi = d, i
nm = names[wi]
o = oracle.run_simulation(*cp[nm], full=True)
print("worst cell", nm)
for row, lab in ((0, "pre"), (1, "run"), (2, "et"), (8, "prc1")):
    print(lab, "gpu", r["raw"][row, 0, :, wi]); print(lab, "orc", o["monthly_in"][0, row])
