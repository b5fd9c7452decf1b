"""DEVELOPER AID: run the kernels through the g++ shim and compare with the oracle bit for bit
(RCRA + the three synthetic grid families).  python scripts/host_emu/check_emu.py [n] [ny]"""
import os, sys, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
subprocess.check_call("g++ -O2 -mfma -ffp-contract=off -std=c++17 -shared -fPIC %s emu.cpp -o libemu.so" % os.environ.get("EMU_FLAGS", ""), shell=True, cwd=HERE)
from pyhelp_b200 import inputs
from support import synth
from oracle import oracle
import run_emu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2
def check(tag, cp):
    b = inputs.batch_from_cellparams(cp)
    r = run_emu.emu_run(b)
    names = list(cp.keys()); worst = 0.0; nbad = 0
    for i, nm in enumerate(names):
        o = oracle.run_simulation(*cp[nm], full=True, daily=True)
        d = np.abs(r["raw"][:, :, :, i] - np.transpose(o["monthly_in"], (1, 0, 2))).max()
        par = np.stack([o[k] for k in oracle.KEYS_PARSED])
        same = np.array_equal(r["monthly"][:, :, :, i], par, equal_nan=True)
        same = same and np.array_equal(r["daily"][i], o["daily"])       # REAL*8 daily columns, bit for bit
        worst = max(worst, d); nbad += (d != 0) or (not same)
    print("%-10s cells %4d  differing cells %d  max raw diff %.3g in" % (tag, len(names), nbad, worst))
    return nbad
ONLY = os.environ.get("EMU_ONLY")
R = os.path.join(ROOT, "tests/golden/rcra/")
d10 = np.char.ljust(open(R + "RCRA.D10").read().splitlines(), 80).astype("S80")
d11 = np.char.ljust(open(R + "RCRA.D11").read().splitlines(), 80).astype("S80")
bad = 0 if ONLY else check("rcra", {"rcra": (R + "RCRA.D4", R + "RCRA.D7", R + "RCRA.D13", d11, d10, "/tmp/x.OUT", 0, 1, 0, 0, 3, 32.0)})
for name, maker in (("natural3", synth.grid_natural_3layer), ("mixed", synth.grid_mixed), ("landfill", synth.grid_landfill),
                    ("recirc", lambda n, seed: synth.grid_landfill(n, seed=seed, recirc=True)), ("deep", synth.grid_deep),
                    ("special", synth.grid_special)):
    if ONLY and name != ONLY: continue
    res = maker(n, seed=11); g, ex = res if isinstance(res, tuple) else (res, None)
    nn = max(2, n // 16)
    wx = synth.synth_weather(nn, 2001, ny, seed=5)
    files = synth.write_weather_files("/tmp/help_b200_dev/" + name, wx)
    node = synth.assign_nodes(n, nn)
    bad += check(name, synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex))
sys.exit(1 if bad else 0)
