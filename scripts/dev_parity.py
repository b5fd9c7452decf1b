"""Developer check: GPU engine vs CPU oracle on RCRA + small synthetic grids.
Usage: python scripts/dev_parity.py [n_cells] [n_years]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, inputs, processing
from support import synth
from oracle import oracle

def cmp(tag, gpu, orc, names):
    keys = processing.KEYS
    worst = {}
    for k in keys:
        a = np.stack([gpu[n][k] for n in names]).astype(np.float64)
        b = np.stack([orc[n][k] for n in names]).astype(np.float64)
        ya, yb = a.sum(axis=2), b.sum(axis=2)
        dy = np.abs(ya - yb)
        rel = dy / np.maximum(np.abs(yb), 1e-9)
        bad = (dy > 0.1) & (rel > 1e-3)
        worst[k] = (np.abs(a - b).max(), dy.max(), int(bad.sum()), float((a == b).mean()))
    print(tag)
    for k, v in worst.items():
        print("  %-8s monthly max|d| %.4f mm  yearly max|d| %.4f mm  yearly-out-of-tol %d  bit-equal months %.4f" % (k, *v))

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    R = os.path.join(ROOT, "tests/golden/rcra/")
    d10 = np.char.ljust(open(R + "RCRA.D10").read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(R + "RCRA.D11").read().splitlines(), 80).astype("S80")
    cp = {"rcra": (R + "RCRA.D4", R + "RCRA.D7", R + "RCRA.D13", d11, d10, "/tmp/x.OUT", 0, 1, 0, 0, 3, 32.0)}
    t = time.time(); gpu = processing.run_help_allcells(cp); print("gpu rcra", time.time() - t)
    orc = oracle.run_help_allcells(cp, ncore=1)
    cmp("RCRA", gpu, orc, ["rcra"])
    print(" gpu yearly sums (in):", {k: (gpu["rcra"][k].sum(axis=1) * 0.0393701).round(4).tolist() for k in processing.KEYS})
    wd = "/tmp/help_b200_dev"
    for name, maker in (("natural3", synth.grid_natural_3layer), ("mixed", synth.grid_mixed), ("landfill", synth.grid_landfill)):
        res = maker(n, seed=11)
        g, ex = res if isinstance(res, tuple) else (res, None)
        nn = max(2, n // 16)
        wx = synth.synth_weather(nn, 2001, ny, seed=5)
        files = synth.write_weather_files(os.path.join(wd, name), wx)
        node = synth.assign_nodes(n, nn)
        cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex)
        t = time.time(); gpu = processing.run_help_allcells(cp); tg = time.time() - t
        t = time.time(); orc = oracle.run_help_allcells(cp); to = time.time() - t
        print("%s: gpu %.2fs oracle %.2fs" % (name, tg, to))
        cmp(name, gpu, orc, list(cp.keys()))
        b = inputs.batch_from_cellparams(cp)
        r = processing.run_batch(b, topo=True)
        print("  flags:", np.unique(r["flags"], return_counts=True), "nseg:", np.unique(r["topo"][0], return_counts=True)[0][:10], r["timing_ms"])

if __name__ == "__main__":
    main()
