"""Parity envelope of HELP against its own libm: the same oracle source built with
(det) the engine's deterministic transcendentals -- the bits the GPU returns --, (glibc) the host libm a gfortran
build of the reference links, (glibc_noise) glibc with 1 ulp of noise on 0.1 % of pow/log/powf/expf/logf results.
python scripts/parity_envelope.py [n_cells] [n_years]  ->  markdown table on stdout."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle
from support import synth

VARS = ("runoff", "evapo", "latdrain", "rechg")


def families(n, ny):
    yield "configs[0] stand-in (example-like grid, sf_edepth=.15, sf_cn=1.15)", synth.grid_example(n, seed=20260101), None, dict(sf_edepth=0.15, sf_cn=1.15), ny + 1
    yield "configs[1] natural 3-layer", synth.grid_natural_3layer(n, seed=1), None, {}, ny
    yield "configs[2] mixed classes", synth.grid_mixed(n, seed=2), None, {}, ny
    g, ex = synth.grid_landfill(n, seed=3)
    yield "configs[3] landfill covers", g, ex, {}, ny


def yearly(out, names):
    y = {}
    for k in ("runoff", "evapo", "rechg"):
        y[k] = np.stack([out[c][k].sum(axis=1) for c in names]).astype(np.float64)
    y["latdrain"] = np.stack([(out[c]["subrun1"] + out[c]["subrun2"]).sum(axis=1) for c in names]).astype(np.float64)
    y["precip"] = np.stack([out[c]["precip"].sum(axis=1) for c in names]).astype(np.float64)
    return y


def compare(a, b):
    """(fraction of cell-years within the north_star tolerance, max error of per-cell multi-year means relative
    to the cell's mean precipitation, relative error of the area mean) per variable"""
    res = {}
    for k in VARS:
        d = np.abs(a[k] - b[k])
        ok = d <= np.maximum(1e-3 * np.abs(b[k]), 0.1)
        cm = np.abs(a[k].mean(axis=1) - b[k].mean(axis=1))
        am = abs(a[k].mean() - b[k].mean())
        res[k] = (float(ok.mean()), float((cm / b["precip"].mean(axis=1)).max()), float(am / max(b[k].mean(), 1e-9)),
                  float(b[k].mean()))
    return res


def run(n=256, ny=10, wd="/tmp/help_envelope"):
    rows = []
    for title, g, ex, kw, years in families(n, ny):
        nn = 8
        wx = synth.synth_weather(nn, 2000, years, seed=5)
        files = synth.write_weather_files(os.path.join(wd, title.split()[0]), wx)
        ok = np.nonzero(np.asarray(g.get("run", np.ones(len(g["cid"])))) == 1)[0][:n].tolist() if "run" in g else None
        node = synth.assign_nodes(len(g["cid"]), nn)
        cp = synth.cellparams_from_grid(g, files, node, node, node, years, tfsoil_c=-3.0, ex=ex, cells=ok, **kw)
        names = list(cp.keys())
        out = {m: yearly(oracle.run_help_allcells(cp, libm=m), names) for m in ("det", "glibc", "glibc_noise")}
        nan = np.isnan(out["glibc"]["rechg"]).any(axis=1) | np.isnan(out["det"]["rechg"]).any(axis=1)
        if nan.any():
            out = {m: {k: v[~nan] for k, v in y.items()} for m, y in out.items()}
        rows.append((title, len(names) - int(nan.sum()), years, compare(out["det"], out["glibc"]),
                     compare(out["glibc_noise"], out["glibc"])))
    return rows


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print("| grid | pair | variable (mean mm/yr) | cell-years within 1e-3 rel or 0.1 mm/yr | worst cell multi-year mean, % of its precipitation | area mean, relative |")
    print("|---|---|---|---|---|---|")
    for title, nc, years, dg, ng in run(n, ny):
        for pair, res in (("engine arithmetic vs glibc", dg), ("glibc + 1 ulp on 0.1 % of calls vs glibc", ng)):
            for k in VARS:
                f, cm, am, mean = res[k]
                if mean < 1e-6:
                    continue
                print(f"| {title} ({nc} cells x {years} yr) | {pair} | {k} ({mean:.0f}) | {100 * f:.1f} % | {100 * cm:.2f} % | {am:.1e} |")
