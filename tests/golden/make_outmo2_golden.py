"""Make tests/golden/outmo2/: .OUT texts written by the oracle (oracle/out_writer.py) and what the REFERENCE's
own parser -- read_monthly_help_output, pyhelp/processing.py:64-147, executed from /root/reference without
importing the package (its __init__ needs the compiled HELP3O, h5py and matplotlib) -- returns for them.
Run here (the reference checkout is not on the GPU box):  python tests/golden/make_outmo2_golden.py"""
import ast
import os
import os.path as osp
import sys

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, out_writer          # noqa: E402
from support import synth                      # noqa: E402

OUT = osp.join(ROOT, "tests", "golden", "outmo2")


def reference_parser(ref_root="/root/reference"):
    """The reference's function object, compiled from its own source file."""
    src = open(osp.join(ref_root, "pyhelp", "processing.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "read_monthly_help_output"][0]
    ns = {"np": np, "csv": __import__("csv")}
    exec(compile(ast.Module([fn], []), "pyhelp/processing.py", "exec"), ns)
    return ns["read_monthly_help_output"]


def cases(workdir):
    """name -> cellparams 12-tuple"""
    R = osp.join(ROOT, "tests", "golden", "rcra")
    d10 = np.char.ljust(open(osp.join(R, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(R, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    out = {"rcra": (osp.join(R, "RCRA.D4"), osp.join(R, "RCRA.D7"), osp.join(R, "RCRA.D13"), d11, d10, "x", 0, 1, 0, 0, 3, 32.0)}
    wx = synth.synth_weather(2, 2003, 2, seed=9)
    files = synth.write_weather_files(workdir, wx)
    node = np.zeros(16, np.int32)
    g = synth.grid_natural_3layer(4, seed=1)
    out["natural"] = synth.cellparams_from_grid(g, files, node, node, node, 2, tfsoil_c=-3.0)["1"]
    g, ex = synth.grid_special(16, seed=8)
    cp = synth.cellparams_from_grid(g, files, node, node, node, 2, tfsoil_c=-3.0, ex=ex)
    out["geotextile"], out["subin_liner"], out["subin_natural"] = cp["0"], cp["3"], cp["5"]
    g, ex = synth.grid_landfill(8, seed=3)
    out["two_liners"] = [v for k, v in synth.cellparams_from_grid(g, files, node, node, node, 2, tfsoil_c=-3.0, ex=ex).items()
                         if g["nlayer"][int(k)] == 8][0]
    # six subprofiles: five drain-over-barrier systems and a soil layer below (the Jul-Dec percolation of the sixth is
    # printed without the x 25.4, HELP3O.FOR:6593-6594)
    n = 1
    rng = np.random.default_rng(5)
    g6 = synth._base_grid(n, rng)
    seq = [1, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 1]
    g6["nlayer"] = np.full(n, len(seq))
    for j, t in enumerate(seq, 1):
        synth._soil_layer(g6, j, n, rng, t)
        s = str(j)
        if t == 2:
            g6["thick" + s] = np.full(n, 30.0); g6["poro" + s] = np.full(n, 0.417); g6["fc" + s] = np.full(n, 0.045)
            g6["wp" + s] = np.full(n, 0.018); g6["ksat" + s] = np.full(n, 0.01)
        if t == 3:
            g6["thick" + s] = np.full(n, 30.0); g6["poro" + s] = np.full(n, 0.427); g6["fc" + s] = np.full(n, 0.418)
            g6["wp" + s] = np.full(n, 0.367); g6["ksat" + s] = np.full(n, 3e-6)
    out["six_subprofiles"] = synth.cellparams_from_grid(g6, files, node, node, node, 2, tfsoil_c=-3.0)["0"]
    # a month with more than 1000 mm of precipitation: F7.3 overflows to '*******'
    wet = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in wx.items()}
    wet["precip"] = wet["precip"].copy()
    wet["precip"][180:211] = 40.0
    files_wet = synth.write_weather_files(osp.join(workdir, "wet"), wet)
    g = synth.grid_natural_3layer(2, seed=2)
    out["overflow"] = synth.cellparams_from_grid(g, files_wet, node, node, node, 2, tfsoil_c=-3.0)["0"]
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    parse = reference_parser()
    exp = {}
    for name, p in cases("/tmp/outmo2_golden").items():
        res = oracle.run_simulation(*p, full=True)
        path = out_writer.write_out(osp.join(OUT, name + ".OUT"), res, p[4])
        try:
            d = parse(path)
            for k, v in d.items():
                exp[name + "/" + k] = v
            print(name, "parsed by the reference:", {k: v.shape for k, v in d.items()})
        except ValueError as e:
            exp[name + "/error"] = np.array(str(e))
            print(name, "the reference's parser raises:", e)
    np.savez_compressed(osp.join(OUT, "expected.npz"), **exp)


if __name__ == "__main__":
    main()
