"""Developer timing of the drop-in text path: python scripts/dropin_timing.py [n_cells] [n_years] -- builds the reference's text
inputs (D10/D11 records, D4/D7/D13 files) for a natural-soil grid and runs pyhelp_b200.run_help_allcells on them under cProfile."""
import cProfile, contextlib, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from pyhelp_b200 import processing
from support import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bench.N_YEARS = ny
nn = max(1, n // 100)
g = synth.grid_natural_3layer(n, seed=1)
wx = synth.synth_weather(nn, bench.YEAR0, ny, seed=5)
node = synth.assign_nodes(n, nn)
wd = tempfile.mkdtemp(prefix="help_dropin_")
cp = bench.build_text_cellparams(g, wx, node, wd)
for rep in range(2):
    with contextlib.redirect_stdout(io.StringIO()):
        pr = cProfile.Profile(); pr.enable()
        t0 = time.perf_counter(); out = processing.run_help_allcells(cp); dt = time.perf_counter() - t0
        pr.disable()
    print("run %d: %.3f s  %s" % (rep, dt, {k: round(v, 3) for k, v in processing.LAST_TIMING.items()}), flush=True)
    del out
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
