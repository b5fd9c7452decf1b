"""Developer timing: kernel / set-up milliseconds of the example configuration's engine run (scripts/example_timing.py)."""
import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import processing
_sim = processing.HelpEngine.simulate
def simulate(self, *a, **k):
    r = _sim(self, *a, **k)
    print("   engine: setup %.1f ms  simulate %.1f ms  (h2d %.1f ms), narrow %s of %s" % (
        self.ms_setup, self.ms_simulate, getattr(self, "ms_h2d", float("nan")), getattr(self, "n_narrow", "?"), self.batch.n_cells), flush=True)
    return r
processing.HelpEngine.simulate = simulate
runpy.run_path(os.path.join(ROOT, "scripts", "example_timing.py"), run_name="__main__")
