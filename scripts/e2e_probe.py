import sys, time, numpy as np
sys.path.insert(0, ".")
import bench, torch
from pyhelp_b200 import _native as N
g, wx, node = bench.make_workload(0)
batch = bench.pin_batch(bench.make_batch(g, wx, node))
n, ny = batch.n_cells, batch.n_years
out_m = torch.empty((N.HR_NROWS, ny, 12, n), dtype=torch.float32).pin_memory()
out_y = torch.empty((N.HY_NROWS, ny, n), dtype=torch.float32).pin_memory()
out_f = torch.empty((n,), dtype=torch.int32).pin_memory()
for it in range(3):
    r = N.Result(); r.monthly, r.yearly, r.flags = out_m.data_ptr(), out_y.data_ptr(), out_f.data_ptr()
    s = batch.as_struct(); t0 = time.perf_counter()
    N.check(N.lib().help_b200_run(s, r), "run"); dt = time.perf_counter() - t0
    print("wall %.1f ms | h2d %.1f setup %.1f sim %.1f d2h %.1f | other %.1f" % (dt*1e3, r.ms_h2d, r.ms_setup, r.ms_simulate, r.ms_d2h, dt*1e3 - r.ms_h2d - r.ms_setup - r.ms_simulate - r.ms_d2h))
