"""Developer calibration of the multi-GPU partition on ONE GPU: plan the shards of the BASELINE configs[2] grid for
`world` ranks with pyhelp_b200.sharding (class-sorted list dealt out, or cut into runs of equal estimated cost) and simulate them one after the other, printing the
kernel time of every shard -- what the ranks of `bench.py --gpus world` would each take.
usage: python scripts/shard_balance.py world [years] [cells] [deal|contiguous]        (HELP_B200_COST_LINER overrides the liner factor)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyhelp_b200 import grid as G, processing, sharding
from support import synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
nn = max(1, n // 100)
mode = sys.argv[4] if len(sys.argv) > 4 else "deal"
g = synth.grid_mixed(n, seed=2)
wx = synth.synth_weather_chunked(nn, 2001, ny, seed=2, chunk=500)
node = synth.assign_nodes(n, nn)
cost, klass = sharding.grid_cost_and_class(g)
order, ranges = sharding.plan_shards(cost, klass, world, node, mode=mode)
ms = []
for r, (a, b) in enumerate(ranges):
    idx = order[a:b]
    sub = {k: np.asarray(v)[idx] for k, v in g.items()}
    used, inv = np.unique(node[idx], return_inverse=True)
    tabs = [np.ascontiguousarray(wx[k][:, used]) for k in ("precip", "airtemp", "solrad")]
    nloc = inv.astype(np.int32)
    batch = G.batch_from_grid(sub, wx["years_of_day"], tabs[0], tabs[1], tabs[2], nloc, nloc, nloc, tfsoil_c=-3.0, weather="resident")
    with processing.HelpEngine(batch, monthly=True) as eng:
        eng.simulate()
        ms.append(eng.ms_simulate)
    print("rank %d: %7d cells  classes %s  est cost %.0f  kernel %.1f ms" % (
        r, idx.size, dict(zip(*[x.tolist() for x in np.unique(np.asarray(g["class"])[idx], return_counts=True)])), cost[idx].sum(), ms[-1]), flush=True)
ms = np.array(ms)
print("world %d  mode %s  liner factor %.3f  max %.1f ms  mean %.1f ms  imbalance %.3f  ->  %d cells x %d yr in %.2f s of kernel time" % (
    world, mode, sharding.COST_LINER, ms.max(), ms.mean(), ms.max() / ms.mean(), n, ny, ms.max() / 1e3))
