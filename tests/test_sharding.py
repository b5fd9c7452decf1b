"""Host-side logic of the multi-GPU path, on CPU: shard arithmetic, weather-node compaction,
and the gather of yearly summaries with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyhelp_b200 import grid as G
from pyhelp_b200 import sharding
from support import synth
from pyhelp_b200 import _native as N


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 100_001):
        for w in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def make_batch(n=40, nn=8, ny=1):
    g = synth.grid_natural_3layer(n, seed=3)
    wx = synth.synth_weather(nn, 2001, ny, seed=4)
    node = synth.assign_nodes(n, nn)
    return G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)


def test_shard_batch_compacts_weather_nodes():
    b = make_batch()
    parts = [sharding.shard_batch(b, 4, r) for r in range(4)]
    assert sum(p.n_cells for p in parts) == b.n_cells
    for r, p in enumerate(parts):
        a, e = sharding.shard_range(b.n_cells, 4, r)
        assert p.pre.shape[0] == 2 and p.tmp.shape[0] == 2 and p.rad.shape[0] == 2      # 8 nodes / 4 ranks
        # every cell still points at ITS weather rows
        for i in range(p.n_cells):
            assert np.array_equal(p.pre[p.cell_i32[N.HB_NODE_P, i]], b.pre[b.cell_i32[N.HB_NODE_P, a + i]])
            assert np.array_equal(p.rad[p.cell_i32[N.HB_NODE_R, i]], b.rad[b.cell_i32[N.HB_NODE_R, a + i]])
            assert np.array_equal(p.idfs[p.cell_i32[N.HB_NODE_R, i]], b.idfs[b.cell_i32[N.HB_NODE_R, a + i]])
        assert np.array_equal(p.cell_f32, b.cell_f32[:, a:e]) and p.names == b.names[a:e]
        p.as_struct()          # all arrays contiguous


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_total, world, rank)
    # a recognisable payload: value = global cell index + 1000*row + 10*year
    idx = torch.arange(a, b, dtype=torch.float32)
    y = idx[None, None, :] + 1000.0 * torch.arange(5)[:, None, None] + 10.0 * torch.arange(3)[None, :, None]
    out = sharding.gather_yearly(y, n_total, world, rank)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [11, 64])
def test_gather_yearly_gloo_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = (np.arange(n_total, dtype=np.float32)[None, None, :] + 1000.0 * np.arange(5)[:, None, None]
            + 10.0 * np.arange(3)[None, :, None])
    assert got.shape == (5, 3, n_total) and np.array_equal(got, want)


def test_cost_balanced_ranges():
    # equal costs: the equal-count split (up to one cell)
    r = sharding.cost_balanced_ranges(np.ones(10), 3)
    assert r[0][0] == 0 and r[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert sorted(b - a for a, b in r) in ([3, 3, 4], [3, 4, 3])
    # expensive cells clustered at the end: the last rank gets fewer cells, every rank ~ the same cost
    cost = np.concatenate([np.ones(900), 5 * np.ones(100)])
    r = sharding.cost_balanced_ranges(cost, 4)
    tot = [cost[a:b].sum() for a, b in r]
    assert max(tot) - min(tot) <= 5.0 and (r[3][1] - r[3][0]) < (r[0][1] - r[0][0])
    assert sharding.cost_balanced_ranges([], 2) == [(0, 0), (0, 0)]
    assert sharding.cost_balanced_ranges([1.0], 3)[-1][1] == 1


def test_estimated_cost_ranks_liner_profiles_above_natural_soil():
    from pyhelp_b200 import grid as G
    from support import synth
    n = 64
    g, ex = synth.grid_landfill(n, seed=3)
    wx = synth.synth_weather(2, 2001, 1, seed=1)
    node = synth.assign_nodes(n, 2)
    bl = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, extra_layers=ex)
    bn = G.batch_from_grid(synth.grid_natural_3layer(n, seed=3), wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"],
                           node, node, node)
    cl, cn = sharding.estimate_cost(bl), sharding.estimate_cost(bn)
    assert cl.shape == cn.shape == (n,) and cl.min() > cn.max() and (cn >= 8).all() and (cn <= 16).all()
    sh = sharding.shard_batch(bn, 2, 1, by_cost=True)
    assert abs(sh.n_cells - n // 2) <= 6


def test_run_batch_devices_joins_shards_in_caller_order(monkeypatch):
    """Host logic of the single-process multi-GPU fan-out (processing.run_batch_devices), with the
    engine stubbed: every shard runs on its own device in its own thread, only with the weather nodes it
    references, and the results come back in the caller's cell order."""
    import threading
    from pyhelp_b200 import processing
    from pyhelp_b200 import _native as N
    n, nn, ny = 53, 7, 2
    g = synth.grid_mixed(n, seed=4)
    wx = synth.synth_weather(nn, 2001, ny, seed=2)
    node = synth.assign_nodes(n, nn)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)
    seen = {}
    together = threading.Barrier(3, timeout=30)           # the three shards must be in flight at the same time

    def fake_set_device(dev):
        seen.setdefault(threading.get_ident(), []).append(dev)

    def fake_run_batch(sub, monthly=True, raw=False, topo=False):
        # a value that identifies the cell: its CN2 and the first precipitation value of ITS node (after compaction)
        tag = sub.cell_f32[N.HB_CN2] + sub.pre[sub.cell_i32[N.HB_NODE_P], 0, 0]
        m = np.broadcast_to(tag.astype(np.float32), (N.HR_NROWS, sub.n_years, 12, sub.n_cells)).copy()
        y = np.broadcast_to(tag.astype(np.float32), (N.HY_NROWS, sub.n_years, sub.n_cells)).copy()
        assert sub.pre.shape[0] == np.unique(sub.cell_i32[N.HB_NODE_P]).size          # compacted weather
        if together is not None:
            together.wait()
        return {"years": sub.years.astype(np.uint16), "flags": np.zeros(sub.n_cells, np.uint32), "monthly": m, "yearly": y,
                "timing_ms": {"h2d": 1.0, "setup": 1.0, "simulate": float(sub.n_cells), "d2h": 1.0}, "n_launches": 4}

    monkeypatch.setattr(N, "set_device", fake_set_device)
    monkeypatch.setattr(processing, "run_batch", fake_run_batch)
    r = processing.run_batch_devices(b, 3)
    want = b.cell_f32[N.HB_CN2] + b.pre[b.cell_i32[N.HB_NODE_P], 0, 0]
    assert r["monthly"].shape == (N.HR_NROWS, ny, 12, n) and r["yearly"].shape == (N.HY_NROWS, ny, n)
    assert np.array_equal(r["monthly"][0, 0, 0], want.astype(np.float32)) and np.array_equal(r["yearly"][2, 1], want.astype(np.float32))
    assert r["flags"].shape == (n,) and r["n_launches"] == 12
    assert sorted(d for v in seen.values() for d in v) == [0, 1, 2] and len(seen) == 3          # one thread per device
    together = None
    # one device: the plain path; more devices than cells: clipped
    assert processing.run_batch_devices(b, 1)["monthly"].shape[-1] == n
    tiny = b.take(np.arange(2))
    assert processing.run_batch_devices(tiny, 8)["yearly"].shape[-1] == 2
    # an error on one device surfaces in the caller
    def boom(sub, monthly=True, raw=False, topo=False):
        raise N.EngineError("device lost")
    monkeypatch.setattr(processing, "run_batch", boom)
    with pytest.raises(N.EngineError):
        processing.run_batch_devices(b, 2)


def test_grid_plan_deals_every_class_evenly_and_contiguous_plan_is_cost_balanced():
    """sharding.grid_cost_and_class / plan_shards: one grid, class-sorted, dealt out over the ranks (every rank the
    same share of every class: SURVEY.md 8e) -- from the grid columns alone, and equal to the cost the packed batch
    gives; the contiguous variant cuts runs of equal estimated cost."""
    n, world = 4000, 4
    g = synth.grid_mixed(n, seed=2)
    node = synth.assign_nodes(n, 40)
    cost, klass = sharding.grid_cost_and_class(g)
    wx = synth.synth_weather(40, 2001, 1, seed=1)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)
    assert np.array_equal(cost, sharding.estimate_cost(b))
    order, ranges = sharding.plan_shards(cost, klass, world, node)
    assert np.array_equal(np.sort(order), np.arange(n))
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    sizes = [b_ - a for a, b_ in ranges]
    assert max(sizes) - min(sizes) <= 1
    for k in np.unique(klass):                                             # every class is shared out evenly ...
        per = [int((klass[order[a:b_]] == k).sum()) for a, b_ in ranges]
        assert max(per) - min(per) <= 1
    for a, b_ in ranges:                                                   # ... and stays sorted inside a rank
        assert (np.diff(klass[order[a:b_]]) >= 0).all()
    tot = np.array([cost[order[a:b_]].sum() for a, b_ in ranges])
    assert tot.max() / tot.mean() < 1.005
    r = sharding.restore_order(np.arange(n, dtype=np.float32)[order][None, :], order)
    assert np.array_equal(r[0], np.arange(n, dtype=np.float32))
    # contiguous runs of equal estimated cost
    order, ranges = sharding.plan_shards(cost, klass, world, node, mode="contiguous")
    assert np.array_equal(np.sort(order), np.arange(n))
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert (np.diff(klass[order]) >= 0).all()                              # class-major
    tot = np.array([cost[order[a:b_]].sum() for a, b_ in ranges])
    assert tot.max() / tot.mean() < 1.01
    a, b_ = ranges[0]                                                      # plain profiles of at most 16 segments first
    lt = np.stack([g[f"lay_type{j}"] for j in range(1, 6)])
    assert (lt[:, order[a:b_]] < 3).all() and g["nlayer"][order[a:b_]].max() <= 3
    assert (ranges[-1][1] - ranges[-1][0]) < (ranges[0][1] - ranges[0][0])  # liner profiles cost more: fewer cells per rank
    with pytest.raises(ValueError):
        sharding.plan_shards(cost, klass, world, node, mode="x")


def _plan_worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.grid_mixed(n_total, seed=5)
    node = synth.assign_nodes(n_total, 9)
    cost, klass = sharding.grid_cost_and_class(g)
    order, ranges = sharding.plan_shards(cost, klass, world, node)
    a, b = ranges[rank]
    idx = torch.as_tensor(order[a:b].astype(np.float32))                    # "result" of a cell = its caller index
    y = idx[None, None, :] + 1000.0 * torch.arange(5)[:, None, None] + 10.0 * torch.arange(2)[None, :, None]
    full = sharding.gather_yearly(y, n_total, world, rank, ranges=ranges)
    if rank == 0:
        q.put((sharding.restore_order(full, order).numpy(), ranges))
    dist.barrier()
    dist.destroy_process_group()


def test_planned_shards_gather_to_caller_order_gloo_world2():
    """the N>1 path of bench.py's `target` on CPU: unequal shards of a class-sorted grid, gathered over gloo and
    put back in the caller's cell order"""
    n_total = 301
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (ranges[0][1] - ranges[0][0]) != (ranges[1][1] - ranges[1][0])      # 301 cells: 151 + 150
    want = (np.arange(n_total, dtype=np.float32)[None, None, :] + 1000.0 * np.arange(5)[:, None, None]
            + 10.0 * np.arange(2)[None, :, None])
    assert np.array_equal(got, want)
