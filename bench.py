#!/usr/bin/env python
"""Throughput of the HELP hot path in cell-years/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic input: the whole daily HELP
water balance (spin-up year + 30 output years) of every cell of the workload.

Headline workload = BASELINE.json configs[1]: 100 000 cells x 30 yr of 3-layer natural-soil
profiles with 1 000 weather nodes, per GPU.  For N>1 (torchrun, one rank per GPU) every rank
simulates its own 100 000-cell shard (weak scaling; cells are independent, no data-path
collective) and the per-cell yearly summaries are gathered to rank 0 over NCCL inside the step.

One JSON line (rank 0):
  value          device-resident throughput (inputs already in HBM when the timed region starts)
  e2e            the same through the public API from the caller's tables: pack (quantise the grid
                 and weather like the reference's text writers) -> H2D -> kernels -> D2H (and, at
                 N>1, the NCCL gather + rank-0 assembly), every step
  e2e_dropin     (N=1) `run_help_allcells(cellparams)`, the reference's own call, on a 100 000-cell
                 text `cellparams` (D10/D11 records + D4/D7/D13 files)
  other_configs  (N=1) configs[2] and [3] at the size one GPU carries: 125 000 mixed-class cells x
                 30 yr and 31 250 landfill covers x 50 yr, each with its own roofline
  target         (N>1) ONE configs[2] grid -- 1 M mixed-class cells, 10 000 weather nodes, 30 yr --
                 partitioned by pyhelp_b200.sharding (class-sorted list dealt out over the ranks), simulated, yearly
                 summaries NCCL-gathered to rank 0 in the caller's order: wall time, per-rank
                 kernel times, imbalance
  roofline, cpu_baseline, clocks, gpu_launches as the contract asks.

`--impl reference` times the reference's CPU algorithm instead: the reference Fortran cannot be
built in this image (no gfortran), so this arm runs the oracle port (oracle/, glibc libm build)
fanned out over all host cores like pyhelp/processing.py:40-59 does, on a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import os.path as osp
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = osp.dirname(osp.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CELLS = 100_000          # per GPU
N_YEARS = 30
N_NODES = 1_000            # per GPU
YEAR0 = 2001
WORKLOAD = "BASELINE configs[1]: synthetic 100k cells x 30 yr, 3-layer natural-soil profiles, 1000 weather nodes, per GPU"
METRIC = "cell_years_per_s"
TARGET = {"cells": 1_000_000, "nodes": 10_000, "years": 30, "seed": 2}


def config_dict(world: int) -> dict:
    """The same keys in both arms."""
    return {"workload": WORKLOAD, "n_cells_per_gpu": N_CELLS, "n_years": N_YEARS, "n_nodes_per_gpu": N_NODES,
            "simulated_years_per_cell": N_YEARS + 1,
            "l2": "no flush needed: each step reads 132 MB of weather and writes 1.0 GB of monthly results (> 126 MB L2)",
            "multi_gpu": ("cells sharded per rank, NCCL gather of yearly summaries to rank 0 inside the step"
                          if world > 1 else "single GPU")}


# ---------------------------------------------------------------------------------------------
def make_workload(rank: int):
    """Grid + weather of one rank's shard (seeded; rank r is shard r of the global grid)."""
    from support import synth
    g = synth.grid_natural_3layer(N_CELLS, seed=1 + 1000 * rank)
    wx = synth.synth_weather(N_NODES, YEAR0, N_YEARS, seed=5 + 1000 * rank)
    node = synth.assign_nodes(N_CELLS, N_NODES)
    return g, wx, node


def make_batch(g, wx, node, ex=None, weather="resident"):
    from pyhelp_b200 import grid as G
    return G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                             tfsoil_c=-3.0, extra_layers=ex, weather=weather)


BATCH_ARRAYS = ("years", "pre", "tmp", "rad", "iu4", "iu7", "iu13", "tm_normals", "idfs",
                "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc")


def pin_batch(batch, into=None):
    """The batch's arrays in pinned host memory (what a caller streaming grids would hold).  ``into``:
    pinned tensors of an earlier call with the same shapes, re-used (a copy, no allocation)."""
    import torch
    keep = into or {}
    for k in BATCH_ARRAYS:
        a = getattr(batch, k)
        if not isinstance(a, np.ndarray):          # a weather table that was packed on the device and stayed there
            continue
        if k in keep and tuple(keep[k].shape) == a.shape:
            keep[k].numpy()[...] = a
        else:
            keep[k] = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        setattr(batch, k, keep[k].numpy())
    batch._pinned = keep
    return batch


def pin_tables(wx):
    """The caller's daily weather tables in pinned memory (the packer uploads them from there)."""
    import torch
    keep = []
    for k in ("precip", "airtemp", "solrad"):
        t = torch.from_numpy(np.ascontiguousarray(wx[k], dtype=np.float64)).pin_memory()
        keep.append(t)
        wx[k] = t.numpy()
    wx["_pinned"] = keep
    return wx


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # "under load": the upper part of the samples (the idle edges of the region pull the median down)
        load = sorted(sm)[len(sm) // 4:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = osp.join(ROOT, "MEASURED_PEAKS.json")
    if osp.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic_json():
    try:
        return json.load(open(osp.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def operative_limits():
    """FP64-pipe / issue / lane utilisation of the simulation kernel from the committed ncu capture
    (profiles/traffic.json): the counters that bound this kernel, reported next to the HBM fraction
    the contract asks for.  Profile-derived, not measured in this run."""
    return _traffic_json().get("operative_limits")


def measured_traffic():
    """dram bytes per launch of the simulation kernel from the committed ncu capture of this
    workload (profiles/traffic.json, written by scripts/ncu_summary.py), else None."""
    d = _traffic_json()
    try:
        return float(d["dram_bytes_per_launch"]) if d.get("workload") == WORKLOAD else None
    except Exception:
        return None


def roofline(alg_bytes: float, kernel_ms: float, headline: bool) -> dict:
    peak, peak_src = peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": measured_traffic() if headline else None, "kernel": "help_sim_kernel", "kernel_ms": kernel_ms,
         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
    if headline:
        r["note"] = "path is FP64-latency bound (DESIGN.md section 3): the HBM fraction is small by construction"
        r["operative_limits"] = operative_limits()
    return r


# ---------------------------------------------------------------------------------------------
def cpu_oracle_rate(g, wx, node, cores: int, cells_per_core: int, libm: str, workdir: str):
    """cell-years/s of the CPU oracle fanned out over `cores` processes on a bounded sample
    (the first cores*cells_per_core cells of the workload, all 30 years), text inputs
    included -- that is the reference's path (pyhelp/processing.py:30-59)."""
    from support import synth
    from oracle import oracle
    ns = min(N_CELLS, cores * cells_per_core)
    cells = list(range(ns))
    nodes = sorted({int(node[i]) for i in cells})
    files = synth.write_weather_files(workdir, wx, nodes=nodes)
    cp = synth.cellparams_from_grid(g, files, node, node, node, N_YEARS, tfsoil_c=-3.0, cells=cells, workdir=workdir)
    oracle.build()
    t0 = time.perf_counter()
    out = oracle.run_help_allcells(cp, ncore=cores, libm=libm)
    dt = time.perf_counter() - t0
    assert len(out) == ns
    return ns * N_YEARS / dt, dt, ns


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    g, wx, node = make_workload(0)
    wd = tempfile.mkdtemp(prefix="help_ref_")
    per_core = 64
    times = []
    for it in range(args.warmup + args.steps):
        r, dt, ns = cpu_oracle_rate(g, wx, node, cores, per_core, "glibc", wd)
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = ns * N_YEARS * len(times) / total
    sample = f"first {ns} cells of the workload x {N_YEARS} yr per step, Pool({cores}), text inputs parsed per cell"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32",
        "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": "cell-years/s", "cores": cores, "kind": "port", "libm": "glibc",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cell-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference Fortran not buildable here (no gfortran): C++ port oracle/help3o_oracle.cpp, glibc libm",
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of a device buffer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _write_weather_chunk(job):
    from support import synth
    workdir, wx, nodes = job
    return synth.write_weather_files(workdir, wx, nodes=nodes)


def build_text_cellparams(g, wx, node, workdir):
    """The reference's inputs for the whole workload: D4/D7/D13 files of every node (written by a
    process pool -- that is the reference's preprocessing, not the timed path) and the D10/D11
    records of every cell."""
    import multiprocessing as mp
    from support import synth
    cores = min(os.cpu_count() or 1, 32)
    nn = wx["precip"].shape[1]
    chunks = [list(range(a, min(a + 25, nn))) for a in range(0, nn, 25)]
    files = {"D4": {}, "D7": {}, "D13": {}}
    with mp.get_context("fork").Pool(cores) as pool:
        for part in pool.imap_unordered(_write_weather_chunk, [(workdir, wx, c) for c in chunks]):
            for k in files:
                files[k].update(part[k])
    return synth.cellparams_from_grid(g, files, node, node, node, N_YEARS, tfsoil_c=-3.0, workdir=workdir)


def one_config(kind: str, n: int, ny: int, seed: int, N):
    """One timed simulation of another BASELINE configuration on this GPU (after one warm-up)."""
    from pyhelp_b200 import processing
    from support import synth
    res = {"mixed": synth.grid_mixed, "landfill": synth.grid_landfill}[kind](n, seed=seed)
    g, ex = res if isinstance(res, tuple) else (res, None)
    nn = max(1, n // 100)
    wx = synth.synth_weather(nn, YEAR0, ny, seed=5)
    node = synth.assign_nodes(n, nn)
    batch = make_batch(g, wx, node, ex=ex)
    alg = N.lib().help_b200_algorithmic_bytes(batch.as_struct())
    with processing.HelpEngine(batch, monthly=True) as eng:
        eng.simulate()
        eng.simulate()
        ms = eng.ms_simulate + eng.ms_setup
        k_ms = eng.ms_simulate
        flags = eng.fetch(monthly=False, yearly=False)["flags"]
    return {"cells": n, "years": ny, "nodes": nn, "value": n * ny / (ms * 1e-3), "unit": "cell-years/s", "ms": ms,
            "cells_flagged_nan_or_bad": int((flags & (N.FLAG_NAN | N.FLAG_BADCELL)).astype(bool).sum()),
            "roofline": roofline(alg, k_ms, headline=False)}


def run_target(world, rank, dev, N, dist, torch):
    """BASELINE configs[2] as `north_star` states the goal: ONE grid of 1 M mixed-class cells with 10 000
    weather nodes, 30 yr, partitioned over the ranks by the product's own sharding code
    (pyhelp_b200.sharding: class-sorted list dealt out over the ranks, every rank the same share of every class), packed per rank, simulated,
    yearly summaries gathered to rank 0 over NCCL and put back in the caller's cell order.
    Every rank holds the caller's tables (as every rank would read the same grid/weather files)."""
    from pyhelp_b200 import grid as G, processing, sharding
    from support import synth
    n, nn, ny = TARGET["cells"], TARGET["nodes"], TARGET["years"]
    g = synth.grid_mixed(n, seed=TARGET["seed"])
    wx = pin_tables(synth.synth_weather_chunked(nn, YEAR0, ny, seed=TARGET["seed"], chunk=500))
    node = synth.assign_nodes(n, nn)
    out_host = torch.empty((N.HY_NROWS, ny, n), dtype=torch.float32).pin_memory() if rank == 0 else None
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # -- plan: every rank computes the same partition from the grid columns
    cost, klass = sharding.grid_cost_and_class(g)
    order, ranges = sharding.plan_shards(cost, klass, world, node)
    a, b = ranges[rank]
    idx = order[a:b]
    t_plan = time.perf_counter()
    # -- pack this rank's cells; weather nodes are compacted when the shard leaves most of them out
    sub = {k: np.asarray(v)[idx] for k, v in g.items()}
    used, inv = np.unique(node[idx], return_inverse=True)
    if used.size * 2 <= nn:
        tabs = [np.ascontiguousarray(wx[k][:, used]) for k in ("precip", "airtemp", "solrad")]
        nloc = inv.astype(np.int32)
    else:
        tabs = [wx[k] for k in ("precip", "airtemp", "solrad")]
        nloc = node[idx]
    batch = G.batch_from_grid(sub, wx["years_of_day"], tabs[0], tabs[1], tabs[2], nloc, nloc, nloc, tfsoil_c=-3.0,
                              weather="resident")
    t_pack = time.perf_counter()
    # -- simulate (monthly results stay on the device of each rank) and gather the yearly summaries
    eng = processing.HelpEngine(batch, monthly=True)
    eng.simulate(stream=torch.cuda.current_stream().cuda_stream)
    k_ms, s_ms = eng.ms_simulate, eng.ms_setup
    t_sim = time.perf_counter()
    ptrs = eng.device_ptrs()
    yearly = torch.as_tensor(_DevArray(ptrs["yearly"], (N.HY_NROWS, ny, batch.n_cells)), device=dev)
    full = sharding.gather_yearly(yearly, n, world, rank, dst=0, ranges=ranges)
    if rank == 0:
        out_host.copy_(sharding.restore_order(full, order), non_blocking=False)
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    flags = eng.fetch(monthly=False, yearly=False)["flags"]
    eng.close()
    t = torch.tensor([t_end - t0, t_plan - t0, t_pack - t_plan, t_sim - t_pack, t_end - t_sim, k_ms, s_ms,
                      float(batch.n_cells), float(cost[idx].sum()),
                      float((flags & (N.FLAG_NAN | N.FLAG_BADCELL)).astype(bool).sum())], device=dev, dtype=torch.float64)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank != 0:
        return None
    A = torch.stack(allt).cpu().numpy()
    wall = float(A[:, 0].max())
    kms = A[:, 5]
    y = out_host.numpy()
    # order check: the precipitation row is the node's weather, whatever rank simulated the cell
    probe = np.random.default_rng(0).choice(n, 64, replace=False)
    yod = wx["years_of_day"]
    ok = all(abs(float(y[N.HY_PRECIP, 0, c]) - float(np.round(wx["precip"][yod == YEAR0, node[c]], 1).sum())) < 0.05
             for c in probe.tolist())
    return {
        "config": "BASELINE configs[2]: synthetic 1M cells x 30 yr, mixed profile classes (50% 3-layer natural, 20% 1-layer, "
                  "20% drain over barrier soil, 10% 5-layer), 10000 weather nodes, ONE grid sharded over the ranks",
        "cells": n, "years": ny, "nodes": nn, "n_gpus": world,
        "wall_s": wall, "value": n * ny / wall, "unit": "cell-years/s",
        "under_10s": bool(wall < 10.0),
        "wall_breakdown_s_max_over_ranks": {"plan": float(A[:, 1].max()), "pack": float(A[:, 2].max()),
                                            "upload_setup_simulate": float(A[:, 3].max()),
                                            "gather_assemble_d2h": float(A[:, 4].max())},
        "per_rank_kernel_ms": [float(x) for x in kms], "per_rank_setup_ms": [float(x) for x in A[:, 6]],
        "per_rank_cells": [int(x) for x in A[:, 7]], "per_rank_estimated_cost": [float(x) for x in A[:, 8]],
        "imbalance": float(kms.max() / kms.mean()),
        "gather_ms": float(A[0, 4] * 1e3), "cells_flagged_nan_or_bad": int(A[:, 9].sum()),
        "caller_order_restored": bool(ok),
        "partition": "pyhelp_b200.sharding.grid_cost_and_class + plan_shards (class-sorted grid dealt out over the ranks: every rank the same share of every class), gather_yearly, restore_order",
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyhelp_b200 import processing, _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the HELP engine has no CPU path")
    torch.cuda.set_device(local)
    N.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    g, wx, node = make_workload(rank)
    pin_tables(wx)
    batch = pin_batch(make_batch(g, wx, node))
    n, ny = batch.n_cells, batch.n_years
    stream = torch.cuda.current_stream()
    eng = processing.HelpEngine(batch, monthly=True)
    ptrs = eng.device_ptrs()
    yearly = torch.as_tensor(_DevArray(ptrs["yearly"], (N.HY_NROWS, ny, n)), device=dev)
    gathered = [torch.empty_like(yearly) for _ in range(world)] if (world > 1 and rank == 0) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.simulate(stream=stream.cuda_stream)
        if world > 1:
            dist.gather(yearly, gathered, dst=0)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim_ms, launches = [], 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        sim_ms.append(eng.ms_simulate)
        launches += eng.n_launches
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n * ny / (ms_step * 1e-3)
    res_check = eng.fetch(monthly=False, yearly=True)
    bad_flags = int((res_check["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL)).astype(bool).sum())
    alg = N.lib().help_b200_algorithmic_bytes(batch.as_struct())
    eng.close()
    del yearly, gathered
    torch.cuda.synchronize()

    # ---- e2e: from the caller's tables to the results on the host, every step ---------------------
    # pack (grid columns + daily weather tables -> quantised SoA batch, the weather round trip on the
    # device) -> pinned staging -> help_b200_run (H2D, set-up, sort, simulation, D2H) -> at N>1 the
    # NCCL gather of the yearly summaries and their assembly on rank 0.
    out_m = torch.empty((N.HR_NROWS, ny, 12, n), dtype=torch.float32).pin_memory()
    out_y = torch.empty((N.HY_NROWS, ny, n), dtype=torch.float32).pin_memory()
    out_f = torch.empty((n,), dtype=torch.int32).pin_memory()
    y_all = [torch.empty((N.HY_NROWS, ny, n), dtype=torch.float32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    y_host = torch.empty((N.HY_NROWS, ny, n * world), dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None
    e2e_steps = max(5, min(args.steps, 8))
    pinned = dict(batch._pinned)
    pack_s = []

    def e2e_step():
        import ctypes as C
        t0 = time.perf_counter()
        b = pin_batch(make_batch(g, wx, node), into=pinned)
        pack_s.append(time.perf_counter() - t0)
        with processing.HelpEngine(b, monthly=True) as e:                     # H2D of the batch
            e.simulate(stream=stream.cuda_stream)
            if world > 1:
                yd = torch.as_tensor(_DevArray(e.device_ptrs()["yearly"], (N.HY_NROWS, ny, n)), device=dev)
                dist.gather(yd, y_all, dst=0)
                if rank == 0:
                    y_host.copy_(torch.cat(y_all, dim=2))
            r = N.Result()
            r.monthly, r.yearly, r.flags = out_m.data_ptr(), out_y.data_ptr(), out_f.data_ptr()
            N.check(N.lib().help_b200_engine_fetch(e._h, C.byref(r)), "engine_fetch")      # D2H into pinned buffers
        return b

    e2e_step()                                   # warm-up
    pack_s.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        b_last = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * ny * e2e_steps / float(t.item())
    raw_tables = sum(wx[k].nbytes for k in ("precip", "airtemp", "solrad"))
    # (weather='resident': the packed tables never leave the device -- no D2H, and only the rest of the batch goes up)
    packed_host = sum(getattr(b_last, k).nbytes for k in ("pre", "tmp", "rad") if isinstance(getattr(b_last, k), np.ndarray))
    packed_all = sum(getattr(b_last, k).nbytes for k in ("pre", "tmp", "rad"))
    h2d = raw_tables + b_last.nbytes() - (packed_all - packed_host)
    d2h = packed_host + out_m.numel() * 4 + out_y.numel() * 4 + out_f.numel() * 4 + \
        (y_host.numel() * 4 if y_host is not None else 0)
    e2e = {"value": e2e_value, "unit": "cell-years/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "pack_s_per_step": float(np.mean(pack_s)),
           "api": "grid.batch_from_grid(weather='resident') -> HelpEngine create/simulate"
                  + (" -> NCCL gather of yearly summaries -> rank-0 host array" if world > 1 else "")
                  + " -> fetch (monthly+yearly+flags into pinned host buffers)",
           "ratio_to_value": e2e_value / value}

    line = None
    if rank == 0:
        k_ms = float(np.mean(sim_ms))
        line = {
            "metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
            "config": config_dict(world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline(alg, k_ms, headline=True),
            "cells_flagged_nan_or_bad": bad_flags,
        }
    if world == 1:
        # ---- the reference's own call on its own input format --------------------------------------
        wd = tempfile.mkdtemp(prefix="help_dropin_")
        t0 = time.perf_counter()
        cp = build_text_cellparams(g, wx, node, wd)
        t_build = time.perf_counter() - t0
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            out = processing.run_help_allcells(cp)
            dt = time.perf_counter() - t0
        assert len(out) == n
        same = all(np.array_equal(out[str(c)]["rechg"][:, :], np.ascontiguousarray(out_m[N.HR_RECHG, :, :, c].numpy()),
                                  equal_nan=True) for c in (0, n // 2, n - 1))
        line["e2e_dropin"] = {"value": n * ny / dt, "unit": "cell-years/s", "wall_s": dt, "cells": n,
                              "api": "pyhelp_b200.run_help_allcells(cellparams): D10/D11 S80 records + D4/D7/D13 files in, dict of per-cell dicts out",
                              "inputs_written_in_s": t_build, "same_bits_as_array_path": bool(same),
                              "stages_s": {k: round(float(v), 4) for k, v in processing.LAST_TIMING.items()}}
        del cp, out
        line["other_configs"] = {
            "mixed_125k_x_30yr": one_config("mixed", 125_000, 30, 2, N),
            "landfill_31k_x_50yr": one_config("landfill", 31_250, 50, 3, N),
        }
        cores = os.cpu_count() or 1
        wd = tempfile.mkdtemp(prefix="help_cpu_")
        rate, cdt, ns = cpu_oracle_rate(g, wx, node, cores, 64, "glibc", wd)
        line["cpu_baseline"] = {"value": rate, "unit": "cell-years/s", "cores": cores, "kind": "port", "libm": "glibc",
                                "sample": f"first {ns} cells x {N_YEARS} yr of the same workload, Pool({cores}), {cdt:.1f} s"}
    else:
        del out_m
        tgt = run_target(world, rank, dev, N, dist, torch)
        if rank == 0:
            line["target"] = tgt
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    args.steps = max(1, args.steps)
    args.warmup = max(0, args.warmup)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(3, args.warmup)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
