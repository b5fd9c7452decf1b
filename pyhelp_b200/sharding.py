"""Multi-GPU sharding of a cell batch: one process per GPU, no data-path collective.

The reference's only parallelism is data parallel over independent cells
(``Pool(ncore).imap_unordered``, reference ``pyhelp/processing.py:40-59``).  Here one grid is
sorted by profile class and dealt out over the ranks (``plan_shards``: every rank the same share of
every class) or, for an already packed batch, cut into contiguous shards (``shard_batch``); each
rank uploads only the weather nodes its cells reference, simulates, and the per-cell yearly
summaries ``[5][Ny][n_local]`` are gathered to rank 0 -- over NCCL (NVLink) on GPUs, over gloo in
the CPU tests -- and put back in the caller's cell order.  Nothing is exchanged during the
simulation itself.
"""
from __future__ import annotations

import os

import numpy as np

from .inputs import CellBatch


def shard_range(n_cells: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous shard; sizes differ by at most one cell."""
    base, rem = divmod(int(n_cells), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


# Relative cost of a cell by profile structure, per modelling segment: a profile with a liner pays head, leakage and
# lateral-drainage iterations every sub-step and runs the larger routing code.  Measured on B200 (scripts/class_cost.py,
# the four classes of the BASELINE configs[2] grid on their own, 100 000 cells each: 0.116 / 0.118 us per cell-year and
# segment for the two natural classes, 0.228 / 0.224 for the two liner classes => 1.9).  The ratio holds when every class
# fills the GPU; a rank that holds fewer liner cells than the SMs keep resident is latency-bound and the ratio drops to
# ~1.55 -- which is why plan_shards deals the classes out instead of relying on this number.
COST_LINER = float(os.environ.get("HELP_B200_COST_LINER", 1.9))


def _segments(nlay, thick_in, on):
    """7 evaporative segments + 1..3 per layer by thickness (reference ``HELP3O.FOR:1503-1616``)."""
    return 7 + np.where(on, np.clip(np.ceil(thick_in / 18.0), 1, 3), 0).sum(axis=0)


def estimate_cost(batch: CellBatch) -> np.ndarray:
    """Relative simulation cost of every cell, from its layer records alone (no GPU work):
    ~ number of modelling segments times :data:`COST_LINER` for a profile with a liner (barrier
    soil or geomembrane)."""
    from . import _native as N
    nlay = batch.cell_i32[N.HB_NLAY].astype(np.int64)
    L = batch.max_layers
    lay = np.arange(L)[:, None] < nlay[None, :]
    thick = batch.layer_f32[N.HL_THICK]                       # cm (IU10 = 2) or inches
    si = batch.cell_i32[N.HB_IU10] == 2
    inches = np.where(si[None, :], thick / 2.54, thick)
    ltype = batch.layer_i32[N.HL_TYPE]
    liner = (lay & (ltype >= 3)).any(axis=0)
    return _segments(nlay, inches, lay) * np.where(liner, COST_LINER, 1.0)


def grid_cost_and_class(grid: dict):
    """(cost, class) of every cell straight from the grid columns (``nlayer``, ``lay_type{i}``,
    ``thick{i}`` in cm; docs/data.rst:122-188), before anything is packed.  ``class`` orders the
    kernel instantiations a shard needs: segment capacity tier of the profile (<= 3 layers: 16
    segments, <= 6: 25, <= 11: 40, else 67 -- the engine sizes its per-thread arrays by the deepest
    profile of a batch) and, inside a tier, profiles without a liner first (they run the
    plain-profile routing)."""
    nlay = np.asarray(grid["nlayer"]).astype(np.int64)
    n = nlay.size
    L = int(nlay.max()) if n else 0
    thick = np.zeros((max(L, 1), n))
    ltype = np.zeros((max(L, 1), n), dtype=np.int64)
    for j in range(L):
        s = str(j + 1)
        if "thick" + s in grid:
            thick[j] = np.maximum(np.asarray(grid["thick" + s], dtype=float), 10.0)     # MINTHICK, preprocessing.py:22
        if "lay_type" + s in grid:
            ltype[j] = np.asarray(grid["lay_type" + s]).astype(np.int64)
    on = np.arange(max(L, 1))[:, None] < nlay[None, :]
    liner = (on & (ltype >= 3)).any(axis=0)
    cost = _segments(nlay, thick / 2.54, on) * np.where(liner, COST_LINER, 1.0)
    tier = np.digitize(nlay, [4, 7, 12])                     # 0: <=3 layers, 1: <=6, 2: <=11, 3: more
    return cost, (2 * tier + liner.astype(np.int64))


def plan_shards(cost, klass, world: int, node=None, mode: str = "deal"):
    """Partition of one grid over ``world`` ranks.  The cells are stably sorted by (class, cost, weather
    node); returns ``(order, ranges)``: rank r simulates the cells ``order[ranges[r][0]:ranges[r][1]]``.

    ``mode="deal"`` (default): the sorted list is dealt out card by card -- rank r takes every ``world``-th cell starting
    at r -- so every rank holds the same number of cells of every class and the same cost distribution: the ranks finish
    together *by construction*, whatever a class costs on the hardware.  Inside a rank the engine sorts again and launches
    the expensive classes first.

    ``mode="contiguous"``: the sorted list is cut into contiguous runs of equal estimated cost (ranks then hold nearly one
    class each).  Measured on B200 with the 1 M-cell BASELINE configs[2] grid cut for 8 ranks: per-shard kernel times
    differ by 15-19 % (max/mean) however the liner factor is set -- the time of a class depends on how full the SMs of its
    rank are, not only on its segments -- while the dealt shards run as fast as the mean of the contiguous ones
    (profiles/r2_notes.md)."""
    cost = np.asarray(cost, dtype=np.float64).reshape(-1)
    keys = [np.zeros(cost.size, dtype=np.int64) if node is None else np.asarray(node).reshape(-1), cost,
            np.asarray(klass).reshape(-1)]
    order = np.lexsort(keys)                                  # last key is the primary one; stable
    if mode == "contiguous":
        return order, cost_balanced_ranges(cost[order], world)
    if mode != "deal":
        raise ValueError("mode must be 'deal' or 'contiguous'")
    hands = [order[r::world] for r in range(world)]
    stops = np.cumsum([h.size for h in hands])
    return np.concatenate(hands) if hands else order, [(int(b - h.size), int(b)) for h, b in zip(hands, stops)]


def restore_order(gathered, order):
    """Results gathered in shard order (``[..., n_total]`` along the last axis) back in the caller's
    cell order -- the reference returns one dict keyed by cell (``pyhelp/processing.py:40-59``)."""
    import torch
    if isinstance(gathered, torch.Tensor):
        out = torch.empty_like(gathered)
        out[..., torch.as_tensor(np.asarray(order), device=gathered.device)] = gathered
        return out
    out = np.empty_like(gathered)
    out[..., np.asarray(order)] = gathered
    return out


def cost_balanced_ranges(cost, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, stop) per rank with (nearly) equal total cost: every rank finishes at
    the same time even when cheap and expensive profiles are clustered in the caller's order.
    Equal costs reproduce :func:`shard_range` up to one cell."""
    c = np.asarray(cost, dtype=np.float64).reshape(-1)
    n = c.size
    if n == 0:
        return [(0, 0)] * world
    pre = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(pre, pre[-1] * r / world, side="left"))
        # the cut that leaves the prefix closest to its target
        if k > 0 and abs(pre[k - 1] - pre[-1] * r / world) <= abs(pre[min(k, n)] - pre[-1] * r / world):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def compact_nodes(batch: CellBatch) -> CellBatch:
    """Drop the weather nodes no cell of this (sub-)batch references and renumber the rest, so a
    rank uploads ~W/world nodes instead of all W."""
    from . import _native as N
    if any(isinstance(t, N.DeviceTable) for t in (batch.pre, batch.tmp, batch.rad)):
        raise TypeError("compact_nodes needs the weather tables on the host (pack with weather='host' or 'device')")
    out = CellBatch(batch.n_cells, batch.n_years, batch.max_layers, batch.years, batch.pre, batch.tmp, batch.rad,
                    batch.iu4, batch.iu7, batch.iu13, batch.tm_normals, batch.idfs,
                    batch.cell_f32, batch.cell_i32.copy(), batch.layer_f32, batch.layer_i32, batch.layer_rc,
                    list(batch.names))
    for row, arrays in ((N.HB_NODE_P, ("pre", "iu4")), (N.HB_NODE_T, ("tmp", "iu7", "tm_normals")),
                        (N.HB_NODE_R, ("rad", "iu13", "idfs"))):
        used, inv = np.unique(batch.cell_i32[row], return_inverse=True)
        if used.size == 0:
            continue
        out.cell_i32[row] = inv.astype(np.int32)
        for name in arrays:
            setattr(out, name, np.ascontiguousarray(getattr(batch, name)[used]))
    return out


def shard_batch(batch: CellBatch, world: int, rank: int, by_cost: bool = False) -> CellBatch:
    """Rank's contiguous shard (equal cell counts, or equal estimated cost with ``by_cost``)
    with only the weather nodes it references."""
    a, b = cost_balanced_ranges(estimate_cost(batch), world)[rank] if by_cost else shard_range(batch.n_cells, world, rank)
    return compact_nodes(batch.take(np.arange(a, b)))


def gather_yearly(yearly_local, n_cells_total: int, world: int, rank: int, dst: int = 0, ranges=None):
    """Gather per-rank ``[5][Ny][n_local]`` tensors to ``dst``; returns ``[5][Ny][n_total]`` there
    (caller's cell order, because shards are contiguous; ``ranges`` = the per-rank [start, stop)
    when they are not the equal-count ones) and None elsewhere.  Works on CUDA
    tensors with the nccl backend and on CPU tensors with gloo."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return yearly_local
    sizes = list(ranges) if ranges is not None else [shard_range(n_cells_total, world, r) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    rows, ny = yearly_local.shape[0], yearly_local.shape[1]
    pad = torch.zeros((rows, ny, nmax), dtype=yearly_local.dtype, device=yearly_local.device)
    pad[:, :, :yearly_local.shape[2]] = yearly_local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:, :, :sizes[r][1] - sizes[r][0]] for r in range(world)], dim=2)
