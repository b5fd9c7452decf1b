"""Multi-GPU sharding of a cell batch: one process per GPU, no data-path collective.

The reference's only parallelism is data parallel over independent cells
(``Pool(ncore).imap_unordered``, reference ``pyhelp/processing.py:40-59``).  Here the cells are cut
into contiguous, equally sized shards (one per rank); each rank uploads only the weather nodes
its cells reference, simulates, and the per-cell yearly summaries ``[5][Ny][n_local]`` are
gathered to rank 0 -- over NCCL (NVLink) on GPUs, over gloo in the CPU tests.  Nothing is
exchanged during the simulation itself.
"""
from __future__ import annotations

import numpy as np

from .inputs import CellBatch


def shard_range(n_cells: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous shard; sizes differ by at most one cell."""
    base, rem = divmod(int(n_cells), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def compact_nodes(batch: CellBatch) -> CellBatch:
    """Drop the weather nodes no cell of this (sub-)batch references and renumber the rest, so a
    rank uploads ~W/world nodes instead of all W."""
    from . import _native as N
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


def shard_batch(batch: CellBatch, world: int, rank: int) -> CellBatch:
    a, b = shard_range(batch.n_cells, world, rank)
    return compact_nodes(batch.take(np.arange(a, b)))


def gather_yearly(yearly_local, n_cells_total: int, world: int, rank: int, dst: int = 0):
    """Gather per-rank ``[5][Ny][n_local]`` tensors to ``dst``; returns ``[5][Ny][n_total]`` there
    (caller's cell order, because shards are contiguous) and None elsewhere.  Works on CUDA
    tensors with the nccl backend and on CPU tensors with gloo."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return yearly_local
    sizes = [shard_range(n_cells_total, world, r) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    rows, ny = yearly_local.shape[0], yearly_local.shape[1]
    pad = torch.zeros((rows, ny, nmax), dtype=yearly_local.dtype, device=yearly_local.device)
    pad[:, :, :yearly_local.shape[2]] = yearly_local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:, :, :sizes[r][1] - sizes[r][0]] for r in range(world)], dim=2)
