"""Nearest weather-grid node of every cell on the GPU.

Mirror of the per-cell loop in ``HelpManager._generate_d4d7d13_input_files``
(reference ``pyhelp/managers.py:307-313``): ``int(np.argmin(calc_dist_from_coord(lat, lon,
node_lats, node_lons)))`` with the haversine distance of ``pyhelp/utils.py:81-96``.  The result is
the index array ``connect_tables[var]`` is built from.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _native as N


def nearest_nodes(cell_lat, cell_lon, node_lat, node_lon) -> np.ndarray:
    """int32[n_cells]: index of the closest node (first one on equal distance)."""
    clat = np.ascontiguousarray(cell_lat, dtype=np.float64).reshape(-1)
    clon = np.ascontiguousarray(cell_lon, dtype=np.float64).reshape(-1)
    nlat = np.ascontiguousarray(node_lat, dtype=np.float64).reshape(-1)
    nlon = np.ascontiguousarray(node_lon, dtype=np.float64).reshape(-1)
    if clat.shape != clon.shape or nlat.shape != nlon.shape:
        raise ValueError("latitude and longitude arrays must have the same length")
    if nlat.size < 1:
        raise ValueError("at least one weather node is required")
    N.require_device()
    out = np.empty(clat.size, dtype=np.int32)
    N.check(N.lib().help_b200_nearest_nodes(clat.size, N.ptr(clat), N.ptr(clon), nlat.size, N.ptr(nlat), N.ptr(nlon),
                                            N.ptr(out)), "nearest_nodes")
    return out
