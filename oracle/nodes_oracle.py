"""CPU restatement of PyHELP's nearest-weather-node assignment (TEST INFRASTRUCTURE).

Follows ``calc_dist_from_coord`` (reference ``pyhelp/utils.py:81-96``) and the per-cell
``int(np.argmin(dist))`` of ``pyhelp/managers.py:307-313`` literally.  Only tests/ import it."""
import numpy as np


def calc_dist_from_coord(lat1, lon1, lat2, lon2):
    lat1, lon1 = np.radians(lat1), np.radians(lon1)
    lat2, lon2 = np.radians(lat2), np.radians(lon2)
    r = 6373
    dlon = lon2 - lon1
    dlat = lat2 - lat1
    a = np.sin(dlat / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin(dlon / 2) ** 2
    c = 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))
    return r * c


def nearest_nodes(cell_lat, cell_lon, node_lat, node_lon, return_margin=False):
    """argmin per cell; with return_margin also the gap between the two smallest distances (km)."""
    node_lat = np.asarray(node_lat, float); node_lon = np.asarray(node_lon, float)
    idx = np.empty(len(cell_lat), dtype=np.int32); gap = np.empty(len(cell_lat))
    for i, (la, lo) in enumerate(zip(cell_lat, cell_lon)):
        d = calc_dist_from_coord(la, lo, node_lat, node_lon)
        idx[i] = int(np.argmin(d))
        if return_margin:
            s = np.partition(d, 1)[:2] if d.size > 1 else np.array([d[0], np.inf])
            gap[i] = s[1] - s[0]
    return (idx, gap) if return_margin else idx
