"""Seeded synthetic inputs for the BASELINE configurations and the
text writers that reproduce the reference's input formats.

Everything here is input GENERATION -- no simulation code.  Two consumers:

* ``bench.py`` builds config 2-5 grids + weather as arrays and packs them with
  :func:`pyhelp_b200.grid.batch_from_grid` (no text);
* the parity tests write the same inputs as the text the reference exchanges
  (D4/D7/D13 files like ``pyhelp/weather_reader.py:22-104``; D10/D11 records
  like ``pyhelp/preprocessing.py:27-185``) so that the CPU oracle and the
  drop-in ``run_help_allcells`` consume exactly what PyHELP would hand them.
"""
from __future__ import annotations

import os
import os.path as osp

import numpy as np


# ---------------------------------------------------------------------------
# weather
# ---------------------------------------------------------------------------
def isleap(y: int) -> bool:
    return y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)


def synth_weather(n_nodes: int, year0: int, n_years: int, seed: int, lat0: float = 46.0):
    """Daily precip (mm), mean air temperature (C) and global radiation (MJ/m2/d)
    for `n_nodes` stations; a cold-temperate climate (snow, frozen soil, summer ET)
    so that every branch of the surface model is exercised.  Values are on the
    grids the reference's D4/D7/D13 writers quantise to (0.1 mm, 0.1 C, 0.01 MJ)."""
    rng = np.random.default_rng(seed)
    years = np.arange(year0, year0 + n_years)
    ndays = np.array([366 if isleap(int(y)) else 365 for y in years])
    yod = np.repeat(years, ndays)
    doy = np.concatenate([np.arange(1, k + 1) for k in ndays]).astype(float)
    T = doy.size
    lat = lat0 + rng.uniform(-1.5, 1.5, n_nodes)
    lon = -74.0 + rng.uniform(-2.0, 2.0, n_nodes)
    # temperature: annual harmonic + AR(1) noise + station offset
    tmean = 5.5 - 0.7 * (lat - lat0) + rng.normal(0, 0.5, n_nodes)
    amp = 15.0 + rng.normal(0, 0.6, n_nodes)
    seas = -np.cos(2 * np.pi * (doy - 20.0) / 365.25)
    eps = rng.normal(0, 1.0, (T, n_nodes)).astype(np.float32)
    ar = np.empty_like(eps)
    ar[0] = eps[0]
    for t in range(1, T):
        ar[t] = 0.72 * ar[t - 1] + 0.69 * eps[t]
    temp = tmean[None, :] + amp[None, :] * seas[:, None] + 3.6 * ar
    # precipitation: two-state Markov occurrence, gamma amounts, wetter in summer
    u = rng.random((T, n_nodes), dtype=np.float32)
    wet = np.zeros((T, n_nodes), dtype=bool)
    wet[0] = u[0] < 0.4
    for t in range(1, T):
        wet[t] = np.where(wet[t - 1], u[t] < 0.58, u[t] < 0.30)
    amount = rng.gamma(0.75, 8.5, (T, n_nodes)) * (1.0 + 0.25 * np.sin(2 * np.pi * (doy - 120) / 365.25))[:, None]
    precip = np.where(wet, np.maximum(amount, 0.1), 0.0)
    # radiation: clear-sky seasonal envelope x cloudiness tied to wet days
    env = 16.0 + 12.5 * np.sin(2 * np.pi * (doy - 81) / 365.25)
    cloud = np.where(wet, rng.uniform(0.25, 0.65, (T, n_nodes)), rng.uniform(0.6, 1.0, (T, n_nodes)))
    solrad = np.maximum(env[:, None] * cloud, 0.3)
    return {
        "years_of_day": yod.astype(np.int32),
        "precip": np.round(precip, 1), "airtemp": np.round(temp.astype(np.float64), 1),
        "solrad": np.round(solrad, 2), "lat": np.round(lat, 4), "lon": np.round(lon, 4),
    }


def synth_weather_chunked(n_nodes: int, year0: int, n_years: int, seed: int, chunk: int = 500, lat0: float = 46.0):
    """:func:`synth_weather` for many nodes, generated `chunk` nodes at a time (seed + chunk index) so
    that the temporaries stay small; the tables come out C-contiguous [day][node]."""
    out = None
    for c, a in enumerate(range(0, n_nodes, chunk)):
        b = min(a + chunk, n_nodes)
        w = synth_weather(b - a, year0, n_years, seed=seed * 7919 + c, lat0=lat0)
        if out is None:
            T = w["years_of_day"].size
            out = {"years_of_day": w["years_of_day"], "lat": np.empty(n_nodes), "lon": np.empty(n_nodes),
                   "precip": np.empty((T, n_nodes)), "airtemp": np.empty((T, n_nodes)), "solrad": np.empty((T, n_nodes))}
        for k in ("precip", "airtemp", "solrad"):
            out[k][:, a:b] = w[k]
        out["lat"][a:b], out["lon"][a:b] = w["lat"], w["lon"]
    return out


# ---------------------------------------------------------------------------
# grids
# ---------------------------------------------------------------------------
def _base_grid(n: int, rng, lat0=46.0):
    g = {
        "cid": np.arange(n),
        "lat_dd": np.round(rng.uniform(lat0 - 1.0, lat0 + 1.0, n), 5),
        "lon_dd": np.round(rng.uniform(-75.0, -73.0, n), 5),
        "wind": np.full(n, 10.0),
        "hum1": np.full(n, 67.9), "hum2": np.full(n, 65.5), "hum3": np.full(n, 74.0), "hum4": np.full(n, 75.7),
        "growth_start": rng.integers(112, 135, n), "growth_end": rng.integers(285, 313, n),
        "LAI": np.round(rng.uniform(0.0, 4.0, n), 2),
        "EZD": rng.integers(10, 61, n).astype(float),
        "CN": rng.integers(40, 91, n).astype(float),
        "context": np.where(rng.random(n) < 0.08, 2, 1),
        "run": np.ones(n, dtype=int),
    }
    return g


def _soil_layer(g, j, n, rng, ltype, on=None, ksat=(1e-5, 1e-2)):
    s = str(j)
    on = np.ones(n, bool) if on is None else on
    poro = np.round(rng.uniform(0.35, 0.50, n), 3)
    fc = np.round(np.minimum(rng.uniform(0.10, 0.30, n), poro - 0.03), 3)
    wp = np.round(np.minimum(rng.uniform(0.03, 0.15, n), fc - 0.02), 3)
    g["lay_type" + s] = np.where(on, ltype, 0)
    g["thick" + s] = np.where(on, rng.integers(20, 151, n), 0).astype(float)
    g["poro" + s] = np.where(on, poro, 0)
    g["fc" + s] = np.where(on, fc, 0)
    g["wp" + s] = np.where(on, wp, 0)
    k = 10 ** rng.uniform(np.log10(ksat[0]), np.log10(ksat[1]), n)
    g["ksat" + s] = np.where(on, np.round(k, 14), 0)
    g["dist_dr" + s] = np.where(on, rng.integers(5, 101, n), 0).astype(float)
    g["slope" + s] = np.where(on, np.round(rng.uniform(0.5, 10.0, n), 2), 0)


def grid_natural_3layer(n: int, seed: int = 1):
    """BASELINE config 2: 3 type-1 layers (natural soil profile)."""
    rng = np.random.default_rng(seed)
    g = _base_grid(n, rng)
    g["nlayer"] = np.full(n, 3)
    for j in (1, 2, 3):
        _soil_layer(g, j, n, rng, 1)
    return g


def grid_example(n: int = 17_178, seed: int = 20260101):
    """BASELINE config 1 stand-in: the reference's example grid
    (``example/input_grid.csv``) is a missing blob in the checkout, so this is a schema-exact
    synthetic grid (columns of ``docs/data.rst:122-188``) with the value ranges of
    ``docs/img/grid_input_data.png``: cells on a 250 m lattice over lat 45.78-46.2, lon -74.5..-74.0
    (the footprint of the example weather nodes), 1-3 type-1 layers, some cells not runnable
    (``run == 0``, ``nlayer == 0``) and a ``Bassin`` column like the example's sub-basin selector."""
    rng = np.random.default_rng(seed)
    ncol = int(np.ceil(np.sqrt(n * 1.19)))              # lattice aspect of the footprint
    nrow = int(np.ceil(n / ncol))
    iy, ix = np.divmod(np.arange(n), ncol)
    lat = 45.78 + (46.2 - 45.78) * iy / max(nrow - 1, 1)
    lon = -74.5 + 0.5 * ix / max(ncol - 1, 1)
    g = {
        "cid": np.array([str(i) for i in range(n)]),
        "lat_dd": np.round(lat, 5), "lon_dd": np.round(lon, 5),
        "wind": np.full(n, 10.0),
        "hum1": np.full(n, 67.9), "hum2": np.full(n, 65.5), "hum3": np.full(n, 74.0), "hum4": np.full(n, 75.7),
        "growth_start": np.full(n, 112), "growth_end": np.full(n, 313),
        "LAI": rng.choice([0.0, 0.2, 3.5], n, p=[0.1, 0.2, 0.7]),
        "EZD": rng.choice([10.0, 20.0, 60.0], n, p=[0.2, 0.3, 0.5]),
        "CN": rng.integers(24, 89, n).astype(float),
        "context": rng.choice([0, 1, 2, 3], n, p=[0.04, 0.82, 0.10, 0.04]),
        "Bassin": (rng.random(n) < 0.6).astype(int),
    }
    nl = rng.choice([0, 1, 2, 3], n, p=[0.02, 0.28, 0.40, 0.30])
    g["nlayer"] = nl
    g["run"] = np.where((g["context"] == 0) | (g["context"] == 3) | (nl == 0), 0, 1)     # water / outside cells do not run
    for j in (1, 2, 3):
        s = str(j)
        on = nl >= j
        g["lay_type" + s] = np.where(on, 1, 0)
        g["thick" + s] = np.where(on, rng.integers(20, 301, n), 0).astype(float)
        g["poro" + s] = np.where(on, np.round(rng.uniform(0.437, 0.473, n), 3), 0)
        g["fc" + s] = np.where(on, np.round(rng.uniform(0.105, 0.222, n), 3), 0)
        g["wp" + s] = np.where(on, np.round(rng.uniform(0.047, 0.104, n), 3), 0)
        g["ksat" + s] = np.where(on, np.round(rng.uniform(5.2e-4, 1.7e-3, n), 14), 0)
        g["dist_dr" + s] = np.where(on, rng.choice([5.0, 25.0], n), 0)
        g["slope" + s] = np.where(on, np.round(rng.uniform(0.5, 4.6, n), 2), 0)
    return g


def grid_mixed(n: int, seed: int = 2):
    """BASELINE config 3: 50 % 3-layer natural, 20 % 1-layer, 20 % 1/2/3 (drain over
    barrier soil), 10 % 5-layer 1/1/2/3/1."""
    rng = np.random.default_rng(seed)
    g = _base_grid(n, rng)
    cls = rng.choice(4, n, p=[0.5, 0.2, 0.2, 0.1])
    nl = np.array([3, 1, 3, 5])[cls]
    g["nlayer"] = nl
    types = {0: [1, 1, 1, 0, 0], 1: [1, 0, 0, 0, 0], 2: [1, 2, 3, 0, 0], 3: [1, 1, 2, 3, 1]}
    for j in range(1, 6):
        _soil_layer(g, j, n, rng, 1, on=nl >= j)
        t = np.zeros(n, int)
        for c, tl in types.items():
            t[cls == c] = tl[j - 1]
        g["lay_type" + str(j)] = t
        drain = t == 2
        barrier = t == 3
        s = str(j)
        g["ksat" + s] = np.where(drain, np.round(10 ** rng.uniform(-2.5, -1.0, n), 14), g["ksat" + s])
        g["ksat" + s] = np.where(barrier, np.round(10 ** rng.uniform(-7.0, -6.0, n), 14), g["ksat" + s])
        g["thick" + s] = np.where(barrier, rng.integers(30, 91, n), g["thick" + s]).astype(float)
        g["thick" + s] = np.where(drain, rng.integers(20, 46, n), g["thick" + s]).astype(float)
        g["poro" + s] = np.where(barrier, 0.427, g["poro" + s])
        g["fc" + s] = np.where(barrier, 0.418, g["fc" + s])
        g["wp" + s] = np.where(barrier, 0.367, g["wp" + s])
        g["poro" + s] = np.where(drain, 0.417, g["poro" + s])
        g["fc" + s] = np.where(drain, 0.045, g["fc" + s])
        g["wp" + s] = np.where(drain, 0.018, g["wp" + s])
    g["class"] = cls
    return g


def grid_landfill(n: int, seed: int = 3, recirc: bool = False):
    """BASELINE config 4: cover = topsoil / lateral drain / geomembrane / clay, half of the
    cells with a second liner system underneath (waste / drain / geomembrane / clay).
    Returns (grid, extra_layers) -- the extra design fields PyHELP's formatter leaves
    blank (pinholes, defects, placement quality).  ``recirc``: the lateral-drain layers send
    10-60 % of their drainage back into a soil layer above (leachate recirculation, D10 fields
    RECIR / LAYR, reference ``HELP3O.FOR:5311-5440``)."""
    rng = np.random.default_rng(seed)
    g = _base_grid(n, rng)
    two = rng.random(n) < 0.5
    nl = np.where(two, 8, 4)
    g["nlayer"] = nl
    ex = {}
    seq = [1, 2, 4, 3, 1, 2, 4, 3]
    for j in range(1, 9):
        on = nl >= j
        _soil_layer(g, j, n, rng, seq[j - 1], on=on)
        s = str(j)
        t = seq[j - 1]
        if t == 2:
            g["thick" + s] = np.where(on, rng.integers(20, 46, n), 0).astype(float)
            g["poro" + s] = np.where(on, 0.417, 0); g["fc" + s] = np.where(on, 0.045, 0); g["wp" + s] = np.where(on, 0.018, 0)
            g["ksat" + s] = np.where(on, np.round(10 ** rng.uniform(-2.5, -1.0, n), 14), 0)
        elif t == 3:
            g["thick" + s] = np.where(on, rng.integers(30, 91, n), 0).astype(float)
            g["poro" + s] = np.where(on, 0.427, 0); g["fc" + s] = np.where(on, 0.418, 0); g["wp" + s] = np.where(on, 0.367, 0)
            g["ksat" + s] = np.where(on, np.round(10 ** rng.uniform(-7.0, -6.0, n), 14), 0)
        elif t == 4:
            g["poro" + s] = np.zeros(n); g["fc" + s] = np.zeros(n); g["wp" + s] = np.zeros(n)
            g["ksat" + s] = np.where(on, np.round(rng.uniform(2e-13, 4e-13, n), 14), 0)
            ex["thick_exact" + s] = np.where(on, np.round(rng.uniform(0.10, 0.15, n), 2), 0)
            ex["phole" + s] = np.where(on, rng.integers(0, 3, n), 0).astype(float)
            ex["defec" + s] = np.where(on, rng.integers(0, 11, n), 0).astype(float)
            ex["ipq" + s] = np.where(on, rng.integers(2, 6, n), 0)
    if recirc:
        rr = np.random.default_rng(seed + 77)
        for j, targets in ((2, (1,)), (6, (1, 5))):              # drain layer -> layer(s) it may recirculate to
            s = str(j)
            on = nl >= j
            ex["recir" + s] = np.where(on, rr.integers(10, 61, n), 0).astype(float)
            ex["layr" + s] = np.where(on, rr.choice(targets, n), 0).astype(int)
    return g, ex


def grid_special(n: int, seed: int = 8):
    """Branches of the reference no PyHELP grid reaches: geomembranes on a geotextile cushion
    (placement quality 6 + transmissivity: FML design cases 5/6, ``HELP3O.FOR:5177-5285``), placement qualities
    1 (perfect contact) and 5 (worst), subsurface inflow into the bottom subprofile with and without a liner
    (IBAR 3/4/5, ``HELP3O.FOR:4164-4210``), user-specified initial moisture (IPRE = 1, ``HELP3O.FOR:986-996``,
    snow water given), and sites at 55 / 65 deg N and 35 deg S (the HYDRO-17 melt-factor branches of SNOW,
    ``HELP3O.FOR:3756-4122``, the June radiation block of the IDFS pre-scan and a growing season that wraps
    around the new year).  Returns (grid, extra_layers)."""
    rng = np.random.default_rng(seed)
    g = _base_grid(n, rng)
    var = np.arange(n) % 8
    lat = np.array([55.3, 65.1, -35.2, 46.0, -35.2, 61.7, 52.4, 46.0])[np.arange(n) % 8 if n < 8 else rng.integers(0, 8, n)]
    g["lat_dd"] = lat
    south = lat < 0
    g["growth_start"] = np.where(south, rng.integers(270, 300, n), g["growth_start"])
    g["growth_end"] = np.where(south, rng.integers(100, 130, n), g["growth_end"])
    # layer sequences: 0, 2, 3: topsoil / drain / geomembrane / clay; 1: topsoil / drain / clay / geomembrane (the
    # geomembrane UNDER its controlling soil: design case 4, 6 with a geotextile); 4: topsoil / drain / clay (barrier
    # soil liner); 5-7: three natural-soil layers
    seqs = {0: [1, 2, 4, 3], 1: [1, 2, 3, 4], 2: [1, 2, 4, 3], 3: [1, 2, 4, 3], 4: [1, 2, 3, 0]}
    g["nlayer"] = np.select([var < 4, var == 4], [4, 3], default=3)
    ex = {"ipre": np.where(var >= 6, 1, 0)}
    for j in range(1, 5):
        s_ = str(j)
        on = g["nlayer"] >= j
        _soil_layer(g, j, n, rng, 1, on=on)
        t = np.ones(n, dtype=int)
        for v, sq in seqs.items():
            t[var == v] = sq[j - 1]
        t = np.where(on, t, 0)
        g["lay_type" + s_] = t
        drain, fml, clay = t == 2, t == 4, t == 3
        g["thick" + s_] = np.where(drain, rng.integers(20, 46, n), np.where(clay, rng.integers(30, 91, n), g["thick" + s_])).astype(float)
        for name, dv, cv in (("poro", 0.417, 0.427), ("fc", 0.045, 0.418), ("wp", 0.018, 0.367)):
            g[name + s_] = np.where(drain, dv, np.where(clay, cv, np.where(fml, 0.0, g[name + s_])))
        g["ksat" + s_] = np.where(drain, np.round(10 ** rng.uniform(-2.5, -1.0, n), 14), g["ksat" + s_])
        g["ksat" + s_] = np.where(clay, np.round(10 ** rng.uniform(-7.0, -6.0, n), 14), g["ksat" + s_])
        g["ksat" + s_] = np.where(fml, np.round(rng.uniform(2e-13, 4e-13, n), 14), g["ksat" + s_])
        ex["thick_exact" + s_] = np.where(fml, np.round(rng.uniform(0.10, 0.15, n), 2), 0)
        ex["phole" + s_] = np.where(fml, rng.integers(0, 3, n), 0).astype(float)
        ex["defec" + s_] = np.where(fml, rng.integers(1, 11, n), 0).astype(float)
        ex["ipq" + s_] = np.where(fml, np.select([var <= 1, var == 2], [6, np.where((np.arange(n) // 8) % 2 == 0, 1, 5)],
                                                 default=rng.integers(2, 5, n)), 0)
        ex["trans" + s_] = np.where(fml & (var <= 1), np.round(10 ** rng.uniform(-1.0, 1.0, n), 3), 0.0)
        # subsurface inflow (mm/yr) into the bottom layer: the barrier soil under a geomembrane (3: IBAR 2 -> 5), a barrier
        # soil liner (4: IBAR 1 -> 4), the last natural layer (5: added to the segment above in REAL*8, F:2060)
        last = on & (g["nlayer"] == j)
        ex["subin" + s_] = np.where(last & np.isin(var, (3, 4, 5)), rng.integers(20, 200, n), 0).astype(float)
        sw = np.round(g["wp" + s_] + rng.uniform(0.2, 0.9, n) * (g["poro" + s_] - g["wp" + s_]), 3)
        ex["sw" + s_] = np.where(on & (var >= 6) & ~fml, sw, 0.0)
    ex["osno"] = np.where(var == 7, np.round(rng.uniform(5, 60, n), 0), 0.0)          # initial snow water, mm
    return g, ex


def grid_deep(n: int, seed: int = 6):
    """The largest profiles HELP accepts (HELP3O.FOR COMMON blocks: 20 layers, 67 segments): 12, 16
    or 20 layers, thick ones (3 segments each), natural soil with up to three drain-over-barrier
    pairs.  Exercises the widest kernel instantiation (segment capacity 67)."""
    rng = np.random.default_rng(seed)
    g = _base_grid(n, rng)
    nl = rng.choice([12, 16, 20], n, p=[0.3, 0.3, 0.4])
    g["nlayer"] = nl
    npairs = rng.integers(0, 4, n)                                    # drain/barrier pairs at layers (5,6), (11,12), (17,18)
    thick_all = rng.random(n) < 0.3                                   # every layer > 36 in: 3 segments each, 67 in all at 20 layers
    for j in range(1, 21):
        on = nl >= j
        _soil_layer(g, j, n, rng, 1, on=on)
        s = str(j)
        g["thick" + s] = np.where(on, np.where(thick_all, rng.integers(95, 151, n), rng.integers(20, 151, n)), 0).astype(float)
        k = {5: 1, 11: 2, 17: 3}.get(j) or {6: 1, 12: 2, 18: 3}.get(j)
        if k is None:
            continue
        pair = on & (nl >= j + (1 if j in (5, 11, 17) else 0)) & (npairs >= k)
        if j in (5, 11, 17):                                          # lateral-drainage layer
            g["lay_type" + s] = np.where(pair, 2, g["lay_type" + s])
            g["thick" + s] = np.where(pair, rng.integers(20, 46, n), g["thick" + s]).astype(float)
            g["poro" + s] = np.where(pair, 0.417, g["poro" + s]); g["fc" + s] = np.where(pair, 0.045, g["fc" + s])
            g["wp" + s] = np.where(pair, 0.018, g["wp" + s])
            g["ksat" + s] = np.where(pair, np.round(10 ** rng.uniform(-2.5, -1.0, n), 14), g["ksat" + s])
        else:                                                         # barrier soil below it
            g["lay_type" + s] = np.where(pair, 3, g["lay_type" + s])
            g["thick" + s] = np.where(pair, rng.integers(30, 91, n), g["thick" + s]).astype(float)
            g["poro" + s] = np.where(pair, 0.427, g["poro" + s]); g["fc" + s] = np.where(pair, 0.418, g["fc" + s])
            g["wp" + s] = np.where(pair, 0.367, g["wp" + s])
            g["ksat" + s] = np.where(pair, np.round(10 ** rng.uniform(-7.0, -6.0, n), 14), g["ksat" + s])
    return g


def assign_nodes(n_cells: int, n_nodes: int, rng=None) -> np.ndarray:
    """Cells in contiguous blocks per weather node (a gridded-weather layout:
    ~n_cells/n_nodes neighbouring cells share a node)."""
    return np.minimum((np.arange(n_cells, dtype=np.int64) * n_nodes) // max(n_cells, 1), n_nodes - 1).astype(np.int32)


# ---------------------------------------------------------------------------
# text writers (the reference's exchange formats)
# ---------------------------------------------------------------------------
def write_weather_files(workdir: str, wx: dict, nodes=None) -> dict:
    """D4/D7/D13 files for the given node indices, formatted like
    pyhelp/weather_reader.py:22-104.  Returns {'D4': [paths], 'D7': [...], 'D13': [...]}."""
    os.makedirs(workdir, exist_ok=True)
    yod = wx["years_of_day"]
    nn = wx["precip"].shape[1]
    nodes = range(nn) if nodes is None else nodes
    out = {"D4": {}, "D7": {}, "D13": {}}
    specs = (("D4", "precip", "{0:>10}", "{0:>5.1f}"), ("D7", "airtemp", "{0:>5}", "{0:>6.1f}"),
             ("D13", "solrad", "{0:>5}", "{0:>6.2f}"))
    for ext, key, yfmt, vfmt in specs:
        for k in nodes:
            lines = ["{0:>2}".format(3), "{0:>2}".format(2), "{0:<40}".format(f"{key} node {k}"[:40])]
            lines.append("{0:>6.2f}".format(float(wx["lat"][k])) if ext == "D13" else "")
            for y in np.unique(yod):
                v = wx[key][yod == y, k]
                v = np.hstack([v, np.zeros(10 - len(v) % 10)]).reshape(37, 10)
                for row in v.tolist():
                    lines.append(yfmt.format(int(y)) + "".join(vfmt.format(x) for x in row))
            path = osp.join(workdir, f"node{k}.{ext}")
            with open(path, "w") as f:
                f.write("\n".join(lines) + "\n")
            out[ext][k] = path
    return out


def format_d11(g: dict, i: int, sf_edepth=1.0, sf_ulai=1.0) -> list:
    """preprocessing.py:27-76 for cell i."""
    edepth = min(max(float(g["EZD"][i]) * sf_edepth, 3), 80)
    return ["{0:>2}".format(2), "{0:<40}".format(str(g["cid"][i])),
            "{0:<10.2f}".format(float(g["lat_dd"][i])) + "{0:>4}".format(int(g["growth_start"][i])) +
            "{0:>4}".format(int(g["growth_end"][i])) + "{0:>7.2f}".format(float(g["LAI"][i]) * sf_ulai) +
            "{0:>8.1f}".format(edepth) + "{0:>5.1f}".format(float(g["wind"][i])) +
            "".join("{0:>5.1f}".format(float(g[f"hum{q}"][i])) for q in (1, 2, 3, 4))]


def format_d10(g: dict, i: int, sf_cn=1.0, ex: dict | None = None) -> list:
    """preprocessing.py:79-185 for cell i (plus the optional liner fields for landfill designs)."""
    ex = ex or {}
    ipre = int(ex["ipre"][i]) if "ipre" in ex else 0
    osno = float(ex["osno"][i]) if "osno" in ex else 0.0
    out = ["{0:<60}".format(str(g["cid"][i])),
           "{0:>2}".format(2) + "{0:>2}".format(ipre) + "{0:>10.0f}".format(osno) + "{0:>10.0f}".format(1) +
           "{0:>6.0f}".format(100) + "{0:>2}".format(1),
           "{0:>7.0f}".format(float(g["CN"][i]) * sf_cn)]
    for j in range(int(g["nlayer"][i])):
        s = str(j + 1)
        thick = max(float(g["thick" + s][i]), 10)
        thick_txt = "{0:>7.0f}".format(thick)
        if "thick_exact" + s in ex and float(ex["thick_exact" + s][i]) > 0:
            thick_txt = "{0:>7.2f}".format(float(ex["thick_exact" + s][i]))
        out.append("{0:>2}".format(int(g["lay_type" + s][i])) + thick_txt + "{0:>4}".format(0) +
                   "{0:>6.3f}".format(float(g["poro" + s][i])) + "{0:>6.3f}".format(float(g["fc" + s][i])) +
                   "{0:>6.3f}".format(float(g["wp" + s][i])) +
                   ("{0:>6.3f}".format(float(ex["sw" + s][i])) if "sw" + s in ex and float(ex["sw" + s][i]) > 0 else "{0:>6}".format("")) +
                   "{0:>16.14f}".format(float(g["ksat" + s][i])))
        def opt(name, fmt, blank_w):
            if name + s in ex:
                return fmt.format(ex[name + s][i])
            return " " * blank_w
        out.append("{0:>7.0f}".format(float(g["dist_dr" + s][i])) + "{0:>6.2f}".format(float(g["slope" + s][i])) +
                   opt("recir", "{0:>6.0f}", 6) + ("{0:>3d}".format(int(ex["layr" + s][i])) if "layr" + s in ex else "{0:>3}".format(0)) +
                   opt("subin", "{0:>13.2f}", 13) +
                   opt("phole", "{0:>7.0f}", 7) + opt("defec", "{0:>7.0f}", 7) + opt("ipq", "{0:>2d}", 2) +
                   opt("trans", "{0:>14.6g}", 14))
    return out


def cellparams_from_grid(g: dict, files: dict, node_p, node_t, node_r, n_years: int, tfsoil_c: float = 0.0,
                         sf_edepth=1.0, sf_ulai=1.0, sf_cn=1.0, ex=None, cells=None, workdir="/tmp") -> dict:
    """The dict `HelpManager.calc_help_cells` hands to run_help_allcells (managers.py:393-428)."""
    cells = range(len(g["cid"])) if cells is None else cells
    tfsoil = (tfsoil_c * 1.8) + 32
    out = {}
    for i in cells:
        d10 = np.char.ljust(np.array([[x] for x in format_d10(g, i, sf_cn, ex)]), 80).astype("S80")
        d11 = np.char.ljust(np.array([[x] for x in format_d11(g, i, sf_edepth, sf_ulai)]), 80).astype("S80")
        name = str(g["cid"][i])
        out[name] = (files["D4"][int(node_p[i])], files["D7"][int(node_t[i])], files["D13"][int(node_r[i])],
                     d11, d10, osp.join(workdir, name + ".OUT"), 0, 1, 0, 0, int(n_years), tfsoil)
    return out
