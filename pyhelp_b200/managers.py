"""The calculation path of the reference's ``pyhelp/managers.py`` on the B200 engine.

Mirrors, for the hot path only:

* :func:`load_grid_from_csv` -- ``pyhelp/managers.py:597-613``
* :func:`load_weather_from_csv` -- ``pyhelp/managers.py:616-656``
* :class:`HelpManager` with ``calc_help_cells(cellnames, tfsoil, sf_edepth, sf_ulai, sf_cn)``
  -- ``pyhelp/managers.py:342-446``: nearest weather node per cell and variable
  (``_generate_d4d7d13_input_files``, ``:280-340``), D10/D11 quantisation
  (``_generate_d10d11_input_data`` -> ``pyhelp/preprocessing.py:27-185``), the simulation of every
  runnable cell, and ``_post_process_output`` (``:448-506``).

What the reference does through text files and one process-pool task per cell happens here in one
GPU batch: nearest nodes by ``help_nodes.cuh``, packing by :func:`pyhelp_b200.grid.batch_from_grid`
(no D4/D7/D13/D10/D11 files are written), the daily water balance by ``help_sim_kernel``, the area
and per-cell means by ``help_post.cuh``.  The returned :class:`HelpOutput` carries the
``HelpOutput.data`` dict of the reference (``pyhelp/output.py:37-60``) unchanged.  HDF5, csv export
and plots are out of scope.  There is no CPU path.
"""
from __future__ import annotations

import csv

import numpy as np

from . import grid as G
from . import nodes
from . import processing

KEYS = ("precip", "runoff", "evapo", "perco", "subrun1", "subrun2", "rechg")


def load_grid_from_csv(path_togrid: str):
    """The grid table as a pandas DataFrame indexed by ``cid`` (kept as a column too), cell ids read as
    strings; the five columns the reference insists on (``pyhelp/managers.py:597-613``) must be there."""
    import pandas as pd
    grid = pd.read_csv(path_togrid, dtype={"cid": "str"})
    missing = [k for k in ("cid", "lat_dd", "lon_dd", "run", "context") if k not in grid.columns]
    if missing:
        raise KeyError("No attribute '{}' found in {}".format(missing[0], path_togrid))
    return grid.set_index("cid", drop=False)


def load_weather_from_csv(filename: str):
    """Daily weather table of a PyHELP weather input file (same result as the reference's loader,
    ``pyhelp/managers.py:616-656``): a DataFrame indexed by date whose columns are the
    ``(lat_dd, lon_dd)`` of the nodes.  File layout: free-text header lines, a ``Latitude (dd),...``
    and a ``Longitude (dd),...`` line, then one ``dd/mm/yyyy,v1,v2,...`` line per day."""
    import pandas as pd
    coords = {}
    days, values = [], []
    with open(filename, "r", newline="") as f:
        for fields in csv.reader(f):
            if not fields or not fields[0].strip():
                continue
            tag = fields[0].strip().lower()
            if tag[:3] in ("lat", "lon") and len(coords) < 2:
                coords[tag[:3]] = np.array([float(x) for x in fields[1:] if x.strip()])
            elif len(coords) == 2 and tag[:1].isdigit():
                days.append(fields[0].strip())
                values.append([float(x) if x.strip() else np.nan for x in fields[1:]])
    if len(coords) < 2:
        raise ValueError(f"{filename}: no latitude/longitude header lines")
    nn = coords["lat"].size
    if coords["lon"].size != nn or any(len(v) != nn for v in values):
        raise ValueError(f"{filename}: rows do not all have one value per node ({nn})")
    index = pd.DatetimeIndex(pd.to_datetime(days, dayfirst=True), name="date")
    columns = pd.MultiIndex.from_arrays([coords["lat"], coords["lon"]], names=["lat_dd", "lon_dd"])
    return pd.DataFrame(np.asarray(values, dtype=np.float64).reshape(len(days), nn), index=index, columns=columns)


class HelpOutput:
    """``data`` = the dict ``HelpManager._post_process_output`` builds (``pyhelp/managers.py:448-506``);
    the two reductions ``pyhelp/output.py`` computes from it come from the device-resident monthly
    buffer (``help_post.cuh``), evaluated once while the engine is alive."""

    def __init__(self, data: dict, area_monthly_avg: dict, cells_yearly_avg: dict):
        self.data = data
        self._area, self._cells = area_monthly_avg, cells_yearly_avg

    def calc_area_monthly_avg(self) -> dict:
        """``HelpOutput.calc_area_monthly_avg`` (``pyhelp/output.py:127-152``): {var: [Ny, 12]} mm."""
        return self._area

    def calc_area_yearly_avg(self) -> dict:
        """``HelpOutput.calc_area_yearly_avg`` (``pyhelp/output.py:154-169``): {var: [Ny]} mm/yr."""
        return {k: np.sum(v, axis=1) for k, v in self._area.items()}

    def calc_cells_yearly_avg(self, year_from=-np.inf, year_to=np.inf) -> dict:
        """``HelpOutput.calc_cells_yearly_avg`` (``pyhelp/output.py:171-203``): {var: [Np]} mm/yr, averaged
        over the years of ``[year_from, year_to]``.  The default window is the device-side reduction
        computed while the engine was alive; a narrower one is evaluated from ``data`` like the
        reference does."""
        years = np.asarray(self.data["years"])
        mask = (years >= year_from) & (years <= year_to)
        if mask.all():
            return self._cells
        return {k: np.mean(np.sum(self.data[k][:, mask, :], axis=2), axis=1) for k in KEYS}


def _canonical_nodes(node: np.ndarray, lat: np.ndarray, lon: np.ndarray, fext: str) -> np.ndarray:
    """The reference names a node's HELP input file '{lat:3.1f}_{lon:3.1f}.{ext}' and writes it only
    if it does not exist yet (``pyhelp/managers.py:296,317-331``): nodes whose coordinates round to the
    same name share the file of the first one a cell asked for.  Returns the node whose data each
    cell really reads."""
    names = np.array(["{:3.1f}_{:3.1f}.{}".format(a, o, fext) for a, o in zip(lat, lon)])
    _, name_id = np.unique(names, return_inverse=True)
    if np.unique(name_id).size == names.size:
        return node
    cell_name = name_id[node]
    _, first_cell = np.unique(cell_name, return_index=True)       # first cell (in run order) per file name
    owner = np.full(name_id.max() + 1, -1, dtype=np.int64)
    owner[cell_name[first_cell]] = node[first_cell]
    return owner[cell_name].astype(np.int32)


class HelpManager:
    """Grid + weather tables and ``calc_help_cells`` (reference ``HelpManager``, calculation path only)."""

    def __init__(self, workdir=None, path_to_grid=None, path_to_precip=None, path_to_airtemp=None,
                 path_to_solrad=None):
        self.workdir = workdir
        self.grid = load_grid_from_csv(path_to_grid) if path_to_grid else None
        self.precip_data = load_weather_from_csv(path_to_precip) if path_to_precip else None
        self.airtemp_data = load_weather_from_csv(path_to_airtemp) if path_to_airtemp else None
        self.solrad_data = load_weather_from_csv(path_to_solrad) if path_to_solrad else None
        self.connect_tables = {}

    @property
    def cellnames(self):
        return [] if self.grid is None else self.grid["cid"].tolist()

    def get_run_cellnames(self, cellnames=None):
        """Cells of ``cellnames`` that are in the grid, flagged ``run == 1`` and have layers
        (``pyhelp/managers.py:564-584``), in grid order."""
        g = self.grid
        keep = np.ones(len(g), dtype=bool) if cellnames is None else np.array(g["cid"].isin(list(cellnames)).values)
        keep = keep & (g["run"].values == 1) & (g["nlayer"].values > 0)
        return g["cid"].values[keep].tolist()

    def calc_help_cells(self, path_to_hdf5=None, cellnames=None, tfsoil=0, sf_edepth: float = 1,
                        sf_ulai: float = 1, sf_cn: float = 1, build_help_input_files: bool = False) -> HelpOutput:
        """Monthly water budget of every runnable cell of ``cellnames`` (``pyhelp/managers.py:342-446``).
        ``tfsoil`` in deg C.  Cells with a -9999 layer field are skipped with the reference's warning."""
        if path_to_hdf5 or build_help_input_files:
            # kept in the signature so that reference call sites bind; the reference's own HelpManager serves
            # these two options (patch_pyhelp() routes only run_help_allcells here)
            raise NotImplementedError("HDF5 output and D10/D11 debug files are outside the hot path (DESIGN.md section 8)")
        for name, v in (("sf_edepth", sf_edepth), ("sf_ulai", sf_ulai), ("sf_cn", sf_cn)):
            if v < 0:                                             # pyhelp/managers.py:246-251
                raise ValueError(f"{name} value cannot be lower than 0.")
        if self.grid is None or self.precip_data is None or self.airtemp_data is None or self.solrad_data is None:
            raise ValueError("grid, precip, airtemp and solrad tables are required")
        run = self.get_run_cellnames(cellnames)
        sub = self.grid.loc[run]
        cols = {c: sub[c].values for c in sub.columns}
        ok = G.runnable_cells(cols)
        skipped = [c for c, k in zip(run, ok) if not k]
        if skipped:
            print("-" * 25)
            print("Warning: calcul for " + ("cell " if len(skipped) == 1 else "cells ") + ", ".join(skipped) +
                  " will be skipped due to problems with the input data.")
            print("-" * 25)
        cols = {c: v[ok] for c, v in cols.items()}
        names = [c for c, k in zip(run, ok) if k]
        if not names:
            raise ValueError("no runnable cell")
        years_of_day = self.precip_data.index.year.values
        for name, tbl in (("airtemp", self.airtemp_data), ("solrad", self.solrad_data)):
            if not np.array_equal(tbl.index.year.values, years_of_day):
                raise ValueError(f"{name} does not cover the same days as precip")       # (F:1130-1137 warns and goes on)
        node = {}
        for var, fext, tbl in (("precip", "D4", self.precip_data), ("airtemp", "D7", self.airtemp_data),
                               ("solrad", "D13", self.solrad_data)):
            lat = tbl.columns.get_level_values("lat_dd").values.astype(float)
            lon = tbl.columns.get_level_values("lon_dd").values.astype(float)
            nearest = nodes.nearest_nodes(cols["lat_dd"], cols["lon_dd"], lat, lon)
            self.connect_tables[var] = dict(zip(names, nearest.tolist()))
            node[var] = _canonical_nodes(nearest, lat, lon, fext)
        batch = G.batch_from_grid(cols, years_of_day, self.precip_data.values, self.airtemp_data.values,
                                  self.solrad_data.values, node["precip"], node["airtemp"], node["solrad"],
                                  tfsoil_c=tfsoil, sf_edepth=sf_edepth, sf_ulai=sf_ulai, sf_cn=sf_cn, weather="resident")
        context = np.asarray(cols["context"]).astype(np.int32)
        with processing.HelpEngine(batch, monthly=True) as eng:
            eng.simulate()
            res = eng.fetch(monthly=True, yearly=False)
            red = eng.post_process(context=context)
        data = processing.post_process_monthly(res["monthly"], res["years"], names, context, cols["lat_dd"], cols["lon_dd"])
        return HelpOutput(data, red["area_monthly_avg"], red["cells_yearly_avg"])
