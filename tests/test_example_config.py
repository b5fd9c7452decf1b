"""BASELINE configs[0]: the reference's example (example/help_example.py) -- the example weather
tables (24 precipitation nodes, 24 air-temperature nodes, 1 solar-radiation node, 2000-2010) with
a schema-exact stand-in for the example grid (its blob is missing from the reference checkout,
SURVEY.md section 8d), through the mirror of ``HelpManager.calc_help_cells``
(``pyhelp_b200/managers.py`` <- reference ``pyhelp/managers.py:342-506``).

The weather comes from ``tests/golden/example_weather.npz`` (made from the reference's csv files by
``tests/golden/make_example_weather.py``); nothing here reads /root/reference.
"""
import os
import os.path as osp

import numpy as np
import pytest

from pyhelp_b200 import managers, processing
from support import synth
from oracle import nodes_oracle, oracle, post_oracle

GOLD = osp.join(osp.dirname(osp.abspath(__file__)), "golden", "example_weather.npz")
KEYS = ("precip", "runoff", "evapo", "perco", "subrun1", "subrun2", "rechg")


def _tables():
    pd = pytest.importorskip("pandas")
    z = np.load(GOLD)
    idx = pd.to_datetime({"year": z["year"].astype(int), "month": z["month"].astype(int), "day": z["day"].astype(int)})
    out = {}
    for var in ("precip", "airtemp", "solrad"):
        cols = pd.MultiIndex.from_tuples(list(zip(z[var + "_lat"], z[var + "_lon"])), names=["lat_dd", "lon_dd"])
        out[var] = pd.DataFrame(z[var].astype(np.float64), index=pd.DatetimeIndex(idx, name="date"), columns=cols)
    return z, out


def _manager(n_cells):
    pd = pytest.importorskip("pandas")
    z, t = _tables()
    g = synth.grid_example(n_cells)
    hm = managers.HelpManager()
    hm.grid = pd.DataFrame(g).set_index("cid", drop=False)
    hm.precip_data, hm.airtemp_data, hm.solrad_data = t["precip"], t["airtemp"], t["solrad"]
    return hm, g, z


# ---- host logic (CPU) ---------------------------------------------------------------------------
def test_example_weather_fixture_shape():
    z = np.load(GOLD)
    assert z["precip"].shape == (4018, 24) and z["airtemp"].shape == (4018, 24) and z["solrad"].shape == (4018, 1)
    assert int(z["year"].min()) == 2000 and int(z["year"].max()) == 2010
    assert float(z["solrad_lat"][0]) == 45.47 and float(z["solrad_lon"][0]) == -73.74       # MONTREAL INTL A
    # stored at the precision the reference prints into the HELP input files
    assert np.array_equal(np.round(z["precip"].astype(np.float64), 1).astype(np.float32), z["precip"])


def test_load_weather_from_csv_reads_the_pyhelp_format(tmp_path):
    """Header block, 'Latitude (dd)' / 'Longitude (dd)' lines, day-first dates (managers.py:616-656)."""
    z = np.load(GOLD)
    head = z["precip_head_exact"]
    p = tmp_path / "precip_input_data.csv"
    with open(p, "w") as f:
        f.write("Precipitation in mm\n,\nCreated by PyHelp 0.1.0.dev\nCreated on 18/10/2018\nCreated from MDDELCC grid\n,\n")
        f.write("Latitude (dd)," + ",".join(repr(float(v)) for v in z["precip_lat"]) + "\n")
        f.write("Longitude (dd)," + ",".join(repr(float(v)) for v in z["precip_lon"]) + "\n,\n")
        for k, row in enumerate(head):
            f.write("%02d/%02d/%04d," % (z["day"][k], z["month"][k], z["year"][k]) + ",".join(repr(float(v)) for v in row) + "\n")
    t = managers.load_weather_from_csv(str(p))
    assert t.shape == head.shape and np.array_equal(t.values, head)
    assert np.array_equal(t.columns.get_level_values("lat_dd").values, z["precip_lat"])
    assert np.array_equal(t.columns.get_level_values("lon_dd").values, z["precip_lon"])
    assert t.index[0].day == 1 and t.index[0].month == 1 and t.index[31].month == 2           # day first
    assert t.index.name == "date"
    with pytest.raises(ValueError):
        q = tmp_path / "bad.csv"
        q.write_text("01/01/2000,1.0\n")
        managers.load_weather_from_csv(str(q))


def test_load_grid_from_csv_requires_the_reference_columns(tmp_path):
    pd = pytest.importorskip("pandas")
    g = synth.grid_example(50)
    p = tmp_path / "input_grid.csv"
    pd.DataFrame(g).to_csv(p, index=False)
    grid = managers.load_grid_from_csv(str(p))
    assert grid.index.name == "cid" and isinstance(grid["cid"].iloc[7], str) and grid.loc["7"]["cid"] == "7"
    pd.DataFrame({k: v for k, v in g.items() if k != "context"}).to_csv(p, index=False)
    with pytest.raises(KeyError):
        managers.load_grid_from_csv(str(p))


def test_example_grid_schema_and_run_cells():
    hm, g, _ = _manager(400)
    for col in ("cid", "lat_dd", "lon_dd", "run", "context", "wind", "hum1", "hum4", "growth_start", "growth_end", "LAI",
                "EZD", "CN", "nlayer", "lay_type3", "thick1", "poro1", "fc1", "wp1", "ksat1", "dist_dr1", "slope1"):
        assert col in g
    assert g["lat_dd"].min() >= 45.78 and g["lat_dd"].max() <= 46.2 and g["lon_dd"].min() >= -74.5 and g["lon_dd"].max() <= -74.0
    run = hm.get_run_cellnames()
    want = [c for c, r, nl in zip(g["cid"], g["run"], g["nlayer"]) if r == 1 and nl > 0]
    assert run == want and 0 < len(run) < 400
    some = hm.get_run_cellnames(["3", "5", "not-a-cell"])
    assert set(some) <= {"3", "5"}


def test_post_process_monthly_equals_the_reference_loop():
    rng = np.random.default_rng(3)
    n, ny = 9, 2
    monthly = rng.uniform(0, 50, (7, ny, 12, n)).astype(np.float32)
    monthly[4, :, :, :5] = 0                    # cells 0-4 have no deep subsurface runoff
    monthly[:, 1, 3, 8] = np.nan                # a cell HELP could not compute
    context = np.array([2, 1, 2, 0, 1, 2, 2, 1, 2])
    names = [f"c{i}" for i in range(n)]
    rows = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")
    per_cell = {nm: dict(years=np.arange(2000, 2000 + ny, dtype=np.uint16),
                         **{k: monthly[r, :, :, i] for r, k in enumerate(rows)}) for i, nm in enumerate(names)}
    want = post_oracle.post_process_output(per_cell, context)
    got = processing.post_process_monthly(monthly, np.arange(2000, 2000 + ny), names, context, np.zeros(n), np.zeros(n))
    for k in KEYS:
        assert got[k].dtype == np.float64 and got[k].shape == (n, ny, 12)
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    assert got["idx_nan"] == want["idx_nan"] == [8] and got["cid"] == names


def test_nodes_that_round_to_one_file_name_share_the_first_one():
    """'{lat:3.1f}_{lon:3.1f}.D4' is written once (managers.py:296,317-331)."""
    lat = np.array([46.04, 45.96, 46.30])
    lon = np.array([-74.0, -74.0, -74.0])                     # nodes 0 and 1 both print as 46.0_-74.0
    node = np.array([1, 0, 2, 1, 0], dtype=np.int32)
    assert managers._canonical_nodes(node, lat, lon, "D4").tolist() == [1, 1, 2, 1, 1]
    lat2 = np.array([46.0, 45.9, 46.3])
    assert managers._canonical_nodes(node, lat2, lon, "D4") is node


# ---- the example run (GPU) ----------------------------------------------------------------------
@pytest.mark.gpu
def test_example_config_end_to_end(tmp_path):
    """help_example.py's call -- cells of the sub-basin, tfsoil=-3, sf_edepth=0.15, sf_cn=1.15 -- on the
    full-size grid: nearest nodes equal the reference formula, a sample of cells equals the CPU oracle
    run through the reference's own text formats bit for bit, the reductions equal numpy on the data."""
    hm, g, z = _manager(17_178)
    sel = [c for c, b in zip(g["cid"], g["Bassin"]) if b == 1]
    out = hm.calc_help_cells(cellnames=sel, tfsoil=-3, sf_edepth=0.15, sf_ulai=1, sf_cn=1.15)
    data = out.data
    names = data["cid"]
    assert names == hm.get_run_cellnames(sel) and 8_000 < len(names) < 11_000
    assert np.array_equal(data["years"], np.arange(2000, 2011))
    # the stand-in grid draws fc and wp independently, so a few soils are inconsistent and HELP itself
    # returns NaN for them (Brooks-Corey set-up, F:1307-1340) -- they exercise the idx_nan bookkeeping
    nan_idx = data["idx_nan"]
    assert 0 < len(nan_idx) < 0.01 * len(names)
    good = np.ones(len(names), bool)
    good[nan_idx] = False
    idx = np.array([int(c) for c in names])
    assert np.array_equal(data["lat_dd"], g["lat_dd"][idx])
    for k in KEYS:
        assert data[k].shape == (len(names), 11, 12) and data[k].dtype == np.float64 and np.isfinite(data[k][good]).all()
    assert np.isnan(data["rechg"][nan_idx]).any(axis=(1, 2)).all()
    # nearest node per variable (connect_tables, managers.py:307-313)
    for var in ("precip", "airtemp", "solrad"):
        want = nodes_oracle.nearest_nodes(g["lat_dd"][idx], g["lon_dd"][idx], z[var + "_lat"], z[var + "_lon"])
        assert [hm.connect_tables[var][c] for c in names] == want.tolist()
    assert len(set(hm.connect_tables["precip"].values())) >= 20            # the grid spans the node lattice
    # a sample of cells through the oracle, fed by the reference's text formats
    rng = np.random.default_rng(5)
    pick = np.unique(np.concatenate([rng.choice(len(names), 40, replace=False), np.asarray(nan_idx[:8], dtype=np.int64)]))
    yod = z["year"].astype(np.int32)
    files = {}
    for var, ext in (("precip", "D4"), ("airtemp", "D7"), ("solrad", "D13")):
        a = z[var].astype(np.float64)
        wx = {"years_of_day": yod, "precip": a, "airtemp": a, "solrad": a, "lat": z[var + "_lat"]}
        files[ext] = synth.write_weather_files(str(tmp_path / var), wx)[ext]
    cells = idx[pick]
    node = {v: np.zeros(len(g["cid"]), dtype=np.int32) for v in ("precip", "airtemp", "solrad")}
    for v in node:
        node[v][idx] = [hm.connect_tables[v][c] for c in names]
    cp = synth.cellparams_from_grid(g, files, node["precip"], node["airtemp"], node["solrad"], 11, tfsoil_c=-3.0,
                                    sf_edepth=0.15, sf_ulai=1.0, sf_cn=1.15, cells=cells.tolist(), workdir=str(tmp_path))
    ref = oracle.run_help_allcells(cp, ncore=1)
    want = post_oracle.post_process_output(ref, g["context"][cells].tolist())
    for k in KEYS:
        assert np.array_equal(data[k][pick], want[k], equal_nan=True), k
    assert [int(pick[i]) for i in want["idx_nan"]] == [i for i in nan_idx if i in set(pick.tolist())]
    # the reductions of pyhelp/output.py on the whole run
    area, cyr = out.calc_area_monthly_avg(), out.calc_cells_yearly_avg()
    want_area, want_cyr = post_oracle.calc_area_monthly_avg(data), post_oracle.calc_cells_yearly_avg(data)
    for k in KEYS:
        np.testing.assert_allclose(area[k], want_area[k], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(cyr[k], want_cyr[k], rtol=1e-12, atol=1e-9)
    # water balance of the area over the 11 years: P = RO + ET + subsurface runoff + recharge + storage change
    tot = {k: out.calc_area_yearly_avg()[k].sum() for k in KEYS}
    resid = tot["precip"] - tot["runoff"] - tot["evapo"] - tot["subrun1"] - tot["subrun2"] - tot["rechg"]
    assert abs(resid) < 0.02 * tot["precip"]
    assert (data["rechg"][(np.asarray(g["context"])[idx] == 2) & good] == 0).all()


@pytest.mark.gpu
def test_calc_help_cells_skips_cells_with_missing_layer_data(capsys):
    hm, g, _ = _manager(300)
    run = hm.get_run_cellnames()
    bad = run[3]
    hm.grid.loc[bad, "poro1"] = -9999
    out = hm.calc_help_cells(tfsoil=-3)
    assert bad not in out.data["cid"] and len(out.data["cid"]) == len(run) - 1
    assert "will be skipped due to problems with the input data" in capsys.readouterr().out
