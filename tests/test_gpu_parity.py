"""Parity of the CUDA engine with the CPU oracle, through the C ABI (`-m gpu`).

Integer/index work and -- because both sides use the deterministic transcendentals of
help_math.h -- also the floating-point results are compared BIT FOR BIT: every monthly value
(float32, mm, 0.001 mm quantised like the reference's F7.3 text) must be equal.  The
tolerance `north_star` allows (1e-3 relative or 0.1 mm/yr on yearly totals) is therefore met
with zero difference; tests/test_libm_sensitivity.py documents what a different libm does."""
import os
import os.path as osp

import numpy as np
import pytest

from oracle import oracle
from pyhelp_b200 import grid as G
from pyhelp_b200 import inputs, processing
from support import synth
from pyhelp_b200 import _native as N

pytestmark = pytest.mark.gpu
MM2IN = 0.0393701


def rcra_params(rcra_dir, ny=3, tfsoil=32.0):
    d10 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    return (osp.join(rcra_dir, "RCRA.D4"), osp.join(rcra_dir, "RCRA.D7"), osp.join(rcra_dir, "RCRA.D13"),
            d11, d10, "/tmp/rcra.OUT", 0, 1, 0, 0, ny, tfsoil)


def assert_bit_equal(got: dict, want: dict):
    assert list(got.keys()) == list(want.keys()) or set(got) == set(want)
    bad = []
    for name, w in want.items():
        assert np.array_equal(got[name]["years"], w["years"])
        for k in processing.KEYS:
            assert got[name][k].dtype == np.float32 and got[name][k].shape == w[k].shape
            if not np.array_equal(got[name][k], w[k], equal_nan=True):
                bad.append((name, k, float(np.nanmax(np.abs(got[name][k] - w[k])))))
    assert not bad, bad[:10]


def test_rcra_known_answers_and_bits(rcra_dir):
    """the reference's own test case (pyhelp/tests/test_help3o.py:73-118) on the GPU"""
    cp = {"rcra": rcra_params(rcra_dir)}
    got = processing.run_help_allcells(cp)
    want = oracle.run_help_allcells(cp, ncore=1)
    assert_bit_equal(got, want)
    kat = {"precip": [48.53, 58.32, 56.71], "runoff": [0.596, 4.130, 2.397], "evapo": [32.178, 33.420, 36.540],
           "subrun1": [15.4971, 22.8698, 19.1360], "subrun2": [0.0543 + 0.0833, 0.1481 + 0.1658, 0.1835 + 0.1935],
           "rechg": [0.000004, 0.000006, 0.000007]}
    for k, v in kat.items():
        assert np.abs(got["rcra"][k].sum(axis=1) * MM2IN - np.array(v)).max() < 0.1, k
    assert got["rcra"]["years"].tolist() == [1, 2, 3] and got["rcra"]["years"].dtype == np.uint16
    name, single = processing.run_help_singlecell(("rcra", cp["rcra"]))
    assert name == "rcra" and np.array_equal(single["rechg"], got["rcra"]["rechg"])


@pytest.mark.parametrize("family,n,ny", [("natural3", 160, 3), ("mixed", 160, 3), ("landfill", 96, 3), ("recirc", 96, 3),
                                         ("deep", 96, 2), ("special", 192, 3)])
def test_grid_families_bit_equal(tmp_path, family, n, ny):
    """text inputs exactly as HelpManager.calc_help_cells hands them over (managers.py:393-428);
    "deep" = the largest profiles HELP accepts (up to 20 layers / 67 segments: the widest kernel tier);
    "special" = the branches no PyHELP grid reaches: geotextile-cushioned geomembranes (FML design cases 5/6,
    HELP3O.FOR:5177-5285), placement qualities 1 and 5, subsurface inflow into the bottom subprofile (IBAR 4/5
    and the REAL*8 inflow sums of HELP3O.FOR:2060-2071), IPRE = 1 with initial moisture and snow given, sites
    at 52-65 deg N and 35 deg S with a growing season across the new year"""
    maker = {"natural3": synth.grid_natural_3layer, "mixed": synth.grid_mixed, "landfill": synth.grid_landfill,
             "recirc": lambda n, seed: synth.grid_landfill(n, seed=seed, recirc=True), "deep": synth.grid_deep,
             "special": synth.grid_special}[family]
    res = maker(n, seed=21)
    g, ex = res if isinstance(res, tuple) else (res, None)
    nn = 6
    wx = synth.synth_weather(nn, 2003, ny, seed=9)           # 2004 is a leap year
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(n, nn)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex)
    got = processing.run_help_allcells(cp)
    want = oracle.run_help_allcells(cp)
    assert list(got.keys()) == list(cp.keys())              # deterministic cell order
    assert_bit_equal(got, want)
    # the vectorised packer (no text) feeds the engine the same numbers
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                          tfsoil_c=-3.0, extra_layers=ex)
    r = processing.run_batch(b)
    direct = processing.split_by_cell(b, r)
    assert_bit_equal(direct, want)
    # yearly summaries = sum of the months
    m = r["monthly"]
    y = r["yearly"]
    np.testing.assert_allclose(y[N.HY_PRECIP], m[N.HR_PRECIP].sum(axis=1), rtol=1e-6)
    np.testing.assert_allclose(y[N.HY_RECHG], m[N.HR_RECHG].sum(axis=1), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(y[N.HY_LATDRAIN], (m[N.HR_SUBRUN1] + m[N.HR_SUBRUN2]).sum(axis=1), rtol=1e-6, atol=1e-4)


def test_century_year_mod4_leap_rule(tmp_path):
    """HELP's LEAP is MOD(year,4)==0 (F:1048): 2100 is simulated with 366 days although the
    weather writer gives it 365; engine and oracle must agree on that too."""
    n, ny = 24, 3
    g = synth.grid_natural_3layer(n, seed=5)
    wx = synth.synth_weather(2, 2099, ny, seed=2)
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(n, 2)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=0.0)
    assert_bit_equal(processing.run_help_allcells(cp), oracle.run_help_allcells(cp))


def test_ragged_sizes_and_order_independence():
    """1, 31, 33, 65 cells (partial warps/blocks); results do not depend on batch composition,
    on the cell order, or on the class sort the engine applies internally."""
    n, ny = 65, 2
    g = synth.grid_mixed(n, seed=3)
    wx = synth.synth_weather(3, 2001, ny, seed=4)
    node = synth.assign_nodes(n, 3)
    full = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0)
    ref = processing.run_batch(full)["monthly"]
    for m in (1, 31, 33):
        sub = processing.run_batch(full.take(np.arange(m)))["monthly"]
        assert np.array_equal(sub, ref[..., :m], equal_nan=True), m
    perm = np.random.default_rng(0).permutation(n)
    shuf = processing.run_batch(full.take(perm))["monthly"]
    assert np.array_equal(shuf, ref[..., perm], equal_nan=True)
    os.environ["HELP_B200_SORT"] = "0"
    try:
        unsorted = processing.run_batch(full)["monthly"]
    finally:
        del os.environ["HELP_B200_SORT"]
    assert np.array_equal(unsorted, ref, equal_nan=True)
    again = processing.run_batch(full)["monthly"]
    assert np.array_equal(again, ref, equal_nan=True)            # run-to-run deterministic


def test_bad_cells_are_flagged_not_fatal():
    n, ny = 8, 1
    g = synth.grid_natural_3layer(n, seed=1)
    wx = synth.synth_weather(1, 2001, ny, seed=1)
    node = np.zeros(n, np.int32)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)
    b.cell_i32[N.HB_NLAY, 2] = 0               # no layers
    b.cell_i32[N.HB_NODE_T, 5] = 7             # weather node out of range
    r = processing.run_batch(b)
    assert r["flags"][2] & N.FLAG_BADCELL and r["flags"][5] & N.FLAG_BADCELL
    assert np.isnan(r["monthly"][:, :, :, 2]).all() and np.isnan(r["monthly"][:, :, :, 5]).all()
    ok = np.setdiff1d(np.arange(n), [2, 5])
    assert (r["flags"][ok] == 0).all() and np.isfinite(r["monthly"][..., ok]).all()
    assert processing.run_help_allcells({}) == {}


def test_post_process_layout(tmp_path):
    """HelpManager._post_process_output layout (managers.py:448-506): (Np, Ny, 12) float64,
    context==2 recharge moved to subsurface runoff."""
    n, ny = 12, 2
    g = synth.grid_mixed(n, seed=8)
    wx = synth.synth_weather(2, 2001, ny, seed=6)
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(n, 2)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0)
    out = processing.run_help_allcells(cp)
    ctx = [2 if i % 3 == 0 else 1 for i in range(n)]
    data = processing.post_process_output(out, context=ctx, lat_dd=g["lat_dd"], lon_dd=g["lon_dd"])
    for k in ("precip", "runoff", "evapo", "perco", "subrun1", "subrun2", "rechg"):
        assert data[k].shape == (n, ny, 12) and data[k].dtype == np.float64
    assert data["cid"] == list(cp.keys()) and data["idx_nan"] == []
    for i, name in enumerate(cp):
        if ctx[i] == 2:
            assert (data["rechg"][i] == 0).all()
            moved = data["subrun1"][i] + data["subrun2"][i] - out[name]["subrun1"] - out[name]["subrun2"]
            np.testing.assert_allclose(moved, out[name]["rechg"], atol=1e-4)
        else:
            assert np.array_equal(data["rechg"][i], out[name]["rechg"].astype(np.float64))


def test_full_size_properties_and_sampled_bits(tmp_path):
    """BASELINE configs[1] at full size (100 000 cells x 30 yr): size-independent properties on
    every cell, and a random sample of cells re-run on the CPU oracle, bit for bit."""
    n, ny, nn = 100_000, 30, 1000
    g = synth.grid_natural_3layer(n, seed=1)
    wx = synth.synth_weather(nn, 2001, ny, seed=5)
    node = synth.assign_nodes(n, nn)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0)
    r = processing.run_batch(b)
    m, y = r["monthly"], r["yearly"]
    assert (r["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL | N.FLAG_OVERFLOW)).sum() == 0
    assert np.isfinite(m).all() and (m >= 0).all()
    # precipitation row = the node's weather summed per month (inches in REAL*4, x25.4, F7.3)
    yod = wx["years_of_day"]
    for cell in (0, 4321, 99_999):
        k = node[cell]
        for iy in (0, 17, 29):
            days = wx["precip"][yod == 2001 + iy, k]
            assert abs(float(m[N.HR_PRECIP, iy, :, cell].sum()) - days.sum()) < 0.05
    # no liner: perco == rechg, no lateral drainage; yearly = sum of months
    assert np.array_equal(m[N.HR_PERCO], m[N.HR_RECHG]) and (m[N.HR_SUBRUN1] == 0).all() and (m[N.HR_SUBRUN2] == 0).all()
    np.testing.assert_allclose(y[N.HY_EVAPO], m[N.HR_EVAPO].sum(axis=1), rtol=1e-6)
    # long-run water balance: P = R + ET + recharge + storage change; storage change over 30 yr
    # is bounded by the profile's capacity (< 0.5 m of water), so the residual per year is small
    P, R, E, Q = (y[i].sum(axis=0).astype(np.float64) for i in (N.HY_PRECIP, N.HY_RUNOFF, N.HY_EVAPO, N.HY_RECHG))
    resid = (P - R - E - Q) / ny
    assert np.abs(resid).max() < 60.0 and abs(resid.mean()) < 5.0, (np.abs(resid).max(), resid.mean())
    # sampled cells on the oracle
    sel = np.sort(np.random.default_rng(1).choice(n, 48, replace=False))
    nodes = sorted({int(node[i]) for i in sel})
    files = synth.write_weather_files(str(tmp_path), wx, nodes=nodes)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, cells=sel.tolist())
    want = oracle.run_help_allcells(cp)
    got = processing.split_by_cell(b.take(sel), {"monthly": m[..., sel], "years": r["years"]})
    assert_bit_equal(got, want)


def test_six_subprofiles_and_f73_overflow(tmp_path):
    """the widest report the reference prints (5 liner systems + a bottom subprofile: the sixth percolation row's
    Jul-Dec half stays in inches, HELP3O.FOR:6593-6594) and a month over 999.999 mm ('*******' in the reference's
    text, on which its parser raises; here NaN + HELP_B200_FLAG_OVERFLOW) -- the cases tests/test_outmo2_parser.py
    pins with the reference's own parser"""
    import sys
    sys.path.insert(0, osp.join(osp.dirname(osp.abspath(__file__)), "golden"))
    import make_outmo2_golden as M
    cases = M.cases(str(tmp_path))
    for name in ("six_subprofiles", "overflow", "subin_liner", "geotextile"):
        cp = {name: cases[name]}
        got = processing.run_help_allcells(cp)
        want = oracle.run_help_allcells(cp, ncore=1)
        assert_bit_equal(got, want)
    b = inputs.batch_from_cellparams({"overflow": cases["overflow"]})
    r = processing.run_batch(b)
    assert r["flags"][0] & N.FLAG_OVERFLOW and np.isnan(r["monthly"][N.HR_PRECIP]).sum() == 1


def test_launch_plan_does_not_change_a_bit(monkeypatch):
    """The launch plan of help_sim_kernel (help_capi.cu: register budget by batch size, block size by profile class,
    blocks in descending cost order) decides which SM simulates a cell and with whom it shares its day barrier, never
    what the cell computes: every override gives the bits of the default plan (which the families above hold to the oracle)."""
    n, nn, ny = 5000, 7, 2                                    # 157 blocks of 32: a second, partial round on 148 SMs
    g = synth.grid_mixed(n, seed=4)
    wx = synth.synth_weather(nn, 2003, ny, seed=6)
    node = synth.assign_nodes(n, nn)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0)
    ref = processing.run_batch(b, monthly=True)
    assert int((ref["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL)).astype(bool).sum()) == 0
    plain = G.batch_from_grid(synth.grid_natural_3layer(n, seed=4), wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"],
                              node, node, node, tfsoil_c=-3.0)
    ref_plain = processing.run_batch(plain, monthly=True)
    for env in ({"HELP_B200_BLOCK": "32"}, {"HELP_B200_BLOCK": "512"}, {"HELP_B200_BLOCK": "96", "HELP_B200_SHAPE": "3"},
                {"HELP_B200_SHAPE": "0"}, {"HELP_B200_SHAPE": "1"}, {"HELP_B200_SORT": "0"}, {"HELP_B200_PLAN": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for batch, want in ((b, ref), (plain, ref_plain)):
            got = processing.run_batch(batch, monthly=True)
            for key in ("monthly", "yearly", "flags"):
                assert np.array_equal(got[key], want[key], equal_nan=True), (env, key)
        for k in env:
            monkeypatch.delenv(k)
