"""Device-side result reductions against the numpy restatement of
pyhelp/managers.py:448-506 and pyhelp/output.py:127-203."""
import numpy as np
import pytest

from oracle import post_oracle


def test_oracle_redistributes_recharge_of_stream_cells():
    z = np.zeros((2, 12), np.float32); one = np.ones((2, 12), np.float32)
    out = {
        "a": dict(years=np.array([2001, 2002]), precip=one * 5, runoff=one, evapo=one, perco=one, subrun1=z, subrun2=z, rechg=one * 2),
        "b": dict(years=np.array([2001, 2002]), precip=one * 5, runoff=one, evapo=one, perco=one, subrun1=one, subrun2=one, rechg=one * 3),
        "c": dict(years=np.array([2001, 2002]), precip=one * 5, runoff=one, evapo=one, perco=one, subrun1=z, subrun2=z, rechg=one * np.nan),
    }
    d = post_oracle.post_process_output(out, [2, 2, 2])
    assert d["idx_nan"] == [2]
    assert (d["subrun1"][0] == 2).all() and (d["rechg"][0] == 0).all()          # shallow: subrun2 all zero
    assert (d["subrun2"][1] == 4).all() and (d["subrun1"][1] == 1).all()         # deep otherwise
    a = post_oracle.calc_area_monthly_avg(d)
    assert a["precip"].shape == (2, 12) and np.allclose(a["precip"], 15 / 2)     # Np excludes the NaN cell, nansum keeps its precip
    y = post_oracle.calc_cells_yearly_avg(d, 2002, 2002)
    assert np.allclose(y["subrun1"][:2], [24.0, 12.0])


def test_library_exports_entry_point():
    from pyhelp_b200 import _native as N
    assert hasattr(N.lib(), "help_b200_engine_post_process")


@pytest.mark.gpu
@pytest.mark.parametrize("family,n,ny", [("natural3", 300, 3), ("mixed", 257, 2)])
def test_gpu_reductions_match_oracle(family, n, ny):
    from pyhelp_b200 import grid as G, processing, _native as N
    from support import synth
    res = {"natural3": synth.grid_natural_3layer, "mixed": synth.grid_mixed}[family](n, seed=21)
    g, ex = res if isinstance(res, tuple) else (res, None)
    wx = synth.synth_weather(3, 2001, ny, seed=4)
    node = synth.assign_nodes(n, 3)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                          tfsoil_c=-3.0, extra_layers=ex)
    b.cell_i32[N.HB_NLAY, 7] = 0                       # one cell HELP cannot run: NaN everywhere
    ctx = np.where(np.arange(n) % 4 == 0, 2, 1).astype(np.int32)
    with processing.HelpEngine(b, monthly=True) as eng:
        eng.simulate()
        r = eng.fetch()
        got = eng.post_process(context=ctx, year_from=2002, year_to=2001 + ny - 1)
        got_all = eng.post_process(context=None)
    names = ["c%d" % i for i in range(n)]
    out = {nm: dict(years=r["years"], **{k: r["monthly"][j, :, :, i] for j, k in enumerate(processing.KEYS)})
           for i, nm in enumerate(names)}
    d = post_oracle.post_process_output(out, ctx.tolist())
    assert got["idx_nan"] == d["idx_nan"] == [7]
    area = post_oracle.calc_area_monthly_avg(d)
    cells = post_oracle.calc_cells_yearly_avg(d, 2002, 2001 + ny - 1)
    for k in post_oracle.KEYS:
        # float64 sums of float32 inputs; only the summation order differs from numpy's pairwise sum
        np.testing.assert_allclose(got["area_monthly_avg"][k], area[k], rtol=1e-12, atol=1e-12)
        ok = np.ones(n, bool); ok[7] = False
        np.testing.assert_allclose(got["cells_yearly_avg"][k][ok], cells[k][ok], rtol=1e-12, atol=1e-12)
        assert np.isnan(got["cells_yearly_avg"][k][7]) == np.isnan(cells[k][7])
    # no context: recharge stays recharge; all years
    d0 = post_oracle.post_process_output(out, [1] * n)
    a0 = post_oracle.calc_area_monthly_avg(d0)
    np.testing.assert_allclose(got_all["area_monthly_avg"]["rechg"], a0["rechg"], rtol=1e-12, atol=1e-12)
    # water balance of the area means closes to the print quantisation: P = RO + ET + SR1 + SR2 + RECHG + storage change
    assert got_all["area_monthly_avg"]["precip"].sum() > 0
