"""BASELINE configs[2] and [3] at the size one GPU carries (`-m gpu`): one eighth of the 1 M-cell
mixed-class grid (125 000 cells x 30 yr) and one eighth of the 250 000 landfill covers x 50 yr.
Size-independent properties on every cell, and a random sample of cells re-run on the CPU oracle
through the reference's text formats, bit for bit.  (configs[1] at full size:
tests/test_gpu_parity.py; configs[0]: tests/test_example_config.py; configs[4]'s share:
scripts/config5_share.py.)"""
import numpy as np
import pytest

from oracle import oracle
from pyhelp_b200 import grid as G
from pyhelp_b200 import processing
from support import synth
from pyhelp_b200 import _native as N

pytestmark = pytest.mark.gpu


def _run_and_check(tmp_path, family, n, ny, n_sample, seed):
    maker = {"mixed": synth.grid_mixed, "landfill": synth.grid_landfill}[family]
    res = maker(n, seed=seed)
    g, ex = res if isinstance(res, tuple) else (res, None)
    nn = n // 100
    wx = synth.synth_weather(nn, 2001, ny, seed=5)
    node = synth.assign_nodes(n, nn)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                          tfsoil_c=-3.0, extra_layers=ex)
    r = processing.run_batch(b)
    m, y = r["monthly"], r["yearly"]
    assert (r["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL | N.FLAG_OVERFLOW)).sum() == 0
    assert np.isfinite(m).all() and (m >= 0).all() and np.isfinite(y).all()
    # precipitation row = the node's weather summed per month
    yod = wx["years_of_day"]
    for cell in (0, n // 3, n - 1):
        for iy in (0, ny // 2, ny - 1):
            days = wx["precip"][yod == 2001 + iy, node[cell]]
            assert abs(float(m[N.HR_PRECIP, iy, :, cell].sum()) - days.sum()) < 0.05
    # one subprofile (no liner): the first PERCOLATION row is also the last (processing.py:118-121);
    # yearly rows are the sums of the months
    has_liner = np.zeros(n, bool)
    for j in range(1, 12):
        if f"lay_type{j}" in g:
            has_liner |= (np.asarray(g[f"lay_type{j}"]) >= 3) & (np.asarray(g["nlayer"]) >= j)
    assert np.array_equal(m[N.HR_PERCO][..., ~has_liner], m[N.HR_RECHG][..., ~has_liner])
    np.testing.assert_allclose(y[N.HY_EVAPO], m[N.HR_EVAPO].sum(axis=1), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(y[N.HY_LATDRAIN], (m[N.HR_SUBRUN1] + m[N.HR_SUBRUN2]).sum(axis=1), rtol=1e-5, atol=1e-3)
    # cells without a lateral-drainage layer drain nothing sideways
    has_drain = np.zeros(n, bool)
    for j in range(1, 12):
        if f"lay_type{j}" in g:
            has_drain |= (np.asarray(g[f"lay_type{j}"]) == 2) & (np.asarray(g["nlayer"]) >= j)
    assert (y[N.HY_LATDRAIN][:, ~has_drain] == 0).all()
    if family == "landfill":
        assert has_drain.all() and has_liner.all()
        assert (y[N.HY_LATDRAIN].sum(axis=0) > 0).mean() > 0.99
    # long-run water balance of every cell: P = R + ET + lateral drainage + recharge + storage change
    P, R, E, D, Q = (y[i].sum(axis=0).astype(np.float64) for i in (N.HY_PRECIP, N.HY_RUNOFF, N.HY_EVAPO, N.HY_LATDRAIN, N.HY_RECHG))
    resid = (P - R - E - D - Q) / ny
    assert np.abs(resid).max() < 80.0 and abs(resid.mean()) < 5.0, (np.abs(resid).max(), resid.mean())
    # sampled cells on the oracle
    sel = np.sort(np.random.default_rng(seed).choice(n, n_sample, replace=False))
    files = synth.write_weather_files(str(tmp_path), wx, nodes=sorted({int(node[i]) for i in sel}))
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex, cells=sel.tolist())
    want = oracle.run_help_allcells(cp)
    got = processing.split_by_cell(b.take(sel), {"monthly": m[..., sel], "years": r["years"]})
    bad = [(name, k) for name, w in want.items() for k in processing.KEYS
           if not np.array_equal(got[name][k], w[k], equal_nan=True)]
    assert not bad, bad[:10]


def test_mixed_classes_one_gpu_share_of_config_3(tmp_path):
    """1 M mixed-class cells x 30 yr over 8 GPUs -> 125 000 cells here (50 % 3-layer natural soil,
    20 % 1-layer, 20 % drain over barrier soil, 10 % 5-layer)."""
    _run_and_check(tmp_path, "mixed", 125_000, 30, 32, seed=2)


def test_landfill_covers_one_gpu_share_of_config_4(tmp_path):
    """250 000 landfill covers x 50 yr (topsoil / lateral-drainage layer / geomembrane / barrier
    soil) -> 31 250 cells here."""
    _run_and_check(tmp_path, "landfill", 31_250, 50, 12, seed=3)


def test_province_scale_share_of_config_4_yearly_only(tmp_path):
    """BASELINE configs[4]: 5 M cells x 30 yr, mixed classes, one weather node per 100 cells, yearly summaries only
    (no 50 GB of monthly arrays) -> one GPU's eighth, 625 000 cells, with a monthly-free engine.  Properties on every
    cell + sampled cells on the oracle (yearly rows = sums of the oracle's months)."""
    n, ny = 625_000, 30
    nn = n // 100
    g = synth.grid_mixed(n, seed=4)
    wx = synth.synth_weather_chunked(nn, 2001, ny, seed=4, chunk=625)
    node = synth.assign_nodes(n, nn)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, tfsoil_c=-3.0,
                          weather="device")
    with processing.HelpEngine(b, monthly=False) as eng:
        eng.simulate()
        r = eng.fetch(monthly=False, yearly=True)
    y = r["yearly"]
    assert y.shape == (N.HY_NROWS, ny, n) and "monthly" not in r
    assert (r["flags"] & (N.FLAG_NAN | N.FLAG_BADCELL | N.FLAG_OVERFLOW)).sum() == 0
    assert np.isfinite(y).all() and (y >= 0).all()
    P, R, E, D, Q = (y[i].sum(axis=0).astype(np.float64) for i in range(5))
    resid = (P - R - E - D - Q) / ny
    assert np.abs(resid).max() < 80.0 and abs(resid.mean()) < 5.0, (np.abs(resid).max(), resid.mean())
    sel = np.sort(np.random.default_rng(4).choice(n, 24, replace=False))
    files = synth.write_weather_files(str(tmp_path), wx, nodes=sorted({int(node[i]) for i in sel}))
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, cells=sel.tolist())
    want = oracle.run_help_allcells(cp)
    for k, c in enumerate(sel.tolist()):
        w = want[str(c)]
        rows = (w["precip"], w["runoff"], w["evapo"], w["subrun1"] + w["subrun2"], w["rechg"])
        for i, m in enumerate(rows):
            # the engine sums the 12 float32 months of a year in order; so does this
            s = np.zeros(ny, dtype=np.float32)
            mm = m if i != 3 else None
            for mo in range(12):
                s = s + (m[:, mo] if i != 3 else (w["subrun1"][:, mo] + w["subrun2"][:, mo]))
            assert np.array_equal(y[i, :, c], s), (c, i)
