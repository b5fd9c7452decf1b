"""Daily channel of the engine (OUTDAY columns, HELP3O.FOR:5984-6029):
bit-equal to the CPU oracle's daily series, and -- through the oracle's pinning -- the
reference's RCRA daily table at its printed precision."""
import json
import os.path as osp

import numpy as np
import pytest

from oracle import oracle
from pyhelp_b200 import inputs, processing
from support import synth
from pyhelp_b200 import _native as N

HERE = osp.dirname(osp.abspath(__file__))
RCRA = osp.join(HERE, "golden", "rcra")


def test_library_exports_daily_entry_points():
    L = N.lib()
    assert hasattr(L, "help_b200_engine_select_daily") and hasattr(L, "help_b200_engine_fetch_daily")
    assert len(processing.DAILY_COLS) == N.DAILY_COLS == len(oracle.DAILY_COLS)


def _rcra_params():
    d10 = np.char.ljust(open(osp.join(RCRA, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(RCRA, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    return (osp.join(RCRA, "RCRA.D4"), osp.join(RCRA, "RCRA.D7"), osp.join(RCRA, "RCRA.D13"),
            d11, d10, "/tmp/rcra.OUT", 0, 1, 0, 0, 3, 32.0)


@pytest.mark.gpu
def test_rcra_daily_table_on_the_gpu():
    p = _rcra_params()
    b = inputs.batch_from_cellparams({"rcra": p})
    with processing.HelpEngine(b, monthly=True) as eng:
        eng.select_daily([0]).simulate()
        got = eng.fetch_daily()[0]
    want = oracle.run_simulation(*p, daily=True)["daily"]
    assert got.shape == want.shape == (1095, 16)
    assert np.array_equal(got, want)                                   # every column of every day, bit for bit
    # the reference's printed daily table (RCRA.OUT:316-1700): rain F5.2, runoff and ET F7.3
    gold = json.load(open(osp.join(RCRA, "rcra_expected.json")))
    G = np.array(gold["daily"]); col = {k: i for i, k in enumerate(gold["daily_columns"])}
    assert np.abs(got[:, 0] - G[:, col["rain"]]).max() < 1e-6
    assert np.array_equal(got[:, 13], G[:, col["air_freeze"]]) and np.array_equal(got[:, 14], G[:, col["soil_freeze"]])
    assert np.mean(np.abs(got[:, 1] - G[:, col["runoff"]]) <= 0.0015) > 0.97      # F7.3 print, 1997 REAL*4 RC build
    assert np.mean(np.abs(got[:, 2] - G[:, col["et"]]) <= 0.0015) > 0.97


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["mixed", "landfill", "recirc"])
def test_liner_profiles_daily_bits(tmp_path, family):
    """barrier soil, geomembrane, lateral drainage: head / drainage / leakage columns too"""
    n, ny, nn = 24, 2, 2
    res = {"mixed": synth.grid_mixed, "landfill": synth.grid_landfill,
           "recirc": lambda n, seed: synth.grid_landfill(n, seed=seed, recirc=True)}[family](n, seed=17)
    g, ex = res if isinstance(res, tuple) else (res, None)
    wx = synth.synth_weather(nn, 2001, ny, seed=2)
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(n, nn)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, ex=ex)
    b = inputs.batch_from_cellparams(cp)
    sel = list(range(0, n, 3))
    with processing.HelpEngine(b, monthly=True) as eng:
        eng.select_daily(sel).simulate()
        daily = eng.fetch_daily()
    names = list(cp)
    heads = 0.0
    for k, c in enumerate(sel):
        want = oracle.run_simulation(*cp[names[c]], daily=True)["daily"]
        assert np.array_equal(daily[k], want), names[c]
        heads += daily[k][:, 4].sum()
    assert heads > 0.0                                   # the liner columns are exercised


@pytest.mark.gpu
def test_natural_soil_daily_bits_and_monthly_consistency(tmp_path):
    n, ny, nn = 40, 2, 3
    g = synth.grid_natural_3layer(n, seed=13)
    wx = synth.synth_weather(nn, 2003, ny, seed=9)          # 2004 is a leap year
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(n, nn)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0)
    b = inputs.batch_from_cellparams(cp)
    sel = [0, 17, 39]
    with processing.HelpEngine(b, monthly=True, raw=True) as eng:
        eng.select_daily(sel).simulate()
        daily = eng.fetch_daily()
        r = eng.fetch()
    assert daily.shape == (3, 365 + 366, 16)
    names = list(cp)
    for k, c in enumerate(sel):
        want = oracle.run_simulation(*cp[names[c]], daily=True)["daily"]
        assert np.array_equal(daily[k], want), names[c]
        # the daily precipitation adds up to the raw (unrounded REAL*4) monthly accumulators' yearly total
        assert abs(daily[k][:365, 0].sum() - r["raw"][0, 0, :, c].sum()) < 1e-3
        assert (daily[k][:, 4:6] == 0).all()                                  # no liner: no head, no lateral drainage
    # deselecting turns the channel off again
    with processing.HelpEngine(b, monthly=False) as eng:
        eng.select_daily([]).simulate()
        with pytest.raises(N.EngineError):
            eng.fetch_daily()
