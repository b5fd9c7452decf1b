"""Pin the CPU oracle against the reference's own golden case (no GPU).

Golden: tests/golden/rcra/ (made by tests/golden/make_rcra_golden.py from the
reference's pyhelp/tests/rca_original_testcase_1997).  The expected numbers were
produced by HELP 3.07 (1997, REAL*4 conductivities); PyHELP's HELP3O.FOR holds
them in REAL*8, and its own acceptance is 0.1 in/yr
(pyhelp/tests/test_help3o.py:87-118), so agreement is at printed precision on
most days, not bitwise.  Tolerances below are the measured envelope + margin.

Both builds of the oracle are pinned: "det" (deterministic transcendentals shared
with the GPU engine, pyhelp_b200/csrc/help_math.h) and "glibc" (the host libm a
gfortran build of the reference would link).
"""
import json
import os.path as osp

import numpy as np
import pytest

from oracle import oracle

MM2IN = 0.0393701


@pytest.fixture(scope="module", params=["det", "glibc"])
def rcra(rcra_dir, request):
    d10 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    out = oracle.run_simulation(
        osp.join(rcra_dir, "RCRA.D4"), osp.join(rcra_dir, "RCRA.D7"), osp.join(rcra_dir, "RCRA.D13"),
        d11, d10, None, 0, 1, 0, 0, 3, 32.0, daily=True, libm=request.param)
    gold = json.load(open(osp.join(rcra_dir, "rcra_expected.json")))
    return out, gold


def test_topology(rcra):
    out, _ = rcra
    info = out["info"]
    # RCRA.OUT:306-314: liners close subprofiles at layers 4, 8, 11; drains 2, 7, 9
    assert info["lay"] == 11
    assert info["lp"][:3] == [4, 8, 11]
    assert info["ld"][:3] == [2, 7, 9]
    assert info["nsegn"][:3] == [10, 6, 2]      # SURVEY.md §7 hand derivation
    assert info["idfs"] == 8
    assert info["nonconv"] == 0


def test_reference_known_answers(rcra):
    """The exact assertions of pyhelp/tests/test_help3o.py:87-118."""
    out, gold = rcra
    kat = gold["kat"]
    for key in ("precip", "runoff", "evapo", "subrun1", "subrun2", "rechg"):
        got = out[key].astype(np.float64).sum(axis=1) * MM2IN
        assert np.abs(got - np.array(kat[key])).max() < kat["tolerance_in"], key
    assert out["years"].tolist() == [1, 2, 3]


def test_annual_totals_tight(rcra):
    out, gold = rcra
    a = out["annual_in"]
    rows = {"PRECIPITATION": 0, "RUNOFF": 1, "EVAPOTRANSPIRATION": 2,
            "DRAINAGE COLLECTED FROM LAYER 2": 3, "DRAINAGE COLLECTED FROM LAYER 7": 4,
            "DRAINAGE COLLECTED FROM LAYER 9": 5, "PERC./LEAKAGE THROUGH LAYER 4": 8,
            "PERC./LEAKAGE THROUGH LAYER 8": 9, "PERC./LEAKAGE THROUGH LAYER 11": 10}
    for y in range(3):
        exp = gold["annual"][str(y + 1)]
        for label, r in rows.items():
            e, g = exp[label], float(a[y, r])
            assert abs(g - e) <= max(0.01, 2e-3 * abs(e)), (y, label, g, e)
    # water budget closes (F:953-968)
    assert np.abs(out["balance_in"]).max() < 5e-4


def test_monthly_totals(rcra):
    out, gold = rcra
    m = out["monthly_in"]
    rows = {"PRECIPITATION": 0, "RUNOFF": 1, "EVAPOTRANSPIRATION": 2,
            "LATERAL DRAINAGE COLLECTED LAYER 2": 3, "LATERAL DRAINAGE COLLECTED LAYER 7": 4,
            "LATERAL DRAINAGE COLLECTED LAYER 9": 5, "PERCOLATION/LEAKAGE THROUGH LAYER 4": 8,
            "PERCOLATION/LEAKAGE THROUGH LAYER 8": 9, "PERCOLATION/LEAKAGE THROUGH LAYER 11": 10}
    for y in range(3):
        for label, r in rows.items():
            e = np.array(gold["monthly"][str(y + 1)][label])
            g = m[y, r].astype(np.float64)
            assert np.abs(g - e).max() <= 0.03, (y, label, np.abs(g - e).max())


def test_initial_and_final_moisture(rcra):
    out, gold = rcra
    init = out["layer_sw"][0][:11]
    fin = out["layer_sw"][1][:11]
    assert np.abs(init - np.array(gold["initial_sw"])).max() < 1e-4   # F9.4 printed
    assert np.abs(fin - np.array(gold["final_sw"])).max() < 5e-4


def test_daily_table(rcra):
    out, gold = rcra
    G = np.array(gold["daily"])
    D = out["daily"]
    assert D.shape[0] == G.shape[0] == 1095
    names = gold["daily_columns"]
    col = {k: names.index(k) for k in names}
    # state machines agree on every single day
    assert np.array_equal(D[:, 13], G[:, col["air_freeze"]])
    assert np.array_equal(D[:, 14], G[:, col["soil_freeze"]])
    assert np.abs(D[:, 0] - G[:, col["rain"]]).max() < 1e-6
    # (oracle column, golden column, max abs, allowed count above print precision, print eps)
    checks = [(1, "runoff", 1e-3, 2, 6e-4), (2, "et", 0.03, 25, 6e-4),
              (3, "ezone_water", 3e-3, 120, 6e-5), (4, "head1", 1.0, 80, 1e-2),
              (5, "drain1", 2e-3, 110, 1e-4), (6, "leak1", 2e-4, 60, 1e-5)]
    for j, name, mx, nbad, eps in checks:
        d = np.abs(D[:, j] - G[:, col[name]])
        assert d.max() <= mx, (name, d.max())
        assert (d > eps).sum() <= nbad, (name, int((d > eps).sum()))
    for j, name in [(8, "drain2"), (9, "leak2"), (11, "drain3"), (12, "leak3")]:
        d = np.abs(D[:, j] - G[:, col[name]])
        assert d.max() <= 1e-4, (name, d.max())


def test_parsed_layout_matches_reference_parser_contract(rcra):
    """processing.py:139-146: dtypes/shapes; row->key mapping."""
    out, _ = rcra
    assert out["years"].dtype == np.uint16
    for k in oracle.KEYS_PARSED:
        assert out[k].shape == (3, 12) and out[k].dtype == np.float32
    m = out["monthly_in"]
    f = lambda a: np.round(a.astype(np.float32) * np.float32(25.4), 3)
    np.testing.assert_allclose(out["subrun1"], f(m[:, 3]), atol=6e-4)
    np.testing.assert_allclose(out["subrun2"], f(m[:, 4]) + f(m[:, 5]), atol=1.2e-3)
    np.testing.assert_allclose(out["perco"], f(m[:, 8]), atol=6e-4)
    np.testing.assert_allclose(out["rechg"], f(m[:, 10]), atol=6e-4)
    # values are multiples of 0.001 mm held in float32
    v = out["precip"].astype(np.float64) * 1000
    assert np.abs(v - np.round(v)).max() < 1e-2
