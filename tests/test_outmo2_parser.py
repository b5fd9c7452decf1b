"""A23 / A24 pinned with the REFERENCE's own parser.

The oracle writes the monthly part of a HELP .OUT file (oracle/out_writer.py: OUTMO frame, OUTMO2 block, OUTSW
trailer -- HELP3O.FOR:6305-6406, 6415-6682, 8543) and the reference's ``read_monthly_help_output``
(pyhelp/processing.py:64-147) reads it.  tests/golden/outmo2/ holds those texts and what the reference's parser
returned for them (made by tests/golden/make_outmo2_golden.py, which executes the function from the reference
checkout).  Here: (1) the writer still produces the committed text byte for byte; (2) the oracle's own ``parsed``
arrays -- the ones the GPU engine must equal bit for bit -- equal what the reference's parser returned: row -> key map
(first 'LAT. DRAINAGE' row -> subrun1, later ones summed -> subrun2, first 'PERCOLATION' -> perco, last -> rechg),
``split()[-12:]``, float32 of the F7.3 text, the inch-valued Jul-Dec percolation of a sixth subprofile
(HELP3O.FOR:6593-6594), rows the parser must ignore ('SUBSURFACE INFLOW INTO LAYER n'); (3) a month over 999.999 mm
prints '*******', on which the reference's parser raises ValueError -- the engine returns NaN and sets
HELP_B200_FLAG_OVERFLOW instead (a documented deviation); (4) with the reference checkout present, the parser is run live."""
import os
import os.path as osp
import sys

import numpy as np
import pytest

from oracle import oracle, out_writer

HERE = osp.dirname(osp.abspath(__file__))
GOLD = osp.join(HERE, "golden", "outmo2")
sys.path.insert(0, osp.join(HERE, "golden"))
import make_outmo2_golden as M          # noqa: E402

EXP = np.load(osp.join(GOLD, "expected.npz"))
KEYS = ("precip", "runoff", "evapo", "subrun1", "subrun2", "perco", "rechg")


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    return M.cases(str(tmp_path_factory.mktemp("outmo2")))


@pytest.mark.parametrize("name", ["rcra", "natural", "geotextile", "subin_liner", "subin_natural", "two_liners", "six_subprofiles", "overflow"])
def test_reference_parser_reads_the_oracle_text_as_the_oracle_parses_it(cases, tmp_path, name):
    p = cases[name]
    res = oracle.run_simulation(*p, full=True)
    path = out_writer.write_out(str(tmp_path / (name + ".OUT")), res, p[4])
    assert open(path).read() == open(osp.join(GOLD, name + ".OUT")).read()          # (1)
    if name == "overflow":                                                            # (3)
        assert "could not convert string to float" in str(EXP[name + "/error"]) and "*******" in open(path).read()
        assert np.isnan(res["precip"]).sum() == 1 and np.isfinite(res["runoff"]).all()
        return
    assert np.array_equal(EXP[name + "/years"], res["years"]) and EXP[name + "/years"].dtype == res["years"].dtype == np.uint16
    for k in KEYS:                                                                    # (2)
        assert EXP[name + "/" + k].dtype == res[k].dtype == np.float32
        assert np.array_equal(EXP[name + "/" + k], res[k]), (name, k)
    if name == "six_subprofiles":
        info = res["info"]
        assert info["lp"][5] > 0 and res["monthly_in"][:, 13, 6:].max() > 0
        raw = res["monthly_in"][:, 13]                        # PRC6M, inches
        want = np.concatenate([raw[:, :6] * np.float32(25.4), raw[:, 6:]], axis=1)
        assert np.allclose(res["rechg"], want, atol=5.1e-4) and not np.allclose(res["rechg"][:, 6:], raw[:, 6:] * 25.4, atol=5e-4)
    if name.startswith("subin"):
        assert "SUBSURFACE INFLOW INTO" in open(path).read()
    if name in ("rcra", "two_liners"):
        assert (res["subrun2"] > 0).any()                     # more than one lateral-drainage row: the later ones are summed


@pytest.mark.skipif(not osp.exists("/root/reference/pyhelp/processing.py"), reason="reference checkout not present (GPU box)")
def test_live_reference_parser(cases, tmp_path):                                       # (4)
    parse = M.reference_parser()
    for name in ("rcra", "six_subprofiles", "subin_liner"):
        p = cases[name]
        res = oracle.run_simulation(*p, full=True)
        got = parse(out_writer.write_out(str(tmp_path / (name + ".OUT")), res, p[4]))
        for k in KEYS:
            assert np.array_equal(got[k], res[k]), (name, k)
    with pytest.raises(ValueError):
        p = cases["overflow"]
        parse(out_writer.write_out(str(tmp_path / "o.OUT"), oracle.run_simulation(*p, full=True), p[4]))
