"""Host packers: Fortran formatted-read semantics and the reference's text quantisation.

The reference hands the engine TEXT (D4/D7/D13 files, D10/D11 records).  Two
packers must produce the same numbers: `inputs.batch_from_cellparams` (parses that text like
HELP3O.FOR's formatted READs) and `grid.batch_from_grid` (quantises arrays like the
formatters, no text)."""
import numpy as np
import pytest

from pyhelp_b200 import grid as G
from pyhelp_b200 import inputs
from support import synth
from pyhelp_b200 import _native as N


@pytest.mark.parametrize("family", ["natural3", "mixed", "landfill", "recirc", "deep"])
def test_vectorised_packer_equals_text_round_trip(tmp_path, family):
    maker = {"natural3": synth.grid_natural_3layer, "mixed": synth.grid_mixed, "landfill": synth.grid_landfill,
             "recirc": lambda n, seed: synth.grid_landfill(n, seed=seed, recirc=True), "deep": synth.grid_deep}[family]
    res = maker(50, seed=13)
    g, ex = res if isinstance(res, tuple) else (res, None)
    ny, nn = 2, 5
    wx = synth.synth_weather(nn, 2003, ny, seed=2)
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(50, nn)
    cp = synth.cellparams_from_grid(g, files, node, node, node, ny, tfsoil_c=-3.0, sf_edepth=0.15, sf_cn=1.15, ex=ex)
    a = inputs.batch_from_cellparams(cp)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node,
                          tfsoil_c=-3.0, sf_edepth=0.15, sf_cn=1.15, extra_layers=ex)
    assert (a.n_cells, a.n_years, a.max_layers) == (b.n_cells, b.n_years, b.max_layers)
    for k in ("years", "pre", "tmp", "rad", "iu4", "iu7", "iu13", "tm_normals", "idfs", "cell_f32", "cell_i32",
              "layer_f32", "layer_i32", "layer_rc"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.names == b.names


def test_fortran_read_semantics():
    rec = inputs._records(["  12 3.5    7", "", " 1.5      25"], 80)
    recd = inputs._records(["1.5D2"], 80)
    # Fw.d: implied decimals only without an explicit point; blank field = 0
    assert inputs._to_real(inputs._field(rec, 0, 4), 1, np.float32)[:3].tolist() == [np.float32(1.2), 0.0, np.float32(1.5)]
    assert inputs._to_real(inputs._field(rec, 4, 4), 2, np.float32)[0] == np.float32(3.5)
    assert inputs._to_int(inputs._field(rec, 8, 5))[:3].tolist() == [7, 0, 25]
    assert inputs._to_real(inputs._field(recd, 0, 5), 0, np.float64)[0] == 150.0      # D exponent


def test_quantisation_matches_python_format():
    x = np.array([0.125, 0.375, 2.5, 1e-15, 123.4567, 0.0005, 0.0015])
    for d in (0, 1, 2, 3, 14):
        want = np.array([float(("{0:.%df}" % d).format(v)) for v in x])
        assert np.array_equal(G.quantize(x, d), want), d


def test_rcra_records_parse(rcra_dir):
    import os.path as osp
    d10 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    cp = {"rcra": (osp.join(rcra_dir, "RCRA.D4"), osp.join(rcra_dir, "RCRA.D7"), osp.join(rcra_dir, "RCRA.D13"),
                   d11.reshape(-1, 1), d10, "x", 0, 1, 0, 0, 3, 32.0)}            # (n,1) and (n,) shapes both accepted
    b = inputs.batch_from_cellparams(cp)
    assert b.n_cells == 1 and b.n_years == 3 and b.max_layers == 11
    assert b.layer_i32[N.HL_TYPE, :, 0].tolist() == [1, 2, 4, 3, 1, 2, 2, 4, 2, 4, 3]      # RCRA.OUT layer list
    assert b.cell_i32[N.HB_IU10, 0] == 1 and b.cell_i32[N.HB_IU11, 0] == 1                # IP units
    assert b.idfs[0, 0] == 8
    assert b.tm_normals[0].max() > 60                                                     # monthly normals present (deg F)


def test_errors():
    with pytest.raises(ValueError):
        inputs.batch_from_cellparams({})
    d = np.char.ljust(np.array(["a", "b", "c"]), 80).astype("S80")
    with pytest.raises(FileNotFoundError):
        inputs.batch_from_cellparams({"x": ("/nonexistent.D4", "/n.D7", "/n.D13", d, d, "o", 0, 1, 0, 0, 1, 32.0)})


def test_deep_family_reaches_the_layer_limit():
    """20 layers is what the reference's COMMON blocks hold (and include/help_b200.h HELP_B200_MAX_LAYERS)."""
    g = synth.grid_deep(200, seed=6)
    assert int(g["nlayer"].max()) == 20
    wx = synth.synth_weather(2, 2001, 1, seed=5)
    node = synth.assign_nodes(200, 2)
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node)
    assert b.max_layers == 20 and b.layer_f32.shape[1] == 20
    types = np.stack([g[f"lay_type{j}"] for j in range(1, 21)])
    assert (types == 2).any() and (types == 3).any()
    # a barrier-soil layer always has its lateral-drainage layer directly above
    jj, cc = np.nonzero(types == 3)
    assert (types[jj - 1, cc] == 2).all()


def test_fixed_length_records_fast_path_equals_the_padded_path(tmp_path):
    """'S80' arrays go to the byte matrix directly (NUL padding read as blanks); shapes (n,) and (n, 1)
    (managers.py:410-411 vs pyhelp/tests/test_help3o.py:42-48) and str arrays give the same records."""
    lines = ["  2", "cell 7", "46.12      112 313   3.50     9.0 10.0 67.9 65.5 74.0 75.7"]
    padded = inputs._records(np.char.ljust(np.array(lines), 80))              # unicode -> encode -> ljust
    short = inputs._records(np.array(lines).astype("S80"))                     # NUL-padded S80, fast path
    col = inputs._records(np.char.ljust(np.array([[x] for x in lines]), 80).astype("S80"))
    assert padded.shape == (3, 80) and np.array_equal(padded, short) and np.array_equal(padded, col)
    assert (padded[1, 6:] == 32).all()
    # a weather file with CRLF line ends and a non-ASCII station name parses like the plain one
    wx = synth.synth_weather(1, 2001, 1, seed=3)
    f = synth.write_weather_files(str(tmp_path), wx)["D7"][0]
    a = inputs.read_weather_file(f, "D7")
    txt = open(f, "rb").read().split(b"\n")
    txt[2] = "Rivi\u00e8re du Nord".encode("latin-1")
    g = tmp_path / "crlf.D7"
    g.write_bytes(b"\r\n".join(txt))
    b = inputs.read_weather_file(str(g), "D7")
    assert np.array_equal(a.data, b.data) and np.array_equal(a.years, b.years) and a.units == b.units
    # the year blocks of a written file are records of one length and are reshaped in one step; a file whose lines differ
    # in length (hand-edited: trailing blanks stripped, an extra line, no final newline) goes the line-by-line way -- the
    # same records either way
    raw = open(f, "rb").read()
    lines = raw.split(b"\n")
    ragged = b"\n".join(x.rstrip() if k % 3 else x for k, x in enumerate(lines))
    for k, variant in enumerate((raw.rstrip(b"\n"), raw + b"one more line\n", ragged, raw.replace(b"\n", b"\r\n"))):
        h = tmp_path / f"variant{k}.D7"
        h.write_bytes(variant)
        c = inputs.read_weather_file(str(h), "D7")
        assert np.array_equal(a.data, c.data) and np.array_equal(a.years, c.years) and np.array_equal(a.header12, c.header12), k


def test_idfs_prescan_reads_f61_whatever_the_units(tmp_path):
    """READIN's pre-scan of the D13 file uses F6.1 (formats 5400/5410, F:1364-1365) although READCD
    reads SI radiation with F6.0 (F:1128): an SI file written without decimal points gives the
    pre-scan values ten times smaller, hence another IDFS than the yearly data would.  The oracle
    reads the file record by record like the Fortran; the vectorised host parser must agree."""
    from oracle import oracle
    wx = synth.synth_weather(1, 2001, 2, seed=3)
    files = synth.write_weather_files(str(tmp_path), wx)
    lines = open(files["D13"][0]).read().split("\n")
    out = lines[:4]
    for ln in lines[4:]:
        if not ln:
            continue
        vals = [float(ln[5 + 6 * k:11 + 6 * k]) for k in range(10)]
        out.append(ln[:5] + "".join("{0:>6d}".format(int(round(v))) for v in vals))       # integer MJ/m2, no point
    p = tmp_path / "int.D13"
    p.write_text("\n".join(out) + "\n")
    plain = inputs.read_weather_file(files["D13"][0], "D13")
    ints = inputs.read_weather_file(str(p), "D13")
    assert np.array_equal(ints.data, np.rint(plain.data))                                  # READCD: F6.0
    assert ints.idfs[0] > plain.idfs[0]                                                    # pre-scan saw a tenth of the radiation
    g = synth.grid_natural_3layer(1, seed=1)
    node = np.zeros(1, np.int32)
    f2 = {"D4": files["D4"], "D7": files["D7"], "D13": {0: str(p)}}
    cp = synth.cellparams_from_grid(g, f2, node, node, node, 2)
    info = oracle.run_simulation(*cp["0"], full=True)["info"]
    assert info["idfs"] == ints.idfs[0]
    # Fw.d without a point: the implied decimals apply to the significand, exponent or not
    assert inputs._to_real(np.array([b"125E1", b"12.5E1", b"125"]), 2, np.float64).tolist() == [12.5, 125.0, 1.25]


def test_host_formatted_read_of_every_field_form():
    """Fw.d input fields as libgfortran reads them (blanks ignored, implied decimals without a point, E/D/bare-sign
    exponents; one correctly rounded conversion) -- the host leg of parse_matrix; tests/test_gpu_pack.py holds the device
    leg to the same values."""
    fields = ["      ", "  12.5", " -0.25", "+3    ", "  1 2 ", "1.5E+3", "1.5D-2", "12E1  ", " 125  ", ".5    ", "5.    ",
              "-.5e-2", "1.5+2 ", "123456", "0.0001", "  -0  ", "9.9999", "000001", "1e22  ", "1e-22 ", "1e23  ", "7E0   "]
    want = [0.0, 12.5, -0.25, 0.03, 0.12, 1500.0, 0.015, 1.2, 1.25, 0.5, 5.0, -0.005, 150.0, 1234.56, 0.0001, -0.0, 9.9999,
            0.01, 1e20, 1e-24, 1e21, 0.07]
    rec = inputs._records(np.array(["".join(fields)]), reclen=6 * len(fields))
    got = inputs.parse_matrix(rec, [(6 * k, 6, 2) for k in range(len(fields))], "host")[:, 0]
    assert np.array_equal(got, np.array(want)) and np.signbit(got[15])
    with pytest.raises(ValueError):
        inputs.parse_matrix(inputs._records(np.array(["  1x3 "]), reclen=6), [(0, 6, 0)], "host")
