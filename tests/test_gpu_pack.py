"""The host packer's device leg (`-m gpu`): help_b200_pack_weather (csrc/help_weather.cuh) must
return the bits of the reference's text round trip -- float32(float(format(x, '.Nf'))) as
pyhelp/weather_reader.py:31,44,57 prints and READCD (HELP3O.FOR:1100-1129) reads -- in the engine's
[node][year][368] layout, i.e. exactly what grid.weather_to_nodes computes on the host."""
import numpy as np
import pytest

from pyhelp_b200 import grid as G
from support import synth

pytestmark = pytest.mark.gpu


def test_device_weather_equals_host_weather():
    wx = synth.synth_weather(70, 2003, 3, seed=9)                    # 2004 is a leap year; 70 nodes: ragged tiles
    for key, d in (("precip", 1), ("airtemp", 1), ("solrad", 2)):
        raw = wx[key] + np.random.default_rng(1).normal(0, 0.04, wx[key].shape)      # off the quantisation grid
        yh, h = G.weather_to_nodes(wx["years_of_day"], raw, d, "host")
        yd, dv = G.weather_to_nodes(wx["years_of_day"], raw, d, "device")
        assert np.array_equal(yh, yd) and h.dtype == dv.dtype == np.float32 and h.shape == dv.shape == (70, 3, 368)
        assert np.array_equal(h.view(np.uint32), dv.view(np.uint32))


def test_device_quantisation_is_pythons_format_on_ties_and_near_ties():
    """'{:.1f}'.format rounds the exact binary value half to even: 0.25 -> 0.2, 0.35 (just below in binary) -> 0.3,
    0.75 -> 0.8, 2.675 -> 2.67 (binary below), 1.005 -> 1.00; products x*10^d that round onto a tie must not be
    treated as one."""
    base = np.array([0.25, 0.35, 0.75, 0.05, 0.15, 0.45, 2.5, 3.5, -0.25, -0.35, 1e-9, 123.45, 2.675, 1.005, 0.125, 0.375,
                     1234.5678, 0.0, 65.55, 99.95])
    rng = np.random.default_rng(3)
    vals = np.concatenate([base, np.nextafter(base, 10), np.nextafter(base, -10), rng.integers(0, 4000, 325) / 8.0,
                           rng.integers(-2000, 2000, 300) / 16.0 + 0.05])
    vals = vals[:365].reshape(365, 1)
    yod = np.full(365, 2001, dtype=np.int32)
    for d in (0, 1, 2, 3):
        want = np.array([float(format(x, f".{d}f")) for x in vals[:, 0]]).astype(np.float32)
        _, got = G.weather_to_nodes(yod, vals, d, "device")
        assert np.array_equal(got[0, 0, :365], want + np.float32(0.0)), d
        assert (got[0, 0, 365:] == 0).all()
        _, host = G.weather_to_nodes(yod, vals, d, "host")
        assert np.array_equal(host[0, 0], got[0, 0])


def test_batch_from_grid_device_weather_gives_the_same_batch():
    n, nn, ny = 64, 5, 2
    g = synth.grid_mixed(n, seed=2)
    wx = synth.synth_weather(nn, 2001, ny, seed=5)
    node = synth.assign_nodes(n, nn)
    a = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, weather="host")
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, weather="device")
    for k in ("years", "pre", "tmp", "rad", "idfs", "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_resident_weather_gives_the_same_batch_and_the_same_results():
    """weather='resident' (help_b200_pack_weather_resident: the packed tables stay in device memory, the engine copies them
    device to device; IDFS from help_b200_idfs_prescan_resident) against weather='host': same IDFS, same simulation bits --
    also when two engines are created from the one resident batch."""
    from pyhelp_b200 import _native as N, processing
    n, nn, ny = 96, 70, 3                                       # 70 nodes: ragged tiles; 2004 is a leap year
    g = synth.grid_mixed(n, seed=2)
    wx = synth.synth_weather(nn, 2003, ny, seed=5)
    node = synth.assign_nodes(n, nn)
    a = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, weather="host")
    b = G.batch_from_grid(g, wx["years_of_day"], wx["precip"], wx["airtemp"], wx["solrad"], node, node, node, weather="resident")
    assert all(isinstance(getattr(b, k), N.DeviceTable) and getattr(b, k).shape == getattr(a, k).shape for k in ("pre", "tmp", "rad"))
    for k in ("years", "idfs", "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert b.nbytes() == a.nbytes()
    want = processing.run_batch(a, monthly=True)
    for _ in range(2):
        got = processing.run_batch(b, monthly=True)
        for key in ("monthly", "yearly", "flags"):
            assert np.array_equal(got[key], want[key], equal_nan=True), key
    sub = b.take(np.arange(0, n, 3))                             # a sub-batch shares the resident tables
    assert np.array_equal(processing.run_batch(sub, monthly=False)["yearly"], want["yearly"][:, :, ::3])
    # southern-hemisphere window and a single year: the pre-scan's other branches
    idfs = np.zeros((nn, 2), dtype=np.int32)
    _, rad1 = G.weather_to_nodes(wx["years_of_day"][:365], wx["solrad"][:365], 2, "resident")
    N.check(N.lib().help_b200_idfs_prescan_resident(rad1.ptr, nn, 1, 2, N.ptr(idfs)), "idfs")
    _, rad1h = G.weather_to_nodes(wx["years_of_day"][:365], wx["solrad"][:365], 2, "host")
    assert np.array_equal(idfs, G.idfs_prescan_nodes(rad1h, 2))
    N.check(N.lib().help_b200_idfs_prescan_resident(rad1.ptr, nn, 1, 1, N.ptr(idfs)), "idfs")
    assert np.array_equal(idfs, G.idfs_prescan_nodes(rad1h, 1))


def test_device_formatted_read_equals_host_read(tmp_path, rcra_dir):
    """help_b200_parse_fields (csrc/help_parse.cuh) vs the host's conversion of the same fixed-width fields: hand-made
    records with every form a Fortran input field takes, then the records and weather files of RCRA and of a synthetic
    landfill grid through batch_from_cellparams(parse='device')."""
    import os.path as osp
    from pyhelp_b200 import inputs
    fields = ["      ", "  12.5", " -0.25", "+3    ", "  1 2 ", "1.5E+3", "1.5D-2", "12E1  ", " 125  ", ".5    ", "5.    ", "-.5e-2", "1.5+2 ",
              "123456", "0.0001", "  -0  ", "9.9999", "000001", "1e22  ", "1e-22 ", "1e23  ", "7E0   "]
    rec = inputs._records(np.array(["".join(fields)]), reclen=6 * len(fields))
    specs = [(6 * k, 6, 2) for k in range(len(fields))]
    host = []
    for k, f in enumerate(fields):                                      # Fortran semantics, field by field
        t = f.replace(" ", "")
        if not t:
            host.append(0.0); continue
        t2 = t.upper().replace("D", "E")
        if "E" not in t2 and ("+" in t2[1:] or "-" in t2[1:]):
            i = max(t2.rfind("+"), t2.rfind("-")); t2 = t2[:i] + "E" + t2[i:]
        mant, _, ex = t2.partition("E")
        # one correctly rounded conversion of significand x 10^(exponent - d), as libgfortran's READ does
        host.append(float(t2) if "." in mant else float(f"{mant}E{int(ex or 0) - 2}"))
    dev = inputs.parse_matrix(rec, specs, "device")[:, 0]
    assert np.array_equal(dev, np.array(host)), list(zip(fields, dev.tolist(), host))
    with pytest.raises(ValueError):
        inputs.parse_matrix(inputs._records(np.array(["  1x3 "]), reclen=6), [(0, 6, 0)], "device")
    # 19 significant digits: flagged by the kernel, converted on the host
    long = inputs.parse_matrix(inputs._records(np.array(["0.1234567890123456789"]), reclen=21), [(0, 21, 0)], "device")
    assert long[0, 0] == 0.1234567890123456789
    # whole inputs: same batch from both paths
    d10 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D10")).read().splitlines(), 80).astype("S80")
    d11 = np.char.ljust(open(osp.join(rcra_dir, "RCRA.D11")).read().splitlines(), 80).astype("S80")
    cps = [{"rcra": (osp.join(rcra_dir, "RCRA.D4"), osp.join(rcra_dir, "RCRA.D7"), osp.join(rcra_dir, "RCRA.D13"), d11, d10, "x", 0, 1, 0, 0, 3, 32.0)}]
    g, ex = synth.grid_special(96, seed=8)
    wx = synth.synth_weather(5, 2003, 2, seed=9)
    files = synth.write_weather_files(str(tmp_path), wx)
    node = synth.assign_nodes(96, 5)
    cps.append(synth.cellparams_from_grid(g, files, node, node, node, 2, tfsoil_c=-3.0, ex=ex))
    for cp in cps:
        a, b = inputs.batch_from_cellparams(cp, parse="host"), inputs.batch_from_cellparams(cp, parse="device")
        for k in ("years", "pre", "tmp", "rad", "iu4", "iu7", "iu13", "tm_normals", "idfs", "cell_f32", "cell_i32", "layer_f32", "layer_i32", "layer_rc"):
            x, y = getattr(a, k), getattr(b, k)
            assert x.dtype == y.dtype and np.array_equal(x, y), k
