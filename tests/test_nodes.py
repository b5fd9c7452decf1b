"""Nearest weather-node assignment: GPU vs the numpy restatement of
pyhelp/utils.py:81-96 + pyhelp/managers.py:307-313."""
import numpy as np
import pytest

from oracle import nodes_oracle


def test_oracle_is_the_reference_formula():
    # one degree of latitude on the reference's sphere (r = 6373 km), and argmin picks the first minimum
    d = nodes_oracle.calc_dist_from_coord(45.0, -74.0, np.array([46.0, 45.0, 46.0]), np.array([-74.0, -74.0, -74.0]))
    assert abs(d[0] - 6373 * np.pi / 180) < 1e-9 and d[1] == 0.0
    idx = nodes_oracle.nearest_nodes([45.6], [-74.0], [46.0, 45.0, 46.0], [-74.0, -74.0, -74.0])
    assert idx.tolist() == [0]                      # nodes 0 and 2 are the same place: the first wins


def test_library_exports_entry_point():
    from pyhelp_b200 import _native as N
    assert hasattr(N.lib(), "help_b200_nearest_nodes")


@pytest.mark.gpu
@pytest.mark.parametrize("n_cells,n_nodes", [(1, 1), (777, 3), (20000, 500), (5000, 2500)])
def test_gpu_matches_oracle(n_cells, n_nodes):
    from pyhelp_b200 import nodes
    rng = np.random.default_rng(n_cells + n_nodes)
    nlat, nlon = rng.uniform(44.0, 49.0, n_nodes), rng.uniform(-76.0, -70.0, n_nodes)
    clat, clon = rng.uniform(44.0, 49.0, n_cells), rng.uniform(-76.0, -70.0, n_cells)
    want, gap = nodes_oracle.nearest_nodes(clat, clon, nlat, nlon, return_margin=True)
    got = nodes.nearest_nodes(clat, clon, nlat, nlon)
    # indices are identical wherever the two closest nodes differ by more than rounding (1e-9 km);
    # the reference's own result depends on numpy's libm inside that band
    clear = gap > 1e-9
    assert np.array_equal(got[clear], want[clear])
    d_got = np.array([nodes_oracle.calc_dist_from_coord(a, b, nlat[i], nlon[i]) for a, b, i in zip(clat, clon, got)])
    d_want = np.array([nodes_oracle.calc_dist_from_coord(a, b, nlat[i], nlon[i]) for a, b, i in zip(clat, clon, want)])
    assert np.abs(d_got - d_want).max() <= 1e-9


@pytest.mark.gpu
def test_gpu_lattice_ties_take_the_first_node():
    """cells exactly between nodes of a regular grid, duplicated nodes: np.argmin's first-minimum rule"""
    from pyhelp_b200 import nodes
    nlat = np.array([45.0, 45.0, 46.0, 46.0, 45.0]); nlon = np.array([-74.0, -73.0, -74.0, -73.0, -74.0])
    clat = np.array([45.0, 45.2, 46.0, 45.0]); clon = np.array([-74.0, -74.0, -73.0, -73.5])
    got = nodes.nearest_nodes(clat, clon, nlat, nlon)
    want = nodes_oracle.nearest_nodes(clat, clon, nlat, nlon)
    assert got[:3].tolist() == want[:3].tolist() == [0, 0, 3]
    assert got[3] in (0, 1)                         # equidistant in exact arithmetic


@pytest.mark.gpu
def test_arguments_are_checked():
    from pyhelp_b200 import nodes
    with pytest.raises(ValueError):
        nodes.nearest_nodes([45.0], [-74.0, -73.0], [45.0], [-74.0])
    with pytest.raises(ValueError):
        nodes.nearest_nodes([45.0], [-74.0], [], [])
    assert nodes.nearest_nodes([], [], [45.0], [-74.0]).size == 0
