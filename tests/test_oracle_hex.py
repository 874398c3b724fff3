"""The H3 restatement (oracle/h3_restated.py) against everything that can be checked without the h3 library:
self-consistency of the recalled constant tables, H3's cell counts, the reference's own hex test
(/root/reference/tests/nodes/test_hex_nodes.py:32 - 122 nodes at resolution 0), grid regularity, and the two cell
centres H3's documentation publishes."""

import numpy as np
import pytest

from oracle import h3_restated as H


def test_constant_tables_form_a_regular_icosahedron():
    chk = H.self_check()
    # a mis-remembered digit in any of the 100 constants would show up here at >= 1e-9
    assert chk["err_centres"] < 1e-15
    assert chk["err_axes"] < 1e-15
    assert chk["err_vertices"] < 1e-15
    assert chk["faces_per_vertex"] == [5] * 12
    assert abs(chk["gnomonic_unit"]) < 1e-15


@pytest.mark.parametrize("res", [0, 1, 2, 3])
def test_cell_counts_and_pentagons(res):
    c, pent = H.cell_centers(res)
    assert c.shape == (2 + 120 * 7**res, 2)  # H3 numHexagons; res 0: tests/nodes/test_hex_nodes.py:32
    assert pent.sum() == 12
    assert np.all(np.abs(c[:, 0]) <= np.pi / 2) and np.all(c[:, 1] > -np.pi) and np.all(c[:, 1] <= np.pi)
    # all centres distinct
    xyz = H._xyz(c[:, 0], c[:, 1])
    from scipy.spatial import cKDTree

    d, _ = cKDTree(xyz).query(xyz, k=2)
    assert d[:, 1].min() > 0.3 * d[:, 1].max()


def test_published_cell_centres():
    """h3ToGeo 85283473fffffff (H3 docs) and h3_to_geo('8928308280fffff') (h3-py README)."""
    for res, lat, lon in H.PUBLISHED_CENTERS:
        got = H.nearest_center(lat, lon, res)
        assert abs(got[0] - lat) < 1e-12 and abs(got[1] - lon) < 1e-12, (res, got, (lat, lon))


def test_published_centre_is_in_the_global_cell_set():
    res, lat, lon = H.PUBLISHED_CENTERS[0]
    c, _ = H.cell_centers(res)
    d = np.abs(np.rad2deg(c) - np.array([lat, lon])).sum(axis=1)
    assert d.min() < 1e-11


@pytest.mark.parametrize("res", [0, 1, 2, 3])
def test_neighbour_tables_are_symmetric_with_h3_degrees(res):
    c, pent = H.cell_centers(res)
    nb = H.neighbours(c, pent)
    deg = (nb >= 0).sum(axis=1)
    assert np.all(deg[pent] == 5) and np.all(deg[~pent] == 6)
    pairs = {(u, int(v)) for u in range(len(c)) for v in nb[u] if v >= 0}
    assert all((v, u) in pairs for (u, v) in pairs)
    assert len(pairs) == 6 * len(c) - 12  # every cell 6 edges, pentagons 5


def test_centre_children_and_aperture_seven():
    for r in range(3):
        coarse, _ = H.cell_centers(r)
        fine, pent = H.cell_centers(r + 1)
        cc = H.center_child_positions(coarse, fine)
        assert len(set(cc.tolist())) == len(cc)
        nb = H.neighbours(fine, pent)
        fam = np.concatenate([cc[:, None], nb[cc]], axis=1)
        members = fam[fam >= 0]
        assert len(set(members.tolist())) == len(members) == len(fine)  # 7 (6) children each, a partition


def test_no_tied_latitudes():
    """generate/utils.py:15-33 on float64 hex centres is 'descending latitude': no two cells share one."""
    for res in (2, 3, 4):
        lat = H.hex_nodes_latlon(res)[:, 0]
        assert np.unique(lat).size == lat.size


def test_multiscale_edge_counts():
    """x_hops = 1 on a global mesh: every level r contributes its 6 N_r - 12 directed neighbour pairs."""
    order = np.argsort(-H.hex_nodes_latlon(2)[:, 0], kind="stable")
    e = H.multiscale_edges_hex([0, 1, 2], 1, order)
    want = sum(6 * (2 + 120 * 7**r) - 12 for r in range(3))
    assert e.shape == (2, want)
    assert np.array_equal(e, e[:, np.lexsort((e[0], e[1]))])
    back = {(int(t), int(s)) for s, t in e.T}
    assert back == {(int(s), int(t)) for s, t in e.T}


def test_restated_hex_edges_match_the_reference_control_flow(golden):
    """tests/golden/hex.npz was written by the UNMODIFIED reference (hex_icosahedron.py, HexNodes /
    LimitedAreaHexNodes, MultiScaleEdges, networkx) running over oracle/shims/h3: the oracle's own restatement of
    that control flow (k_ring & nodes, compact / uncompact, centre children) must give the same nodes and edges."""
    from oracle import ref_path as R

    g = golden("hex")
    coords = H.hex_nodes_latlon(2)
    order = np.argsort(-coords[:, 0], kind="stable")
    np.testing.assert_array_equal(coords[order].astype(np.float32), g["res2_x"])
    np.testing.assert_array_equal(g["res2_x"], g["res_0_2_x"])
    for hops in (1, 2):
        np.testing.assert_array_equal(H.multiscale_edges_hex([0, 1, 2], hops, order), g[f"res2_hops{hops}_edge_index"])
    np.testing.assert_array_equal(H.multiscale_edges_hex([0, 2], 3, order), g["res_0_2_hops3_edge_index"])
    # limited area
    coords = H.hex_nodes_latlon(3)
    order = np.argsort(-coords[:, 0], kind="stable")
    mask = R.knn_area_mask(g["lam_data_x"], coords, 150.0)
    order = order[mask[order]]
    np.testing.assert_array_equal(coords[order].astype(np.float32), g["lam_x"])
    for hops in (1, 2):
        want = g[f"lam_hops{hops}_edge_index"]
        np.testing.assert_array_equal(H.multiscale_edges_hex([0, 1, 2, 3], hops, order, in_graph=mask), want)
