"""GPU parity of the hexagonal (H3) hidden mesh - HexNodes, LimitedAreaHexNodes, MultiScaleEdges on them - against the
numpy restatement of H3's geometry (oracle/h3_restated.py; pinned to H3's published cell centres, not to the h3
library itself, which is not installable here).  Mirrors /root/reference/tests/nodes/test_hex_nodes.py and
tests/edges/test_multiscale_edges.py:50-70 but checks values."""

import numpy as np
import pytest
import torch

from oracle import h3_restated as H
from oracle import ref_path as R

pytestmark = pytest.mark.gpu

CENTRE_ATOL = 1e-12  # radians: CUDA float64 sin/cos/atan2/asin differ from numpy by a few ulp, amplified by 1/cos(lat) in the longitude


def canon(ei):
    ei = ei.cpu().numpy() if isinstance(ei, torch.Tensor) else np.asarray(ei)
    return R.canonical_sort(ei)


@pytest.mark.parametrize("res", [0, 1, 2, 3, 4])
def test_hex_cells_match_the_restatement(res):
    from anemoi_graphs_b200 import ops

    cells = ops.HexCells(res)
    assert cells.n == 2 + 120 * 7**res == ops.hex_num_cells(res)
    got = cells.latlon.cpu().numpy()
    want, pent = H.cell_centers(res)
    want = np.deg2rad(want * (180.0 / np.pi))
    np.testing.assert_allclose(got, want, rtol=0, atol=CENTRE_ATOL)  # same (face, i, j) order
    np.testing.assert_array_equal(cells.pentagon.cpu().numpy().astype(bool), pent)
    # the float32 coordinates the graph stores: identical except where a float64 ulp straddles a float32 rounding boundary
    same = (got.astype(np.float32) == want.astype(np.float32)).all(axis=1).mean()
    assert same > 0.9999


def test_published_h3_centre_is_generated():
    from anemoi_graphs_b200 import ops

    res, lat, lon = H.PUBLISHED_CENTERS[0]  # h3ToGeo 85283473fffffff
    got = np.rad2deg(ops.HexCells(res).latlon.cpu().numpy())
    assert np.abs(got - np.array([lat, lon])).sum(axis=1).min() < 1e-11


def test_hex_cells_bad_resolution():
    from anemoi_graphs_b200 import ops

    with pytest.raises(ValueError):
        ops.hex_num_cells(16)
    with pytest.raises(ValueError):
        ops.HexCells(-1)


@pytest.mark.parametrize("res", [0, 2, 3])
def test_hex_adjacency(res):
    from anemoi_graphs_b200 import ops

    cells = ops.HexCells(res)
    nb, deg = (t.cpu().numpy() for t in cells.neighbours())
    want, pent = H.cell_centers(res)
    ref = H.neighbours(want, pent)
    np.testing.assert_array_equal(deg, np.where(pent, 5, 6))
    np.testing.assert_array_equal(np.sort(nb, axis=1), np.sort(ref, axis=1))


def _oracle_order(res):
    coords = H.hex_nodes_latlon(res)
    order = R.coordinates_ordering(coords)
    # distinct latitudes: the reference's two unstable argsorts reduce to "descending latitude"
    np.testing.assert_array_equal(coords[order, 0], np.sort(coords[:, 0])[::-1])
    return coords, order


@pytest.mark.parametrize("resolution", [0, 2])
def test_hex_nodes(resolution):
    """reference tests/nodes/test_hex_nodes.py, with values."""
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HexNodes
    from anemoi_graphs_b200.nodes.builders.base import BaseNodeBuilder

    node_builder = HexNodes(resolution, "test_nodes")
    assert isinstance(node_builder, BaseNodeBuilder)
    coords = node_builder.get_coordinates()
    assert isinstance(coords, torch.Tensor) and coords.dtype == torch.float32
    assert coords.shape == (2 + 120 * 7**resolution, 2)
    graph = HexNodes(resolution, "test_nodes").update_graph(HeteroData(), {})
    for hidden in ("_resolutions", "_nx_graph", "_node_ordering"):
        assert hidden in graph["test_nodes"]
    assert len(graph["test_nodes"]["_node_ordering"]) == graph["test_nodes"].num_nodes
    want, order = _oracle_order(resolution)
    np.testing.assert_array_equal(np.asarray(graph["test_nodes"]["_node_ordering"]), order)
    np.testing.assert_allclose(graph["test_nodes"].x.cpu().numpy(), want[order].astype(np.float32), rtol=0, atol=2.5e-7)


@pytest.mark.parametrize("resolutions,hops", [(1, 1), (2, 1), (2, 2), ([0, 2], 3), (3, 1)])
def test_multiscale_edges_on_hex_nodes(resolutions, hops):
    from anemoi_graphs_b200.edges import MultiScaleEdges
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HexNodes

    graph = HexNodes(resolutions, "hex").update_graph(HeteroData(), {})
    graph = MultiScaleEdges("hex", "hex", hops).update_graph(graph)
    assert ("hex", "to", "hex") in graph.edge_types
    ei = graph[("hex", "to", "hex")].edge_index
    assert ei.dtype == torch.int32
    levels = list(range(resolutions + 1)) if isinstance(resolutions, int) else resolutions
    _, order = _oracle_order(max(levels))
    np.testing.assert_array_equal(canon(ei), H.multiscale_edges_hex(levels, hops, order))


def test_multiscale_edges_hex_fail_nodes():
    from anemoi_graphs_b200.edges import MultiScaleEdges
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HexNodes

    graph = HexNodes(1, "test_hex_nodes").update_graph(HeteroData(), {})
    graph["fail_nodes"].x = [1, 2, 3]
    graph["fail_nodes"].node_type = "FailNodes"
    with pytest.raises(AssertionError):
        MultiScaleEdges("fail_nodes", "fail_nodes", 1).update_graph(graph)


def _patch_graph():
    from anemoi_graphs_b200.graph import HeteroData

    lat, lon = np.meshgrid(np.linspace(35.0, 65.0, 61), np.linspace(-10.0, 30.0, 81), indexing="ij")
    x = np.deg2rad(np.stack([lat.reshape(-1), lon.reshape(-1)], axis=1)).astype(np.float32)
    graph = HeteroData()
    graph["data"].x = torch.from_numpy(x)
    graph["data"].node_type = "LatLonNodes"
    return graph, x


@pytest.mark.parametrize("hops", [1, 2])
def test_limited_area_hex_nodes_and_edges(hops):
    """LimitedAreaHexNodes + MultiScaleEdges: only cells whose every descendant is a node join a coarse level
    (h3.compact, hex_icosahedron.py:206-210), the disk is walked through absent cells (k_ring & nodes)."""
    from anemoi_graphs_b200.edges import MultiScaleEdges
    from anemoi_graphs_b200.nodes import LimitedAreaHexNodes

    graph, data_x = _patch_graph()
    res = 3
    graph = LimitedAreaHexNodes(res, "data", "lam", margin_radius_km=150.0).update_graph(graph, {})
    assert graph["lam"].node_type == "LimitedAreaHexNodes"
    coords, order = _oracle_order(res)
    mask = R.knn_area_mask(data_x, coords, 150.0)
    order = order[mask[order]]
    assert 200 < len(order) < len(coords) // 4
    np.testing.assert_array_equal(np.asarray(graph["lam"]["_node_ordering"]), order)
    MultiScaleEdges("lam", "lam", hops).update_graph(graph)
    want = H.multiscale_edges_hex(list(range(res + 1)), hops, order, in_graph=mask)
    assert want.shape[1] > 6 * len(order)  # the finest level alone has ~6 per node; coarser complete cells add more
    np.testing.assert_array_equal(canon(graph[("lam", "to", "lam")].edge_index), want)


def test_hex_against_the_reference_control_flow(golden):
    """tests/golden/hex.npz: the UNMODIFIED reference hexagonal path over the h3 shim (oracle/make_golden.py make_hex)."""
    from anemoi_graphs_b200.edges import MultiScaleEdges
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HexNodes, LimitedAreaHexNodes

    g = golden("hex")
    for tag, resolution, hops_list in (("res2", 2, (1, 2)), ("res_0_2", [0, 2], (3,))):
        for hops in hops_list:
            graph = HexNodes(resolution, "h").update_graph(HeteroData(), {})
            np.testing.assert_allclose(graph["h"].x.cpu().numpy(), g[f"{tag}_x"], rtol=0, atol=2.5e-7)
            MultiScaleEdges("h", "h", hops).update_graph(graph)
            np.testing.assert_array_equal(canon(graph[("h", "to", "h")].edge_index), g[f"{tag}_hops{hops}_edge_index"])
    for hops in (1, 2):
        graph = HeteroData()
        graph["data"].x = torch.from_numpy(g["lam_data_x"])
        graph["data"].node_type = "LatLonNodes"
        graph = LimitedAreaHexNodes(3, "data", "lam", margin_radius_km=150.0).update_graph(graph, {})
        np.testing.assert_allclose(graph["lam"].x.cpu().numpy(), g["lam_x"], rtol=0, atol=2.5e-7)
        MultiScaleEdges("lam", "lam", hops).update_graph(graph)
        np.testing.assert_array_equal(canon(graph[("lam", "to", "lam")].edge_index), g[f"lam_hops{hops}_edge_index"])


def test_hex_recipe_through_graph_creator():
    """encoder / processor / decoder recipe with a hexagonal hidden mesh (docs recipe with HexNodes)."""
    from anemoi_graphs_b200 import grids
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    T = "anemoi.graphs."
    lat, lon = grids.octahedral_grid(32)
    attrs = {
        "edge_length": {"_target_": T + "edges.attributes.EdgeLength", "norm": "unit-std"},
        "edge_dirs": {"_target_": T + "edges.attributes.EdgeDirection", "norm": "unit-std"},
    }
    recipe = {
        "nodes": {
            "data": {"node_builder": {"_target_": T + "nodes.LatLonNodes", "latitudes": lat, "longitudes": lon}},
            "hidden": {"node_builder": {"_target_": T + "nodes.HexNodes", "resolution": 2}},
        },
        "edges": [
            {"source_name": "data", "target_name": "hidden", "attributes": attrs,
             "edge_builders": [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}]},
            {"source_name": "hidden", "target_name": "hidden", "attributes": attrs,
             "edge_builders": [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}]},
            {"source_name": "hidden", "target_name": "data", "attributes": attrs,
             "edge_builders": [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}]},
        ],
    }  # fmt: skip
    graph = GraphCreator(recipe).update_graph(HeteroData())
    dx, hx = graph["data"].x.cpu().numpy(), graph["hidden"].x.cpu().numpy()
    assert hx.shape == (5882, 2)
    np.testing.assert_array_equal(canon(graph[("data", "to", "hidden")].edge_index), R.canonical_sort(R.cutoff_edges(dx, hx, 0.6)))
    want, _ = R.knn_edges_canonical(hx, dx, 3)
    np.testing.assert_array_equal(canon(graph[("hidden", "to", "data")].edge_index), want)
    _, order = _oracle_order(2)
    np.testing.assert_array_equal(canon(graph[("hidden", "to", "hidden")].edge_index), H.multiscale_edges_hex([0, 1, 2], 1, order))
    key = ("hidden", "to", "hidden")
    ei = graph[key].edge_index.cpu().numpy()
    np.testing.assert_allclose(
        graph[key]["edge_length"].cpu().numpy(), R.edge_length(hx, hx, ei, norm="unit-std"), rtol=1e-6, atol=0
    )


def test_hex_multiscale_extreme_hops():
    """x_hops = 8 on the 122 base cells reaches most of the sphere from every cell; 9 is refused."""
    from anemoi_graphs_b200.edges import MultiScaleEdges
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HexNodes

    graph = HexNodes([0], "hex").update_graph(HeteroData(), {})
    MultiScaleEdges("hex", "hex", 8).update_graph(graph)
    _, order = _oracle_order(0)
    np.testing.assert_array_equal(canon(graph[("hex", "to", "hex")].edge_index), H.multiscale_edges_hex([0], 8, order))
    with pytest.raises(NotImplementedError):
        MultiScaleEdges("hex", "hex", 9).update_graph(HexNodes([0], "hex").update_graph(HeteroData(), {}))
