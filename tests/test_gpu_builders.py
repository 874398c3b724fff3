"""GPU parity of the reference-facing builder API (GraphCreator, KNNEdges, CutOffEdges, MultiScaleEdges,
EdgeLength, EdgeDirection, TriNodes) against golden fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py) and against the oracle.  Reads like the reference's own
tests (tests/edges/*, tests/test_create.py) but checks values, not just types."""

import numpy as np
import pytest
import torch

from anemoi_graphs_b200 import grids
from oracle import ref_path as R

pytestmark = pytest.mark.gpu

T = "anemoi.graphs."
ATTR_RTOL = 1e-6


def canon(ei):
    ei = ei.cpu().numpy() if isinstance(ei, torch.Tensor) else np.asarray(ei)
    return R.canonical_sort(ei)


def attr_cfg(norm="unit-std"):
    return {
        "edge_length": {"_target_": T + "edges.attributes.EdgeLength", "norm": norm},
        "edge_dirs": {"_target_": T + "edges.attributes.EdgeDirection", "norm": norm},
    }


def latlon_nodes(lat, lon):
    return {"node_builder": {"_target_": T + "nodes.LatLonNodes", "latitudes": lat, "longitudes": lon}}


def tri_nodes(res):
    return {"node_builder": {"_target_": T + "nodes.TriNodes", "resolution": res}}


def edges(src, dst, builders, attributes=None):
    return {"source_name": src, "target_name": dst, "edge_builders": builders, "attributes": attributes or {}}


def build(nodes, edge_cfgs):
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    return GraphCreator({"nodes": nodes, "edges": edge_cfgs}).update_graph(HeteroData())


def by_canonical_order(graph, key, name):
    ei = graph[key].edge_index.numpy()
    order = np.lexsort((ei[0], ei[1]))
    return graph[key][name].numpy()[order]


def test_tri_nodes_match_reference(golden):
    g = golden("tri_nodes")
    for res in range(5):
        graph = build({"hidden": tri_nodes(res)}, [])
        x = graph["hidden"].x
        assert x.dtype == torch.float32 and not x.is_cuda
        np.testing.assert_array_equal(x.numpy().view(np.int32), g[f"res{res}_x"].view(np.int32))
        np.testing.assert_array_equal(np.asarray(graph["hidden"]["_node_ordering"]), g[f"res{res}_node_ordering"])
        assert graph["hidden"].node_type == "TriNodes"
        for hidden in ("_resolutions", "_nx_graph", "_node_ordering", "_area_mask_builder"):
            assert hidden in graph["hidden"]  # reference: tests/nodes/test_tri_nodes.py:35-43
    assert build({"hidden": tri_nodes(2)}, [])["hidden"].x.shape == (162, 2)  # tests/nodes/test_tri_nodes.py:32


def test_toy_recipe_matches_reference(golden):
    g = golden("toy")
    nodes = {"data": latlon_nodes(g["data_lat_deg"], g["data_lon_deg"]), "hidden": tri_nodes(2)}
    graph = build(
        nodes,
        [
            edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg("l2")),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg("l2")),
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg("l2")),
        ],
    )
    np.testing.assert_array_equal(graph["data"].x.numpy().view(np.int32), g["data_x"].view(np.int32))
    np.testing.assert_array_equal(graph["hidden"].x.numpy().view(np.int32), g["hidden_x"].view(np.int32))
    cases = (
        (("data", "to", "hidden"), "cutoff", "CutOffEdges"),
        (("hidden", "to", "hidden"), "multiscale1", "MultiScaleEdges"),
        (("hidden", "to", "data"), "knn3", "KNNEdges"),
    )
    for key, tag, edge_type in cases:
        store = graph[key]
        assert store.edge_index.dtype == torch.int32 and store.edge_type == edge_type
        ref = g[f"{tag}_edge_index"]
        np.testing.assert_array_equal(canon(store.edge_index), canon(ref))
        order = np.lexsort((ref[0], ref[1]))
        short = {"cutoff": "cutoff", "multiscale1": "ms1", "knn3": "knn3"}[tag]
        want = g[f"{short}_len_l2"][order]
        np.testing.assert_allclose(by_canonical_order(graph, key, "edge_length"), want, rtol=ATTR_RTOL, atol=0)
        want = g[f"{short}_dir_rot_l2"][order]
        got = by_canonical_order(graph, key, "edge_dirs")
        np.testing.assert_allclose(got, want, rtol=ATTR_RTOL, atol=ATTR_RTOL * np.abs(want).max())
        assert store.edge_length.dtype == torch.float32 and store.edge_length.shape == (ref.shape[1], 1)
        assert store.edge_dirs.shape == (ref.shape[1], 2)


def test_x_hops_2_and_merged_builders(golden):
    g = golden("toy")
    nodes = {"data": latlon_nodes(g["data_lat_deg"], g["data_lon_deg"]), "hidden": tri_nodes(2)}
    graph = build(nodes, [edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 2}])])
    np.testing.assert_array_equal(canon(graph[("hidden", "to", "hidden")].edge_index), canon(g["multiscale2_edge_index"]))
    graph = build(
        nodes,
        [
            edges(
                "data",
                "hidden",
                [
                    {"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6},
                    {"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 5},
                ],
            ),
        ],
    )
    store = graph[("data", "to", "hidden")]
    # concat_edges: columns sorted lexicographically and de-duplicated - ORDER included (utils.py:66-81)
    np.testing.assert_array_equal(store.edge_index.numpy(), g["cutoff_plus_knn5_edge_index"])
    assert store.edge_type == str(g["cutoff_plus_knn5_edge_type"])


def test_masked_builders_match_reference(golden):
    from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges
    from anemoi_graphs_b200.graph import HeteroData

    g = golden("toy")
    graph = HeteroData()
    graph["data"].x = torch.from_numpy(g["data_x"])
    graph["hidden"].x = torch.from_numpy(g["hidden_x"])
    graph["data"]["m"] = torch.from_numpy(g["data_mask"])
    graph["hidden"]["m"] = torch.from_numpy(g["hidden_mask"])
    KNNEdges("hidden", "data", 3, source_mask_attr_name="m", target_mask_attr_name="m").update_graph(graph)
    CutOffEdges("data", "hidden", 0.6, source_mask_attr_name="m", target_mask_attr_name="m").update_graph(graph)
    np.testing.assert_array_equal(canon(graph[("hidden", "to", "data")].edge_index), canon(g["masked_knn3_edge_index"]))
    np.testing.assert_array_equal(canon(graph[("data", "to", "hidden")].edge_index), canon(g["masked_cutoff_edge_index"]))


def test_o96_res5_full_recipe(golden):
    """Config 1 of BASELINE.json: the documented 62 980 / 81 900 / 120 960 edges, bit-exact index sets."""
    g = golden("o96_res5")
    lat, lon = grids.octahedral_grid(96)
    graph = build(
        {"data": latlon_nodes(lat, lon), "hidden": tri_nodes(5)},
        [
            edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg()),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg()),
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg()),
        ],
    )
    np.testing.assert_array_equal(graph["hidden"].x.numpy().view(np.int32), g["hidden_x"].view(np.int32))
    cut = canon(graph[("data", "to", "hidden")].edge_index)
    ms = canon(graph[("hidden", "to", "hidden")].edge_index)
    knn = canon(graph[("hidden", "to", "data")].edge_index)
    assert (cut.shape[1], ms.shape[1], knn.shape[1]) == (62980, 81900, 120960)
    np.testing.assert_array_equal(cut, g["cutoff_edge_index"])
    np.testing.assert_array_equal(ms, g["multiscale_edge_index"])
    want, info = R.knn_edges_canonical(g["hidden_x"], graph["data"].x.numpy(), 3)
    np.testing.assert_array_equal(knn, want)
    ref = g["knn3_edge_index"]
    tied = info["tied_queries"]
    np.testing.assert_array_equal(knn[:, ~np.isin(knn[1], tied)], ref[:, ~np.isin(ref[1], tied)])
    stride = int(g["attr_sample_stride"])
    for key, tag in ((("data", "to", "hidden"), "cutoff"), (("hidden", "to", "hidden"), "multiscale")):
        got = by_canonical_order(graph, key, "edge_length")[::stride]
        np.testing.assert_allclose(got, g[f"{tag}_edge_length_sample"], rtol=ATTR_RTOL, atol=0)
        want_d = g[f"{tag}_edge_dirs_sample"]
        got = by_canonical_order(graph, key, "edge_dirs")[::stride]
        np.testing.assert_allclose(got, want_d, rtol=ATTR_RTOL, atol=ATTR_RTOL * np.abs(want_d).max())


def test_device_resident_graph_stays_on_gpu(golden):
    from anemoi_graphs_b200 import device as agx_device

    g = golden("toy")
    prev = agx_device.set_resident(True)
    try:
        nodes = {"data": latlon_nodes(g["data_lat_deg"], g["data_lon_deg"]), "hidden": tri_nodes(2)}
        graph = build(
            nodes, [edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg())]
        )
    finally:
        agx_device.set_resident(prev)
    store = graph[("hidden", "to", "data")]
    assert graph["data"].x.is_cuda and store.edge_index.is_cuda and store.edge_length.is_cuda and store.edge_dirs.is_cuda
    np.testing.assert_array_equal(canon(store.edge_index), canon(g["knn3_edge_index"]))


def test_clean_save_reload(tmp_path, golden):
    """tests/test_create.py:20-56 of the reference: dtypes, no private attributes after clean, file reloads."""
    from anemoi_graphs_b200.create import GraphCreator

    g = golden("toy")
    recipe = {
        "nodes": {"data": latlon_nodes(g["data_lat_deg"], g["data_lon_deg"]), "hidden": tri_nodes(2)},
        "edges": [
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg()),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg()),
        ],
    }
    path = tmp_path / "graph.pt"
    graph = GraphCreator(recipe).create(save_path=path)
    for name in graph.node_types:
        assert graph[name].x.dtype == torch.float32
        assert not [a for a in graph[name] if a.startswith("_")]
    for key in graph.edge_types:
        assert graph[key].edge_index.dtype == torch.int32
        assert not [a for a in graph[key] if a.startswith("_")]
        for a in ("edge_length", "edge_dirs"):
            assert graph[key][a].dtype == torch.float32
    loaded = torch.load(path, weights_only=False)
    assert loaded.node_types == graph.node_types and loaded.edge_types == graph.edge_types
    np.testing.assert_array_equal(
        loaded[("hidden", "to", "data")].edge_index.numpy(), graph[("hidden", "to", "data")].edge_index.numpy()
    )


def test_attribute_api(golden):
    """tests/edges/test_edge_attributes.py of the reference + values."""
    from anemoi_graphs_b200.edges.attributes import EdgeDirection, EdgeLength
    from anemoi_graphs_b200.graph import HeteroData

    g = golden("toy")
    graph = HeteroData()
    graph["data"].x = torch.from_numpy(g["data_x"])
    graph["hidden"].x = torch.from_numpy(g["hidden_x"])
    graph[("hidden", "to", "data")].edge_index = torch.from_numpy(g["knn3_edge_index"])
    key = ("hidden", "to", "data")
    for norm in ["l1", "l2", "unit-max", "unit-std"]:
        for rot in (True, False):
            v = EdgeDirection(norm=norm, luse_rotated_features=rot).compute(graph, key)
            assert isinstance(v, torch.Tensor) and v.dtype == torch.float32
        v = EdgeLength(norm=norm).compute(graph, key)
        n = norm.replace("-", "_")
        np.testing.assert_allclose(v.numpy(), g[f"knn3_len_{n}"], rtol=ATTR_RTOL, atol=0)
    v = EdgeLength(norm="unit-max", invert=True).compute(graph, key)
    np.testing.assert_allclose(v.numpy(), g["knn3_len_inv_unit_max"], rtol=0, atol=ATTR_RTOL)
    with pytest.raises(AssertionError):
        EdgeLength().compute(graph, ("hidden", "to", "nope"))
    with pytest.raises(ValueError):
        EdgeLength(norm="bogus").compute(graph, key)


def _lam_graph(g):
    from anemoi_graphs_b200.graph import HeteroData

    graph = HeteroData()
    graph["data"].x = torch.from_numpy(g["data_x"])
    graph["data"].node_type = "LatLonNodes"
    graph["data"]["cutout"] = torch.from_numpy(g["cutout"])
    return graph


@pytest.mark.parametrize("hops", [1, 2])
def test_limited_area_tri_nodes_and_edges(golden, hops):
    """BASELINE config 4 in miniature: LimitedAreaTriNodes + MultiScaleEdges + masked KNN / CutOff vs the reference."""
    from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges, MultiScaleEdges
    from anemoi_graphs_b200.nodes import LimitedAreaTriNodes

    g = golden("lam")
    graph = _lam_graph(g)
    graph = LimitedAreaTriNodes(6, "data", "lam", mask_attr_name="cutout", margin_radius_km=100.0).update_graph(graph, {})
    assert graph["lam"].node_type == "LimitedAreaTriNodes"
    np.testing.assert_array_equal(graph["lam"].x.numpy().view(np.int32), g["lam_x"].view(np.int32))
    np.testing.assert_array_equal(np.asarray(graph["lam"]["_node_ordering"]), g["lam_node_ordering"])
    MultiScaleEdges("lam", "lam", hops).update_graph(graph)
    np.testing.assert_array_equal(canon(graph[("lam", "to", "lam")].edge_index), g[f"lam_hops{hops}_edge_index"])
    KNNEdges("lam", "data", 4, target_mask_attr_name="cutout").update_graph(graph)
    CutOffEdges("data", "lam", 0.6, source_mask_attr_name="cutout").update_graph(graph)
    np.testing.assert_array_equal(canon(graph[("lam", "to", "data")].edge_index), canon(g["lam_knn4_edge_index"]))
    np.testing.assert_array_equal(canon(graph[("data", "to", "lam")].edge_index), canon(g["lam_cutoff_edge_index"]))


@pytest.mark.parametrize("hops", [1, 2])
def test_stretched_tri_nodes_and_edges(golden, hops):
    from anemoi_graphs_b200.edges import CutOffEdges, KNNEdges, MultiScaleEdges
    from anemoi_graphs_b200.nodes import StretchedTriNodes
    from anemoi_graphs_b200.utils import get_grid_reference_distance

    g = golden("lam")
    graph = _lam_graph(g)
    graph = StretchedTriNodes(2, 6, "str", "data", "cutout", margin_radius_km=100.0).update_graph(graph, {})
    np.testing.assert_array_equal(graph["str"].x.numpy().view(np.int32), g["str_x"].view(np.int32))
    np.testing.assert_array_equal(np.asarray(graph["str"]["_node_ordering"]), g["str_node_ordering"])
    MultiScaleEdges("str", "str", hops).update_graph(graph)
    np.testing.assert_array_equal(canon(graph[("str", "to", "str")].edge_index), g[f"str_hops{hops}_edge_index"])
    assert get_grid_reference_distance(graph["str"].x) == float(g["str_reference_distance"])
    KNNEdges("str", "data", 4).update_graph(graph)
    CutOffEdges("data", "str", 0.6).update_graph(graph)
    np.testing.assert_array_equal(canon(graph[("str", "to", "data")].edge_index), canon(g["str_knn4_edge_index"]))
    np.testing.assert_array_equal(canon(graph[("data", "to", "str")].edge_index), canon(g["str_cutoff_edge_index"]))


# ------------------------------------------------------------------------------------------------
# post-processor: RemoveUnconnectedNodes (reference tests/processors/test_post_process.py + a randomised check
# against the reference's own dict-based algorithm)
# ------------------------------------------------------------------------------------------------
def _graph_with_isolated_nodes():
    from anemoi_graphs_b200.graph import HeteroData

    graph = HeteroData()
    graph["test_nodes"].x = torch.tensor([[1], [2], [3], [4], [5], [6]])
    graph["test_nodes"]["mask_attr"] = torch.tensor([[1], [1], [1], [0], [0], [0]], dtype=torch.bool)
    graph["test_nodes", "to", "test_nodes"].edge_index = torch.tensor([[2, 3, 4], [1, 2, 3]])
    return graph


def test_remove_unconnected_nodes_reference_cases():
    from anemoi_graphs_b200.processors import RemoveUnconnectedNodes

    graph = RemoveUnconnectedNodes(nodes_name="test_nodes", ignore=None, save_mask_indices_to_attr=None).update_graph(
        _graph_with_isolated_nodes()
    )
    assert graph["test_nodes"].num_nodes == 4
    assert torch.equal(graph["test_nodes"].x, torch.tensor([[2], [3], [4], [5]]))
    assert "original_indices" not in graph["test_nodes"]

    graph = RemoveUnconnectedNodes(
        nodes_name="test_nodes", ignore=None, save_mask_indices_to_attr="original_indices"
    ).update_graph(_graph_with_isolated_nodes())
    assert graph["test_nodes"].num_nodes == 4
    assert torch.equal(graph["test_nodes", "to", "test_nodes"].edge_index, torch.tensor([[1, 2, 3], [0, 1, 2]]))
    assert torch.equal(graph["test_nodes"].original_indices, torch.tensor([[1], [2], [3], [4]]))

    graph = RemoveUnconnectedNodes(nodes_name="test_nodes", ignore="mask_attr", save_mask_indices_to_attr=None).update_graph(
        _graph_with_isolated_nodes()
    )
    assert graph["test_nodes"].num_nodes == 5
    assert torch.equal(graph["test_nodes"].x, torch.tensor([[1], [2], [3], [4], [5]]))
    assert torch.equal(graph["test_nodes", "to", "test_nodes"].edge_index, torch.tensor([[2, 3, 4], [1, 2, 3]]))


@pytest.mark.parametrize("resident", [False, True])
def test_remove_unconnected_nodes_random_graph_matches_reference_algorithm(resident):
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.processors import RemoveUnconnectedNodes

    rng = np.random.default_rng(7)
    n_a, n_b = 5000, 300
    used = rng.choice(n_a, size=1800, replace=False)
    e_ab = np.stack([rng.choice(used, size=20000), rng.integers(0, n_b, size=20000)]).astype(np.int32)
    e_ba = np.stack([rng.integers(0, n_b, size=9000), rng.choice(used[:900], size=9000)]).astype(np.int32)
    e_aa = np.stack([rng.choice(used, size=4000), rng.choice(used, size=4000)]).astype(np.int32)
    keep_anyway = np.zeros((n_a, 1), dtype=bool)
    keep_anyway[rng.choice(n_a, size=50, replace=False)] = True
    where = (lambda t: t.cuda()) if resident else (lambda t: t)
    graph = HeteroData()
    graph["a"].x = where(torch.from_numpy(rng.random((n_a, 2)).astype(np.float32)))
    graph["a"]["w"] = where(torch.arange(n_a, dtype=torch.float32).reshape(-1, 1))
    graph["a"]["keep"] = where(torch.from_numpy(keep_anyway))
    graph["b"].x = where(torch.from_numpy(rng.random((n_b, 2)).astype(np.float32)))
    for key, e in ((("a", "to", "b"), e_ab), (("b", "to", "a"), e_ba), (("a", "to", "a"), e_aa)):
        graph[key].edge_index = where(torch.from_numpy(e.copy()))
    # the reference's algorithm (post_process.py:45-60,133-149), restated with numpy
    mask = keep_anyway[:, 0].copy()
    mask[e_ab[0]] = True
    mask[e_ba[1]] = True
    mask[e_aa[0]] = True
    mask[e_aa[1]] = True
    mapping = dict(zip(np.where(mask)[0].tolist(), range(int(mask.sum()))))
    remap = np.vectorize(mapping.get)
    graph = RemoveUnconnectedNodes("a", save_mask_indices_to_attr="orig", ignore="keep").update_graph(graph)
    assert graph["a"].num_nodes == int(mask.sum())
    assert graph["a"].x.is_cuda == resident and graph["a", "to", "b"].edge_index.is_cuda == resident
    np.testing.assert_array_equal(graph["a"]["w"].cpu().numpy()[:, 0], np.where(mask)[0].astype(np.float32))
    np.testing.assert_array_equal(graph["a"]["orig"].cpu().numpy()[:, 0], np.where(mask)[0])
    np.testing.assert_array_equal(graph["a", "to", "b"].edge_index.cpu().numpy(), np.stack([remap(e_ab[0]), e_ab[1]]))
    np.testing.assert_array_equal(graph["b", "to", "a"].edge_index.cpu().numpy(), np.stack([e_ba[0], remap(e_ba[1])]))
    np.testing.assert_array_equal(graph["a", "to", "a"].edge_index.cpu().numpy(), np.stack([remap(e_aa[0]), remap(e_aa[1])]))
    assert graph["a", "to", "b"].edge_index.dtype == torch.int32
    assert graph["b"].num_nodes == n_b


@pytest.mark.parametrize("hidden_first", [False, True])
def test_early_target_row_is_adopted_and_equal(golden, monkeypatch, hidden_first):
    """Host-resident graph, KNN decoder over a node set the graph arrives with: its target row is sent to the host at the
    top of the build and adopted by the builder (provisional source: hidden_first False; final source: True)."""
    from anemoi_graphs_b200 import device as agx_device
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    g = golden("toy")
    monkeypatch.setattr(GraphCreator, "EARLY_ROW_MIN_TARGETS", 0)
    recipe = {
        "nodes": {"hidden": tri_nodes(2)},
        "edges": [edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg("l2"))],
    }

    def run(early):
        monkeypatch.setattr(agx_device, "EARLY_TARGET_ROW", early)
        monkeypatch.setattr(agx_device, "LAZY_NODE_ORDER", not hidden_first)
        graph = HeteroData()
        graph["data"].x = torch.from_numpy(g["data_x"]).pin_memory()
        graph["data"].node_type = "LatLonNodes"
        adopted = []
        real = agx_device.edge_index_like_input

        def spy(edge_dev, ref):
            before = len(agx_device._early_rows)
            out = real(edge_dev, ref)
            adopted.append(before - len(agx_device._early_rows))
            return out

        monkeypatch.setattr(agx_device, "edge_index_like_input", spy)
        graph = GraphCreator(recipe).update_graph(graph)
        monkeypatch.setattr(agx_device, "edge_index_like_input", real)
        assert not agx_device._early_rows
        return graph, sum(adopted)

    (a, n_a), (b, n_b) = run(True), run(False)
    assert (n_a, n_b) == (1, 0)
    key = ("hidden", "to", "data")
    ea, eb = a[key].edge_index, b[key].edge_index
    assert ea.dtype == torch.int32 and not ea.is_cuda and ea.is_pinned()
    np.testing.assert_array_equal(ea.numpy(), eb.numpy())
    np.testing.assert_array_equal(canon(ea), canon(g["knn3_edge_index"]))
    np.testing.assert_array_equal(a[key].edge_length.numpy(), b[key].edge_length.numpy())


def test_failed_node_order_sort_opens_the_gates_and_is_reported(monkeypatch):
    """Pre-launched tail: the device waits behind stream gates for index arrays the host is sorting.  A sort that fails
    must not leave it waiting: the gates open over valid (identity) arrays and the failure is raised by the build."""
    from anemoi_graphs_b200 import device as agx_device
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.generate import tri_icosahedron
    from anemoi_graphs_b200.graph import HeteroData

    def broken_sort(lat, lon, emit=None):
        raise FloatingPointError("sort failed on purpose")

    monkeypatch.setattr(agx_device, "PRELAUNCH_TAIL", True)
    monkeypatch.setattr(tri_icosahedron, "_sort_columns_host", broken_sort)
    recipe = {
        "nodes": {"hidden": tri_nodes(3)},
        "edges": [edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg("unit-max"))],
    }
    prev = agx_device.set_resident(True)
    try:
        with pytest.raises(FloatingPointError):
            GraphCreator(recipe).update_graph(HeteroData())
        torch.cuda.synchronize()  # returns: nothing is left waiting on the device
    finally:
        agx_device.set_resident(prev)
        agx_device.flush()


# ------------------------------------------------------------------------------------------------
# provisional node numbering (device.Provisional): building while the node order is still being sorted must give
# the same graph as sorting first
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prelaunch", [False, True])
@pytest.mark.parametrize("resident", [False, True])
def test_provisional_numbering_gives_the_same_graph(golden, resident, prelaunch, monkeypatch):
    from anemoi_graphs_b200 import device as agx_device

    # prelaunch: everything behind the node order is queued behind stream gates before the sort has ended
    monkeypatch.setattr(agx_device, "PRELAUNCH_TAIL", prelaunch)
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    g = golden("toy")
    mask = torch.from_numpy(g["data_mask"])
    recipe = {
        "nodes": {"hidden": tri_nodes(3)},
        "edges": [
            edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg("unit-std")),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 2}], attr_cfg("unit-range")),
            # hidden as KNN *target* may stay provisional, as *source* it may not (lower-index tie rule)
            edges("data", "hidden", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 2}], attr_cfg("l2")),
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg("unit-max")),
            edges("hidden", "data", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.4,
                                      "target_mask_attr_name": "m"}], attr_cfg(None)),
        ],
    }  # fmt: skip

    def run(lazy):
        prev_lazy, agx_device.LAZY_NODE_ORDER = agx_device.LAZY_NODE_ORDER, lazy
        prev_res = agx_device.set_resident(resident)
        try:
            graph = HeteroData()
            x = torch.from_numpy(g["data_x"])
            graph["data"].x = x.cuda() if resident else x
            graph["data"].node_type = "LatLonNodes"
            graph["data"]["m"] = mask.cuda() if resident else mask
            graph = GraphCreator(recipe).update_graph(graph)
            torch.cuda.synchronize()
            return graph
        finally:
            agx_device.LAZY_NODE_ORDER = prev_lazy
            agx_device.set_resident(prev_res)

    a, b = run(True), run(False)
    assert a["hidden"].x.is_cuda == resident
    np.testing.assert_array_equal(a["hidden"].x.cpu().numpy().view(np.int32), b["hidden"].x.cpu().numpy().view(np.int32))
    np.testing.assert_array_equal(np.asarray(a["hidden"]["_node_ordering"]), np.asarray(b["hidden"]["_node_ordering"]))
    for key in (("data", "to", "hidden"), ("hidden", "to", "hidden"), ("hidden", "to", "data")):
        ea, eb = a[key].edge_index.cpu().numpy(), b[key].edge_index.cpu().numpy()
        oa, ob = np.lexsort((ea[0], ea[1])), np.lexsort((eb[0], eb[1]))
        np.testing.assert_array_equal(ea[:, oa], eb[:, ob])
        assert a[key].edge_type == b[key].edge_type
        for name in ("edge_length", "edge_dirs"):
            va, vb = a[key][name].cpu().numpy()[oa], b[key][name].cpu().numpy()[ob]
            # the same edges with the same coordinates: identical raw values; the float64 statistics are folded in
            # edge order, which differs, so the normalised float32 values may differ in the last bit
            np.testing.assert_allclose(va, vb, rtol=3e-7, atol=1e-7 * np.abs(vb).max())


# ------------------------------------------------------------------------------------------------
# advisor findings, round 1
# ------------------------------------------------------------------------------------------------
def test_resident_graph_built_under_provisional_numbering_can_be_saved(tmp_path):
    """``torch.save`` of a device-resident graph whose TriNodes order was computed on the worker thread: nothing
    un-picklable (the Provisional, its Future) travels with the tensors, ``clean`` not required."""
    from anemoi_graphs_b200 import device as D
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    lat, lon = grids.uniform_sphere(3000, seed=3)
    prev = D.set_resident(True)
    try:
        creator = GraphCreator(
            {
                "nodes": {"data": latlon_nodes(lat, lon), "hidden": tri_nodes(3)},
                "edges": [
                    edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg()),
                    edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg()),
                    edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg()),
                ],
            }
        )
        graph = creator.update_graph(HeteroData())
        assert graph[("hidden", "to", "data")].edge_index.is_cuda
        torch.save(graph, tmp_path / "raw.pt")  # before clean: private attributes included
        creator.save(creator.clean(graph), tmp_path / "graph.pt")
    finally:
        D.set_resident(prev)
    back = torch.load(tmp_path / "graph.pt", weights_only=False)
    for key in graph.edge_types:
        assert torch.equal(back[key].edge_index.cpu(), graph[key].edge_index.cpu())


def test_reference_contract_edge_builder_plugin(golden):
    """An edge builder written for the REFERENCE's plugin contract (subclass providing ``get_adjacency_matrix`` ->
    scipy COO, edges/builder.py:63) is instantiable and gives the reference's edge list; it sees complete host
    tensors even inside GraphCreator's deferred scope."""
    from scipy.sparse import coo_matrix

    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.edges.builder import BaseEdgeBuilder
    from anemoi_graphs_b200.graph import HeteroData

    seen = {}

    class EveryThird(BaseEdgeBuilder):
        def get_adjacency_matrix(self, source_nodes, target_nodes):
            sx, tx = source_nodes["x"].numpy(), target_nodes["x"].numpy()  # host reads, like the reference's builders
            seen["src_x"] = sx.copy()
            rows = np.arange(tx.shape[0])
            cols = (3 * rows) % sx.shape[0]
            return coo_matrix((np.ones(rows.size), (rows, cols)), shape=(tx.shape[0], sx.shape[0]))

    import anemoi_graphs_b200.edges as pkg

    pkg.EveryThird = EveryThird
    try:
        lat, lon = grids.uniform_sphere(500, seed=11)
        graph = GraphCreator(
            {
                "nodes": {"hidden": tri_nodes(2), "data": latlon_nodes(lat, lon)},
                "edges": [edges("hidden", "data", [{"_target_": "anemoi_graphs_b200.edges.EveryThird"}], attr_cfg())],
            }
        ).update_graph(HeteroData())
    finally:
        del pkg.EveryThird
    hx = golden("tri_nodes")["res2_x"]
    np.testing.assert_array_equal(seen["src_x"].view(np.int32), hx.view(np.int32))  # final order, not a blank buffer
    ei = graph[("hidden", "to", "data")].edge_index.numpy()
    assert ei.dtype == np.int32
    np.testing.assert_array_equal(ei[1], np.arange(500))
    np.testing.assert_array_equal(ei[0], (3 * np.arange(500)) % 162)
    want = R.edge_length(hx, graph["data"].x.numpy(), ei, "unit-std")
    np.testing.assert_allclose(graph[("hidden", "to", "data")]["edge_length"].numpy(), want, rtol=ATTR_RTOL)


def test_describe_a_graph_built_on_the_gpu(tmp_path, capsys, golden):
    """N4 (describe.py:20-225) on the device: a graph built by GraphCreator.create, saved, described - the descriptor's
    reductions run on the GPU; sizes, isolated-node counts and attribute statistics equal the reference's expressions
    (``num_nodes - len(torch.unique(row))``, min / mean / max / std per attribute) evaluated with numpy."""
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.describe import GraphDescriptor

    g = golden("toy")
    path = tmp_path / "graph.pt"
    graph = GraphCreator(
        {
            "nodes": {"data": latlon_nodes(g["data_lat_deg"], g["data_lon_deg"]), "hidden": tri_nodes(2)},
            "edges": [
                edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg("unit-std")),
                edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg("unit-max")),
            ],
        }
    ).create(save_path=path)
    d = GraphDescriptor(path)
    assert d._device.type == "cuda"
    want_size = sum(v.numel() * v.element_size() for s in list(graph.node_stores) + list(graph.edge_stores)
                    for v in s.values() if isinstance(v, torch.Tensor))  # fmt: skip
    assert d.total_size == want_size
    for row in d.get_edge_summary():
        src, dst, n_edges, iso_src, iso_dst, dim, names = row
        ei = graph[(src, "to", dst)].edge_index.numpy()
        assert n_edges == ei.shape[1] and dim == 3 and names == "edge_length(1D), edge_dirs(2D)"
        assert iso_src == graph[src].num_nodes - np.unique(ei[0]).size
        assert iso_dst == graph[dst].num_nodes - np.unique(ei[1]).size
    assert {r[0]: r[1] for r in d.get_node_summary()} == {"data": 2000, "hidden": 162}
    for kind, where, name, dtype, lo, mean, hi, std in d.get_attribute_table():
        s, t = where.split("-->")
        v = graph[(s, "to", t)][name].numpy().astype(np.float64)
        np.testing.assert_allclose([lo, mean, hi, std], [v.min(), v.mean(), v.max(), v.std(ddof=1)], rtol=2e-5, atol=1e-6)
    d.describe()
    assert "Graph ready." in capsys.readouterr().out
