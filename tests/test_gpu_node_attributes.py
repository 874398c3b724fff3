"""GPU parity of the node attributes: SphericalAreaWeights - scipy's SphericalVoronoi areas -
against golden fixtures written by the UNMODIFIED reference class (tests/golden/area_weights.npz,
oracle/make_golden.py) and against the oracle's scipy call at larger sizes.  Mirrors
/root/reference/tests/nodes/test_node_attributes.py."""

import numpy as np
import pytest
import torch

from oracle import ref_path as R

pytestmark = pytest.mark.gpu

NORMS = [None, "l1", "l2", "unit-max", "unit-range", "unit-std"]
RAW_RTOL = 1e-9  # float64 areas: same half-spaces and the same solid-angle formula as scipy, different summation order
ATTR_RTOL = 1e-6  # float32 attribute values (north star)


def graph_of(x):
    from anemoi_graphs_b200.graph import HeteroData

    graph = HeteroData()
    graph["test_nodes"].x = torch.as_tensor(x)
    graph["test_nodes"].node_type = "LatLonNodes"
    return graph


@pytest.fixture
def graph_with_nodes():
    """reference tests/conftest.py: 12 nodes over the globe."""
    lats, lons = [-0.15, 0, 0.15], [0, 0.25, 0.5, 0.75]
    coords = np.array([[lat, lon] for lat in lats for lon in lons])
    return graph_of(torch.tensor(2 * torch.pi * coords, dtype=torch.float32))


@pytest.mark.parametrize("name", ["o24", "tri3", "random"])
def test_spherical_area_weights_match_reference(golden, name):
    from anemoi_graphs_b200.nodes.attributes import SphericalAreaWeights

    g = golden("area_weights")
    graph = graph_of(g[f"{name}_x"])
    raw = SphericalAreaWeights(norm=None, dtype="float64").compute(graph, "test_nodes")
    assert raw.dtype == torch.float64 and raw.shape == (g[f"{name}_x"].shape[0], 1)
    np.testing.assert_allclose(raw.numpy(), g[f"{name}_raw64"], rtol=RAW_RTOL, atol=0)
    for norm in NORMS:
        got = SphericalAreaWeights(norm=norm).compute(graph, "test_nodes")
        assert got.dtype == torch.float32
        # unit-range maps the smallest cell to 0: absolute tolerance (1e-6 of the range, which is 1) for that norm
        atol = ATTR_RTOL if norm == "unit-range" else 0
        np.testing.assert_allclose(got.numpy(), g[f"{name}_{norm}"], rtol=ATTR_RTOL, atol=atol)


def test_area_weights_o96_vs_scipy_and_total():
    """40 320 generators: every cell against scipy, and the areas tile the sphere."""
    from anemoi_graphs_b200 import grids, ops

    lat, lon = grids.octahedral_grid(96)
    x = grids.latlon_deg_to_x(lat, lon)
    got = ops.voronoi_areas(x.cuda()).cpu().numpy()
    want = R.spherical_area_weights(x.numpy(), None, "float64")[:, 0]
    np.testing.assert_allclose(got, want, rtol=RAW_RTOL, atol=0)
    np.testing.assert_allclose(got.sum(), 4 * np.pi, rtol=1e-7)  # float32 generators: the hull is not quite the sphere


def test_area_weights_o1280_tile_the_sphere():
    from anemoi_graphs_b200 import grids, ops

    lat, lon = grids.octahedral_grid(1280)
    x = grids.latlon_deg_to_x(lat, lon).cuda()
    areas = ops.voronoi_areas(x)
    assert areas.shape == (6599680,) and bool((areas > 0).all())
    np.testing.assert_allclose(float(areas.sum()), 4 * np.pi, rtol=1e-7)
    # a sample of cells against scipy on the generators around them is not possible (SphericalVoronoi is global);
    # the O96 test above pins cell-by-cell parity, this one the full-size bookkeeping (retries, no cell lost)


@pytest.mark.parametrize("norm", [None, "l1", "l2", "unit-max", "unit-std"])
def test_uniform_weights(graph_with_nodes, norm):
    from anemoi_graphs_b200.nodes.attributes import UniformWeights

    weights = UniformWeights(norm=norm).compute(graph_with_nodes, "test_nodes")
    assert isinstance(weights, torch.Tensor)
    assert weights.shape[0] == graph_with_nodes["test_nodes"].x.shape[0]
    want = R.normalise(np.ones((12, 1)), norm).astype(np.float32)
    np.testing.assert_allclose(weights.numpy(), want, rtol=ATTR_RTOL)


@pytest.mark.parametrize("norm", ["l3", "invalide"])
def test_uniform_weights_fail(graph_with_nodes, norm):
    from anemoi_graphs_b200.nodes.attributes import UniformWeights

    with pytest.raises(ValueError):
        UniformWeights(norm=norm).compute(graph_with_nodes, "test_nodes")


def test_area_weights(graph_with_nodes):
    """reference test_area_weights: the 12-node graph, where every cell needs every other generator."""
    from anemoi_graphs_b200.nodes.attributes import AreaWeights

    weights = AreaWeights().compute(graph_with_nodes, "test_nodes")
    assert isinstance(weights, torch.Tensor)
    assert weights.shape[0] == graph_with_nodes["test_nodes"].x.shape[0]
    want = R.spherical_area_weights(graph_with_nodes["test_nodes"].x.numpy())
    np.testing.assert_allclose(weights.numpy(), want, rtol=ATTR_RTOL)


@pytest.mark.parametrize("radius", [-1.0, "hello", None])
def test_area_weights_fail(graph_with_nodes, radius):
    from anemoi_graphs_b200.nodes.attributes import AreaWeights

    with pytest.raises(ValueError):
        AreaWeights(radius=radius).compute(graph_with_nodes, "test_nodes")


def test_area_weights_duplicate_generators():
    from anemoi_graphs_b200.nodes.attributes import SphericalAreaWeights

    x = np.deg2rad(np.array([[0, 0], [0, 90], [0, 180], [0, 270], [80, 0], [-80, 0], [0, 90]], dtype=np.float32))
    with pytest.raises(ValueError, match="Duplicate generators"):
        SphericalAreaWeights().compute(graph_of(x), "test_nodes")


def test_boolean_masks(graph_with_nodes):
    from anemoi_graphs_b200.nodes.attributes import BooleanAndMask, BooleanNot, BooleanOrMask

    g = graph_with_nodes
    a = torch.tensor([True, False] * 6)
    b = torch.tensor([True, True, False] * 4)
    g["test_nodes"]["a"], g["test_nodes"]["b"] = a, b
    np.testing.assert_array_equal(BooleanNot("a").compute(g, "test_nodes").numpy()[:, 0], ~a.numpy())
    np.testing.assert_array_equal(BooleanAndMask(["a", "b"]).compute(g, "test_nodes").numpy()[:, 0], (a & b).numpy())
    np.testing.assert_array_equal(BooleanOrMask(["a", BooleanNot("b")]).compute(g, "test_nodes").numpy()[:, 0], (a | ~b).numpy())


def test_area_weights_through_graph_creator():
    """the recipe route: ``attributes: {area_weight: {_target_: anemoi.graphs.nodes.attributes.AreaWeights, norm: unit-max}}``."""
    from anemoi_graphs_b200 import grids
    from anemoi_graphs_b200.create import GraphCreator
    from anemoi_graphs_b200.graph import HeteroData

    T = "anemoi.graphs."
    lat, lon = grids.octahedral_grid(16)
    recipe = {
        "nodes": {
            "data": {
                "node_builder": {"_target_": T + "nodes.LatLonNodes", "latitudes": lat, "longitudes": lon},
                "attributes": {"area_weight": {"_target_": T + "nodes.attributes.AreaWeights", "norm": "unit-max"}},
            },
            "hidden": {
                "node_builder": {"_target_": T + "nodes.TriNodes", "resolution": 2},
                "attributes": {"area_weight": {"_target_": T + "nodes.attributes.SphericalAreaWeights", "norm": "l1"}},
            },
        },
        "edges": [],
    }
    graph = GraphCreator(recipe).update_graph(HeteroData())
    for name, norm in (("data", "unit-max"), ("hidden", "l1")):
        want = R.spherical_area_weights(graph[name].x.numpy(), norm)
        np.testing.assert_allclose(graph[name]["area_weight"].numpy(), want, rtol=ATTR_RTOL)


def test_area_weights_too_few_generators():
    from anemoi_graphs_b200 import ops

    x = torch.tensor([[0.0, 0.0], [0.5, 1.0], [-0.5, 2.0]], dtype=torch.float32).cuda()
    with pytest.raises(ValueError):
        ops.voronoi_areas(x)


def test_area_weights_random_cloud_needs_retries():
    """A random cloud: a quarter of the cells are not closed by 16 neighbours and go through the k = 32 / 63 passes."""
    from anemoi_graphs_b200 import ops

    rng = np.random.default_rng(11)
    x = np.stack([np.arcsin(rng.uniform(-1, 1, 20000)), rng.uniform(0, 2 * np.pi, 20000)], 1).astype(np.float32)
    got = ops.voronoi_areas(torch.from_numpy(x).cuda()).cpu().numpy()
    want = R.spherical_area_weights(x, None, "float64")[:, 0]
    np.testing.assert_allclose(got, want, rtol=RAW_RTOL, atol=0)


@pytest.mark.parametrize("n", [4, 5, 8, 20])
def test_area_weights_very_few_generators(n):
    """Cells as large as a hemisphere: every generator needs every other one (the exhaustive pass)."""
    from anemoi_graphs_b200 import ops

    rng = np.random.default_rng(100 + n)
    while True:  # scipy needs the origin inside the hull; redraw until it is
        x = np.stack([np.arcsin(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)], 1).astype(np.float32)
        try:
            want = R.spherical_area_weights(x, None, "float64")[:, 0]
        except Exception:
            continue
        if np.isfinite(want).all() and abs(want.sum() - 4 * np.pi) < 1e-6:
            break
    # cells wider than the gnomonic hemisphere go through the exhaustive spherical form (agx_voronoi_areas_small)
    got = ops.voronoi_areas(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)
    np.testing.assert_allclose(got.sum(), 4 * np.pi, rtol=1e-8)
