"""GPU parity of HEALPixNodes / LimitedAreaHEALPixNodes (reference tests/nodes/test_healpix.py, with values)."""

import numpy as np
import pytest
import torch

from oracle import healpix_restated as P
from oracle import ref_path as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("resolution", [1, 2, 3, 5, 7])
def test_healpix_nodes(golden, resolution):
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HEALPixNodes
    from anemoi_graphs_b200.nodes.builders.base import BaseNodeBuilder

    node_builder = HEALPixNodes(resolution, "test_nodes")
    assert isinstance(node_builder, BaseNodeBuilder)
    graph = node_builder.register_nodes(HeteroData())
    x = graph["test_nodes"].x
    assert isinstance(x, torch.Tensor) and x.dtype == torch.float32 and x.shape == (12 * 4**resolution, 2)
    assert graph["test_nodes"].node_type == "HEALPixNodes"
    want = P.healpix_nodes_x(resolution)
    got = x.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=2.5e-7)  # CUDA acos / atan2 vs numpy: float64 ulps
    assert (got == want).all(axis=1).mean() > 0.999
    if resolution in (1, 3):
        np.testing.assert_allclose(got, golden("healpix")[f"res{resolution}_x"], rtol=0, atol=2.5e-7)


@pytest.mark.parametrize("resolution", ["2", 4.3, -7])
def test_healpix_fail_init(resolution):
    from anemoi_graphs_b200.nodes import HEALPixNodes

    with pytest.raises(AssertionError):
        HEALPixNodes(resolution, "test_nodes")


def test_limited_area_healpix_nodes():
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import LimitedAreaHEALPixNodes

    lat, lon = np.meshgrid(np.linspace(35.0, 65.0, 61), np.linspace(-10.0, 30.0, 81), indexing="ij")
    data_x = np.deg2rad(np.stack([lat.reshape(-1), lon.reshape(-1)], axis=1)).astype(np.float32)
    graph = HeteroData()
    graph["data"].x = torch.from_numpy(data_x)
    graph["data"].node_type = "LatLonNodes"
    graph = LimitedAreaHEALPixNodes(5, "data", "lam", margin_radius_km=150.0).update_graph(graph, {})
    full = P.healpix_nodes_x(5)
    mask = R.knn_area_mask(data_x, full, 150.0)
    assert 50 < mask.sum() < full.shape[0] // 4
    np.testing.assert_allclose(graph["lam"].x.cpu().numpy(), full[mask], rtol=0, atol=2.5e-7)


def test_healpix_area_weights_are_nearly_equal():
    """HEALPix is an equal-area pixelisation: the Voronoi areas of its centres scatter by a few percent only."""
    from anemoi_graphs_b200.graph import HeteroData
    from anemoi_graphs_b200.nodes import HEALPixNodes
    from anemoi_graphs_b200.nodes.attributes import SphericalAreaWeights

    graph = HEALPixNodes(4, "n").update_graph(HeteroData(), {"w": {"_target_": "anemoi.graphs.nodes.attributes.SphericalAreaWeights", "dtype": "float64"}})
    w = graph["n"]["w"].cpu().numpy()[:, 0]
    want = R.spherical_area_weights(graph["n"].x.cpu().numpy(), None, "float64")[:, 0]
    np.testing.assert_allclose(w, want, rtol=1e-9)
    assert abs(w.sum() - 4 * np.pi) < 1e-6 and w.std() / w.mean() < 0.1
