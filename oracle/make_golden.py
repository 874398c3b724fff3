"""ORACLE - TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz``.

Runs the UNMODIFIED reference (``/root/reference/src``) in this container behind the import
shims in ``oracle/shims`` and stores its outputs as fixtures.  ``/root/reference`` does not
exist on the GPU box, so tests read only the committed ``.npz`` files.

    python oracle/make_golden.py            # rewrites tests/golden/

Fixtures (all produced by the reference's own classes through ``GraphCreator`` /
``directional_edge_features`` / ``haversine_distance`` - nothing from this repository's
product code is on that path, only ``anemoi_graphs_b200.grids`` for the synthetic inputs):

* ``toy.npz``        2 000 random data points -> TriNodes(2): CutOff 0.6, MultiScale x_hops 1 and 2,
                     KNN k=3, masked KNN / CutOff, every attribute x norm; full arrays.
* ``o96_res5.npz``   config 1 (O96 -> TriNodes 5): full edge_index of the three edge sets
                     (canonical (dst, src) order), hidden nodes, cut-off radius, strided attribute
                     samples + float64 sums.
* ``tri_nodes.npz``  TriNodes coordinates + node ordering for resolutions 0-4, multi-scale edges
                     for resolution 3 with x_hops 1, 2, 3.
* ``lam.npz``        config 4 in miniature: limited-area patch + global points with a ``cutout`` mask;
                     LimitedAreaTriNodes / StretchedTriNodes coordinates, multi-scale edges (x_hops 1, 2),
                     masked KNN and cut-off edges.
* ``attr_vectors.npz`` SURVEY appendix-B style edge cases for EdgeLength / EdgeDirection.
* ``hex.npz``        the reference's hexagonal path (generate/hex_icosahedron.py, HexNodes / LimitedAreaHexNodes,
                     MultiScaleEdges) executed UNMODIFIED over ``oracle/shims/h3`` - the h3 calls answered by the
                     restatement of H3's geometry, everything else (k_ring & nodes, compact / uncompact, centre
                     children, networkx, ordering) the reference's own code.
* ``healpix.npz``    HEALPixNodes (resolutions 1, 3) of the UNMODIFIED reference over
                     ``oracle/shims/healpy`` (healpy's pix2ang answered by the restated HEALPix formulas).
* ``area_weights.npz`` SphericalAreaWeights of the reference (scipy SphericalVoronoi) on an O24 grid, TriNodes(3)
                     and 3 000 random points, raw and for every norm.
"""

from __future__ import annotations

import hashlib
import pathlib
import sys
import warnings

REPO = pathlib.Path(__file__).resolve().parents[1]
sys.path[:0] = [str(REPO / "oracle" / "shims"), "/root/reference/src", str(REPO)]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from anemoi.graphs.create import GraphCreator  # noqa: E402
from anemoi.graphs.edges.directional import directional_edge_features  # noqa: E402
from anemoi.graphs.utils import haversine_distance  # noqa: E402
from anemoi.utils.config import DotDict  # noqa: E402

from anemoi_graphs_b200 import grids  # noqa: E402

OUT = REPO / "tests" / "golden"
NORMS = [None, "l1", "l2", "unit-max", "unit-range", "unit-std"]
T = "anemoi.graphs."


def canon(ei: np.ndarray) -> np.ndarray:
    order = np.lexsort((ei[0], ei[1]))
    return np.ascontiguousarray(ei[:, order])


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def attr_cfg(norm="unit-std", rotated=True, invert=False):
    return {
        "edge_length": {"_target_": T + "edges.attributes.EdgeLength", "norm": norm, "invert": invert},
        "edge_dirs": {"_target_": T + "edges.attributes.EdgeDirection", "norm": norm, "luse_rotated_features": rotated},
    }


def build(nodes: dict, edges: list) -> object:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return GraphCreator(DotDict({"nodes": nodes, "edges": edges})).update_graph(
            __import__("torch_geometric.data", fromlist=["HeteroData"]).HeteroData()
        )


def latlon_nodes(lat, lon, attributes=None):
    return {
        "node_builder": {"_target_": T + "nodes.LatLonNodes", "latitudes": lat, "longitudes": lon},
        "attributes": attributes or {},
    }


def tri_nodes(res):
    return {"node_builder": {"_target_": T + "nodes.TriNodes", "resolution": res}, "attributes": {}}


def edges(src, dst, builders, attributes=None):
    return {"source_name": src, "target_name": dst, "edge_builders": builders, "attributes": attributes or {}}


def make_toy() -> None:
    lat, lon = grids.uniform_sphere(2000, seed=7)
    out: dict[str, np.ndarray] = {"data_lat_deg": lat, "data_lon_deg": lon}
    nodes = {"data": latlon_nodes(lat, lon), "hidden": tri_nodes(2)}
    g = build(
        nodes,
        [
            edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}]),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}]),
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}]),
        ],
    )
    out["data_x"] = g["data"].x.numpy()
    out["hidden_x"] = g["hidden"].x.numpy()
    out["hidden_node_ordering"] = np.asarray(g["hidden"]["_node_ordering"], dtype=np.int64)
    out["cutoff_edge_index"] = g[("data", "to", "hidden")].edge_index.numpy()
    out["multiscale1_edge_index"] = g[("hidden", "to", "hidden")].edge_index.numpy()
    out["knn3_edge_index"] = g[("hidden", "to", "data")].edge_index.numpy()

    g2 = build(nodes, [edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 2}])])
    out["multiscale2_edge_index"] = g2[("hidden", "to", "hidden")].edge_index.numpy()

    # several builders on one node pair (concat_edges: sorted unique columns) + self KNN on hidden
    g3 = build(
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
            edges("hidden", "hidden", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 4}]),
        ],
    )
    out["cutoff_plus_knn5_edge_index"] = g3[("data", "to", "hidden")].edge_index.numpy()
    out["cutoff_plus_knn5_edge_type"] = np.array(g3[("data", "to", "hidden")].edge_type)
    out["hidden_self_knn4_edge_index"] = g3[("hidden", "to", "hidden")].edge_index.numpy()

    # masks: boolean node attributes registered by hand, then masked builders
    from torch_geometric.data import HeteroData  # shim

    from anemoi.graphs.edges import CutOffEdges, KNNEdges

    g4 = HeteroData()
    g4["data"].x = g["data"].x
    g4["hidden"].x = g["hidden"].x
    rng = np.random.default_rng(11)
    dmask = torch.from_numpy(rng.random(2000) < 0.5)[:, None]
    hmask = torch.from_numpy(rng.random(out["hidden_x"].shape[0]) < 0.7)[:, None]
    g4["data"]["m"] = dmask
    g4["hidden"]["m"] = hmask
    out["data_mask"] = dmask.numpy()
    out["hidden_mask"] = hmask.numpy()
    KNNEdges("hidden", "data", 3, source_mask_attr_name="m", target_mask_attr_name="m").update_graph(g4)
    CutOffEdges("data", "hidden", 0.6, source_mask_attr_name="m", target_mask_attr_name="m").update_graph(g4)
    out["masked_knn3_edge_index"] = g4[("hidden", "to", "data")].edge_index.numpy()
    out["masked_cutoff_edge_index"] = g4[("data", "to", "hidden")].edge_index.numpy()

    # attributes on the three base edge sets for every norm / mode
    from anemoi.graphs.edges.attributes import EdgeDirection, EdgeLength

    for tag, key in (("cutoff", ("data", "to", "hidden")), ("ms1", ("hidden", "to", "hidden")), ("knn3", ("hidden", "to", "data"))):
        for norm in NORMS:
            n = "none" if norm is None else norm.replace("-", "_")
            out[f"{tag}_len_{n}"] = EdgeLength(norm=norm).compute(g, key).numpy()
            out[f"{tag}_dir_rot_{n}"] = EdgeDirection(norm=norm).compute(g, key).numpy()
        out[f"{tag}_len_inv_unit_max"] = EdgeLength(norm="unit-max", invert=True).compute(g, key).numpy()
        out[f"{tag}_dir_norot_unit_std"] = EdgeDirection(norm="unit-std", luse_rotated_features=False).compute(g, key).numpy()
        out[f"{tag}_dir_norot_none"] = EdgeDirection(norm=None, luse_rotated_features=False).compute(g, key).numpy()
    np.savez_compressed(OUT / "toy.npz", **out)
    print("toy.npz", {k: v.shape for k, v in out.items() if "edge_index" in k})


def make_o96() -> None:
    lat, lon = grids.octahedral_grid(96)
    nodes = {"data": latlon_nodes(lat, lon), "hidden": tri_nodes(5)}
    g = build(
        nodes,
        [
            edges("data", "hidden", [{"_target_": T + "edges.CutOffEdges", "cutoff_factor": 0.6}], attr_cfg()),
            edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}], attr_cfg()),
            edges("hidden", "data", [{"_target_": T + "edges.KNNEdges", "num_nearest_neighbours": 3}], attr_cfg()),
        ],
    )
    from anemoi.graphs.utils import get_grid_reference_distance

    out: dict[str, np.ndarray] = {}
    out["data_x_sha256"] = np.array(sha(g["data"].x.numpy()))
    out["hidden_x"] = g["hidden"].x.numpy()
    out["hidden_node_ordering"] = np.asarray(g["hidden"]["_node_ordering"], dtype=np.int32)
    out["reference_distance"] = np.array(get_grid_reference_distance(g["hidden"].x), dtype=np.float64)
    stride = 37
    for tag, key in (("cutoff", ("data", "to", "hidden")), ("multiscale", ("hidden", "to", "hidden")), ("knn3", ("hidden", "to", "data"))):
        ei = g[key].edge_index.numpy()
        order = np.lexsort((ei[0], ei[1]))
        out[f"{tag}_edge_index"] = np.ascontiguousarray(ei[:, order])
        for a in ("edge_length", "edge_dirs"):
            v = g[key][a].numpy()[order]
            out[f"{tag}_{a}_sample"] = v[::stride]
            out[f"{tag}_{a}_sum64"] = np.array(v.astype(np.float64).sum())
            out[f"{tag}_{a}_abs_sum64"] = np.array(np.abs(v.astype(np.float64)).sum())
    out["attr_sample_stride"] = np.array(stride)
    np.savez_compressed(OUT / "o96_res5.npz", **out)
    print("o96_res5.npz", {k: v.shape for k, v in out.items() if k.endswith("edge_index")})


def make_tri() -> None:
    out: dict[str, np.ndarray] = {}
    for res in range(5):
        g = build({"hidden": tri_nodes(res)}, [])
        out[f"res{res}_x"] = g["hidden"].x.numpy()
        out[f"res{res}_node_ordering"] = np.asarray(g["hidden"]["_node_ordering"], dtype=np.int32)
    for hops in (1, 2, 3):
        g = build({"hidden": tri_nodes(3)}, [edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": hops}])])
        out[f"res3_hops{hops}_edge_index"] = canon(g[("hidden", "to", "hidden")].edge_index.numpy())
    # a resolution LIST (only levels 1 and 3)
    g = build(
        {"hidden": {"node_builder": {"_target_": T + "nodes.TriNodes", "resolution": [1, 3]}, "attributes": {}}},
        [edges("hidden", "hidden", [{"_target_": T + "edges.MultiScaleEdges", "x_hops": 1}])],
    )
    out["res_1_3_hops1_edge_index"] = canon(g[("hidden", "to", "hidden")].edge_index.numpy())
    np.savez_compressed(OUT / "tri_nodes.npz", **out)
    print("tri_nodes.npz", {k: v.shape for k, v in out.items()})


def make_lam() -> None:
    """Config 4 in miniature: a 40 x 40 limited-area patch (25 km spacing, centred 50N 10E) plus 1 500 global
    points; ``cutout`` marks the patch.  LimitedAreaTriNodes(6) and StretchedTriNodes(2 -> 6) hidden meshes with
    MultiScaleEdges (x_hops 1 and 2), KNN k=4 decoder and CutOff encoder through the reference's own classes."""
    from torch_geometric.data import HeteroData  # shim

    from anemoi.graphs.edges import CutOffEdges, KNNEdges, MultiScaleEdges
    from anemoi.graphs.nodes import LimitedAreaTriNodes, StretchedTriNodes

    lam_lat, lam_lon = grids.lam_patch(40, 40, 25.0)
    glob_lat, glob_lon = grids.uniform_sphere(1500, seed=3)
    lat = np.concatenate([lam_lat, glob_lat])
    lon = np.concatenate([lam_lon, glob_lon])
    x = grids.latlon_deg_to_x(lat, lon)
    cutout = torch.zeros((x.shape[0], 1), dtype=torch.bool)
    cutout[: lam_lat.size] = True
    out: dict[str, np.ndarray] = {"data_x": x.numpy(), "cutout": cutout.numpy()}

    def base_graph():
        g = HeteroData()
        g["data"].x = x
        g["data"].node_type = "LatLonNodes"
        g["data"]["cutout"] = cutout
        return g

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # limited-area mesh
        g = base_graph()
        g = LimitedAreaTriNodes(6, "data", "lam", mask_attr_name="cutout", margin_radius_km=100.0).update_graph(g, {})
        out["lam_x"] = g["lam"].x.numpy()
        out["lam_node_ordering"] = np.asarray(g["lam"]["_node_ordering"], dtype=np.int64)
        for hops in (1, 2):
            gg = base_graph()
            gg = LimitedAreaTriNodes(6, "data", "lam", mask_attr_name="cutout", margin_radius_km=100.0).update_graph(gg, {})
            MultiScaleEdges("lam", "lam", hops).update_graph(gg)
            out[f"lam_hops{hops}_edge_index"] = canon(gg[("lam", "to", "lam")].edge_index.numpy())
        KNNEdges("lam", "data", 4, target_mask_attr_name="cutout").update_graph(g)
        CutOffEdges("data", "lam", 0.6, source_mask_attr_name="cutout").update_graph(g)
        out["lam_knn4_edge_index"] = g[("lam", "to", "data")].edge_index.numpy()
        out["lam_cutoff_edge_index"] = g[("data", "to", "lam")].edge_index.numpy()
        # stretched mesh
        g = base_graph()
        g = StretchedTriNodes(2, 6, "str", "data", "cutout", margin_radius_km=100.0).update_graph(g, {})
        out["str_x"] = g["str"].x.numpy()
        out["str_node_ordering"] = np.asarray(g["str"]["_node_ordering"], dtype=np.int64)
        for hops in (1, 2):
            gg = base_graph()
            gg = StretchedTriNodes(2, 6, "str", "data", "cutout", margin_radius_km=100.0).update_graph(gg, {})
            MultiScaleEdges("str", "str", hops).update_graph(gg)
            out[f"str_hops{hops}_edge_index"] = canon(gg[("str", "to", "str")].edge_index.numpy())
        KNNEdges("str", "data", 4).update_graph(g)
        CutOffEdges("data", "str", 0.6).update_graph(g)
        out["str_knn4_edge_index"] = g[("str", "to", "data")].edge_index.numpy()
        out["str_cutoff_edge_index"] = g[("data", "to", "str")].edge_index.numpy()
        from anemoi.graphs.utils import get_grid_reference_distance

        out["str_reference_distance"] = np.array(get_grid_reference_distance(g["str"].x), dtype=np.float64)
        out["lam_reference_distance"] = np.array(get_grid_reference_distance(torch.from_numpy(out["lam_x"])), dtype=np.float64)
    np.savez_compressed(OUT / "lam.npz", **out)
    print("lam.npz", {k: v.shape for k, v in out.items()})


def make_attr_vectors() -> None:
    h = np.float32(np.pi / 2)
    src = np.array(
        [(0.10, 0.20), (0.5, 6.28), (1.5, 1.0), (-1.5, 2.0), (h, 0.0), (0.1, 1.0), (0.0, 1.0), (0.3, 0.4), (h, 0.0),
         (0.3, 0.4), (-0.7, 3.0), (1.2, 5.9), (0.0, 0.0), (-h, 1.0)],
        dtype=np.float32,
    )  # fmt: skip
    dst = np.array(
        [(0.12, 0.25), (0.5, 0.01), (h, 0.0), (-h, 0.0), (1.5, 1.0), (0.2, 1.0), (0.0, 1.1), (0.3, 0.4), (h, 0.0),
         (-0.3, 0.4 + np.pi), (-0.7001, 3.0001), (1.2001, 0.1), (1e-4, 1e-4), (-1.5, 1.0)],
        dtype=np.float32,
    )  # fmt: skip
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rot = directional_edge_features(src.T.copy(), dst.T.copy(), True).T
        norot = directional_edge_features(src.T.copy(), dst.T.copy(), False).T
        length = haversine_distance(src, dst)
    np.savez_compressed(OUT / "attr_vectors.npz", src=src, dst=dst, dir_rotated=rot, dir_nonrotated=norot, length=length)
    print("attr_vectors.npz", rot.dtype, length.dtype)


def make_hex() -> None:
    from anemoi.graphs.edges import MultiScaleEdges
    from anemoi.graphs.nodes import HexNodes, LimitedAreaHexNodes
    from torch_geometric.data import HeteroData

    out: dict[str, np.ndarray] = {}
    for tag, resolution, hops_list in (("res2", 2, (1, 2)), ("res_0_2", [0, 2], (3,))):
        for hops in hops_list:
            g = HexNodes(resolution, "h").update_graph(HeteroData(), {})
            out[f"{tag}_x"] = g["h"].x.numpy()
            g = MultiScaleEdges("h", "h", hops).update_graph(g)
            out[f"{tag}_hops{hops}_edge_index"] = canon(g[("h", "to", "h")].edge_index.numpy())
    # limited area: a lat/lon patch as reference nodes, 150 km margin, resolution 3
    lat, lon = np.meshgrid(np.linspace(35.0, 65.0, 61), np.linspace(-10.0, 30.0, 81), indexing="ij")
    data_x = np.deg2rad(np.stack([lat.reshape(-1), lon.reshape(-1)], axis=1)).astype(np.float32)
    for hops in (1, 2):
        g = HeteroData()
        g["data"].x = torch.from_numpy(data_x)
        g["data"].node_type = "LatLonNodes"
        g = LimitedAreaHexNodes(3, "data", "lam", margin_radius_km=150.0).update_graph(g, {})
        out["lam_data_x"] = data_x
        out["lam_x"] = g["lam"].x.numpy()
        g = MultiScaleEdges("lam", "lam", hops).update_graph(g)
        out[f"lam_hops{hops}_edge_index"] = canon(g[("lam", "to", "lam")].edge_index.numpy())
    np.savez_compressed(OUT / "hex.npz", **out)
    print("hex.npz", {k: v.shape for k, v in out.items()})


def make_healpix() -> None:
    from anemoi.graphs.nodes import HEALPixNodes
    from torch_geometric.data import HeteroData

    out: dict[str, np.ndarray] = {}
    for res in (1, 3):
        out[f"res{res}_x"] = HEALPixNodes(res, "h").update_graph(HeteroData(), {})["h"].x.numpy()
    # LimitedAreaHEALPixNodes cannot be run: its constructor assigns ``area_mask_builder`` BEFORE calling the base
    # constructor, which resets it to None (from_healpix.py:84-87, nodes/builders/base.py:38), so ``register_nodes``
    # fails with AttributeError in the reference itself.
    np.savez_compressed(OUT / "healpix.npz", **out)
    print("healpix.npz", {k: v.shape for k, v in out.items()})


def make_area_weights() -> None:
    from anemoi.graphs.nodes.attributes import SphericalAreaWeights, UniformWeights
    from torch_geometric.data import HeteroData

    rng = np.random.default_rng(7)
    lat, lon = grids.octahedral_grid(24)
    sets = {
        "o24": grids.latlon_deg_to_x(lat, lon).numpy(),
        "random": np.stack([np.arcsin(rng.uniform(-1, 1, 3000)), rng.uniform(0, 2 * np.pi, 3000)], 1).astype(np.float32),
    }
    g = build({"hidden": tri_nodes(3)}, [])
    sets["tri3"] = g["hidden"].x.numpy()
    out = {}
    for name, x in sets.items():
        graph = HeteroData()
        graph["n"].x = torch.from_numpy(x)
        graph["n"].node_type = "LatLonNodes"
        out[f"{name}_x"] = x
        out[f"{name}_raw64"] = SphericalAreaWeights(norm=None, dtype="float64").compute(graph, "n").numpy()
        for norm in NORMS:
            out[f"{name}_{norm}"] = SphericalAreaWeights(norm=norm).compute(graph, "n").numpy()
        out[f"{name}_uniform"] = UniformWeights(norm="l1").compute(graph, "n").numpy()
    np.savez_compressed(OUT / "area_weights.npz", **out)
    print("area_weights.npz", {k: v.shape for k, v in out.items() if k.endswith("_raw64")})


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("attr_vectors", make_attr_vectors), ("tri", make_tri), ("toy", make_toy), ("o96", make_o96),
                     ("lam", make_lam), ("area_weights", make_area_weights), ("hex", make_hex), ("healpix", make_healpix)):  # fmt: skip
        if not only or name in only:
            fn()
