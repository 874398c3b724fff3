"""ORACLE - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``anemoi_graphs_b200``) never does.

CPU restatement of the reference's edge-construction hot path.  The reference is pure Python
glue over scikit-learn / scipy / networkx / numpy; those libraries are present in this image
(scikit-learn 1.9.0, scipy 1.18.1, networkx 3.6.1, numpy 2.3.5 - unpinned in
/root/reference/pyproject.toml:42-57), so every function below issues the SAME third-party
calls, in the same order and with the same arguments, as the reference lines it cites.
trimesh is absent: its algorithm is restated in ``oracle/trimesh_icosphere.py``.

Parity status: PINNED against the unmodified reference executed in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``)
and against the reference's published counts (docs/_static/hetero_data_graph.txt:13,19,25;
docs/graphs/edges/tri_refined_edges.csv; tests/nodes/test_tri_nodes.py:32).
HexNodes: the h3 library is absent; ``oracle/h3_restated.py`` restates its geometry (parity with h3 itself unpinned).

Citations are relative to /root/reference/src/anemoi/graphs/.
"""

from __future__ import annotations

import math

import networkx as nx
import numpy as np
import scipy.sparse as sp
from scipy.spatial.transform import Rotation
from sklearn.neighbors import BallTree, NearestNeighbors

try:  # loaded by file path from bench.py / tests as well as ``import oracle.ref_path``
    from . import trimesh_icosphere as _tm
except ImportError:  # pragma: no cover
    import importlib.util
    import pathlib

    _spec = importlib.util.spec_from_file_location(
        "_oracle_trimesh_icosphere", pathlib.Path(__file__).with_name("trimesh_icosphere.py")
    )
    _tm = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_tm)

EARTH_RADIUS = 6371.0  # __init__.py:10
TIE_TAU = 2.0**-40  # relative width of an "ulp-level" tie in float64 rdist (parity rules)
REFERENCE_N_JOBS = 4  # hard-coded in edges/builder.py:259,364, utils.py:37, generate/masks.py:51


# ----------------------------------------------------------------------------------------
# canonical forms
# ----------------------------------------------------------------------------------------
def canonical_sort(edge_index: np.ndarray) -> np.ndarray:
    """Sort columns of a (2, E) edge_index by (dst, src): row 1 major, row 0 minor."""
    edge_index = np.asarray(edge_index)
    order = np.lexsort((edge_index[0], edge_index[1]))
    return edge_index[:, order]


def rdist64(lat1, lon1, lat2, lon2) -> np.ndarray:
    """sklearn ``HaversineDistance64.rdist`` (sklearn/metrics/_dist_metrics.pyx.tp:2639-2648)
    in numpy float64; x1 = query, x2 = tree point."""
    lat1 = np.asarray(lat1, dtype=np.float64)
    lon1 = np.asarray(lon1, dtype=np.float64)
    lat2 = np.asarray(lat2, dtype=np.float64)
    lon2 = np.asarray(lon2, dtype=np.float64)
    sin_0 = np.sin(0.5 * (lat1 - lat2))
    sin_1 = np.sin(0.5 * (lon1 - lon2))
    return sin_0 * sin_0 + np.cos(lat1) * np.cos(lat2) * sin_1 * sin_1


# ----------------------------------------------------------------------------------------
# KNN edges                                                    edges/builder.py:235-270, 69-87
# ----------------------------------------------------------------------------------------
def knn_edges(source_x: np.ndarray, target_x: np.ndarray, k: int, n_jobs: int = REFERENCE_N_JOBS) -> np.ndarray:
    """The reference's KNNEdges.get_edge_index on unmasked nodes: (2, E) int32, row0 = source."""
    nn = NearestNeighbors(metric="haversine", n_jobs=n_jobs)
    nn.fit(source_x)
    adj = nn.kneighbors_graph(target_x, n_neighbors=k, mode="distance").tocoo()
    return np.stack([adj.col, adj.row], axis=0).astype(np.int32)


def knn_edges_canonical(
    source_x: np.ndarray, target_x: np.ndarray, k: int, extra: int = 8, tau: float = TIE_TAU, n_jobs: int = -1
):
    """KNN edge set under the north-star tie rule.

    Decisions are sklearn's float64 haversine ``rdist`` on the float32 inputs; candidates whose
    ``rdist`` lies within ``tau`` (relative) of the k-th smallest form a tie group, from which the
    LOWEST SOURCE INDICES are kept.  Returns ``(edge_index sorted by (dst, src), info)`` where
    ``info["tied_queries"]`` lists the queries whose k-th boundary is such a tie and
    ``info["differs_from_reference"]`` those where sklearn's own traversal-order choice
    (utils/_heap.pyx:46 keeps the first visited) is a different set.
    """
    source_x = np.asarray(source_x, dtype=np.float32)
    target_x = np.asarray(target_x, dtype=np.float32)
    n_src = source_x.shape[0]
    kk = min(n_src, k + extra)
    nn = NearestNeighbors(metric="haversine", n_jobs=n_jobs)
    nn.fit(source_x)
    _, ind = nn.kneighbors(target_x, n_neighbors=kk, return_distance=True)
    lat_q = target_x[:, 0:1].astype(np.float64)
    lon_q = target_x[:, 1:2].astype(np.float64)
    cand = source_x[ind]  # (nq, kk, 2)
    rd = rdist64(lat_q, lon_q, cand[..., 0], cand[..., 1])  # (nq, kk)
    # order candidates by recomputed rdist (order inside a tie group is irrelevant: the group is
    # re-sorted by source index below)
    order = np.argsort(rd, axis=1, kind="stable")
    rd_s = np.take_along_axis(rd, order, axis=1)
    ind_s = np.take_along_axis(ind, order, axis=1)
    chosen = ind_s[:, :k].copy()
    tied_queries: list[int] = []
    overflow: list[int] = []
    if kk > k:
        r_k = rd_s[:, k - 1]
        in_group = np.abs(rd_s - r_k[:, None]) <= tau * r_k[:, None]
        has_tie = in_group[:, k]  # the (k+1)-th candidate is inside the group
        for q in np.nonzero(has_tie)[0]:
            g = in_group[q]
            if g[-1] and kk < n_src:
                overflow.append(int(q))  # tie group may extend past the candidates fetched
            below = ind_s[q][(rd_s[q] < r_k[q]) & ~g]
            group = np.sort(ind_s[q][g])
            chosen[q] = np.concatenate([below, group[: k - below.size]])
            tied_queries.append(int(q))
    if overflow:
        raise RuntimeError(f"tie group wider than extra={extra} for queries {overflow[:5]}...; raise `extra`")
    chosen.sort(axis=1)
    nq = target_x.shape[0]
    dst = np.repeat(np.arange(nq, dtype=np.int64), k)
    edge_index = np.stack([chosen.reshape(-1), dst], axis=0).astype(np.int32)
    # which of the tied queries did sklearn resolve differently?  (its own k-sized heap decides,
    # so ask it for exactly k)
    ref_ind = nn.kneighbors(target_x, n_neighbors=k, return_distance=False)
    ref_sets = np.sort(ref_ind, axis=1)
    differs = [q for q in tied_queries if not np.array_equal(ref_sets[q], chosen[q])]
    untied_mismatch = np.nonzero((ref_sets != chosen).any(axis=1))[0]
    untied_mismatch = sorted(set(untied_mismatch.tolist()) - set(tied_queries))
    info = {
        "tied_queries": np.asarray(tied_queries, dtype=np.int64),
        "differs_from_reference": np.asarray(differs, dtype=np.int64),
        "untied_mismatch": np.asarray(untied_mismatch, dtype=np.int64),  # must be empty
    }
    return edge_index, info


# ----------------------------------------------------------------------------------------
# Cut-off edges                                   edges/builder.py:312-371, utils.py:17-63
# ----------------------------------------------------------------------------------------
def grid_reference_distance(x: np.ndarray, n_jobs: int = REFERENCE_N_JOBS) -> float:
    """utils.get_grid_reference_distance: max positive distance in a k=2 self query (float64)."""
    nn = NearestNeighbors(metric="haversine", n_jobs=n_jobs)
    nn.fit(x)
    dists, _ = nn.kneighbors(x, n_neighbors=2, return_distance=True)
    return float(dists[dists > 0].max())


def cutoff_radius(target_x: np.ndarray, cutoff_factor: float, n_jobs: int = REFERENCE_N_JOBS) -> float:
    """CutOffEdges.get_cutoff_radius (edges/builder.py:312-334): reference distance of the TARGET nodes."""
    return grid_reference_distance(target_x, n_jobs) * cutoff_factor


def cutoff_edges(
    source_x: np.ndarray, target_x: np.ndarray, cutoff_factor: float, radius: float | None = None,
    n_jobs: int = REFERENCE_N_JOBS,
) -> np.ndarray:  # fmt: skip
    """CutOffEdges.get_edge_index on unmasked nodes (edges/builder.py:341-371, 69-87)."""
    if radius is None:
        radius = cutoff_radius(target_x, cutoff_factor, n_jobs)
    nn = NearestNeighbors(metric="haversine", n_jobs=n_jobs)
    nn.fit(source_x)
    adj = nn.radius_neighbors_graph(target_x, radius=radius).tocoo()
    return np.stack([adj.col, adj.row], axis=0).astype(np.int32)


def cutoff_boundary_pairs(source_x, target_x, edge_index, radius: float, tau: float = TIE_TAU) -> np.ndarray:
    """Edges of ``edge_index`` whose float64 rdist is within ``tau`` (relative) of sin^2(r/2)."""
    thr = math.sin(0.5 * radius) ** 2
    s, t = edge_index[0], edge_index[1]
    rd = rdist64(target_x[t, 0], target_x[t, 1], source_x[s, 0], source_x[s, 1])
    return np.nonzero(np.abs(rd - thr) <= tau * thr)[0]


# ----------------------------------------------------------------------------------------
# node masking                                                edges/builder.py:162-193
# ----------------------------------------------------------------------------------------
def masked_edges(kind: str, source_x, target_x, source_mask, target_mask, param, n_jobs: int = REFERENCE_N_JOBS):
    """KNN / cut-off on row-selected coordinates, compact indices mapped back (undo_masking).

    For cut-off the radius still comes from ALL target nodes: ``get_cutoff_radius`` is called
    without a mask (edges/builder.py:338) and ``get_nearest_neighbour`` ignores it anyway
    (utils.py:32-39).
    """
    src_sel = np.arange(source_x.shape[0]) if source_mask is None else np.where(np.asarray(source_mask).squeeze())[0]
    dst_sel = np.arange(target_x.shape[0]) if target_mask is None else np.where(np.asarray(target_mask).squeeze())[0]
    if kind == "knn":
        ei = knn_edges(source_x[src_sel], target_x[dst_sel], int(param), n_jobs)
    elif kind == "cutoff":
        radius = cutoff_radius(target_x, param, n_jobs)
        ei = cutoff_edges(source_x[src_sel], target_x[dst_sel], param, radius=radius, n_jobs=n_jobs)
    else:
        raise ValueError(kind)
    return np.stack([src_sel[ei[0]], dst_sel[ei[1]]], axis=0).astype(np.int32)


# ----------------------------------------------------------------------------------------
# area mask                                                     generate/masks.py:23-99
# ----------------------------------------------------------------------------------------
def knn_area_mask(reference_x: np.ndarray, coords_rad: np.ndarray, margin_radius_km: float) -> np.ndarray:
    nn = NearestNeighbors(metric="haversine", n_jobs=REFERENCE_N_JOBS)
    nn.fit(reference_x)
    d, _ = nn.kneighbors(coords_rad, n_neighbors=1)
    return d[:, 0] * EARTH_RADIUS <= margin_radius_km


# ----------------------------------------------------------------------------------------
# TriNodes                generate/tri_icosahedron.py:24-58,108-123; generate/utils.py:15-33;
#                         generate/transforms.py:34-52; nodes/builders/from_refined_icosahedron.py:30-69
# ----------------------------------------------------------------------------------------
def cartesian_to_latlon_rad(xyz: np.ndarray) -> np.ndarray:
    lat = np.arcsin(xyz[..., 2] / (xyz**2).sum(axis=1))
    lon = np.arctan2(xyz[..., 1], xyz[..., 0])
    return np.array((lat, lon), dtype=np.float32).transpose()


def coordinates_ordering(coords: np.ndarray) -> np.ndarray:
    """generate/utils.py:30-33 - NOTE both argsorts are numpy's default (unstable) kind."""
    index_latitude = np.argsort(coords[:, 1])
    index_longitude = np.argsort(coords[index_latitude][:, 0])[::-1]
    return np.arange(coords.shape[0])[index_latitude][index_longitude]


def tri_nodes(resolution: int) -> tuple[np.ndarray, np.ndarray]:
    """``(x, node_ordering)``: float32 (N, 2) coordinates in graph order, and the vertex id at each position."""
    verts, _ = _tm.icosphere(resolution)
    coords = cartesian_to_latlon_rad(verts)
    order = coordinates_ordering(coords)
    return coords[order], order


def lam_tri_nodes(resolution: int, reference_x: np.ndarray, margin_radius_km: float):
    """LimitedAreaTriNodes (nodes/builders/from_refined_icosahedron.py:72-136, generate/tri_icosahedron.py:24-58):
    ``(x, node_ordering, all vertex coords)`` keeping the vertices within the margin of the reference nodes."""
    verts, _ = _tm.icosphere(resolution)
    coords = cartesian_to_latlon_rad(verts)
    order = coordinates_ordering(coords)
    mask = knn_area_mask(reference_x, coords, margin_radius_km)
    order = order[mask[order]]
    return coords[order], order, coords


def stretched_tri_nodes(base_resolution: int, lam_resolution: int, reference_x: np.ndarray, margin_radius_km: float):
    """StretchedTriNodes (generate/tri_icosahedron.py:61-105): base-level vertices outside the area of
    interest + lam-level vertices inside it."""
    base = cartesian_to_latlon_rad(_tm.icosphere(base_resolution)[0])
    lam = cartesian_to_latlon_rad(_tm.icosphere(lam_resolution)[0])
    base_mask = ~knn_area_mask(reference_x, base, margin_radius_km)
    lam_mask = knn_area_mask(reference_x, lam, margin_radius_km)
    coords = np.concatenate([base[base_mask], lam[lam_mask]])
    order = coordinates_ordering(coords)
    return coords[order], order, coords


def multiscale_edges_tri_masked(resolutions, x_hops: int, x: np.ndarray, mask_reference_x: np.ndarray, margin_km: float):
    """``add_edges_to_nx_graph`` with an area mask (generate/tri_icosahedron.py:138-224,274-310) for limited-area
    (mask reference = the cut-out data nodes, margin = the builder's) and stretched (mask reference = the graph
    nodes themselves, margin 1 km, edges/builder.py:422-432) tri nodes; graph indices are positions in ``x``.
    networkx ego graphs on the valid sub-mesh of every level, vertices mapped to nodes by haversine 1-NN."""
    tree = BallTree(x, metric="haversine")
    pairs = set()
    for resolution in resolutions:
        verts, faces = _tm.icosphere(resolution)
        r_vertices_rad = cartesian_to_latlon_rad(verts)
        valid = np.where(knn_area_mask(mask_reference_x, r_vertices_rad, margin_km))[0]
        edges = _tm.edges_unique(faces)
        edges = edges[np.isin(edges, valid).all(axis=1)]
        g = nx.from_edgelist(edges)
        _, vmap = tree.query(r_vertices_rad, k=1)
        for i in valid:
            if i not in g:
                continue
            for nb in nx.ego_graph(g, i, radius=x_hops, center=False):
                if nb != i:
                    pairs.add((int(vmap[i][0]), int(vmap[nb][0])))  # (target, source)
    ei = np.array(sorted(pairs), dtype=np.int64).reshape(-1, 2)
    return np.stack([ei[:, 1], ei[:, 0]], axis=0).astype(np.int32)


# ----------------------------------------------------------------------------------------
# MultiScaleEdges (tri)     edges/builder.py:412-455; generate/tri_icosahedron.py:138-310
# ----------------------------------------------------------------------------------------
def multiscale_edges_tri_networkx(resolutions, x_hops: int, node_ordering: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Faithful (slow) restatement: networkx ego graphs, BallTree vertex mapping, DiGraph, COO."""
    node_ordering = np.asarray(node_ordering)
    graph = nx.DiGraph()
    for pos, node_id in enumerate(node_ordering):
        graph.add_node(int(node_id), hcoords_rad=x[pos])
    graph_nodes_idx = list(sorted(graph.nodes))
    graph_vertices = np.array([graph.nodes[i]["hcoords_rad"] for i in graph_nodes_idx])
    tree = BallTree(graph_vertices, metric="haversine")
    for resolution in resolutions:
        verts, faces = _tm.icosphere(resolution)
        r_vertices_rad = cartesian_to_latlon_rad(verts)
        edges = _tm.edges_unique(faces)
        g = nx.from_edgelist(edges)
        neighbours = {
            i: set(nx.ego_graph(g, i, radius=x_hops, center=False) if i in g else []) for i in range(len(verts))
        }
        _, vmap = tree.query(r_vertices_rad, k=1)
        pairs = [
            (graph_nodes_idx[vmap[nb][0]], graph_nodes_idx[vmap[node][0]])
            for node, nbs in neighbours.items()
            for nb in nbs
            if node != nb
        ]
        graph.add_edges_from(pairs)
    adj = nx.to_scipy_sparse_array(graph, format="coo")
    return np.stack([adj.col, adj.row], axis=0).astype(np.int32)


def multiscale_edges_tri(resolutions, x_hops: int, node_ordering: np.ndarray) -> np.ndarray:
    """Fast restatement for GLOBAL TriNodes, canonical (dst, src) order.

    For global TriNodes level-r vertices are bit-identical prefixes of the finest level
    (trimesh appends midpoints after the old vertices), so the BallTree 1-NN mapping
    (generate/tri_icosahedron.py:185) is the identity on vertex ids and the graph index of a
    vertex is its position in ``node_ordering`` (edges/builder.py:453).  x_hops reachability is a
    boolean sparse matrix power.  Checked against ``multiscale_edges_tri_networkx`` in tests.
    """
    node_ordering = np.asarray(node_ordering)
    n = node_ordering.shape[0]
    rank = np.empty(n, dtype=np.int64)
    rank[node_ordering] = np.arange(n)
    rows, cols = [], []
    for resolution in resolutions:
        _, faces = _tm.icosphere(resolution)
        e = _tm.edges_unique(faces)
        nv = int(e.max()) + 1
        a = sp.coo_matrix((np.ones(2 * len(e), dtype=np.int8), (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])),
                          shape=(nv, nv)).tocsr()  # fmt: skip
        reach = a.copy()
        frontier = a.copy()
        for _ in range(x_hops - 1):
            frontier = (frontier @ a).astype(bool).astype(np.int8)
            reach = (reach + frontier).astype(bool).astype(np.int8)
        reach = reach.tocoo()
        keep = reach.row != reach.col
        # edge (neighbour -> node): source = neighbour, target = node
        rows.append(rank[reach.row[keep]])  # target
        cols.append(rank[reach.col[keep]])  # source
    ei = np.unique(np.stack([np.concatenate(rows), np.concatenate(cols)], axis=1), axis=0)  # sorted by (dst, src)
    return np.stack([ei[:, 1], ei[:, 0]], axis=0).astype(np.int32)


# ----------------------------------------------------------------------------------------
# attributes         edges/attributes.py:24-157, edges/directional.py:19-94,
#                    generate/transforms.py:91-140, utils.py:84-103, normalise.py:20-55
# ----------------------------------------------------------------------------------------
def haversine_distance(source_coords: np.ndarray, target_coords: np.ndarray) -> np.ndarray:
    dlat = target_coords[:, 0] - source_coords[:, 0]
    dlon = target_coords[:, 1] - source_coords[:, 1]
    a = np.sin(dlat / 2) ** 2 + np.cos(source_coords[:, 0]) * np.cos(target_coords[:, 0]) * np.sin(dlon / 2) ** 2
    return 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))


def latlon_rad_to_cartesian(loc, radius: float = 1.0) -> np.ndarray:
    latr, lonr = loc[0], loc[1]
    x = radius * np.cos(latr) * np.cos(lonr)
    y = radius * np.cos(latr) * np.sin(lonr)
    z = radius * np.sin(latr)
    return np.array((x, y, z)).T


def direction_vec(points: np.ndarray, reference: np.ndarray, epsilon: float = 10e-11) -> np.ndarray:
    v = np.cross(points, reference)
    vnorm1 = np.power(v, 2).sum(axis=-1)
    redo_idx = np.where(vnorm1 < epsilon)[0]
    if len(redo_idx) > 0:
        points[redo_idx] += epsilon
        v = np.cross(points, reference)
        vnorm1 = np.power(v, 2).sum(axis=-1)
    return v.T / np.sqrt(vnorm1)


def edge_directions_raw(source_coords_t: np.ndarray, target_coords_t: np.ndarray, rotated: bool = True) -> np.ndarray:
    """``directional_edge_features``: inputs (2, E) float32 [lat; lon], output (2, E)."""
    if not rotated:
        return target_coords_t - source_coords_t
    pole_vec = np.array([0, 0, 1])
    loc1_xyz = latlon_rad_to_cartesian(source_coords_t, 1.0)
    loc2_xyz = latlon_rad_to_cartesian(target_coords_t, 1.0)
    v_unit = direction_vec(loc2_xyz, pole_vec)
    theta = np.arccos(np.dot(loc2_xyz, pole_vec))
    r = Rotation.from_rotvec(np.transpose(v_unit * theta))
    direction = direction_vec(r.apply(loc1_xyz), pole_vec)
    direction = direction / np.sqrt(np.power(direction, 2).sum(axis=0))
    assert np.allclose(direction[2], 0)
    return direction[:2]


def normalise(values: np.ndarray, norm: str | None) -> np.ndarray:
    if norm is None:
        return values
    if norm == "l1":
        return values / np.sum(values)
    if norm == "l2":
        return values / np.linalg.norm(values)
    if norm == "unit-max":
        return values / np.amax(values)
    if norm == "unit-range":
        return (values - np.amin(values)) / (np.amax(values) - np.amin(values))
    if norm == "unit-std":
        std = np.std(values)
        return values if std == 0 else values / std
    raise ValueError(f'Attribute normalisation "{norm}" is not valid.')


def edge_length(source_x, target_x, edge_index, norm: str | None = None, invert: bool = False) -> np.ndarray:
    """EdgeLength.compute -> float32 (E, 1)."""
    v = haversine_distance(source_x[edge_index[0]], target_x[edge_index[1]])[:, np.newaxis]
    out = normalise(v, norm).astype(np.float32)
    return 1 - out if invert else out


def edge_direction(source_x, target_x, edge_index, norm: str | None = None, rotated: bool = True) -> np.ndarray:
    """EdgeDirection.compute -> float32 (E, 2)."""
    v = edge_directions_raw(source_x[edge_index[0]].T, target_x[edge_index[1]].T, rotated).T
    return normalise(v, norm).astype(np.float32)


def concat_edges(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """utils.concat_edges: ``torch.unique(cat, dim=1)`` = columns sorted lexicographically, de-duplicated."""
    return np.unique(np.concatenate([e1, e2], axis=1), axis=1)


# ----------------------------------------------------------------------------------------
# SphericalAreaWeights                                  nodes/attributes.py:165-221, 44-52
# ----------------------------------------------------------------------------------------
def spherical_area_weights(x: np.ndarray, norm: str | None = None, dtype: str = "float32") -> np.ndarray:
    """The reference's ``SphericalAreaWeights(norm, dtype=dtype).compute``: (N, 1)."""
    from scipy.spatial import SphericalVoronoi

    latitudes, longitudes = np.asarray(x[:, 0]), np.asarray(x[:, 1])
    points = latlon_rad_to_cartesian((latitudes, longitudes))  # float32 in, float32 out
    sv = SphericalVoronoi(points, 1.0, np.array([0, 0, 0]))
    mask = np.array([bool(i) for i in sv.regions])
    sv.regions = [region for region in sv.regions if region]
    area_weights = sv.calculate_areas()
    result = np.ones(points.shape[0]) * 0.0
    result[mask] = area_weights
    return normalise(result[:, np.newaxis], norm).astype(dtype)
