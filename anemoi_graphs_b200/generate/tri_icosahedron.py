"""Triangular refined-icosahedron nodes and multi-scale edges
(/root/reference/src/anemoi/graphs/generate/tri_icosahedron.py).

Vertices and faces of every refinement level come from the device subdivision kernels
(``ops.Icosphere``, which restates ``trimesh.creation.icosphere``); the multi-scale edges from the CSR
frontier-expansion kernels.  No networkx graph is built: where the reference threads an ``nx.DiGraph``
through the hidden node attribute ``_nx_graph``, this package keeps the device mesh there.
"""

from __future__ import annotations

import numpy as np
import torch

from .. import device as _device
from .. import ops


class DeviceMesh:
    """What the hidden ``_nx_graph`` attribute holds here: the icosphere of a node set on the device."""

    def __init__(self, icosphere: ops.Icosphere, order_dev: torch.Tensor) -> None:
        self.icosphere = icosphere
        self.order_dev = order_dev  # CUDA int64: icosphere vertex (or candidate row) at every graph position

    def number_of_nodes(self) -> int:
        return int(self.order_dev.shape[0])

    def __getstate__(self):  # never pickled with device memory (``clean`` removes it anyway)
        return {"icosphere": None, "order_dev": None}


def get_icosphere(resolution: int) -> ops.Icosphere:
    """Device icosphere of ``resolution`` (all levels 0..resolution)."""
    return ops.Icosphere(int(resolution))


def get_latlon_coords_icosphere(resolution: int) -> np.ndarray:
    """float32 (lat, lon) radians of the icosphere vertices (tri_icosahedron.py:108-123)."""
    return get_icosphere(resolution).latlon.cpu().numpy()


def _ordering_of(latlon_dev: torch.Tensor) -> torch.Tensor:
    """Node ordering of device coordinates (generate/utils.py:15-33) as a CUDA int64 tensor.

    The two argsorts are numpy's, on the host - they DEFINE the order (unstable sorts over tens of thousands of
    tied keys, ``generate.utils.get_coordinates_ordering``).  The gathers around them (``lat[index]`` between the
    sorts, ``index_latitude[index_longitude]`` after) run on the device, so the host only ever touches one
    contiguous float32 column per sort."""
    lat_dev = latlon_dev[:, 0].contiguous()
    lon = latlon_dev[:, 1].contiguous().cpu().numpy()
    index_latitude = torch.from_numpy(np.argsort(lon)).to(latlon_dev.device)
    lat_sorted = lat_dev[index_latitude].cpu().numpy()
    index_longitude = torch.from_numpy(np.argsort(lat_sorted)).to(latlon_dev.device)
    return index_latitude[index_longitude.flip(0)]


def _sort_columns_host(lat: np.ndarray, lon: np.ndarray, emit=None):
    """The host half of ``get_coordinates_ordering`` for a provisional node set (runs on the worker thread): the
    reference's two numpy argsorts on contiguous float32 columns and the gather between them.  Same calls on the
    same values as ``_ordering_of`` - the permutation is numpy's.  ``emit`` receives each index array the moment it
    is final (the first one is uploaded while the second sort runs)."""
    index_latitude = np.argsort(lon)
    if emit is not None:
        emit(index_latitude)
    index_longitude = np.argsort(lat[index_latitude])
    if emit is not None:
        emit(index_longitude)
    return index_latitude, index_longitude


def _combine_order(index_latitude: torch.Tensor, index_longitude: torch.Tensor) -> torch.Tensor:
    """``arange(n)[index_latitude][index_longitude[::-1]]`` on the device."""
    return index_latitude[index_longitude.flip(0)]


def create_tri_nodes_provisional(resolution: int):
    """Global mesh nodes whose order is computed on a host thread WHILE the edge kernels run
    (``device.Provisional``): returns ``(mesh, provisional)``; ``mesh.order_dev`` is set when the order resolves."""
    ico = get_icosphere(resolution)
    mesh = DeviceMesh(ico, None)
    prov = _device.Provisional(ico.latlon, _sort_columns_host, "latlon")  # combined by agx_order_resolve
    prov.on_resolved = lambda p: setattr(mesh, "order_dev", p.order_dev)
    return mesh, prov


def create_tri_nodes(resolution: int, area_mask_builder=None):
    """Global (or area-limited) mesh nodes from a refined icosahedron (tri_icosahedron.py:24-58).

    Returns ``(mesh, coords_rad, node_ordering)``: the device mesh, the float32 vertex coordinates (CUDA tensor,
    not ordered) and the order that sorts them by latitude and longitude (CUDA int64)."""
    ico = get_icosphere(resolution)
    node_ordering = _ordering_of(ico.latlon)

    if area_mask_builder is not None:
        area_mask = area_mask_builder.get_mask_device(ico.latlon)
        node_ordering = node_ordering[area_mask[node_ordering]]

    return DeviceMesh(ico, node_ordering), ico.latlon, node_ordering


def create_stretched_tri_nodes(base_resolution: int, lam_resolution: int, area_mask_builder=None):
    """Global mesh with two resolution levels (tri_icosahedron.py:61-105): ``base_resolution`` outside the
    area of interest, ``lam_resolution`` inside."""
    assert area_mask_builder is not None, "AOI mask builder must be provided to build refined grid."
    lam = get_icosphere(lam_resolution)
    # lower levels are prefixes of the finer one; a base level above the lam level (unusual) is generated apart
    if base_resolution <= lam_resolution:
        base_latlon = lam.latlon[: ops.ico_num_vertices(base_resolution)]
    else:
        base_latlon = get_icosphere(base_resolution).latlon
    base_area_mask = ~area_mask_builder.get_mask_device(base_latlon)
    lam_area_mask = area_mask_builder.get_mask_device(lam.latlon)

    coords_rad = torch.cat([base_latlon[base_area_mask], lam.latlon[lam_area_mask]])
    node_ordering = _ordering_of(coords_rad)
    return DeviceMesh(lam, node_ordering), coords_rad, node_ordering


def multiscale_edges(
    nodes, resolutions, x_hops: int = 1, area_mask_builder=None, allow_provisional: bool = False
) -> torch.Tensor:
    """Multi-scale connections of a tri-node set: CUDA int32 (2, E) sorted by (target, source).

    Replaces ``add_edges_to_nx_graph`` + ``nx.to_scipy_sparse_array`` (tri_icosahedron.py:138-224,
    edges/builder.py:412-455): for every level ``r`` in ``resolutions`` the directed pairs (u -> v), u != v,
    within ``x_hops`` hops on the level-r mesh restricted to valid vertices; level vertices are identified with
    graph nodes by nearest neighbour (tri_icosahedron.py:185)."""
    assert x_hops > 0, "x_hops == 0, graph would have no edges ..."
    resolutions = [int(r) for r in resolutions]
    mesh = nodes.get("_nx_graph", None)
    ico = mesh.icosphere if isinstance(mesh, DeviceMesh) and mesh.icosphere is not None else None
    if ico is None or ico.max_level < max(resolutions):
        ico = get_icosphere(max(resolutions))
    node_type = nodes["node_type"]
    st = _device.node_state(nodes, provisional_ok=(allow_provisional and node_type == "TriNodes" and area_mask_builder is None))
    n_nodes = int(st.x.shape[0])
    if node_type == "TriNodes" and area_mask_builder is None and n_nodes == ops.ico_num_vertices(ico.max_level):
        if st.prov is not None:
            # the final order is not known yet: edges in the icosphere's own numbering, relabelled when it is
            identity = torch.arange(n_nodes, dtype=torch.int32, device=st.x.device)
            edges = ops.multiscale_tri_edges(ico, resolutions, x_hops, identity)
            return _device.tag_rows(edges, st.prov, st.prov)
        mesh = nodes.get("_nx_graph", None)
        if isinstance(mesh, DeviceMesh) and mesh.order_dev is not None:
            order = mesh.order_dev
        else:
            order = torch.as_tensor(np.asarray(nodes["_node_ordering"]))
        return ops.multiscale_tri_edges(ico, resolutions, x_hops, order)
    if st.prov is not None:  # not reachable for the node types that defer their order; never mix numberings
        st = _device.node_state(nodes)
    # limited-area / stretched: valid vertices by the area mask, vertex -> node by 1-NN
    nv = ops.ico_num_vertices(max(resolutions))
    coords = ico.latlon[:nv]
    with ops.NeighbourIndex(st.x, hint_k=1) as index:
        nearest = index.knn(coords, 1, tag="knn_vertex_map")[0]
    if area_mask_builder is not None:
        valid = area_mask_builder.get_mask_device(coords)
        vertex_map = torch.where(valid, nearest, torch.full_like(nearest, -1))
    else:
        vertex_map = nearest
    if ico.max_level > max(resolutions):
        pad = torch.full((ops.ico_num_vertices(ico.max_level) - nv,), -1, dtype=torch.int32, device=vertex_map.device)
        vertex_map = torch.cat([vertex_map, pad])
    return ops.multiscale_tri_edges_mapped(ico, resolutions, x_hops, n_nodes, vertex_map)
