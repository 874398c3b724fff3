"""Device-level operations: thin, typed wrappers over the C ABI working on CUDA tensors.

These are the calls the builders (``edges/builder.py``, ``edges/attributes.py``, ``nodes/builders.py``)
make; PyTorch is used only to own device memory and streams.  Everything here requires a CUDA
device - there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
import math
from ctypes import byref, c_double, c_int, c_int32, c_int64, c_void_p

import torch

from . import _cabi
from ._cabi import NORM_CODES, check, current_stream, load_library, ptr


def _dev_x(x: torch.Tensor) -> torch.Tensor:
    """float32 (n, 2) contiguous CUDA coordinates."""
    _cabi.require_cuda()
    if not x.is_cuda:
        x = x.cuda(non_blocking=True)
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    assert x.dim() == 2 and x.shape[1] == 2, f"coordinates must have shape (N, 2), got {tuple(x.shape)}"
    return x.contiguous()


class NeighbourIndex:
    """Cell-binned reference point set - the stand-in for a fitted ``NearestNeighbors(metric="haversine")``
    (/root/reference/src/anemoi/graphs/edges/builder.py:259-260, 364-365)."""

    def __init__(self, x: torch.Tensor, cells_per_face: int = 0, hint_k: int = 0, hint_radius: float = 0.0) -> None:
        self.lib = load_library()
        self.x = _dev_x(x)  # kept alive: the index reads it during float64 refinement
        self.n = int(self.x.shape[0])
        handle = c_void_p()
        check(
            self.lib.agx_index_build(
                ptr(self.x), self.n, int(cells_per_face), int(hint_k), float(hint_radius), current_stream(), byref(handle)
            )
        )
        self.handle = handle

    @property
    def cells_per_face(self) -> int:
        n, c = c_int64(), c_int()
        check(self.lib.agx_index_info(self.handle, byref(n), byref(c)))
        return c.value

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.agx_index_free(self.handle, current_stream())
            self.handle = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "NeighbourIndex":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ------------------------------------------------------------------------------------------
    def knn(
        self,
        q: torch.Tensor,
        k: int,
        dst_base: int = 0,
        return_rdist: bool = False,
        stats: torch.Tensor | None = None,
        out: torch.Tensor | None = None,
    ):
        """``edge_index`` (2, nq*k) int32 - row 0 the k nearest reference points of each query, row 1
        ``dst_base + query`` - and optionally the float64 ``rdist`` (nq, k)."""
        q = _dev_x(q)
        nq = int(q.shape[0])
        if out is None:
            out = torch.empty((2, nq * k), dtype=torch.int32, device=q.device)
        assert out.shape == (2, nq * k) and out.dtype == torch.int32 and out.is_contiguous()
        rdist = torch.empty((nq, k), dtype=torch.float64, device=q.device) if return_rdist else None
        check(
            self.lib.agx_knn(
                self.handle, ptr(q), nq, int(k), out[0].data_ptr(), out[1].data_ptr(), int(dst_base), ptr(rdist),
                ptr(stats), current_stream(),
            )
        )  # fmt: skip
        return (out, rdist) if return_rdist else out

    def radius(self, q: torch.Tensor, radius: float, dst_base: int = 0, stats: torch.Tensor | None = None) -> torch.Tensor:
        """``edge_index`` (2, E) int32 of every (reference, query) pair within ``radius`` (inclusive)."""
        q = _dev_x(q)
        nq = int(q.shape[0])
        counts = torch.empty(nq, dtype=torch.int32, device=q.device)
        offsets = torch.empty(nq + 1, dtype=torch.int64, device=q.device)
        stream = current_stream()
        check(self.lib.agx_radius_count(self.handle, ptr(q), nq, float(radius), ptr(counts), stream))
        total = c_int64()
        check(self.lib.agx_exclusive_scan(ptr(counts), nq, ptr(offsets), byref(total), stream))
        out = torch.empty((2, total.value), dtype=torch.int32, device=q.device)
        if total.value:
            check(
                self.lib.agx_radius_fill(
                    self.handle, ptr(q), nq, float(radius), ptr(offsets), out[0].data_ptr(), out[1].data_ptr(),
                    int(dst_base), ptr(stats), stream,
                )
            )  # fmt: skip
        return out


def new_stats(device) -> torch.Tensor:
    return torch.zeros(4, dtype=torch.int64, device=device)


def haversine_rdist_host(lat1: float, lon1: float, lat2: float, lon2: float) -> float:
    """sklearn ``HaversineDistance64.rdist`` with the C library's sin/cos - the same libm calls sklearn's
    compiled code makes, so the value is bit-identical to the reference's."""
    s0 = math.sin(0.5 * (lat1 - lat2))
    s1 = math.sin(0.5 * (lon1 - lon2))
    return s0 * s0 + math.cos(lat1) * math.cos(lat2) * s1 * s1


def grid_reference_distance(x: torch.Tensor) -> float:
    """``utils.get_grid_reference_distance`` (/root/reference/src/anemoi/graphs/utils.py:44-63): the largest
    strictly positive distance in a k=2 self query.  The search, the float64 distances and the max run on
    the GPU; the ONE winning pair is then re-evaluated with libm so the returned float64 carries the same
    bits sklearn produces (it becomes the cut-off radius)."""
    xd = _dev_x(x)
    lib = load_library()
    with NeighbourIndex(xd, hint_k=2) as index:
        ei, rdist = index.knn(xd, 2, return_rdist=True)
    value, flat = c_double(), c_int64()
    check(lib.agx_max_positive(ptr(rdist), rdist.numel(), byref(value), byref(flat), current_stream()))
    if flat.value < 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    pair = torch.stack([ei[0, flat.value], ei[1, flat.value]]).cpu()
    a = xd[pair[1].item()].cpu().tolist()  # query
    b = xd[pair[0].item()].cpu().tolist()  # its neighbour
    rd = haversine_rdist_host(a[0], a[1], b[0], b[1])
    return 2.0 * math.asin(math.sqrt(rd))


class NodeTables:
    """Per-node tables for the attribute kernel: float32 (x, y, z, cos lat) and, for target nodes, the
    float64 rotation quaternion."""

    def __init__(self, x: torch.Tensor, with_rotation: bool) -> None:
        self.x = _dev_x(x)
        n = int(self.x.shape[0])
        self.xyzc = torch.empty((n, 4), dtype=torch.float32, device=self.x.device)
        self.quat = torch.empty((n, 4), dtype=torch.float64, device=self.x.device) if with_rotation else None
        check(load_library().agx_node_tables(ptr(self.x), n, ptr(self.xyzc), ptr(self.quat), current_stream()))


_workspace: dict = {}


def _attr_workspace(device) -> torch.Tensor:
    key = (device.type, device.index)
    if key not in _workspace:
        n = int(load_library().agx_edge_attrs_workspace())
        _workspace[key] = torch.empty(n, dtype=torch.float64, device=device)
    return _workspace[key]


def edge_attributes(
    edge_index: torch.Tensor,
    src: NodeTables,
    dst: NodeTables,
    length: bool = True,
    length_norm: str | None = None,
    length_invert: bool = False,
    direction: bool = True,
    direction_norm: str | None = None,
    direction_rotated: bool = True,
):
    """Fused EdgeLength / EdgeDirection: returns ``(len (E, 1) float32 | None, dir (E, 2) float32 | None)``."""
    for norm in (length_norm, direction_norm):
        if norm not in NORM_CODES:
            raise ValueError(
                f"Attribute normalisation \"{norm}\" is not valid. Options are: 'l1', 'l2', 'unit-max' or 'unit-std'."
            )
    assert edge_index.is_cuda and edge_index.dtype == torch.int32 and edge_index.dim() == 2 and edge_index.shape[0] == 2
    edge_index = edge_index.contiguous()
    n_edges = int(edge_index.shape[1])
    dev = edge_index.device
    out_len = torch.empty((n_edges, 1), dtype=torch.float32, device=dev) if length else None
    out_dir = torch.empty((n_edges, 2), dtype=torch.float32, device=dev) if direction else None
    if direction and direction_rotated and dst.quat is None:
        raise ValueError("rotated directions need target NodeTables built with with_rotation=True")
    check(
        load_library().agx_edge_attrs(
            edge_index[0].data_ptr(), edge_index[1].data_ptr(), n_edges, ptr(src.x), ptr(src.xyzc), ptr(dst.x),
            ptr(dst.xyzc), ptr(dst.quat), NORM_CODES[length_norm] if length else -1, int(bool(length_invert)),
            ptr(out_len), NORM_CODES[direction_norm] if direction else -1, int(bool(direction_rotated)), ptr(out_dir),
            ptr(_attr_workspace(dev)), current_stream(),
        )
    )  # fmt: skip
    return out_len, out_dir


# ----------------------------------------------------------------------------------------------
# icosphere + multi-scale edges
# ----------------------------------------------------------------------------------------------
def ico_num_vertices(level: int) -> int:
    return 10 * 4**level + 2


def ico_num_faces(level: int) -> int:
    return 20 * 4**level


def ico_face_offset(level: int) -> int:
    return 20 * ((4**level - 1) // 3)


class Icosphere:
    """All levels 0..max_level of ``trimesh.creation.icosphere`` on the device."""

    def __init__(self, max_level: int, device=None) -> None:
        _cabi.require_cuda()
        device = torch.device("cuda") if device is None else device
        self.max_level = int(max_level)
        nv = ico_num_vertices(self.max_level)
        self.vertices = torch.empty((nv, 3), dtype=torch.float64, device=device)
        self.faces_all = torch.empty((ico_face_offset(self.max_level + 1), 3), dtype=torch.int32, device=device)
        self.latlon = torch.empty((nv, 2), dtype=torch.float32, device=device)
        check(
            load_library().agx_icosphere(
                self.max_level, ptr(self.vertices), ptr(self.faces_all), ptr(self.latlon), current_stream()
            )
        )

    def faces(self, level: int) -> torch.Tensor:
        o = ico_face_offset(level)
        return self.faces_all[o : o + ico_num_faces(level)]


def multiscale_tri_edges(ico: Icosphere, levels: list[int], x_hops: int, node_ordering: torch.Tensor) -> torch.Tensor:
    """MultiScaleEdges for global TriNodes: (2, E) int32 sorted by (dst, src).

    ``node_ordering[p]`` is the icosphere vertex at graph position ``p``."""
    lib = load_library()
    dev = ico.vertices.device
    nv = ico_num_vertices(ico.max_level)
    order = node_ordering.to(device=dev, dtype=torch.int32).contiguous()
    assert order.shape == (nv,), f"node_ordering must list all {nv} vertices of level {ico.max_level}"
    rank = torch.empty(nv, dtype=torch.int32, device=dev)
    rank[order.long()] = torch.arange(nv, dtype=torch.int32, device=dev)
    lv = (c_int32 * len(levels))(*[int(v) for v in levels])
    per_node = int(lib.agx_multiscale_scratch_per_node(len(levels), int(x_hops)))
    if x_hops > 8:
        raise NotImplementedError(f"x_hops = {x_hops} > 8 is not built yet")
    scratch = torch.empty(per_node * nv, dtype=torch.int32, device=dev)
    counts = torch.empty(nv, dtype=torch.int32, device=dev)
    offsets = torch.empty(nv + 1, dtype=torch.int64, device=dev)
    stream = current_stream()
    check(
        lib.agx_multiscale_tri_count(
            ico.max_level, ptr(ico.faces_all), lv, len(levels), int(x_hops), ptr(order), ptr(rank), ptr(counts),
            ptr(scratch), stream,
        )
    )  # fmt: skip
    total = c_int64()
    check(lib.agx_exclusive_scan(ptr(counts), nv, ptr(offsets), byref(total), stream))
    out = torch.empty((2, total.value), dtype=torch.int32, device=dev)
    if total.value:
        check(
            lib.agx_multiscale_tri_fill(
                nv, ptr(counts), ptr(offsets), ptr(scratch), per_node, out[0].data_ptr(), out[1].data_ptr(), stream
            )
        )
    return out


__all__ = [
    "NeighbourIndex",
    "NodeTables",
    "Icosphere",
    "edge_attributes",
    "grid_reference_distance",
    "multiscale_tri_edges",
    "haversine_rdist_host",
    "new_stats",
]

_ = ctypes  # keep the import for type users
