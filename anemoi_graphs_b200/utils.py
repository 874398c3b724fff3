"""Graph utilities on the device (/root/reference/src/anemoi/graphs/utils.py)."""

from __future__ import annotations

import torch

from . import device as _device
from . import ops


def get_grid_reference_distance(coords_rad: torch.Tensor, mask: torch.Tensor | None = None, state=None) -> float:
    """Largest nearest-neighbour distance of a node set, float64 radians (utils.py:44-63).

    Like the reference, ``mask`` is only shape-checked and otherwise ignored (utils.py:32-39).  ``state`` (the node
    set's ``device.NodeState``, when ``coords_rad`` is its coordinate tensor) lets the search share the node set's
    cached neighbour index with the other builders of the recipe."""
    assert mask is None or mask.shape == (
        coords_rad.shape[0],
        1,
    ), "Mask must have the same shape as the number of nodes."
    return ops.grid_reference_distance(_device.to_device(coords_rad, torch.float32), state=state)


def concat_edges_device(e1: torch.Tensor, e2: torch.Tensor, n_src_nodes: int | None = None, n_dst_nodes: int | None = None) -> torch.Tensor:
    """``torch.unique(cat, dim=1)`` of two CUDA (2, E) int32 edge lists: columns sorted by (src, dst), unique
    (utils.py:66-81) - ``agx_concat_edges_*``: both lists packed into 64-bit keys ``src << 32 | dst`` by one kernel, a
    radix sort over the key bits the node counts can set, the distinct keys unpacked straight into the rows of the
    result.  Peak extra memory is two key buffers (twice the result); no concatenated int64 list, no ``torch.unique``."""
    from ctypes import byref, c_int64, c_void_p

    from ._cabi import check, current_stream, load_library

    _device.wait_for(e1)  # a sharded builder's all-gather may still be filling its result
    _device.wait_for(e2)
    for e in (e1, e2):
        assert e.is_cuda and e.dtype == torch.int32 and e.dim() == 2 and e.shape[0] == 2
    e1, e2 = e1.contiguous(), e2.contiguous()
    lib = load_library()
    handle, n_unique = c_void_p(), c_int64()
    big = 2**31 - 1
    check(
        lib.agx_concat_edges_begin(
            e1[0].data_ptr(), e1[1].data_ptr(), int(e1.shape[1]), e2[0].data_ptr(), e2[1].data_ptr(), int(e2.shape[1]),
            int(n_src_nodes or big), int(n_dst_nodes or big), byref(handle), byref(n_unique), current_stream(),
        )
    )
    out = torch.empty((2, n_unique.value), dtype=torch.int32, device=e1.device)
    check(lib.agx_concat_edges_finish(handle, out[0].data_ptr(), out[1].data_ptr(), current_stream()))
    return out


def concat_edges(edge_indices1: torch.Tensor, edge_indices2: torch.Tensor) -> torch.Tensor:
    """utils.py:66-81.  Result lives where ``edge_indices1`` lives."""
    out = concat_edges_device(
        _device.to_device(edge_indices1, torch.int32), _device.to_device(edge_indices2, torch.int32)
    )
    res = _device.like_input(out, edge_indices1)
    _device.maybe_flush()
    return res
