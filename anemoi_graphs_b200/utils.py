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
    result.  The result allocation doubles as the sort's alternate buffer: the scratch besides it is one key buffer (8 bytes
    per input edge); no concatenated int64 list, no ``torch.unique``."""
    from ctypes import byref, c_int64

    from ._cabi import check, current_stream, load_library

    _device.wait_for(e1)  # a sharded builder's all-gather may still be filling its result
    _device.wait_for(e2)
    for e in (e1, e2):
        assert e.is_cuda and e.dtype == torch.int32 and e.dim() == 2 and e.shape[0] == 2
    e1, e2 = e1.contiguous(), e2.contiguous()
    n = int(e1.shape[1]) + int(e2.shape[1])
    # the result allocation, sized for "no duplicates": it doubles as the sort's alternate key buffer
    buf = torch.empty(2 * max(n, 1), dtype=torch.int32, device=e1.device)
    n_unique = c_int64()
    big = 2**31 - 1
    check(
        load_library().agx_concat_edges(
            e1[0].data_ptr(), e1[1].data_ptr(), int(e1.shape[1]), e2[0].data_ptr(), e2[1].data_ptr(), int(e2.shape[1]),
            int(n_src_nodes or big), int(n_dst_nodes or big), buf.data_ptr(), byref(n_unique), current_stream(),
        )
    )
    u = n_unique.value
    out = buf[: 2 * u].view(2, u)
    if n > 0 and u < 0.75 * n:  # many duplicates: do not keep the worst-case allocation alive behind a small result
        out = out.clone()
    return out


def concat_edges(edge_indices1: torch.Tensor, edge_indices2: torch.Tensor) -> torch.Tensor:
    """utils.py:66-81.  Result lives where ``edge_indices1`` lives."""
    out = concat_edges_device(
        _device.to_device(edge_indices1, torch.int32), _device.to_device(edge_indices2, torch.int32)
    )
    res = _device.like_input(out, edge_indices1)
    _device.maybe_flush()
    return res
