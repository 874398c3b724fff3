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


def concat_edges_device(e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    """``torch.unique(cat, dim=1)`` of two CUDA (2, E) int32 edge lists: columns sorted by (src, dst), unique.

    Column-wise unique of a 2-row int32 array is a sort-unique of the packed 64-bit key ``src << 32 | dst``
    (indices are non-negative), done with the device sort (plumbing)."""
    _device.wait_for(e1)  # a sharded builder's all-gather may still be filling its result
    _device.wait_for(e2)
    cat = torch.cat([e1, e2], dim=1).to(torch.int64)
    key = torch.unique((cat[0] << 32) | cat[1], sorted=True)
    return torch.stack([key >> 32, key & 0xFFFFFFFF]).to(torch.int32)


def concat_edges(edge_indices1: torch.Tensor, edge_indices2: torch.Tensor) -> torch.Tensor:
    """utils.py:66-81.  Result lives where ``edge_indices1`` lives."""
    out = concat_edges_device(
        _device.to_device(edge_indices1, torch.int32), _device.to_device(edge_indices2, torch.int32)
    )
    res = _device.like_input(out, edge_indices1)
    _device.maybe_flush()
    return res
