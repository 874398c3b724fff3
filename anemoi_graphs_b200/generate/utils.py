"""Node ordering (/root/reference/src/anemoi/graphs/generate/utils.py:15-33)."""

import numpy as np


def get_coordinates_ordering(coords: np.ndarray | None = None, lat: np.ndarray | None = None, lon: np.ndarray | None = None):
    """Order that sorts node coordinates by latitude (descending) and longitude.

    Evaluated ON THE HOST with the reference's numpy sorts: both argsorts are numpy's default (unstable) kind
    and the coordinates contain tens of thousands of ties, so the resulting order is defined by numpy's own
    sort implementation on this machine - any re-implementation would label the nodes
    differently from the reference.  The node order is an INPUT of the GPU path (it only relabels indices).

    The reference writes ``argsort(coords[:, 1])``, ``argsort(coords[index_latitude][:, 0])[::-1]`` and
    ``arange(n)[index_latitude][index_longitude]``; numpy copies a strided column into a contiguous buffer
    before sorting, so sorting contiguous ``lat`` / ``lon`` columns (and gathering one column instead of whole
    rows) gives the identical permutation at a fifth of the host time (tests/test_api_cpu.py pins this)."""
    if coords is not None:
        lat = np.ascontiguousarray(coords[:, 0])
        lon = np.ascontiguousarray(coords[:, 1])
    index_latitude = np.argsort(lon)
    index_longitude = np.argsort(lat[index_latitude])[::-1]
    node_ordering = index_latitude[index_longitude]
    return node_ordering
