"""Node ordering (/root/reference/src/anemoi/graphs/generate/utils.py:15-33)."""

import numpy as np


def get_coordinates_ordering(coords: np.ndarray) -> np.ndarray:
    """Order that sorts node coordinates by latitude (descending) and longitude.

    Evaluated ON THE HOST with the reference's exact numpy calls: both argsorts are numpy's default
    (unstable) kind and the coordinates contain tens of thousands of ties, so the resulting order is
    defined by numpy's own sort implementation on this machine - any re-implementation
    would label the nodes differently from the reference.  O(N log N) on <= 1e6 nodes; the node order is an
    INPUT of the GPU path (it only relabels indices)."""
    index_latitude = np.argsort(coords[:, 1])
    index_longitude = np.argsort(coords[index_latitude][:, 0])[::-1]
    node_ordering = np.arange(coords.shape[0])[index_latitude][index_longitude]
    return node_ordering
