"""ORACLE (test infrastructure - never imported by the product path).

CPU restatement of ``trimesh.creation.icosphere`` for the one use the reference makes of it
(/root/reference/src/anemoi/graphs/generate/tri_icosahedron.py:121,173,212).

trimesh (``trimesh>=4.1``, /root/reference/pyproject.toml:55) is a third-party dependency
whose source is NOT under /root/reference and which is not installed in this image, so its
published algorithm is restated here:

* ``trimesh.creation.icosahedron``: the 12-vertex / 20-face golden-ratio table, vertices
  scaled by ``1/sqrt(2+t)``, ``t=(1+sqrt 5)/2``;
* ``Trimesh.subdivide`` (``trimesh.remesh.subdivide``): ``edges = sort(faces_to_edges(faces))``,
  one midpoint per unique edge, unique edges found with ``grouping.unique_rows`` = ``np.unique``
  on the row packed as ``v_min | v_max << 32`` (ascending => midpoints are numbered by
  (max, min) vertex), ``mid = vertices[edge].mean(axis=1)``, new vertices appended after the old
  ones, each face replaced by 4 faces ``[a, ab, ca], [ab, b, bc], [ca, bc, c], [ab, bc, ca]``;
* ``icosphere``'s ``refine_spherical`` after every subdivision:
  ``scalar = sqrt(dot(v**2, [1,1,1]))``, ``v += (v / scalar) * (radius - scalar)`` on ALL vertices.

Parity status: pinned only by counts (docs tri_refined_edges.csv / tri_nodes.csv,
tests/nodes/test_tri_nodes.py:32) - vertex *numbering* cannot be checked against real trimesh
here.  It influences the reference only through argsort tie-breaks.
"""

from __future__ import annotations

import numpy as np


def icosahedron() -> tuple[np.ndarray, np.ndarray]:
    t = (1.0 + 5.0**0.5) / 2.0
    vertices = [
        -1, t, 0, 1, t, 0, -1, -t, 0, 1, -t, 0,
        0, -1, t, 0, 1, t, 0, -1, -t, 0, 1, -t,
        t, 0, -1, t, 0, 1, -t, 0, -1, -t, 0, 1,
    ]  # fmt: skip
    faces = [
        0, 11, 5, 0, 5, 1, 0, 1, 7, 0, 7, 10, 0, 10, 11,
        1, 5, 9, 5, 11, 4, 11, 10, 2, 10, 7, 6, 7, 1, 8,
        3, 9, 4, 3, 4, 2, 3, 2, 6, 3, 6, 8, 3, 8, 9,
        4, 9, 5, 2, 4, 11, 6, 2, 10, 8, 6, 7, 9, 8, 1,
    ]  # fmt: skip
    v = np.reshape(np.array(vertices, dtype=np.float64), (-1, 3)) / np.sqrt(2.0 + t)
    f = np.reshape(np.array(faces, dtype=np.int64), (-1, 3))
    return v, f


def faces_to_edges(faces: np.ndarray) -> np.ndarray:
    return faces[:, [0, 1, 1, 2, 2, 0]].reshape((-1, 2))


def unique_rows_packed(edges_sorted: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``trimesh.grouping.unique_rows`` for 2-column non-negative integer rows."""
    key = edges_sorted[:, 0].astype(np.uint64) | (edges_sorted[:, 1].astype(np.uint64) << np.uint64(32))
    _, unique, inverse = np.unique(key, return_index=True, return_inverse=True)
    return unique, inverse


def subdivide(vertices: np.ndarray, faces: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    edges = np.sort(faces_to_edges(faces), axis=1)
    unique, inverse = unique_rows_packed(edges)
    mid = vertices[edges[unique]].mean(axis=1)
    mid_idx = inverse.reshape((-1, 3)) + len(vertices)
    f = np.column_stack(
        [
            faces[:, 0], mid_idx[:, 0], mid_idx[:, 2],
            mid_idx[:, 0], faces[:, 1], mid_idx[:, 1],
            mid_idx[:, 2], mid_idx[:, 1], faces[:, 2],
            mid_idx[:, 0], mid_idx[:, 1], mid_idx[:, 2],
        ]
    ).reshape((-1, 3))  # fmt: skip
    return np.vstack((vertices, mid)), f


def refine_spherical(vertices: np.ndarray, radius: float = 1.0) -> np.ndarray:
    scalar = np.sqrt(np.dot(vertices**2, [1, 1, 1]))
    unit = vertices / scalar.reshape((-1, 1))
    return vertices + unit * (radius - scalar).reshape((-1, 1))


def icosphere(subdivisions: int = 3, radius: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
    """Vertices (float64, (10*4**s+2, 3)) and faces (int64, (20*4**s, 3))."""
    v, f = icosahedron()
    for _ in range(subdivisions):
        v, f = subdivide(v, f)
        v = refine_spherical(v, radius)
    return v, f


def edges_unique(faces: np.ndarray) -> np.ndarray:
    """``Trimesh.edges_unique``: sorted vertex pairs, one row per undirected edge."""
    edges = np.sort(faces_to_edges(faces), axis=1)
    unique, _ = unique_rows_packed(edges)
    return edges[unique]
