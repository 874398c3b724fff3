"""Pins ``oracle/exact_search.c`` (float64 haversine with libm over a latitude-band grid) to sklearn itself: the k
nearest (indices AND float64 rdist bits), radius sets, poles, the date line, duplicate points, and the tie report
against ``oracle.ref_path.knn_edges_canonical`` (which asks sklearn)."""

import math

import numpy as np
import pytest
from sklearn.neighbors import NearestNeighbors

from anemoi_graphs_b200 import grids
from oracle import exact_search as X
from oracle import ref_path as R


def _x(lat_deg, lon_deg) -> np.ndarray:
    return grids.latlon_deg_to_x(lat_deg, lon_deg).numpy()


def _sk_rdist(dist: np.ndarray) -> np.ndarray:
    return np.sin(0.5 * dist) ** 2


@pytest.mark.parametrize("n_src,n_q,k", [(5000, 3000, 3), (20000, 4000, 11), (300, 500, 16)])
def test_knn_matches_sklearn_on_random_clouds(n_src, n_q, k):
    src = _x(*grids.uniform_sphere(n_src, seed=n_src))
    q = _x(*grids.uniform_sphere(n_q, seed=n_q + 1))
    ind, rd = X.Grid(src).knn(q, k)
    nn = NearestNeighbors(metric="haversine").fit(src)
    dist, want = nn.kneighbors(q, n_neighbors=k)
    np.testing.assert_array_equal(ind, want)  # random clouds have no ties: the order is unique
    # the same float64 values sklearn's compiled rdist produces (dist = 2 asin sqrt(rdist) is monotone)
    np.testing.assert_array_equal(2.0 * np.arcsin(np.sqrt(rd)), dist)
    # and as sorted as claimed
    assert (np.diff(rd, axis=1) >= 0).all()


def test_knn_poles_dateline_duplicates():
    lat = np.array([90.0, -90.0, 89.999, 0.0, 0.0, 0.0, 45.0, 45.0, 10.0, 10.0])
    lon = np.array([0.0, 0.0, 123.0, 359.999, 0.001, 180.0, -179.999, 179.999, 20.0, 20.0])  # last two: duplicates
    extra = _x(*grids.uniform_sphere(2000, seed=5))
    src = np.concatenate([_x(lat, lon), extra])
    q = np.concatenate([_x(lat, lon), _x(*grids.uniform_sphere(500, seed=6))])
    ind, rd = X.Grid(src).knn(q, 5)
    nn = NearestNeighbors(metric="haversine").fit(src)
    dist, want = nn.kneighbors(q, n_neighbors=5)
    np.testing.assert_array_equal(2.0 * np.arcsin(np.sqrt(rd)), dist)
    same = (ind == want).all(axis=1)
    # rows may differ only where two candidates are exactly equidistant (the duplicate pair)
    for row in np.nonzero(~same)[0]:
        assert sorted(ind[row]) == sorted(want[row]) or len(set(rd[row])) < 5


def test_knn_whole_o96_res5_vs_sklearn_route():
    dx = _x(*grids.octahedral_grid(96))
    hx, _ = R.tri_nodes(5)
    want, info = R.knn_edges_canonical(hx, dx, 3)
    got, ginfo = X.knn_edges_canonical(hx, dx, 3)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(ginfo["tied_queries"], info["tied_queries"])
    assert ginfo["tied_queries"].size == 88
    assert all(r["rdist_bit_equal"] for r in ginfo["report"])  # mirror-image sources: bit-equal float64 rdist


def test_radius_matches_sklearn():
    dx = _x(*grids.octahedral_grid(48))
    hx, _ = R.tri_nodes(4)
    radius = R.cutoff_radius(hx, 0.6)
    got, near = X.cutoff_edges(dx, hx, radius)
    np.testing.assert_array_equal(got, R.canonical_sort(R.cutoff_edges(dx, hx, 0.6)))
    assert near == 0
    # a large radius that covers the poles and wraps the date line
    src = _x(*grids.uniform_sphere(3000, seed=9))
    q = np.concatenate([_x(np.array([90.0, -90.0, 0.0]), np.array([0.0, 10.0, 359.9])), src[:200]])
    for r in (0.05, 0.7, 2.0):
        off, s, _ = X.Grid(src, cell_rad=0.05).radius(q, r)
        nn = NearestNeighbors(metric="haversine").fit(src)
        ind = nn.radius_neighbors(q, radius=r, return_distance=False)
        for i in range(q.shape[0]):
            np.testing.assert_array_equal(s[off[i] : off[i + 1]], np.sort(ind[i]))


def test_pair_rdist_is_libm():
    rng = np.random.default_rng(0)
    a = rng.uniform(-1.5, 1.5, (1000, 2)).astype(np.float32)
    b = rng.uniform(-1.5, 1.5, (1000, 2)).astype(np.float32)
    got = X.pair_rdist(a, b)
    for i in range(0, 1000, 37):
        lat1, lon1, lat2, lon2 = float(a[i, 0]), float(a[i, 1]), float(b[i, 0]), float(b[i, 1])
        s0, s1 = math.sin(0.5 * (lat1 - lat2)), math.sin(0.5 * (lon1 - lon2))
        assert got[i] == s0 * s0 + math.cos(lat1) * math.cos(lat2) * s1 * s1
