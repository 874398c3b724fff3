"""Deterministic synthetic node sets for tests and ``bench.py``.

Every generator returns latitudes / longitudes in DEGREES (float64), the unit the reference's
file / array node builders take (/root/reference/src/anemoi/graphs/nodes/builders/base.py:84-101);
``latlon_deg_to_x`` then applies the reference's ``reshape_coords`` semantics
(stack -> ``np.deg2rad`` in float64 -> ``torch.float32``).
"""

from __future__ import annotations

import numpy as np
import torch


def gaussian_latitudes_deg(n_between_pole_and_equator: int) -> np.ndarray:
    """The 2N Gaussian latitudes in degrees, north to south (Gauss-Legendre nodes)."""
    nodes, _ = np.polynomial.legendre.leggauss(2 * n_between_pole_and_equator)
    return np.rad2deg(np.arcsin(nodes))[::-1].copy()


def _rows_to_points(lat_rows: np.ndarray, n_per_row: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    lats = np.repeat(lat_rows, n_per_row)
    lons = np.concatenate([np.arange(n, dtype=np.float64) * (360.0 / n) for n in n_per_row])
    return lats, lons


def octahedral_grid(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Octahedral reduced Gaussian grid O<n>: row i (1-based from the nearer pole) has 4i+16 points.

    O96 -> 40 320 points, O320 -> 421 120, O1280 -> 6 599 680.
    """
    lat_rows = gaussian_latitudes_deg(n)
    i = np.arange(1, n + 1)
    per_hemisphere = 4 * i + 16
    n_per_row = np.concatenate([per_hemisphere, per_hemisphere[::-1]])
    return _rows_to_points(lat_rows, n_per_row)


def _fft_friendly(n: int) -> int:
    """Smallest integer >= n of the form 2^a 3^b 5^c with a >= 1."""
    m = max(int(n), 2)
    while True:
        k = m
        if k % 2 == 0:
            for p in (2, 3, 5):
                while k % p == 0:
                    k //= p
            if k == 1:
                return m
        m += 1


def reduced_gaussian_grid(n: int) -> tuple[np.ndarray, np.ndarray]:
    """Synthetic classic reduced Gaussian grid N<n>.

    ECMWF's real ``pl`` table for N320 (542 080 points) is neither in the reference nor in this
    image, so a documented rule stands in: each of the 2n Gaussian latitudes gets the smallest
    FFT-friendly (2^a 3^b 5^c, even) count >= 4n*cos(lat), but at least 18.  N320 -> see
    ``tests/test_grids.py`` for the pinned count.
    """
    lat_rows = gaussian_latitudes_deg(n)
    want = np.ceil(4 * n * np.cos(np.deg2rad(lat_rows)) - 1e-9).astype(np.int64)
    n_per_row = np.array([_fft_friendly(max(18, w)) for w in want], dtype=np.int64)
    n_per_row = np.minimum(n_per_row, 4 * n)
    return _rows_to_points(lat_rows, n_per_row)


def uniform_sphere(n: int, seed: int = 1234) -> tuple[np.ndarray, np.ndarray]:
    """Uniform random points on the sphere: lat = arcsin(U(-1,1)), lon = U(0, 360)."""
    rng = np.random.default_rng(seed)
    lat = np.rad2deg(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    lon = rng.uniform(0.0, 360.0, n)
    return lat, lon


def lam_patch(
    n_lat: int = 1000, n_lon: int = 1000, spacing_km: float = 2.5, centre_lat: float = 50.0, centre_lon: float = 10.0
) -> tuple[np.ndarray, np.ndarray]:
    """Regular lat/lon limited-area patch with ``spacing_km`` grid length (row-major, north to south)."""
    dlat = np.rad2deg(spacing_km / 6371.0)
    dlon = dlat / np.cos(np.deg2rad(centre_lat))
    lat_rows = centre_lat + (np.arange(n_lat, dtype=np.float64)[::-1] - (n_lat - 1) / 2.0) * dlat
    lon_cols = centre_lon + (np.arange(n_lon, dtype=np.float64) - (n_lon - 1) / 2.0) * dlon
    lats = np.repeat(lat_rows, n_lon)
    lons = np.tile(np.mod(lon_cols, 360.0), n_lat)
    return lats, lons


def latlon_deg_to_x(latitudes: np.ndarray, longitudes: np.ndarray) -> torch.Tensor:
    """``BaseNodeBuilder.reshape_coords`` (/root/reference/.../nodes/builders/base.py:84-101)."""
    coords = np.stack([latitudes, longitudes], axis=-1).reshape((-1, 2))
    coords = np.deg2rad(coords)
    return torch.tensor(coords, dtype=torch.float32)


def named_grid(name: str) -> tuple[np.ndarray, np.ndarray]:
    """``"o96"``, ``"o1280"``, ``"n320"``, ``"sphere:1000000"``, ``"lam"`` ..."""
    key = name.lower()
    if key.startswith("o") and key[1:].isdigit():
        return octahedral_grid(int(key[1:]))
    if key.startswith("n") and key[1:].isdigit():
        return reduced_gaussian_grid(int(key[1:]))
    if key.startswith("sphere:"):
        return uniform_sphere(int(key.split(":")[1]))
    if key == "lam":
        return lam_patch()
    raise ValueError(f"Unknown synthetic grid {name!r}")
