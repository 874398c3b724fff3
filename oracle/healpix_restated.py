"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy float64) of the one healpy call the reference makes for HEALPix nodes
(/root/reference/src/anemoi/graphs/nodes/builders/from_healpix.py:61-66):
``hp.pix2ang(nside, range(npix), nest=True, lonlat=True)``.  healpy is not installable here; the algorithm is
HEALPix's published ``pix2loc`` for the NESTED scheme (Gorski et al. 2005; healpix_cxx healpix_base.cc): the pixel
number splits into a base face and bit-interleaved (ix, iy) inside the face, the ring index is
``jr = jrll[face] * nside - ix - iy - 1``, and z / phi follow from the polar-cap or equatorial-belt formulas.

Parity with healpy itself is UNPINNED; what is checked (tests/test_oracle_healpix.py): the nested centres are
exactly the textbook RING-scheme centres (an independent formula), 12 nside^2 of them, 4 nside - 1 iso-latitude
rings with the right populations, nside = 1 reproduces the analytic base-pixel centres (z = 2/3, 0, -2/3).
"""

from __future__ import annotations

import numpy as np

JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)
HALFPI = 1.570796326794896619231321691639751442099


def nside2npix(nside: int) -> int:
    return 12 * nside * nside


def nside2resol(nside: int, arcmin: bool = False) -> float:
    """healpy.nside2resol: sqrt of the pixel area, radians (or arc minutes)."""
    resol = np.sqrt(4.0 * np.pi / nside2npix(nside))
    return float(np.rad2deg(resol) * 60.0) if arcmin else float(resol)


def _compress_bits(v: np.ndarray) -> np.ndarray:
    """every second bit of v (bit 0, 2, 4, ...) packed together."""
    out = np.zeros_like(v)
    for b in range(32):
        out |= ((v >> (2 * b)) & 1) << b
    return out


def pix2zphi_nest(nside: int, ipix) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """healpix_base.cc ``pix2loc`` (NEST): z, phi, and sin(theta) where the library computes it separately near the
    poles (``have_sth``)."""
    ipix = np.asarray(ipix, dtype=np.int64)
    order = int(np.log2(nside))
    assert 1 << order == nside, "nside must be a power of two in the nested scheme"
    npface = nside * nside
    npix = 12 * npface
    fact2 = 4.0 / npix
    fact1 = (nside << 1) * fact2
    face = ipix >> (2 * order)
    p = ipix & (npface - 1)
    ix, iy = _compress_bits(p), _compress_bits(p >> 1)
    jr = (JRLL[face] << order) - ix - iy - 1
    north, south = jr < nside, jr > 3 * nside
    nr = np.where(north, jr, np.where(south, 4 * nside - jr, nside))
    tmp = (nr * nr) * fact2
    z = np.where(north, 1.0 - tmp, np.where(south, tmp - 1.0, (2 * nside - jr) * fact1))
    have_sth = (north & (z > 0.99)) | (south & (z < -0.99))
    sth = np.where(have_sth, np.sqrt(tmp * (2.0 - tmp)), 0.0)
    t = JPLL[face] * nr + ix - iy
    t = np.where(t < 0, t + 8 * nr, t)
    phi = np.where(nr == nside, 0.75 * HALFPI * t * fact1, (0.5 * HALFPI * t) / nr)
    return z, phi, sth, have_sth


def pix2ang_nest_lonlat(nside: int, ipix=None) -> tuple[np.ndarray, np.ndarray]:
    """``hp.pix2ang(nside, ipix, nest=True, lonlat=True)``: (lon, lat) in degrees.  theta = atan2(sth, z) where the
    library has sin(theta), acos(z) elsewhere; lon = degrees(phi), lat = 90 - degrees(theta)."""
    if ipix is None:
        ipix = np.arange(nside2npix(nside))
    z, phi, sth, have = pix2zphi_nest(nside, ipix)
    theta = np.where(have, np.arctan2(sth, z), np.arccos(np.clip(z, -1.0, 1.0)))
    return np.degrees(phi), 90.0 - np.degrees(theta)


def ring_centres(nside: int) -> tuple[np.ndarray, np.ndarray]:
    """(z, phi) of all pixel centres from the textbook RING-scheme formulas (Gorski et al. 2005, eqs. 2-9) - an
    independent route to the same point set."""
    zs, phis = [], []
    for i in range(1, 4 * nside):
        if i < nside:  # north polar cap
            n_in_ring, z = 4 * i, 1.0 - i * i / (3.0 * nside * nside)
            phi = (np.arange(1, n_in_ring + 1) - 0.5) * np.pi / (2.0 * i)
        elif i <= 3 * nside:  # equatorial belt
            n_in_ring, z = 4 * nside, 4.0 / 3.0 - 2.0 * i / (3.0 * nside)
            s = (i - nside + 1) % 2
            phi = (np.arange(1, n_in_ring + 1) - s / 2.0) * np.pi / (2.0 * nside)
        else:  # south polar cap
            ii = 4 * nside - i
            n_in_ring, z = 4 * ii, -1.0 + ii * ii / (3.0 * nside * nside)
            phi = (np.arange(1, n_in_ring + 1) - 0.5) * np.pi / (2.0 * ii)
        zs.append(np.full(n_in_ring, z))
        phis.append(np.mod(phi, 2.0 * np.pi))
    return np.concatenate(zs), np.concatenate(phis)


def healpix_nodes_x(resolution: int) -> np.ndarray:
    """The reference's ``HEALPixNodes(resolution).get_coordinates()``: float32 (N, 2) radians
    (from_healpix.py:61-66 + nodes/builders/base.py:84-101)."""
    lon, lat = pix2ang_nest_lonlat(2**resolution)
    coords = np.stack([lat, lon], axis=-1).reshape((-1, 2))
    return np.deg2rad(coords).astype(np.float32)
