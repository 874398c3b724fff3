"""The HEALPix restatement (oracle/healpix_restated.py) against what can be checked without healpy: the independent
RING-scheme formulas, pixel counts, analytic base pixels, and the fixture written by the UNMODIFIED reference
HEALPixNodes over the healpy shim."""

import numpy as np
import pytest

from oracle import healpix_restated as P


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 64])
def test_nested_centres_are_the_ring_scheme_centres(nside):
    z, phi, _, _ = P.pix2zphi_nest(nside, np.arange(12 * nside * nside))
    zr, pr = P.ring_centres(nside)
    assert z.size == zr.size == P.nside2npix(nside)
    a = np.lexsort((np.round(phi, 11), np.round(z, 11)))
    b = np.lexsort((np.round(pr, 11), np.round(zr, 11)))
    np.testing.assert_allclose(z[a], zr[b], rtol=0, atol=1e-15)
    np.testing.assert_allclose(phi[a], pr[b], rtol=0, atol=2e-15)
    assert np.unique(np.round(z, 12)).size == 4 * nside - 1  # iso-latitude rings


def test_base_pixels_and_nested_order():
    lon, lat = P.pix2ang_nest_lonlat(1)
    np.testing.assert_allclose(lat, np.repeat([np.degrees(np.arcsin(2 / 3)), 0.0, -np.degrees(np.arcsin(2 / 3))], 4), atol=1e-12)
    np.testing.assert_allclose(lon, [45, 135, 225, 315, 0, 90, 180, 270, 45, 135, 225, 315], atol=1e-12)
    # nside 2, face 0: pixel 0 is the southern corner, 1 / 2 its eastern / western neighbours, 3 the northern corner
    lon, lat = P.pix2ang_nest_lonlat(2, np.arange(4))
    np.testing.assert_allclose(lon, [45.0, 67.5, 22.5, 45.0], atol=1e-12)
    np.testing.assert_allclose(lat, np.degrees(np.arcsin([1 / 3, 2 / 3, 2 / 3, 11 / 12])), atol=1e-12)


def test_oracle_matches_reference_healpix_nodes(golden):
    g = golden("healpix")
    for res in (1, 3):
        np.testing.assert_array_equal(P.healpix_nodes_x(res), g[f"res{res}_x"])
    assert abs(P.nside2resol(8, arcmin=True) - np.degrees(np.sqrt(4 * np.pi / 768)) * 60) < 1e-12
