"""ORACLE SHIM - TEST INFRASTRUCTURE ONLY: the three healpy calls the reference makes
(/root/reference/src/anemoi/graphs/nodes/builders/from_healpix.py:61-66) answered by ``oracle/healpix_restated.py``,
so that the UNMODIFIED ``HEALPixNodes`` / ``LimitedAreaHEALPixNodes`` run (``tests/golden/healpix.npz``)."""

from __future__ import annotations

import importlib.util
import pathlib

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "_oracle_healpix_restated", pathlib.Path(__file__).resolve().parents[2] / "healpix_restated.py"
)
_P = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_P)

nside2npix = _P.nside2npix
nside2resol = _P.nside2resol


def pix2ang(nside, ipix, nest=False, lonlat=False):
    if not (nest and lonlat):
        raise NotImplementedError("shim: only nest=True, lonlat=True (what the reference asks for)")
    return _P.pix2ang_nest_lonlat(int(nside), np.asarray(list(ipix) if isinstance(ipix, range) else ipix))
