"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes front end of ``oracle/exact_search.c``.

Whole-graph checks at O1280 size need an oracle that finishes in seconds; sklearn's ball tree takes minutes
for 6.6 M queries.  ``exact_search.c`` evaluates sklearn's float64 haversine ``rdist`` with the C library's
``sin`` / ``cos`` (the calls sklearn's compiled ``HaversineDistance64`` makes) over a latitude-band grid and
returns, per query, the ``k + extra`` nearest sources sorted by (rdist, index), or every source within a radius.
It is pinned against sklearn itself in ``tests/test_oracle_exact_search.py`` and - at full size - against
``NearestNeighbors.kneighbors`` / ``radius_neighbors`` in ``tests/test_gpu_full_size.py``.

The functions at the bottom restate, on top of that, what ``oracle/ref_path.py`` does with sklearn:
``knn_edges_canonical`` (north-star tie rule: lower source index inside a tie group of relative width tau) and
``cutoff_edges``; they return the same canonical arrays.
"""

from __future__ import annotations

import ctypes
import math
import os
import pathlib
import subprocess

import numpy as np

ORACLE = pathlib.Path(__file__).resolve().parent
SRC = ORACLE / "exact_search.c"
LIB = ORACLE / "_build" / "libexact_search.so"
CFLAGS = ["-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC"]
TIE_TAU = 2.0**-40

_lib = None


def build(force: bool = False) -> pathlib.Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(exist_ok=True)
        tmp = LIB.with_suffix(".so.tmp")
        subprocess.run(["gcc", *CFLAGS, str(SRC), "-o", str(tmp), "-lm"], check=True)
        os.replace(tmp, LIB)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        vp, i64, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        _lib.oracle_grid_build.restype = vp
        _lib.oracle_grid_build.argtypes = [vp, i64, f64]
        _lib.oracle_grid_free.argtypes = [vp]
        _lib.oracle_knn.argtypes = [vp, vp, i64, ctypes.c_int, f64, vp, vp]
        _lib.oracle_radius_count.argtypes = [vp, vp, i64, f64, f64, vp, vp]
        _lib.oracle_radius_fill.argtypes = [vp, vp, i64, f64, vp, vp]
        _lib.oracle_pair_rdist.argtypes = [vp, vp, i64, vp]
    return _lib


def _f32(x) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2 and x.shape[1] == 2
    return x


class Grid:
    """Latitude-band grid over a source set; ``cell_rad`` ~ the expected neighbour distance."""

    def __init__(self, source_x, cell_rad: float | None = None) -> None:
        self.x = _f32(source_x)
        n = self.x.shape[0]
        if cell_rad is None:  # ~2 sources per bin
            cell_rad = math.sqrt(4.0 * math.pi / max(n, 1) * 2.0)
        self.cell_rad = float(min(max(cell_rad, 1e-4), 0.5))
        self.handle = lib().oracle_grid_build(self.x.ctypes.data, n, self.cell_rad)

    def __del__(self) -> None:  # pragma: no cover
        if getattr(self, "handle", None):
            lib().oracle_grid_free(self.handle)
            self.handle = None

    def knn(self, q, kk: int) -> tuple[np.ndarray, np.ndarray]:
        """``(index (nq, kk) int32, rdist (nq, kk) float64)`` ascending by (rdist, index)."""
        q = _f32(q)
        nq = q.shape[0]
        idx = np.empty((nq, kk), dtype=np.int32)
        rd = np.empty((nq, kk), dtype=np.float64)
        r0 = self.cell_rad * max(1.0, math.sqrt(kk / 2.0))
        rc = lib().oracle_knn(self.handle, q.ctypes.data, nq, kk, r0, idx.ctypes.data, rd.ctypes.data)
        if rc != 0:
            raise ValueError(f"kk = {kk} > number of sources {self.x.shape[0]}")
        return idx, rd

    def radius(self, q, radius: float, tau: float = TIE_TAU):
        """``(offsets (nq+1,) int64, sources int32 ascending per query, near-threshold pair count)``."""
        q = _f32(q)
        nq = q.shape[0]
        counts = np.empty(nq, dtype=np.int64)
        near = np.empty(nq, dtype=np.int64)
        lib().oracle_radius_count(self.handle, q.ctypes.data, nq, float(radius), float(tau), counts.ctypes.data, near.ctypes.data)
        offsets = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        src = np.empty(int(offsets[-1]), dtype=np.int32)
        lib().oracle_radius_fill(self.handle, q.ctypes.data, nq, float(radius), offsets.ctypes.data, src.ctypes.data)
        return offsets, src, int(near.sum())


def pair_rdist(q, s) -> np.ndarray:
    """sklearn ``rdist(query, source)`` of explicit pairs with libm (float64)."""
    q, s = _f32(q), _f32(s)
    out = np.empty(q.shape[0], dtype=np.float64)
    lib().oracle_pair_rdist(q.ctypes.data, s.ctypes.data, q.shape[0], out.ctypes.data)
    return out


# ----------------------------------------------------------------------------------------------------------------
# the reference's edge sets on top of the exact search
# ----------------------------------------------------------------------------------------------------------------
def knn_candidates(source_x, target_x, k: int, extra: int = 8):
    kk = min(int(np.asarray(source_x).shape[0]), k + extra)
    return Grid(source_x).knn(target_x, kk)


def knn_tie_groups(ind: np.ndarray, rd: np.ndarray, k: int, tau: float = TIE_TAU):
    """From the (rdist, index)-sorted candidates: the k chosen under the lower-index rule and the tie report.

    A tie group = the candidates whose rdist lies within ``tau`` (relative) of the k-th smallest; when the (k+1)-th
    candidate is inside it the query is TIED and the lowest source indices of the group are kept.  Returns
    ``(chosen (nq, k) int32 sorted by index, tied query ids, per tied query: dict)``."""
    nq, kk = ind.shape
    chosen = np.sort(ind[:, :k], axis=1)
    tied = np.empty(0, dtype=np.int64)
    report = []
    if kk > k:
        r_k = rd[:, k - 1]
        has_tie = np.abs(rd[:, k] - r_k) <= tau * r_k
        tied = np.nonzero(has_tie)[0]
        for q in tied:
            g = np.abs(rd[q] - r_k[q]) <= tau * r_k[q]
            if g[-1] and kk < 2**31:  # group may extend past the candidates fetched
                pass
            below = ind[q][(rd[q] < r_k[q]) & ~g]
            group = np.sort(ind[q][g])
            pick = np.concatenate([below, group[: k - below.size]])
            chosen[q] = np.sort(pick)
            grd = rd[q][g]
            report.append(
                {
                    "query": int(q),
                    "tied_sources": [int(v) for v in group],
                    "tied_rdist_hex": [float(v).hex() for v in grd[np.argsort(ind[q][g], kind="stable")]],
                    "rdist_bit_equal": bool((grd == grd[0]).all()),
                    "relative_spread": float((grd.max() - grd.min()) / r_k[q]) if r_k[q] > 0 else 0.0,
                    "group_open_ended": bool(g[-1]),
                    "chosen": [int(v) for v in np.sort(pick)],
                }
            )
    return chosen.astype(np.int32), tied, report


def knn_edges_canonical(source_x, target_x, k: int, extra: int = 8, tau: float = TIE_TAU):
    """Same contract as ``ref_path.knn_edges_canonical`` minus the sklearn comparison: ``(edge_index sorted by
    (dst, src), info)`` with ``info["tied_queries"]``, ``info["report"]``, and the candidate arrays."""
    ind, rd = knn_candidates(source_x, target_x, k, extra)
    chosen, tied, report = knn_tie_groups(ind, rd, k, tau)
    open_ended = [r["query"] for r in report if r["group_open_ended"] and ind.shape[1] < np.asarray(source_x).shape[0]]
    if open_ended:
        raise RuntimeError(f"tie group wider than extra={extra} for queries {open_ended[:5]}...; raise `extra`")
    nq = chosen.shape[0]
    dst = np.repeat(np.arange(nq, dtype=np.int32), k)
    edge_index = np.stack([chosen.reshape(-1), dst], axis=0)
    return edge_index, {"tied_queries": tied, "report": report, "ind": ind, "rdist": rd}


def cutoff_edges(source_x, target_x, radius: float, tau: float = TIE_TAU):
    """All (source, target) pairs with ``rdist <= sin(r/2)^2`` sorted by (dst, src) + the count of pairs within
    ``tau`` (relative) of the threshold."""
    offsets, src, near = Grid(source_x, cell_rad=max(radius, 1e-4)).radius(target_x, radius, tau)
    nq = np.asarray(target_x).shape[0]
    dst = np.repeat(np.arange(nq, dtype=np.int32), np.diff(offsets))
    return np.stack([src, dst], axis=0), near
