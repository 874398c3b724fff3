"""ORACLE SHIM - TEST INFRASTRUCTURE ONLY: the slice of the ``h3`` (v3) Python API the reference calls
(/root/reference/src/anemoi/graphs/generate/hex_icosahedron.py:47,79,99,147-151,190-194,202-210), backed by the
numpy restatement of H3's geometry in ``oracle/h3_restated.py``.

The point is NOT to be h3 (the real library is not installable here) but to let the UNMODIFIED reference code run
end to end - ``create_hex_nodes``, ``add_edges_to_nx_graph``, networkx, scipy - so that the way the reference USES
h3 (``k_ring(idx, k) & nodes``, ``compact`` / ``uncompact`` of the graph's nodes, centre children, ``add_edge``'s
membership rules, the node ordering) is pinned by golden fixtures from the reference's own control flow
(``oracle/make_golden.py make_hex``), not by a second restatement.

Cell identifiers are synthetic strings ``"<res>:<position in h3_restated.cell_centers(res)>"``; real H3 indices never
reach the graph (nodes are re-ordered by coordinates, edges are positions).  ``uncompact`` returns a ``set`` like
h3-py v3 does, so the reference sees the same arbitrary iteration order it sees with the real library.
"""

from __future__ import annotations

import functools
import importlib.util
import pathlib

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "_oracle_h3_restated", pathlib.Path(__file__).resolve().parents[2] / "h3_restated.py"
)
_H = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_H)


@functools.lru_cache(maxsize=None)
def _level(res: int):
    centers, pent = _H.cell_centers(res)
    return centers, pent, _H.neighbours(centers, pent)


@functools.lru_cache(maxsize=None)
def _children_of(res: int):
    """(n_res, 7) positions at res + 1 of every cell's children (-1 pads the 6-child pentagons): the centre child and
    the cells around it (aperture 7)."""
    coarse, _, _ = _level(res)
    fine, _, nb = _level(res + 1)
    cc = _H.center_child_positions(coarse, fine)
    return np.concatenate([cc[:, None], nb[cc]], axis=1)


@functools.lru_cache(maxsize=None)
def _parent_of(res: int):
    """position at res - 1 of every cell of ``res``."""
    fam = _children_of(res - 1)
    parent = np.full(_level(res)[0].shape[0], -1, dtype=np.int64)
    for p, kids in enumerate(fam):
        parent[kids[kids >= 0]] = p
    assert (parent >= 0).all()
    return parent


def _split(idx: str) -> tuple[int, int]:
    r, i = idx.split(":")
    return int(r), int(i)


def _make(res: int, i: int) -> str:
    return f"{res}:{int(i)}"


def get_res0_indexes() -> set:
    return {_make(0, i) for i in range(_H.num_cells(0))}


def h3_get_resolution(idx: str) -> int:
    return _split(idx)[0]


def h3_to_geo(idx: str) -> tuple[float, float]:
    """(lat, lng) in DEGREES, as h3-py returns them (h3api radsToDegs)."""
    res, i = _split(idx)
    lat, lon = _level(res)[0][i]
    return float(lat * _H_180_PI), float(lon * _H_180_PI)


_H_180_PI = 57.29577951308232087679815481410517033240547


def h3_to_children(idx: str, res: int | None = None) -> set:
    r, i = _split(idx)
    res = r + 1 if res is None else res
    cells = np.array([i], dtype=np.int64)
    for level in range(r, res):
        fam = _children_of(level)[cells].reshape(-1)
        cells = fam[fam >= 0]
    return {_make(res, c) for c in cells}


def h3_to_center_child(idx: str, res: int | None = None) -> str:
    r, i = _split(idx)
    res = r + 1 if res is None else res
    for level in range(r, res):
        i = int(_children_of(level)[i, 0])
    return _make(res, i)


def h3_to_parent(idx: str, res: int | None = None) -> str:
    r, i = _split(idx)
    res = r - 1 if res is None else res
    for level in range(r, res, -1):
        i = int(_parent_of(level)[i])
    return _make(res, i)


def k_ring(idx: str, k: int = 1) -> set:
    res, i = _split(idx)
    nb = _level(res)[2]
    seen = {i}
    frontier = [i]
    for _ in range(k):
        nxt = []
        for w in frontier:
            for x in nb[w]:
                x = int(x)
                if x >= 0 and x not in seen:
                    seen.add(x)
                    nxt.append(x)
        frontier = nxt
    return {_make(res, c) for c in seen}


def uncompact(cells, res: int) -> set:
    out = set()
    for idx in cells:
        r, _ = _split(idx)
        if r > res:
            raise ValueError("uncompact: a cell is finer than the target resolution")
        out |= h3_to_children(idx, res) if r < res else {idx}
    return out


def compact(cells) -> set:
    """h3 ``compact``: replace every complete set of siblings by its parent until nothing changes."""
    current = set(cells)
    while True:
        by_res: dict[int, set] = {}
        for idx in current:
            r, i = _split(idx)
            by_res.setdefault(r, set()).add(i)
        changed = False
        for res in sorted(by_res, reverse=True):
            if res == 0:
                continue
            here = by_res[res]
            parent = _parent_of(res)
            fam = _children_of(res - 1)
            for p in {int(parent[i]) for i in here}:
                kids = [int(c) for c in fam[p] if c >= 0]
                if all(c in here for c in kids):
                    for c in kids:
                        current.discard(_make(res, c))
                    current.add(_make(res - 1, p))
                    changed = True
            if changed:
                break
        if not changed:
            return current
