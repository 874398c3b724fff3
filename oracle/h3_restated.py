"""ORACLE - TEST INFRASTRUCTURE ONLY (see ``oracle/ref_path.py`` for who may import this).

CPU restatement (numpy float64) of the part of the H3 C library (``h3>=3.7.6,<4``, pinned in
/root/reference/pyproject.toml:45; NOT under /root/reference, NOT installed, no network) that the
reference's hexagonal hidden mesh needs (generate/hex_icosahedron.py:20-236):

* ``h3.uncompact(h3.get_res0_indexes(), res)`` + ``h3.h3_to_geo``  -> ``cell_centers(res)``
* ``h3.k_ring(idx, k)``                                            -> ``grid_disk`` over ``neighbours``
* ``h3.h3_to_center_child(idx, res)``                              -> the finer cell with the same centre
* ``h3.compact`` / ``h3.uncompact`` of a node set                  -> ``complete_cells``

H3's published geometry (h3lib/lib/faceijk.c, coordijk.c, geoCoord.c of H3 3.7): the sphere is cut
into the 20 faces of an icosahedron in a fixed orientation (face centres ``FACE_CENTER_GEO``, azimuth
of each face's Class II i-axis ``FACE_AXES_AZ_CII``); on every face, cells of resolution ``r`` are the
points of a hexagonal lattice in the face's gnomonic projection, unit length
``RES0_U_GNOMONIC / sqrt(7)**r``, rotated by ``asin(sqrt(3/28))`` counter-clockwise at odd
("Class III") resolutions; the 12 icosahedron vertices are pentagon centres at every resolution.
A cell centre is ``_hex2dToGeo`` of its lattice point: inverse gnomonic scaling of the radius, the
azimuth from the face centre, then ``_geoAzDistanceRads``.

Parity status: **UNPINNED against h3 itself** (the library cannot be run here).  The restatement is
checked against everything that CAN be checked without it (``self_check`` / tests/test_oracle_hex.py):
the 20 face centres and 60 axis azimuths, recalled from the H3 source, form a regular icosahedron to
1e-15 and every axis hits an icosahedron vertex (a mis-remembered digit would break this); cell counts
equal H3's 2 + 120 * 7**r; every cell has 6 neighbours, the 12 pentagons 5; and the two cell centres
H3's documentation publishes (``85283473fffffff`` -> 37.3457933754, -121.9763759726 and
``8928308280fffff`` -> 37.77670234943567, -122.41845932318311) are reproduced.  H3 cell *indices* are
not modelled: the reference never lets them reach the graph (nodes are re-ordered by
``get_coordinates_ordering``, edges are adjacency positions).  Which face evaluates a centre lying
exactly on a face edge follows the lowest face number here and H3's base-cell home face there; the
two evaluations agree to ~1e-16 rad.
"""

from __future__ import annotations

import numpy as np

M_SQRT7 = 2.6457513110645905905016157536392604257102
M_SQRT3_2 = 0.8660254037844386467637231707529361834714
M_AP7_ROT_RADS = 0.333473172251832115336090755351601070065900389  # asin(sqrt(3/28))
RES0_U_GNOMONIC = 0.38196601125010500003
EPSILON = 0.0000000000000001

# faceijk.c: faceCenterGeo (lat, lon radians)
FACE_CENTER_GEO = np.array(
    [
        [0.803582649718989942, 1.248397419617396099],
        [1.307747883455638156, 2.536945009877921159],
        [1.054751253523952054, -1.347517358900396623],
        [0.600191595538186799, -0.450603909469755746],
        [0.491715428198773866, 0.401988202911306943],
        [0.172745327415618701, 1.678146885280433686],
        [0.605929321571350690, 2.953923329812411617],
        [0.427370518328979641, -1.888876200336285401],
        [-0.079066118549212831, -0.733429513380867741],
        [-0.230961644455383637, 0.506495587332349035],
        [0.079066118549212831, 2.408163140208925497],
        [0.230961644455383637, -2.635097066257444203],
        [-0.172745327415618701, -1.463445768309359553],
        [-0.605929321571350690, -0.187669323777381622],
        [-0.427370518328979641, 1.252716453253507838],
        [-0.600191595538186799, 2.690988744120037492],
        [-0.491715428198773866, -2.739604450678486295],
        [-0.803582649718989942, -1.893195233972397139],
        [-1.307747883455638156, -0.604647643711872080],
        [-1.054751253523952054, 1.794075294689396615],
    ]
)

# faceijk.c: faceAxesAzRadsCII - azimuth (clockwise from north) of the Class II i, j, k axes at each face centre
FACE_AXES_AZ_CII = np.array(
    [
        [5.619958268523939882, 3.525563166130744542, 1.431168063737548730],
        [5.760339081714187279, 3.665943979320991689, 1.571548876927796127],
        [0.780213654393430055, 4.969003859179821079, 2.874608756786625655],
        [0.430469363979999913, 4.619259568766391033, 2.524864466373195467],
        [6.130269123335111400, 4.035874020941915804, 1.941478918548720291],
        [2.692877706530642877, 0.598482604137447119, 4.787272808923838195],
        [2.982963003477243874, 0.888567901084048369, 5.077358105870439581],
        [3.532912002790141181, 1.438516900396945656, 5.627307105183336758],
        [3.494305004259568154, 1.399909901866372864, 5.588700106652763840],
        [3.003214169499538391, 0.908819067106342928, 5.097609271892733906],
        [5.930472956509811562, 3.836077854116615875, 1.741682751723420374],
        [0.138378484090254847, 4.327168688876645809, 2.232773586483450311],
        [0.448714947059150361, 4.637505151845541521, 2.543110049452346120],
        [0.158629650112549365, 4.347419854898940135, 2.253024752505744869],
        [5.891865957979238535, 3.797470855586042958, 1.703075753192847583],
        [2.711123289609793325, 0.616728187216597771, 4.805518392002988683],
        [3.294508837434268316, 1.200113735041072948, 5.388903939827463911],
        [3.804819692245439833, 1.710424589852244509, 5.899214794638635174],
        [3.664438879055192436, 1.570043776661997111, 5.758833981448388027],
        [2.361378999196363184, 0.266983896803167583, 4.455774101589558636],
    ]
)

# the two cell centres H3's documentation publishes (degrees): (resolution, lat, lon)
PUBLISHED_CENTERS = [
    (5, 37.34579337536848, -121.97637597255124),  # h3ToGeo 85283473fffffff
    (9, 37.77670234943567, -122.41845932318311),  # h3_to_geo('8928308280fffff'), h3-py README
]


def num_cells(res: int) -> int:
    """H3's ``numHexagons(res)``: 2 + 120 * 7**res."""
    return 2 + 120 * 7**res


def _pos_angle(a):
    """geoCoord.c ``_posAngleRads``."""
    a = np.where(a < 0.0, a + 2.0 * np.pi, a)
    return np.where(a >= 2.0 * np.pi, a - 2.0 * np.pi, a)


def _constrain_lng(lng):
    """geoCoord.c ``constrainLng``: into (-pi, pi]."""
    lng = np.where(lng > np.pi, lng - 2.0 * np.pi, lng)
    return np.where(lng < -np.pi, lng + 2.0 * np.pi, lng)


def geo_az_distance(lat1, lon1, az, distance):
    """geoCoord.c ``_geoAzDistanceRads``: the point at ``distance`` radians along azimuth ``az`` from p1."""
    lat1, lon1, az, distance = np.broadcast_arrays(
        *(np.asarray(v, dtype=np.float64) for v in (lat1, lon1, az, distance))
    )
    az = _pos_angle(az)
    with np.errstate(invalid="ignore", divide="ignore"):
        sinlat = np.clip(np.sin(lat1) * np.cos(distance) + np.cos(lat1) * np.sin(distance) * np.cos(az), -1.0, 1.0)
        lat2 = np.arcsin(sinlat)
        sinlon = np.clip(np.sin(az) * np.sin(distance) / np.cos(lat2), -1.0, 1.0)
        coslon = np.clip((np.cos(distance) - np.sin(lat1) * np.sin(lat2)) / np.cos(lat1) / np.cos(lat2), -1.0, 1.0)
        lon2 = _constrain_lng(lon1 + np.arctan2(sinlon, coslon))
    north = np.abs(lat2 - 0.5 * np.pi) < EPSILON
    south = np.abs(lat2 + 0.5 * np.pi) < EPSILON
    lat2 = np.where(north, 0.5 * np.pi, np.where(south, -0.5 * np.pi, lat2))
    lon2 = np.where(north | south, 0.0, lon2)
    # due north / due south branch
    due_n, due_s = az < EPSILON, np.abs(az - np.pi) < EPSILON
    if np.any(due_n | due_s):
        lat_ns = np.where(due_n, lat1 + distance, lat1 - distance)
        pole_n, pole_s = np.abs(lat_ns - 0.5 * np.pi) < EPSILON, np.abs(lat_ns + 0.5 * np.pi) < EPSILON
        lat_ns = np.where(pole_n, 0.5 * np.pi, np.where(pole_s, -0.5 * np.pi, lat_ns))
        lon_ns = np.where(pole_n | pole_s, 0.0, _constrain_lng(lon1))
        lat2 = np.where(due_n | due_s, lat_ns, lat2)
        lon2 = np.where(due_n | due_s, lon_ns, lon2)
    same = distance < EPSILON
    return np.where(same, lat1, lat2), np.where(same, lon1, lon2)


def hex2d_to_geo(x, y, face, res: int):
    """faceijk.c ``_hex2dToGeo`` (substrate = 0): lattice-plane point of ``face`` at ``res`` -> (lat, lon) radians."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    face = np.asarray(face)
    r = np.sqrt(x * x + y * y)
    theta = np.arctan2(y, x)
    for _ in range(res):
        r = r / M_SQRT7
    r = r * RES0_U_GNOMONIC
    r = np.arctan(r)
    if res % 2 == 1:
        theta = _pos_angle(theta + M_AP7_ROT_RADS)
    theta = _pos_angle(FACE_AXES_AZ_CII[face, 0] - theta)
    lat, lon = geo_az_distance(FACE_CENTER_GEO[face, 0], FACE_CENTER_GEO[face, 1], theta, r)
    centre = np.sqrt(x * x + y * y) < EPSILON
    return np.where(centre, FACE_CENTER_GEO[face, 0], lat), np.where(centre, FACE_CENTER_GEO[face, 1], lon)


def axial_to_hex2d(a, b):
    """coordijk.c ``_ijkToHex2d`` with k = 0: (i, j) -> (x, y)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a - 0.5 * b, b * M_SQRT3_2


def face_vertices_axial(res: int) -> np.ndarray:
    """Lattice coordinates (i, j with k = 0) of the three icosahedron vertices of a face at ``res``: res 0 has
    them two units out on the i, j, k axes (faceijk.c base cell table: pentagons sit at {2,0,0} of their home
    face); each finer level applies coordijk.c ``_downAp7`` (odd = Class III, counter-clockwise:
    i -> (3,0,1), j -> (1,3,0)) or ``_downAp7r`` (even, clockwise: i -> (3,1,0), j -> (0,3,1))."""
    v = np.array([[2, 0], [0, 2], [-2, -2]], dtype=np.int64)
    for level in range(1, res + 1):
        if level % 2 == 1:
            iv, jv = np.array([2, -1]), np.array([1, 3])
        else:
            iv, jv = np.array([3, 1]), np.array([-1, 2])
        v = v[:, :1] * iv[None, :] + v[:, 1:] * jv[None, :]
    return v


def _xyz(lat, lon):
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def icosahedron_tables():
    """(vertex id of each face corner (20, 3), owner face of each corner (20, 3), owner face of each edge (20, 3)):
    corner k of a face is where its k-th Class II axis meets the icosahedron vertex; edge k joins corners k and
    k + 1.  The owner of a shared lattice point is the lowest-numbered face that contains it."""
    dist = np.arctan(2.0 * RES0_U_GNOMONIC)
    f = np.repeat(np.arange(20), 3)
    lat, lon = geo_az_distance(FACE_CENTER_GEO[f, 0], FACE_CENTER_GEO[f, 1], FACE_AXES_AZ_CII.reshape(-1), dist)
    p = _xyz(lat, lon)
    ids = -np.ones(60, dtype=np.int64)
    n = 0
    for i in range(60):
        if ids[i] < 0:
            ids[np.linalg.norm(p - p[i], axis=1) < 1e-9] = n
            n += 1
    assert n == 12, n
    corner = ids.reshape(20, 3)
    corner_owner = np.zeros((20, 3), dtype=np.int64)
    edge_owner = np.zeros((20, 3), dtype=np.int64)
    for face in range(20):
        for k in range(3):
            corner_owner[face, k] = min(g for g in range(20) if corner[face, k] in corner[g])
            pair = {corner[face, k], corner[face, (k + 1) % 3]}
            edge_owner[face, k] = min(g for g in range(20) if pair <= set(corner[g]))
    return corner, corner_owner, edge_owner


def self_check() -> dict:
    """Geometric consistency of the recalled constant tables (no h3 needed)."""
    c = _xyz(FACE_CENTER_GEO[:, 0], FACE_CENTER_GEO[:, 1])
    gram = c @ c.T
    # face centres of a regular icosahedron: 3 nearest neighbours at cos = sqrt(5)/3, the antipode at -1
    s = np.sort(gram, axis=1)
    err_centres = max(np.abs(s[:, -4:-1] - np.sqrt(5.0) / 3.0).max(), np.abs(s[:, 0] + 1.0).max())
    dist = np.arctan(2.0 * RES0_U_GNOMONIC)
    f = np.repeat(np.arange(20), 3)
    lat, lon = geo_az_distance(FACE_CENTER_GEO[f, 0], FACE_CENTER_GEO[f, 1], FACE_AXES_AZ_CII.reshape(-1), dist)
    p = _xyz(lat, lon)
    corner, _, _ = icosahedron_tables()
    verts = np.stack([p[corner.reshape(-1) == v].mean(axis=0) for v in range(12)])
    err_axes = np.abs(p - verts[corner.reshape(-1)]).max()
    vg = np.sort(verts @ verts.T, axis=1)
    err_vertices = max(np.abs(vg[:, 1:6] + 1.0 / np.sqrt(5.0)).max(), np.abs(vg[:, 6:11] - 1.0 / np.sqrt(5.0)).max())
    counts = np.bincount(corner.reshape(-1), minlength=12)
    return {
        "err_centres": float(err_centres),
        "err_axes": float(err_axes),
        "err_vertices": float(err_vertices),
        "faces_per_vertex": counts.tolist(),
        "gnomonic_unit": float(np.tan(np.arccos(np.sqrt((5.0 + 2.0 * np.sqrt(5.0)) / 15.0))) / 2.0 - RES0_U_GNOMONIC),
    }


def face_cells(res: int, face: int, corner_owner, edge_owner):
    """Lattice points (i, j) of ``face`` at ``res`` that the face owns, in (i, j) row-major order, and a pentagon flag."""
    v = face_vertices_axial(res)
    lo, hi = v.min(axis=0), v.max(axis=0)
    a, b = np.meshgrid(np.arange(lo[0], hi[0] + 1), np.arange(lo[1], hi[1] + 1), indexing="ij")
    a, b = a.reshape(-1), b.reshape(-1)
    cross = []
    for k in range(3):
        e = v[(k + 1) % 3] - v[k]
        cross.append(e[0] * (b - v[k][1]) - e[1] * (a - v[k][0]))
    cross = np.stack(cross)  # >= 0 inside (corners are counter-clockwise: i, j, k axes 120 degrees apart)
    inside = np.all(cross >= 0, axis=0)
    keep = inside.copy()
    pent = np.zeros_like(inside)
    for k in range(3):
        on_edge = inside & (cross[k] == 0)
        at_corner = inside & (cross[k] == 0) & (cross[(k + 2) % 3] == 0)  # corner k = edges k and k-1
        keep &= ~(on_edge & ~at_corner & (edge_owner[face, k] != face))
        keep &= ~(at_corner & (corner_owner[face, k] != face))
        pent |= at_corner
    # a corner lies on two edges: the per-edge rule above must not have dropped a corner its face owns
    for k in range(3):
        at_corner = inside & (cross[k] == 0) & (cross[(k + 2) % 3] == 0)
        keep = np.where(at_corner, corner_owner[face, k] == face, keep)
    # points on an edge but not at a corner: re-apply (the corner fix above only touches corners)
    return a[keep], b[keep], pent[keep]


def cell_centers(res: int):
    """All H3 cell centres of resolution ``res``: (N, 2) float64 (lat, lon) radians, N = 2 + 120 * 7**res, in
    (face, i, j) order, and the pentagon flags.  ``h3.uncompact(h3.get_res0_indexes(), res)`` +
    ``h3.h3_to_geo`` (generate/hex_icosahedron.py:47,99) up to the (arbitrary, set-iteration) order."""
    _, corner_owner, edge_owner = icosahedron_tables()
    lats, lons, pents = [], [], []
    for face in range(20):
        a, b, pent = face_cells(res, face, corner_owner, edge_owner)
        x, y = axial_to_hex2d(a, b)
        lat, lon = hex2d_to_geo(x, y, np.full(a.shape, face), res)
        lats.append(lat)
        lons.append(lon)
        pents.append(pent)
    out = np.stack([np.concatenate(lats), np.concatenate(lons)], axis=1)
    assert out.shape[0] == num_cells(res), (out.shape, num_cells(res))
    return out, np.concatenate(pents)


def hex_nodes_latlon(res: int) -> np.ndarray:
    """What the reference feeds to ``get_coordinates_ordering``: ``np.deg2rad(h3.h3_to_geo(...))`` - H3 returns
    degrees (``radsToDegs``: x * 180/pi), numpy converts back (generate/hex_icosahedron.py:47)."""
    c, _ = cell_centers(res)
    return np.deg2rad(c * (180.0 / np.pi))


def nearest_center(lat_deg: float, lon_deg: float, res: int):
    """``h3_to_geo(geo_to_h3(lat, lon, res))`` in degrees: project on the nearest face (faceijk.c ``_geoToHex2d``),
    round to the nearest lattice point, evaluate its centre."""
    lat, lon = np.deg2rad(lat_deg), np.deg2rad(lon_deg)
    p = _xyz(np.array(lat), np.array(lon))
    face = int(np.argmax(_xyz(FACE_CENTER_GEO[:, 0], FACE_CENTER_GEO[:, 1]) @ p))
    clat, clon = FACE_CENTER_GEO[face]
    r = np.arccos(np.clip(np.sin(clat) * np.sin(lat) + np.cos(clat) * np.cos(lat) * np.cos(lon - clon), -1, 1))
    az = np.arctan2(
        np.cos(lat) * np.sin(lon - clon), np.cos(clat) * np.sin(lat) - np.sin(clat) * np.cos(lat) * np.cos(lon - clon)
    )
    theta = _pos_angle(FACE_AXES_AZ_CII[face, 0] - _pos_angle(az))
    if res % 2 == 1:
        theta = _pos_angle(theta - M_AP7_ROT_RADS)
    r = np.tan(r) / RES0_U_GNOMONIC * M_SQRT7**res
    x, y = r * np.cos(theta), r * np.sin(theta)
    b0 = y / M_SQRT3_2
    a0 = x + 0.5 * b0
    best = None
    for a in (np.floor(a0), np.floor(a0) + 1):
        for b in (np.floor(b0), np.floor(b0) + 1):
            cx, cy = axial_to_hex2d(a, b)
            d = (cx - x) ** 2 + (cy - y) ** 2
            if best is None or d < best[0]:
                best = (d, a, b)
    cx, cy = axial_to_hex2d(best[1], best[2])
    la, lo = hex2d_to_geo(cx, cy, face, res)
    return float(np.rad2deg(la)), float(np.rad2deg(lo))


def neighbours(centers: np.ndarray, pentagon: np.ndarray) -> np.ndarray:
    """(N, 6) int64 table of the cells sharing an edge with each cell (-1 in the 6th slot of a pentagon):
    on an aperture-7 hexagonal grid these are the 6 (5) nearest centres - the second ring is sqrt(3) times
    farther, the gnomonic distortion inside an icosahedron face is at most 1.26."""
    from scipy.spatial import cKDTree

    p = _xyz(centers[:, 0], centers[:, 1])
    _, idx = cKDTree(p).query(p, k=7)
    nb = idx[:, 1:].astype(np.int64)
    assert np.all(idx[:, 0] == np.arange(len(p)))
    nb[pentagon, 5] = -1
    return nb


def grid_disk(nb: np.ndarray, k: int) -> list[np.ndarray]:
    """``h3.k_ring(idx, k)`` for every cell: all cells within ``k`` grid steps, the cell itself included."""
    n = nb.shape[0]
    out = []
    for u in range(n):
        seen = {u}
        frontier = [u]
        for _ in range(k):
            nxt = []
            for w in frontier:
                for x in nb[w]:
                    if x >= 0 and x not in seen:
                        seen.add(int(x))
                        nxt.append(int(x))
            frontier = nxt
        out.append(np.fromiter(seen, dtype=np.int64))
    return out


def center_child_positions(coarse: np.ndarray, fine: np.ndarray) -> np.ndarray:
    """``h3.h3_to_center_child(idx, res)``: index (into ``fine``) of the finer cell with the same centre."""
    from scipy.spatial import cKDTree

    d, idx = cKDTree(_xyz(fine[:, 0], fine[:, 1])).query(_xyz(coarse[:, 0], coarse[:, 1]), k=1)
    assert d.max() < 1e-9, d.max()
    return idx.astype(np.int64)


def complete_cells(level_centers: dict, level_pent: dict, res_max: int, in_graph: np.ndarray) -> dict:
    """``select_nodes_from_graph_at_resolution`` (generate/hex_icosahedron.py:206-210): a cell of level r < res_max
    survives ``h3.compact`` of the graph's nodes iff ALL its descendants at ``res_max`` are graph nodes; the
    children of a cell are its centre child and that child's neighbours (aperture 7).  Returns {level: bool mask}."""
    valid = {res_max: np.asarray(in_graph, dtype=bool)}
    for r in range(res_max - 1, -1, -1):
        fine, pent = level_centers[r + 1], level_pent[r + 1]
        nb = neighbours(fine, pent)
        cc = center_child_positions(level_centers[r], fine)
        fam = np.concatenate([cc[:, None], nb[cc]], axis=1)
        ok = np.where(fam >= 0, valid[r + 1][np.maximum(fam, 0)], True)
        valid[r] = ok.all(axis=1)
    return valid


def multiscale_edges_hex(resolutions, x_hops: int, node_ordering: np.ndarray, in_graph: np.ndarray | None = None):
    """``hex_icosahedron.add_edges_to_nx_graph`` (depth_children = 0) + ``nx.to_scipy_sparse_array`` + ``get_edge_index``
    (generate/hex_icosahedron.py:103-154, edges/builder.py:434-455, 69-87): canonical (2, E) int32.  ``node_ordering``
    lists the cells (positions in ``cell_centers(max(resolutions))``) that are graph nodes, in graph order."""
    res_max = max(resolutions)
    levels = range(res_max + 1) if in_graph is not None else sorted(set(resolutions))
    cen, pen = {}, {}
    for r in levels:
        cen[r], pen[r] = cell_centers(r)
    n_fine = cen[res_max].shape[0]
    rank = -np.ones(n_fine, dtype=np.int64)
    rank[np.asarray(node_ordering)] = np.arange(len(node_ordering))
    valid = None
    if in_graph is not None:
        valid = complete_cells(cen, pen, res_max, rank >= 0)
    edges = set()
    for r in sorted(set(resolutions)):
        nb = neighbours(cen[r], pen[r])
        cc = center_child_positions(cen[r], cen[res_max]) if r != res_max else np.arange(n_fine)
        ok = valid[r] if valid is not None else np.ones(cen[r].shape[0], dtype=bool)
        disks = grid_disk(nb, x_hops)
        for u in np.nonzero(ok)[0]:
            for v in disks[u]:
                if v != u and ok[v]:
                    s, t = rank[cc[u]], rank[cc[v]]
                    if s >= 0 and t >= 0 and s != t:
                        edges.add((int(s), int(t)))
                        edges.add((int(t), int(s)))
    e = np.array(sorted(edges, key=lambda st: (st[1], st[0])), dtype=np.int32).reshape(-1, 2).T
    return e
