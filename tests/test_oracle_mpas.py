"""The MPAS part of the oracle (oracle/orc_mpas.c) against things that do not depend on it: brute-force nearest nodes, an
independent literal Python transliteration of search_for_cells with the module's queue and dictionary (remapper.F:1124-1326),
hand-derivable Wachspress values, numpy statements of remap_field and of process_mpas_fields' per-level extraction."""
import collections

import numpy as np
import pytest

from wps_b200 import synth


def target(nlat=40, nlon=50, lat0=20.0, lat1=55.0, lon0=-130.0, lon1=-60.0):
    lat = np.deg2rad(np.linspace(lat0, lat1, nlat, dtype=np.float64))[:, None] + 0.01 * np.sin(np.arange(nlon))[None, :]
    lon = np.deg2rad(np.linspace(lon0, lon1, nlon, dtype=np.float64))[None, :] + 0.01 * np.cos(np.arange(nlat))[:, None]
    return np.ascontiguousarray(lat, dtype=np.float32), np.ascontiguousarray(lon, dtype=np.float32)


def sphere_distance64(lat1, lon1, lat2, lon2):
    return 2.0 * np.arcsin(np.sqrt(np.sin(0.5 * (lat2 - lat1)) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin(0.5 * (lon2 - lon1)) ** 2))


@pytest.fixture(scope="module")
def mesh():
    return synth.mpas_mesh(2562, seed=1)


@pytest.fixture(scope="module")
def tables(orc, mesh):
    lat, lon = target()
    return lat, lon, orc.mpas_remap_setup(mesh, lat, lon)


def test_synthetic_mesh_is_a_consistent_voronoi_mesh(mesh):
    nc, me = mesh["nCells"], mesh["maxEdges"]
    assert mesh["nEdges"] == 3 * (nc - 2) and mesh["nVertices"] == 2 * (nc - 2)          # Euler, as mpas_mesh.F:57 assumes
    coc, voc, eoc = mesh["cellsOnCell"], mesh["verticesOnCell"], mesh["edgesOnCell"]
    for c in range(0, nc, 37):
        n = mesh["nEdgesOnCell"][c]
        assert (coc[c, n:] == 0).all() and (coc[c, :n] >= 1).all()
        for k in range(n):
            nb = coc[c, k] - 1
            assert c + 1 in coc[nb, :mesh["nEdgesOnCell"][nb]]                              # neighbour relation is symmetric
            assert set(mesh["cellsOnEdge"][eoc[c, k] - 1]) == {c + 1, nb + 1}
            assert c + 1 in mesh["cellsOnVertex"][voc[c, k] - 1]


def test_nearest_cell_vertex_edge_equal_brute_force(mesh, tables):
    """The greedy walks end at the globally nearest cell; the vertex / edge walks end at the nearest vertex / edge of it."""
    lat, lon, t = tables
    la, lo = lat.ravel().astype(np.float64)[:, None], lon.ravel().astype(np.float64)[:, None]
    d = sphere_distance64(mesh["latCell"][None, :].astype(np.float64), mesh["lonCell"][None, :].astype(np.float64), la, lo)
    assert (d.argmin(1) + 1 == t["nearestCell"].ravel()).mean() > 0.999
    # the vertex / edge walks end at a FIXED POINT of the rule (:851-895): no vertex / edge of the nearest of the node's own
    # cells is closer.  For vertices that is the vertex of the nearest cell closest to the target (> 99.9 % of the points); for
    # edges 3 % of the points have a second fixed point on a neighbouring cell, so the result depends on where the walk came from.
    nc = t["nearestCell"].ravel() - 1
    nn_all = mesh["nEdgesOnCell"]
    for where, key, conn, con2 in (("Vertex", "nearestVertex", "verticesOnCell", "cellsOnVertex"), ("Edge", "nearestEdge", "edgesOnCell", "cellsOnEdge")):
        latn, lonn = mesh["lat" + where].astype(np.float64), mesh["lon" + where].astype(np.float64)
        res = t[key].ravel() - 1
        cells = mesh[con2][res] - 1                                        # [npts][degree]
        dc = sphere_distance64(mesh["latCell"].astype(np.float64)[cells], mesh["lonCell"].astype(np.float64)[cells], la, lo)
        ic = cells[np.arange(len(res)), dc.argmin(1)]
        nodes = mesh[conn][ic]
        d = sphere_distance64(latn[np.maximum(nodes, 1) - 1], lonn[np.maximum(nodes, 1) - 1], la, lo)
        d[np.arange(mesh["maxEdges"])[None, :] >= nn_all[ic][:, None]] = np.inf
        dres = sphere_distance64(latn[res], lonn[res], la[:, 0], lo[:, 0])
        assert (d.min(1) >= dres - 1e-7).all(), where
        nodes = mesh[conn][nc]
        d = sphere_distance64(latn[np.maximum(nodes, 1) - 1], lonn[np.maximum(nodes, 1) - 1], la, lo)
        d[np.arange(mesh["maxEdges"])[None, :] >= nn_all[nc][:, None]] = np.inf
        on_nearest_cell = (nodes[np.arange(len(nc)), d.argmin(1)] - 1 == res).mean()
        assert on_nearest_cell > (0.999 if where == "Vertex" else 0.95), (where, on_nearest_cell)


def test_wachspress_weights_known_answers(orc, mesh, tables):
    lat, lon, t = tables
    w = t["cellWeights"]
    assert np.abs(w.sum(-1) - 1.0).max() < 1e-5
    # sourceCells are the three cells of the nearest vertex, in cellsOnVertex order
    assert (t["sourceCells"] == mesh["cellsOnVertex"][t["nearestVertex"] - 1]).all()
    # at a cell centre the weight of that cell is 1; at the centroid of a small triangle every weight is 1/3
    import ctypes as C
    L = orc.lib()

    def lx(la, lo, r=6371229.0):
        return np.array([r * np.cos(lo) * np.cos(la), r * np.sin(lo) * np.cos(la), r * np.sin(la)], np.float32)
    cs = mesh["cellsOnVertex"][100] - 1
    vc = np.ascontiguousarray(np.stack([lx(mesh["latCell"][c].astype(np.float64), mesh["lonCell"][c].astype(np.float64)) for c in cs]))
    out = np.zeros(3, np.float32)
    for k in range(3):
        p = vc[k].copy()
        L.orc_mpas_wachspress3(C.c_void_p(vc.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data))
        assert abs(out[k] - 1.0) < 1e-5 and np.abs(np.delete(out, k)).max() < 1e-5
    p = vc.astype(np.float64).mean(0)
    p = (p / np.linalg.norm(p) * 6371229.0).astype(np.float32)
    L.orc_mpas_wachspress3(C.c_void_p(vc.ctypes.data), C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data))
    # spherical areas of the three sub-triangles of the centroid agree to second order in the triangle size
    assert np.abs(out - 1.0 / 3.0).max() < 2e-2 and abs(out.sum() - 1.0) < 1e-6
    # vertex / edge tables: three Wachspress weights (the reference calls the routine with nVertices = 3), zeros beyond
    for key, src in (("vertexWeights", "sourceVertices"), ("edgeWeights", "sourceEdges")):
        assert (t[key][..., 3:] == 0.0).all() and np.abs(t[key][..., :3].sum(-1) - 1.0).max() < 1e-4
        nn = mesh["nEdgesOnCell"][t["nearestCell"] - 1]
        conn = mesh["verticesOnCell" if "V" in src else "edgesOnCell"][t["nearestCell"] - 1]
        k = np.arange(mesh["maxEdges"])[None, None, :]
        assert (t[src] == np.where(k < nn[..., None], conn, 1)).all()


def py_search_for_cells(mesh, tlat, tlon, start):
    """remapper.F:1124-1210, written from the Fortran independently of orc_mpas.c: ring-buffer queue of 2700 whose inserts are
    dropped when full, dictionary as a Python set bounded at 82000 * 32 cells, binary32 distances."""
    f = np.float32

    def dist(la1, lo1, la2, lo2):
        s1 = f(np.sin(np.float64(f(0.5) * f(la2 - la1))))
        s2 = f(np.sin(np.float64(f(0.5) * f(lo2 - lo1))))
        c1, c2 = f(np.cos(np.float64(la1))), f(np.cos(np.float64(la2)))
        arg = f(np.sqrt(f(f(s1 * s1) + f(f(c1 * c2) * f(s2 * s2)))))
        return f(f(2.0) * f(np.arcsin(np.float64(arg))))
    q = collections.deque()
    seen = set()
    cells, w = [1, 1, 1], [0.0, 0.0, 0.0]

    def insert(i):
        if len(q) < 2700:
            q.append(i)
        if (i - 1) // 32 + 1 <= 82000:
            seen.add(i)
    insert(start)
    no_more, unscanned, d_min = False, 0, f(1000.0)
    while q:
        scan = q.popleft()
        if mesh["landmask"][scan - 1] == 1:
            d = dist(tlat, tlon, mesh["latCell"][scan - 1], mesh["lonCell"][scan - 1])
            if d < d_min:
                d_min, cells[0], w[0] = d, scan, 1.0
                if not no_more:
                    unscanned = len(q)
                no_more = True
        if (not no_more) or unscanned > 0:
            for i in range(mesh["nEdgesOnCell"][scan - 1]):
                nb = int(mesh["cellsOnCell"][scan - 1, i])
                if nb not in seen:
                    insert(nb)
        if no_more:
            unscanned -= 1
    return cells, w


def test_masked_cells_and_search_for_cells(orc, mesh, tables):
    lat, lon, t = tables
    sc, w = t["sourceMaskedCells"], t["cellMaskedWeights"]
    # search_for_cells results are (cell, 1, 1) / (1, 0, 0); every other point keeps the three cells of a vertex
    is_search = (sc[..., 1] == 1) & (sc[..., 2] == 1)
    assert (w[is_search] == np.array([1.0, 0.0, 0.0], np.float32)).all()
    # The masked scan is a walk sequence of its own (after a search point idx holds a cell index, :272), so its vertex can differ
    # from nearestVertex at start-sensitive points; where it does not, the weights are cellWeights with water cells zeroed and
    # the rest renormalised (:264-270)
    same = ~is_search & (sc == t["sourceCells"]).all(-1)
    assert same.sum() > 0.95 * (~is_search).sum()
    land = mesh["landmask"][t["sourceCells"] - 1] == 1
    wl = np.where(land, t["cellWeights"], np.float32(0.0))
    ssum = ((wl[..., 0] + wl[..., 1]) + wl[..., 2]).astype(np.float32)
    assert (ssum[same] > 0).all()
    assert np.array_equal(w[same], (wl[same] / ssum[same][:, None]).astype(np.float32))
    assert (w[same][~land[same]] == 0.0).all()
    rest = np.argwhere(is_search)
    assert len(rest) > 50
    orc.set_transc_mode(1)
    for iy, ix in rest[:: max(1, len(rest) // 60)]:
        start = orc.lib().orc_mpas_nearest_cell  # noqa: F841  (the start is the nearest cell: brute force below)
        d = sphere_distance64(mesh["latCell"].astype(np.float64), mesh["lonCell"].astype(np.float64), np.float64(lat[iy, ix]), np.float64(lon[iy, ix]))
        cells, ww = py_search_for_cells(mesh, lat[iy, ix], lon[iy, ix], int(d.argmin()) + 1)
        assert list(sc[iy, ix]) == cells and list(w[iy, ix]) == ww, (iy, ix)
        assert mesh["landmask"][sc[iy, ix, 0] - 1] == 1


def test_search_for_cells_single_land_cell(orc):
    """One land cell on the whole sphere: every search must return it, whatever the start."""
    import ctypes as C
    m = synth.mpas_mesh(162, seed=3)
    m["landmask"][:] = 0
    m["landmask"][77] = 1
    d, keep = orc._mpas_mesh(m)
    cells = (C.c_int * 3)()
    w = (C.c_float * 3)()
    for start in (1, 40, 78, 162):
        orc.lib().orc_mpas_search_for_cells(C.byref(d), C.c_float(0.3), C.c_float(1.0), start, cells, w)
        assert list(cells) == [78, 1, 1] and list(w) == [1.0, 0.0, 0.0]


def test_remap_field_and_plane_extraction_are_the_stated_formulas(orc, mesh, tables):
    lat, lon, t = tables
    f = synth.mpas_field(mesh, 6, seed=4)
    out = orc.mpas_remap_field(t["sourceCells"], t["cellWeights"], f)
    s, w = t["sourceCells"] - 1, t["cellWeights"]
    ref = np.zeros_like(out)
    for j in range(3):                      # dst = dst + w(j) * src(:, node(j)), binary32, in this order (remapper.F:533-539)
        ref = (ref + (w[..., j:j + 1] * f[s[..., j]]).astype(np.float32)).astype(np.float32)
    assert np.array_equal(out, ref)
    f64 = f.astype(np.float64) * 1.000000123
    out64 = orc.mpas_remap_field(t["sourceCells"], t["cellWeights"], f64)
    ref64 = np.zeros_like(out64)
    for j in range(3):
        ref64 = ref64 + w[..., j:j + 1].astype(np.float64) * f64[s[..., j]]
    assert np.array_equal(out64, ref64)
    fi = (f * 10).astype(np.int32)
    assert np.array_equal(orc.mpas_remap_field_int(t["nearestCell"], fi), fi[t["nearestCell"] - 1])
    # 1-d source: every table entry takes part (:506)
    v1 = synth.mpas_field(mesh, 1, seed=5, where="Vertex")[:, 0]
    o1 = orc.mpas_remap_field(t["sourceVertices"], t["vertexWeights"], v1)
    r1 = np.zeros(o1.shape[:2], np.float32)
    for j in range(mesh["maxEdges"]):
        r1 = (r1 + (t["vertexWeights"][..., j] * v1[t["sourceVertices"][..., j] - 1]).astype(np.float32)).astype(np.float32)
    assert np.array_equal(o1[..., 0], r1)
    # planes: mode 0 = transpose; mode 1 = surface level then interface means (process_domain_module.F:1733-1776); landmask fill
    lm = (np.arange(out.shape[0] * out.shape[1]).reshape(out.shape[:2]) % 3 != 0).astype(np.float32)
    p0 = orc.mpas_extract_planes(out, 0)
    assert np.array_equal(p0, np.moveaxis(out, 2, 0))
    p1 = orc.mpas_extract_planes(out, 1, landmask=lm, fill=-7.0)
    ref1 = np.moveaxis(out, 2, 0).copy()
    ref1[1:] = (np.float32(0.5) * (np.moveaxis(out, 2, 0)[:-1] + np.moveaxis(out, 2, 0)[1:])).astype(np.float32)
    ref1[:, lm == 0] = -7.0
    assert np.array_equal(p1, ref1)
