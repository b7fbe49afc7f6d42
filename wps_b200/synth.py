"""Synthetic GFS/ERA5-shaped metgrid workloads: seeded, analytic, regenerable at any size.

This module only manufactures INPUTS (source slabs, masks, destination grids); every operator it calls goes
through libwpsgpu.so.  The field list follows ungrib/Variable_Tables/Vtable.GFS and the interp options are the
default METGRID.TBL.ARW entries (metgrid/METGRID.TBL.ARW)."""
import numpy as np

from . import capi

# name, n_levels, output stagger, METGRID.TBL.ARW entry (interp_option, missing_value, fill_missing, masked, interp_mask)
C3D = "sixteen_pt+four_pt+average_4pt"
SOIL = "sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search"


def gfs_slab_table(nlev3d=34):
    """One time of a GFS-shaped FILE: -> list of dict(field, nlev, stagger, chain, missing, fill, masked, mask, kind)."""
    t = []

    def add(field, nlev, stagger, chain, missing=1e20, fill=1e20, masked=capi.NOT_MASKED, mask=None, kind="smooth"):
        t.append(dict(field=field, nlev=nlev, stagger=stagger, chain=chain, missing=missing, fill=fill, masked=masked,
                      mask=mask, kind=kind))
    add("TT", nlev3d + 1, capi.M, C3D, fill=0.0)
    add("RH", nlev3d + 1, capi.M, C3D, fill=0.0)
    add("UU", nlev3d + 1, capi.U, C3D, fill=0.0, kind="wind")
    add("VV", nlev3d + 1, capi.V, C3D, fill=0.0, kind="wind")
    add("GHT", nlev3d, capi.M, C3D, fill=0.0)
    add("PSFC", 1, capi.M, "four_pt+average_4pt")
    add("PMSL", 1, capi.M, C3D)
    add("SOILHGT", 1, capi.M, "four_pt+average_4pt")
    add("LANDSEA", 1, capi.M, "nearest_neighbor", fill=-1.0, kind="landsea")
    add("SKINTEMP", 1, capi.M, SOIL, fill=0.0, masked=capi.MASKED_BOTH, mask=("both", 1.0, 0.0))
    add("SEAICE", 1, capi.M, "four_pt+average_4pt", missing=-1e30, fill=0.0, masked=capi.MASKED_LAND, mask=("mask", 1.0), kind="water")
    add("SNOW", 1, capi.M, "four_pt+average_4pt", fill=0.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    add("SNOWH", 1, capi.M, "four_pt+average_4pt", fill=0.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    for nm in ("ST000010", "ST010040", "ST040100", "ST100200"):
        add(nm, 1, capi.M, SOIL, missing=-1e30, fill=285.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    for nm in ("SM000010", "SM010040", "SM040100", "SM100200"):
        add(nm, 1, capi.M, SOIL, missing=-1e30, fill=1.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    return t


def n_slabs(table):
    return sum(e["nlev"] for e in table)


LAND_THRESHOLD = -1.2  # ~59 % land over the CONUS domain, coastlines with ~100-km structure


def landsea_fn(xp, lam, phi):
    """Synthetic land/sea pattern (xp = numpy or torch); land where the value exceeds LAND_THRESHOLD."""
    return xp.sin(5 * lam) + xp.cos(4 * phi) + 0.5 * xp.sin(19 * lam) * xp.cos(13 * phi)


def landsea_np(nx, ny):
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny)[:, None]
    return (landsea_fn(np, lam, phi) > LAND_THRESHOLD).astype(np.float32)


def slab_np(nx, ny, seed, kind, landsea=None, missing=1e20):
    """v = A + B sin(3 lam) cos(2 phi) + C lev + eps; exact zeros at 0.1 % of the cells."""
    rng = np.random.default_rng(seed)
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny)[:, None]
    if kind == "landsea":
        return landsea.copy()
    if kind == "wind":
        f = 12.0 * np.sin(3 * lam + 0.1 * seed) * np.cos(2 * phi) + rng.uniform(-0.012, 0.012, (ny, nx))
    else:
        f = 280.0 + 30.0 * np.sin(3 * lam) * np.cos(2 * phi) + 0.5 * (seed % 40) + rng.uniform(-0.03, 0.03, (ny, nx))
    f = f.astype(np.float32)
    f[rng.random((ny, nx)) < 0.001] = 0.0
    if kind == "land":
        f[landsea == 0] = missing
    elif kind == "water":
        f[landsea == 1] = missing
    return f


def gfs_step_numpy(config, seed0=20261019):
    """The slabs of one FILE time of `config` as numpy arrays -- the SAME table, masks and values for the device arm and the
    CPU reference arm of bench.py.  -> (landsea, [dict(e=table entry, slabs=[nlev][sy][sx + 6] halo'd float32)])"""
    sx, sy, _, _ = config["src"]
    key = (sx, sy, config["nlev"], seed0)
    if key in _STEP_CACHE:
        return _STEP_CACHE[key]
    ls = landsea_np(sx, sy)
    out, seed = [], seed0
    for e in gfs_slab_table(config["nlev"]):
        lev = [add_halo(slab_np(sx, sy, seed + k, e["kind"], ls, np.float32(e["missing"]))) for k in range(e["nlev"])]
        seed += e["nlev"]
        out.append(dict(e=e, slabs=np.stack(lev)))
    _STEP_CACHE.clear()          # one step at a time: 0.8 GB for config 2
    _STEP_CACHE[key] = (ls, out)
    return ls, out


_STEP_CACHE = {}


def add_halo(slab, bdr=3):
    """halo_slab of process_domain_module.F:1121-1135 (global lat-lon source); works for numpy and torch."""
    nx = slab.shape[-1]
    if isinstance(slab, np.ndarray):
        return np.ascontiguousarray(np.concatenate([slab[..., nx - bdr:], slab, slab[..., :bdr]], axis=-1))
    import torch
    return torch.cat([slab[..., nx - bdr:], slab, slab[..., :bdr]], dim=-1).contiguous()


def lambert_domain(nx, ny, dx, truelat1, truelat2, stand_lon, ref_lat, ref_lon):
    return capi.map_set(capi.PROJ_LC, truelat1=truelat1, truelat2=truelat2, stdlon=stand_lon, lat1=ref_lat,
                        lon1=ref_lon, knowni=(nx + 1) / 2.0, knownj=(ny + 1) / 2.0, dx=dx)


CONFIGS = {
    # BASELINE.json configs[0]: 1-degree GFS-shaped source -> 100x100 30-km Lambert (namelist.wps.all_options geometry)
    "config1": dict(src=(360, 181, -1.0, 1.0), nlev=32, dom=(100, 100, 30000.0, 30.0, 60.0, -98.0, 34.83, -81.03)),
    # BASELINE.json configs[1]: GFS 0.25 deg, 34 levels -> 3-km CONUS Lambert 1799x1059 (HRRR-like)
    "config2": dict(src=(1440, 721, -0.25, 0.25), nlev=34, dom=(1799, 1059, 3000.0, 38.5, 38.5, -97.5, 38.5, -97.5)),
    # BASELINE.json configs[2]: geogrid 30-arc-second topography + 21-category land use -> the same 3-km CONUS domain, smth-desmth
    "config3": dict(dom=(1799, 1059, 3000.0, 38.5, 38.5, -97.5, 38.5, -97.5), dx_deg=0.00833333, known_lat=-89.99583, known_lon=-179.99583,
                    tile=1200, ncat=21, smooth_passes=1),
    # BASELINE.json configs[3]: 3-domain nest 27/9/3 km (parent_grid_ratio 3, namelist.wps-style i/j_parent_start), masked SST / soil
    # fields, Lambert-grid source winds rotated to earth-relative and back (map_to_met + met_to_map)
    "config4": dict(src=(720, 361, -0.5, 0.5), nlev=8, dom=(150, 130, 27000.0, 30.0, 60.0, -98.0, 38.0, -98.0),
                    nests=[(1, 1, 1, 150, 130), (1, 3, (40, 35), 160, 151), (2, 3, (50, 45), 172, 163)]),   # (parent, ratio, (i, j)_parent_start, e_we, e_sn)
    # BASELINE.json configs[4]: ERA5-shaped 0.25 deg, 137 model levels -> 1-km 4000x4000 Lambert, sharded field x level
    "config5": dict(src=(1440, 721, -0.25, 0.25), nlev=137, dom=(4000, 4000, 1000.0, 38.5, 38.5, -97.5, 38.5, -97.5)),
}


def mpas_mesh(ncells=2562, seed=0, jitter=0.25):
    """A synthetic MPAS-like mesh: spherical Voronoi tessellation (scipy) of a jittered Fibonacci lattice, in the arrays
    mpas_mesh_setup reads (mpas_mesh.F:36-250): 1-based connectivity padded to maxEdges, counter-clockwise vertex order,
    edge i of a cell between its vertices i and i+1, cellsOnCell(i) across that edge, radians, lonCell in [0, 2 pi).
    The cell graph is the Delaunay triangulation, as on a real SCVT mesh, so nEdges = 3 (nCells - 2)."""
    from scipy.spatial import SphericalVoronoi
    rng = np.random.default_rng(seed)
    i = np.arange(ncells) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / ncells)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    pts = np.c_[np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)]
    pts += jitter * np.sqrt(4.0 * np.pi / ncells) * 0.5 * rng.standard_normal(pts.shape)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    sv = SphericalVoronoi(pts, 1.0, threshold=1e-12)
    sv.sort_vertices_of_regions()
    verts = sv.vertices
    nv = len(verts)
    regions = []
    for c, reg in enumerate(sv.regions):
        reg = list(reg)
        a, b = verts[reg[0]], verts[reg[1]]
        if np.dot(np.cross(a - pts[c], b - pts[c]), pts[c]) < 0.0:      # make the order counter-clockwise seen from outside
            reg = reg[::-1]
        regions.append(reg)
    max_edges = max(len(r) for r in regions)
    edge_of = {}
    for c, reg in enumerate(regions):
        n = len(reg)
        for k in range(n):
            key = (min(reg[k], reg[(k + 1) % n]), max(reg[k], reg[(k + 1) % n]))
            edge_of.setdefault(key, []).append(c)
    assert all(len(v) == 2 for v in edge_of.values()), "degenerate tessellation"
    edge_id = {key: e for e, key in enumerate(sorted(edge_of))}
    ne = len(edge_id)
    assert ne == 3 * (ncells - 2)
    n_edges_on_cell = np.array([len(r) for r in regions], np.int32)
    cells_on_cell = np.zeros((ncells, max_edges), np.int32)
    vertices_on_cell = np.zeros((ncells, max_edges), np.int32)
    edges_on_cell = np.zeros((ncells, max_edges), np.int32)
    cells_on_edge = np.zeros((ne, 2), np.int32)
    cells_on_vertex = [[] for _ in range(nv)]
    for c, reg in enumerate(regions):
        n = len(reg)
        for k in range(n):
            key = (min(reg[k], reg[(k + 1) % n]), max(reg[k], reg[(k + 1) % n]))
            pair = edge_of[key]
            vertices_on_cell[c, k] = reg[k] + 1
            cells_on_cell[c, k] = (pair[1] if pair[0] == c else pair[0]) + 1
            edges_on_cell[c, k] = edge_id[key] + 1
            cells_on_vertex[reg[k]].append(c)
    for key, e in edge_id.items():
        cells_on_edge[e] = [edge_of[key][0] + 1, edge_of[key][1] + 1]
    assert all(len(v) == 3 for v in cells_on_vertex), "vertex of degree /= 3"
    cov = np.zeros((nv, 3), np.int32)
    for v, cs in enumerate(cells_on_vertex):
        a, b, c3 = (pts[k] for k in cs)
        if np.dot(np.cross(b - a, c3 - a), verts[v]) < 0.0:
            cs = [cs[0], cs[2], cs[1]]
        cov[v] = [k + 1 for k in cs]
    emid = pts[cells_on_edge[:, 0] - 1] + pts[cells_on_edge[:, 1] - 1]
    emid /= np.linalg.norm(emid, axis=1, keepdims=True)

    def latlon(x):
        lon = np.arctan2(x[:, 1], x[:, 0])
        return np.arcsin(np.clip(x[:, 2], -1.0, 1.0)).astype(np.float32), np.where(lon < 0.0, lon + 2.0 * np.pi, lon).astype(np.float32)
    lat_c, lon_c = latlon(pts)
    lat_v, lon_v = latlon(verts)
    lat_e, lon_e = latlon(emid)
    landmask = (np.sin(5.0 * lon_c.astype(np.float64)) + np.cos(4.0 * lat_c.astype(np.float64)) > 0.3).astype(np.int32)
    return dict(nCells=ncells, nVertices=nv, nEdges=ne, maxEdges=max_edges, landmask=landmask, nEdgesOnCell=n_edges_on_cell,
                cellsOnCell=cells_on_cell, verticesOnCell=vertices_on_cell, edgesOnCell=edges_on_cell, cellsOnVertex=cov,
                cellsOnEdge=cells_on_edge, latCell=lat_c, lonCell=lon_c, latVertex=lat_v, lonVertex=lon_v, latEdge=lat_e, lonEdge=lon_e)


def mpas_field(mesh, nlev, seed=0, where="Cell"):
    """[nNodes][nlev] binary32 values on the cells / vertices / edges of `mesh`: smooth in space plus noise, level index fastest."""
    lat = mesh["lat" + where].astype(np.float64)[:, None]
    lon = mesh["lon" + where].astype(np.float64)[:, None]
    lev = np.arange(nlev, dtype=np.float64)[None, :]
    rng = np.random.default_rng(seed)
    v = 250.0 + 30.0 * np.sin(3.0 * lon) * np.cos(2.0 * lat) + 0.7 * lev + 0.05 * rng.standard_normal((lat.shape[0], nlev))
    return np.ascontiguousarray(v, dtype=np.float32)
