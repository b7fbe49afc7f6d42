"""ctypes binding of libwpsgpu.so (include/wpsgpu.h) -- the C-ABI drop-in boundary.

Every call goes through the shared library; there is no Python/CPU implementation of
any operator behind it.  Importing this module fails loudly when the library is
missing, and every entry point fails when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WPSGPU_LIB") or os.path.join(_HERE, "libwpsgpu.so")   # WPSGPU_LIB: another build of the same library (A/B measurements)

NAN = np.float32(1.0e20)
NOT_MASKED, MASKED_BOTH, MASKED_WATER, MASKED_LAND = -2.0, -1.0, 0.0, 1.0
SIXTEEN_POINT, FOUR_POINT, N_NEIGHBOR, AVERAGE4, AVERAGE16, W_AVERAGE4, W_AVERAGE16, SEARCH = range(1, 9)
M, U, V, HH, VV, CORNER = range(1, 7)
ONETWOONE, SMTHDESMTH, SMTHDESMTH_SPECIAL = 1, 2, 3
CONTINUOUS, CATEGORICAL, SP_CONTINUOUS = 0, 1, 2
PROJ_LATLON, PROJ_LC, PROJ_PS, PROJ_MERC, PROJ_GAUSS, PROJ_CYL, PROJ_CASSINI = 0, 1, 2, 3, 4, 5, 6
PROJ_PS_WGS84, PROJ_ALBERS_NAD83, PROJ_ROTLL = 102, 105, 203
PTR_HOST, PTR_DEVICE = 0, 1
PRECISION_EXACT, PRECISION_TOLERANCE = 0, 1
MAX_OPS = 16
MAX_GAUSS = 2048

_A = dict(lat1=0, lon1=1, lat0=2, lon0=3, knowni=4, knownj=5, dx=6, dy=7, latinc=8, loninc=9, stdlon=10,
          truelat1=11, truelat2=12, nlat=13, ixdim=15, jydim=16, nxmin=17, nxmax=18, stagger=19, phi=20, r_earth=22)
_A["lambda"] = 21


class Proj(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("code", "nlat", "nxmin", "nxmax", "stagger", "comp_ll", "init", "n_gauss")] + \
               [(n, C.c_float) for n in ("lat1", "lon1", "lat0", "lon0", "dx", "dy", "latinc", "loninc", "dlon",
                                         "stdlon", "truelat1", "truelat2", "hemi", "cone", "polei", "polej", "rsw",
                                         "rebydx", "knowni", "knownj", "re_m", "rho0", "nc", "bigc")] + \
               [("gauss_lat", C.c_float * MAX_GAUSS), ("ixdim", C.c_int32), ("jydim", C.c_int32), ("phi", C.c_float), ("lambda", C.c_float)]


class MapArgs(C.Structure):
    _fields_ = [("present", C.c_uint32)] + \
               [(n, C.c_float) for n in ("lat1", "lon1", "lat0", "lon0", "knowni", "knownj", "dx", "dy", "latinc",
                                         "loninc", "stdlon", "truelat1", "truelat2", "r_earth")] + \
               [(n, C.c_int32) for n in ("nlat", "nxmin", "nxmax", "stagger", "ixdim", "jydim")] + [("phi", C.c_float), ("lambda", C.c_float)]


class FieldOpts(C.Structure):
    _fields_ = [("ops", C.c_int32 * MAX_OPS), ("opts", C.c_int32 * MAX_OPS), ("missing_value", C.c_float),
                ("fill_missing", C.c_float), ("masked", C.c_float), ("mask_val", C.c_float * 3),
                ("mask_rel", C.c_char * 4)]


class Slab(C.Structure):
    _fields_ = [("slab", C.c_void_p), ("minx", C.c_int32), ("maxx", C.c_int32), ("miny", C.c_int32),
                ("maxy", C.c_int32), ("bdr", C.c_int32), ("mask", C.c_void_p * 3), ("grid_id", C.c_int32),
                ("field_stagger", C.c_int32), ("use_landmask", C.c_int32), ("process_bdy_width", C.c_int32),
                ("r_arr", C.c_void_p), ("valid_mask", C.c_void_p), ("new_pts", C.c_void_p),
                ("raw_nx", C.c_int32), ("raw_big_endian", C.c_int32)]


class GeoField(C.Structure):
    _fields_ = [("dom_proj", Proj), ("istagger", C.c_int32), ("xlat", C.c_void_p), ("xlon", C.c_void_p),
                ("start_i", C.c_int32), ("end_i", C.c_int32), ("start_j", C.c_int32), ("end_j", C.c_int32),
                ("start_k", C.c_int32), ("end_k", C.c_int32), ("landmask", C.c_void_p), ("field", C.c_void_p)]


class GeoLevel(C.Structure):
    _fields_ = [("src_proj", Proj), ("ilevel", C.c_int32), ("itype", C.c_int32), ("ops", C.c_int32 * MAX_OPS),
                ("opts", C.c_int32 * MAX_OPS), ("msg_val", C.c_float), ("msg_fill_val", C.c_float),
                ("mask_val", C.c_float), ("has_scale", C.c_int32), ("scale_factor", C.c_float),
                ("sr_x", C.c_int32), ("sr_y", C.c_int32)]


MPAS_CELLS, MPAS_VERTICES, MPAS_EDGES, MPAS_MASKED_CELLS = 0, 1, 2, 3
MPAS_REAL, MPAS_DOUBLE, MPAS_INTEGER = 0, 1, 2
_MESH_INT = ("landmask", "nEdgesOnCell", "cellsOnCell", "verticesOnCell", "edgesOnCell", "cellsOnVertex", "cellsOnEdge")
_MESH_REAL = ("latCell", "lonCell", "latVertex", "lonVertex", "latEdge", "lonEdge")


class MpasMesh(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nCells", "nVertices", "nEdges", "maxEdges")] + \
               [(n, C.c_void_p) for n in _MESH_INT + _MESH_REAL]


class WpsGpuError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libwpsgpu.so; there is no fallback if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WpsGpuError("%s not found: build it with wps_b200/csrc/build.sh (or __graft_entry__.build()); "
                              "this package has no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.wpsgpu_last_error.restype = C.c_char_p
        _lib.wpsgpu_stream.restype = C.c_void_p
        _lib.wpsgpu_stream.argtypes = [C.c_void_p]
        _lib.wpsgpu_launch_count.restype = C.c_int64
        _lib.wpsgpu_launch_count.argtypes = [C.c_void_p]
    return _lib


def check(rc):
    if rc != 0:
        raise WpsGpuError("wpsgpu error %d: %s" % (rc, lib().wpsgpu_last_error().decode(errors="replace")))


def ptr(a):
    """Raw address of a numpy array (host mode) / torch tensor (device mode) / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "arrays must be contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


def map_set(code, **kw):
    """map_set (module_map_utils.F:243) with keyword OPTIONAL arguments."""
    a = MapArgs()
    present = 0
    for k, v in kw.items():
        present |= 1 << _A[k]
        setattr(a, k, v)
    a.present = present
    p = Proj()
    check(lib().wpsgpu_map_set(int(code), C.byref(a), C.byref(p)))
    return p


def compute_nest_locations(iproj, nests, **coarse_kw):
    """compute_nest_locations (llxy_module.F:331): nests = [(parent_id, parent_grid_ratio, i_parent_start, j_parent_start,
    e_we, e_sn), ...] with entry 0 the coarse domain; coarse_kw = the map_set keywords of domain 1.  -> list of Proj."""
    a = MapArgs()
    present = 0
    for k, v in coarse_kw.items():
        present |= 1 << _A[k]
        setattr(a, k, v)
    a.present = present
    n = len(nests)
    arr = lambda col: (C.c_int32 * n)(*[int(t[col]) for t in nests])
    out = (Proj * n)()
    check(lib().wpsgpu_compute_nest_locations(int(iproj), C.byref(a), n, arr(0), arr(1), arr(2), arr(3), arr(4), arr(5), out))
    return [out[k] for k in range(n)]


def push_source_projection(iproj, stand_lon=0., truelat1=0., truelat2=0., dxkm=0., dykm=0., dlat=0., dlon=0.,
                           known_x=1., known_y=1., known_lat=0., known_lon=0., cassini=None, earth_radius=None):
    """push_source_projection (llxy_module.F:39)."""
    p = Proj()
    f = C.c_float
    cas = None if cassini is None else (C.c_float * 6)(*cassini)
    er = None if earth_radius is None else C.byref(C.c_float(earth_radius))
    check(lib().wpsgpu_push_source_projection(int(iproj), f(stand_lon), f(truelat1), f(truelat2), f(dxkm), f(dykm),
                                              f(dlat), f(dlon), f(known_x), f(known_y), f(known_lat), f(known_lon),
                                              cas, er, C.byref(p)))
    return p


def parse_interp_option(s):
    """-> (codes, opts, has_gcell, gcell_threshold); interp_module.F:20-295, interp_option_module.F:692."""
    codes = (C.c_int32 * MAX_OPS)()
    opts = (C.c_int32 * MAX_OPS)()
    n = C.c_int()
    hg = C.c_int()
    thr = C.c_float()
    check(lib().wpsgpu_parse_interp_option(s.encode(), codes, opts, MAX_OPS, C.byref(n), C.byref(hg), C.byref(thr)))
    return list(codes[:n.value]), list(opts[:n.value]), bool(hg.value), thr.value


def field_opts(chain, missing_value=NAN, fill_missing=NAN, masked=NOT_MASKED, mask_val=(NAN, NAN, NAN),
               mask_rel="   "):
    fo = FieldOpts()
    if isinstance(chain, str):
        codes, opts, _, _ = parse_interp_option(chain)
    else:
        codes, opts = chain
    for i, (c, o) in enumerate(zip(codes, opts)):
        fo.ops[i] = int(c)
        fo.opts[i] = int(o)
    fo.missing_value = missing_value
    fo.fill_missing = fill_missing
    fo.masked = masked
    for i in range(3):
        fo.mask_val[i] = mask_val[i]
    fo.mask_rel = mask_rel.encode()
    return fo


def nwords(nx):
    return (nx + 31) // 32


class Handle:
    """One handle per GPU (wpsgpu_create / wpsgpu_destroy)."""

    def __init__(self, device=0, ptr_mode=PTR_HOST):
        self._h = C.c_void_p()
        check(lib().wpsgpu_create(int(device), C.byref(self._h)))
        self.device = device
        self.ptr_mode = PTR_HOST
        self._keep = []
        if ptr_mode != PTR_HOST:
            self.set_pointer_mode(ptr_mode)

    def close(self):
        if self._h:
            lib().wpsgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_pointer_mode(self, mode):
        check(lib().wpsgpu_set_pointer_mode(self._h, int(mode)))
        self.ptr_mode = mode

    def synchronize(self):
        check(lib().wpsgpu_synchronize(self._h))

    @property
    def stream(self):
        return lib().wpsgpu_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().wpsgpu_launch_count(self._h))

    # ---- projections ----
    def lltoxy(self, proj, lat, lon, stagger=M, out=None):
        n = int(np.prod(lat.shape))
        if out is None:
            out = (_like(lat), _like(lat))
        check(lib().wpsgpu_lltoxy(self._h, C.byref(proj), int(stagger), C.c_void_p(ptr(lat)), C.c_void_p(ptr(lon)),
                                  C.c_int64(n), C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1]))))
        return out

    def xytoll(self, proj, x, y, stagger=M, out=None):
        n = int(np.prod(x.shape))
        if out is None:
            out = (_like(x), _like(x))
        check(lib().wpsgpu_xytoll(self._h, C.byref(proj), int(stagger), C.c_void_p(ptr(x)), C.c_void_p(ptr(y)),
                                  C.c_int64(n), C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1]))))
        return out

    # ---- metgrid ----
    def metgrid_set_grid(self, grid_id, xlat, xlon, sm1, sm2):
        ny, nx = xlat.shape
        check(lib().wpsgpu_metgrid_set_grid(self._h, int(grid_id), C.c_void_p(ptr(xlat)), C.c_void_p(ptr(xlon)),
                                            int(sm1), int(sm1 + nx - 1), int(sm2), int(sm2 + ny - 1)))
        if self.ptr_mode == PTR_DEVICE:
            self._keep.append((xlat, xlon))

    def metgrid_set_landmask(self, grid_id, landmask):
        check(lib().wpsgpu_metgrid_set_landmask(self._h, int(grid_id), C.c_void_p(ptr(landmask))))
        if self.ptr_mode == PTR_DEVICE:
            self._keep.append(landmask)

    def metgrid_set_domain(self, sd1, ed1, sd2, ed2):
        check(lib().wpsgpu_metgrid_set_domain(self._h, int(sd1), int(ed1), int(sd2), int(ed2)))

    @staticmethod
    def slab_desc(slab, minx, miny, bdr, grid_id, field_stagger, r_arr, new_pts, valid_mask=None,
                  masks=(None, None, None), use_landmask=False, process_bdy_width=0, out=None, raw_big_endian=None):
        """Fill a wpsgpu_slab descriptor (a caller that re-uses its buffers can build these once).
        raw_big_endian not None: `slab` is the record's [ny][nx] array as stored in the file (4-byte words, big-endian
        if True) and the library adds the `bdr` wrap-around columns itself; minx then still names the halo'd bound."""
        sny, snx = slab.shape
        s = Slab() if out is None else out
        s.slab = ptr(slab)
        s.raw_nx, s.raw_big_endian = 0, 0
        if raw_big_endian is not None:
            s.raw_nx, s.raw_big_endian = int(snx), int(bool(raw_big_endian))
            snx = snx + 2 * int(bdr)
        s.minx, s.maxx, s.miny, s.maxy, s.bdr = int(minx), int(minx + snx - 1), int(miny), int(miny + sny - 1), int(bdr)
        for i in range(3):
            s.mask[i] = ptr(masks[i])
        s.grid_id = int(grid_id)
        s.field_stagger = int(field_stagger)
        s.use_landmask = int(bool(use_landmask))
        s.process_bdy_width = int(process_bdy_width)
        s.r_arr = ptr(r_arr)
        s.valid_mask = ptr(valid_mask)
        s.new_pts = ptr(new_pts)
        return s

    def metgrid_enqueue_slabs(self, proj, fo, descs, n=None):
        """descs: a ctypes array of Slab descriptors (levels of one field); one C call queues them all."""
        n = len(descs) if n is None else int(n)
        check(lib().wpsgpu_metgrid_enqueue_slabs(self._h, C.byref(proj), C.byref(fo), descs, n))

    def metgrid_enqueue_slab(self, proj, fo, slab, minx, miny, bdr, grid_id, field_stagger, r_arr, new_pts,
                             valid_mask=None, masks=(None, None, None), use_landmask=False, process_bdy_width=0,
                             raw_big_endian=None):
        s = self.slab_desc(slab, minx, miny, bdr, grid_id, field_stagger, r_arr, new_pts, valid_mask=valid_mask, masks=masks,
                           use_landmask=use_landmask, process_bdy_width=process_bdy_width, raw_big_endian=raw_big_endian)
        check(lib().wpsgpu_metgrid_enqueue_slab(self._h, C.byref(proj), C.byref(fo), C.byref(s)))

    def metgrid_enqueue_gcell_slab(self, proj, dom_proj, fo, slab, minx, miny, bdr, grid_id, field_stagger, r_arr, new_pts,
                                   valid_mask=None, mask=None, use_landmask=False, process_bdy_width=0):
        """average_gcell branch of interp_met_field (process_domain_module.F:2331-2478); results after metgrid_flush."""
        s = self.slab_desc(slab, minx, miny, bdr, grid_id, field_stagger, r_arr, new_pts, valid_mask=valid_mask,
                           masks=(mask, None, None), use_landmask=use_landmask, process_bdy_width=process_bdy_width)
        check(lib().wpsgpu_metgrid_enqueue_gcell_slab(self._h, C.byref(proj), C.byref(dom_proj), C.byref(fo), C.byref(s)))

    def metgrid_flush(self, wait=True):
        """wait=False: wpsgpu_metgrid_flush_async -- results are complete after synchronize()."""
        check(lib().wpsgpu_metgrid_flush(self._h) if wait else lib().wpsgpu_metgrid_flush_async(self._h))

    def metgrid_copy_bytes(self):
        a, b = C.c_int64(), C.c_int64()
        check(lib().wpsgpu_metgrid_copy_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def metgrid_last_kernel(self):
        ms = C.c_float()
        pts = C.c_double()
        check(lib().wpsgpu_metgrid_last_kernel(self._h, C.byref(ms), C.byref(pts)))
        return ms.value, pts.value

    def metgrid_last_timings(self):
        ms = (C.c_float * 7)()
        pts = (C.c_double * 7)()
        check(lib().wpsgpu_metgrid_last_timings(self._h, ms, pts))
        return list(ms), list(pts)

    def fill_missing_levels(self, levels, r_arrs, valid_masks):
        """Vertical interpolation of fill_missing_levels (process_domain_module.F:2862-2921) over the levels of one
        3-d field, in place.  r_arrs / valid_masks: sequences of per-level arrays [ny][nx] / [ny][nwords]."""
        n = len(levels)
        ny, nx = r_arrs[0].shape
        lv = (C.c_int32 * n)(*[int(l) for l in levels])
        pr = (C.c_void_p * n)(*[ptr(a) for a in r_arrs])
        pv = (C.c_void_p * n)(*[ptr(a) for a in valid_masks])
        check(lib().wpsgpu_fill_missing_levels(self._h, n, lv, pr, pv, int(nx), int(ny)))

    def create_level(self, r_arr, valid_mask, fill_const=None, src=None, src_valid=None):
        """create_level (process_domain_module.F:3191-3332): constant fill, or copy of (src, src_valid)."""
        ny, nx = r_arr.shape
        check(lib().wpsgpu_create_level(self._h, int(nx), int(ny), C.c_float(0.0 if fill_const is None else fill_const),
                                        C.c_void_p(ptr(src)), C.c_void_p(ptr(src_valid)), C.c_void_p(ptr(r_arr)),
                                        C.c_void_p(ptr(valid_mask))))

    def metgrid_set_timing(self, on=True):
        """Record the CUDA events metgrid_last_kernel / metgrid_last_timings read (off by default: ~2 % of a step)."""
        check(lib().wpsgpu_metgrid_set_timing(self._h, int(bool(on))))

    def metgrid_set_precision(self, mode):
        """PRECISION_EXACT (default, bit-identical sixteen_pt) or PRECISION_TOLERANCE (north_star's 1e-5, ~3x faster)."""
        check(lib().wpsgpu_metgrid_set_precision(self._h, int(mode)))

    def debug_disable_fast(self, disable=True):
        check(lib().wpsgpu_debug_disable_fast(self._h, int(disable)))

    def rotate_winds(self, idir, proj, u, u_mask, v, v_mask, xlon_u, xlon_v, us=(1, 1), vs=(1, 1), xlat_u=None, xlat_v=None):
        """u [nlev][uny][unx], v [nlev][vny][vnx] rotated in place (metmap_xform); xlat_u / xlat_v for PROJ_CASSINI."""
        nlev, uny, unx = u.shape
        _, vny, vnx = v.shape
        check(lib().wpsgpu_rotate_winds_ll(self._h, int(idir), C.byref(proj), C.c_void_p(ptr(u)), C.c_void_p(ptr(u_mask)),
                                           C.c_void_p(ptr(v)), C.c_void_p(ptr(v_mask)), us[0], us[1], us[0] + unx - 1,
                                           us[1] + uny - 1, vs[0], vs[1], vs[0] + vnx - 1, vs[1] + vny - 1,
                                           C.c_void_p(ptr(xlon_u)), C.c_void_p(ptr(xlon_v)), C.c_void_p(ptr(xlat_u)),
                                           C.c_void_p(ptr(xlat_v)), int(nlev)))

    # ---- geogrid ----
    def rotate_winds_nmm(self, idir, proj, u, u_mask, v, v_mask, xlat_v, xlon_v):
        """metmap_xform_nmm (rotate_winds_module.F:455-631): u, v [nlev][ny][nx] on the E-grid's velocity points, in place."""
        nlev, ny, nx = u.shape
        check(lib().wpsgpu_rotate_winds_nmm(self._h, int(idir), C.byref(proj), C.c_void_p(ptr(u)), C.c_void_p(ptr(u_mask)), C.c_void_p(ptr(v)),
                                            C.c_void_p(ptr(v_mask)), 1, 1, int(nx), int(ny), C.c_void_p(ptr(xlat_v)), C.c_void_p(ptr(xlon_v)), int(nlev)))

    def geogrid_field_begin(self, dom_proj, istagger, xlat, xlon, start_i, start_j, start_k, field, landmask=None):
        nk, ny, nx = field.shape
        f = GeoField()
        f.dom_proj = dom_proj
        f.istagger = int(istagger)
        f.xlat, f.xlon = ptr(xlat), ptr(xlon)
        f.start_i, f.end_i = int(start_i), int(start_i + nx - 1)
        f.start_j, f.end_j = int(start_j), int(start_j + ny - 1)
        f.start_k, f.end_k = int(start_k), int(start_k + nk - 1)
        f.landmask = ptr(landmask)
        f.field = ptr(field)
        self._keep.append((xlat, xlon, field, landmask))
        check(lib().wpsgpu_geogrid_field_begin(self._h, C.byref(f)))

    def geogrid_level_begin(self, src_proj, ilevel, itype, chain, msg_val=NAN, msg_fill_val=NAN, mask_val=-1.0,
                            scale_factor=None, sr_x=1, sr_y=1):
        l = GeoLevel()
        l.src_proj = src_proj
        l.ilevel, l.itype = int(ilevel), int(itype)
        codes, opts, _, _ = parse_interp_option(chain) if isinstance(chain, str) else (chain[0], chain[1], 0, 0)
        for i, (c, o) in enumerate(zip(codes, opts)):
            l.ops[i], l.opts[i] = int(c), int(o)
        l.msg_val, l.msg_fill_val, l.mask_val = msg_val, msg_fill_val, mask_val
        l.has_scale = int(scale_factor is not None)
        l.scale_factor = 1.0 if scale_factor is None else scale_factor
        l.sr_x, l.sr_y = int(sr_x), int(sr_y)
        check(lib().wpsgpu_geogrid_level_begin(self._h, C.byref(l)))

    def geogrid_add_tile(self, tile, src_min_x, src_min_y, src_min_z, bdr):
        nz, ny, nx = tile.shape
        check(lib().wpsgpu_geogrid_add_tile(self._h, C.c_void_p(ptr(tile)), int(src_min_x), int(src_min_x + nx - 1),
                                            int(src_min_y), int(src_min_y + ny - 1), int(src_min_z),
                                            int(src_min_z + nz - 1), int(bdr)))

    def geogrid_add_tile_raw(self, raw, shape, wordsize, is_signed, endian, row_order, src_min_x, src_min_y,
                             src_min_z, bdr):
        nz, ny, nx = shape
        check(lib().wpsgpu_geogrid_add_tile_raw(self._h, C.c_void_p(ptr(raw)), int(wordsize), int(is_signed),
                                                int(endian), int(row_order), int(src_min_x), int(src_min_x + nx - 1),
                                                int(src_min_y), int(src_min_y + ny - 1), int(src_min_z),
                                                int(src_min_z + nz - 1), int(bdr)))

    def geogrid_decode_tile(self, raw, shape, wordsize, is_signed, endian, row_order=1, scalefactor=1.0, out=None):
        """read_geogrid.c on the device: raw bytes -> float32 [nz][ny][nx]."""
        nz, ny, nx = shape
        if out is None:
            out = np.empty((nz, ny, nx), np.float32)
        check(lib().wpsgpu_geogrid_decode_tile(self._h, C.c_void_p(ptr(raw)), int(wordsize), int(is_signed), int(endian),
                                               int(row_order), int(nx), int(ny), int(nz), C.c_float(scalefactor),
                                               C.c_void_p(ptr(out))))
        return out

    def geogrid_level_end(self):
        check(lib().wpsgpu_geogrid_level_end(self._h))

    def geogrid_field_finish(self, processed_out=None):
        inv = C.c_int64()
        check(lib().wpsgpu_geogrid_field_finish(self._h, C.c_void_p(ptr(processed_out)), C.byref(inv)))
        self._keep = []
        return inv.value

    def fractions_dominant(self, field, min_cat, msg_fill_val=NAN, lm_vals=(), is_water_mask=True,
                           landmask_out=None, dominant_out=None):
        ncat = field.shape[0]
        npts = int(np.prod(field.shape[1:]))
        lm = (C.c_int32 * max(1, len(lm_vals)))(*lm_vals)
        check(lib().wpsgpu_fractions_dominant(self._h, C.c_void_p(ptr(field)), C.c_int64(npts), int(min_cat),
                                              int(min_cat + ncat - 1), C.c_float(msg_fill_val), lm, len(lm_vals),
                                              int(is_water_mask), C.c_void_p(ptr(landmask_out)),
                                              C.c_void_p(ptr(dominant_out))))

    def smooth(self, array, opt, npass):
        a3 = array if array.ndim == 3 else array.reshape((1,) + tuple(array.shape))
        nz, ny, nx = a3.shape
        check(lib().wpsgpu_smooth(self._h, int(opt), int(npass), C.c_void_p(ptr(array)), 1, nx, 1, ny, 1, nz))

    def smooth_egrid(self, array, opt, npass, msgval=1.0e20, hflag=1.0, sx=1, sy=1):
        """one_two_one_egrid / smth_desmth_egrid (smooth_module.F:252-585) in place on array [nz][ny][nx] with lower bounds (sx, sy)."""
        nz, ny, nx = array.shape
        check(lib().wpsgpu_smooth_egrid(self._h, int(opt), int(npass), C.c_void_p(ptr(array)), int(sx), int(sx + nx - 1), int(sy), int(sy + ny - 1),
                                        1, int(nz), C.c_float(msgval), C.c_float(hflag)))

    def xytoll_grid(self, dom, stagger, si, sj, nx, ny, sub_x=1.0, sub_y=1.0, out=None):
        if out is None:
            out = (np.empty((ny, nx), np.float32), np.empty((ny, nx), np.float32))
        check(lib().wpsgpu_xytoll_grid(self._h, C.byref(dom), int(stagger), int(si), int(sj), int(si + nx - 1),
                                       int(sj + ny - 1), C.c_float(sub_x), C.c_float(sub_y),
                                       C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1]))))
        return out

    def map_factor(self, iproj, truelat1, truelat2, xlat, xlon, pole_lat=90.0, pole_lon=0.0, stand_lon=0.0, out=None):
        if out is None:
            out = (_like(xlat), _like(xlat))
        f = C.c_float
        check(lib().wpsgpu_map_factor(self._h, int(iproj), f(truelat1), f(truelat2), f(pole_lat), f(pole_lon),
                                      f(stand_lon), C.c_void_p(ptr(xlat)), C.c_void_p(ptr(xlon)),
                                      C.c_int64(int(np.prod(xlat.shape))), C.c_void_p(ptr(out[0])),
                                      C.c_void_p(ptr(out[1]))))
        return out

    def coriolis(self, xlat, out=None):
        if out is None:
            out = (_like(xlat), _like(xlat))
        check(lib().wpsgpu_coriolis(self._h, C.c_void_p(ptr(xlat)), C.c_int64(int(np.prod(xlat.shape))),
                                    C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1]))))
        return out

    def rotang(self, iproj, truelat1, truelat2, stand_lon, xlat, xlon, out=None):
        ny, nx = xlat.shape
        if out is None:
            out = (_like(xlat), _like(xlat))
        f = C.c_float
        check(lib().wpsgpu_rotang(self._h, int(iproj), f(truelat1), f(truelat2), f(stand_lon), C.c_void_p(ptr(xlat)),
                                  C.c_void_p(ptr(xlon)), 1, 1, nx, ny, C.c_void_p(ptr(out[0])), C.c_void_p(ptr(out[1]))))
        return out  # (cosa, sina)

    def dfdx(self, src, sr_x, dom_dx, mapfac=None, out=None):
        return self._dfd(lib().wpsgpu_dfdx, src, sr_x, dom_dx, mapfac, out)

    def dfdy(self, src, sr_y, dom_dy, mapfac=None, out=None):
        return self._dfd(lib().wpsgpu_dfdy, src, sr_y, dom_dy, mapfac, out)

    def _dfd(self, fn, src, sr, dom_d, mapfac, out):
        a3 = src if src.ndim == 3 else src.reshape((1,) + tuple(src.shape))
        nz, ny, nx = a3.shape
        if out is None:
            out = _like(src)
        check(fn(self._h, C.c_void_p(ptr(src)), C.c_void_p(ptr(out)), 1, 1, 1, nx, ny, nz, int(sr),
                 C.c_void_p(ptr(mapfac)), C.c_float(dom_d)))
        return out

    # ---- MPAS remapping (remapper.F) ----
    def mpas_set_mesh(self, mesh):
        """mpas_mesh_setup: `mesh` is a dict of numpy arrays (1-based connectivity, radians) as synth.mpas_mesh makes it."""
        d = MpasMesh(nCells=mesh["nCells"], nVertices=mesh["nVertices"], nEdges=mesh["nEdges"], maxEdges=mesh["maxEdges"])
        keep = []
        for n in _MESH_INT + _MESH_REAL:
            a = np.ascontiguousarray(mesh[n], dtype=np.int32 if n in _MESH_INT else np.float32)
            keep.append(a)
            setattr(d, n, a.ctypes.data)
        check(lib().wpsgpu_mpas_set_mesh(self._h, C.byref(d)))
        self._mpas_dims = (mesh["nCells"], mesh["nVertices"], mesh["nEdges"], mesh["maxEdges"])
        self._mpas_grids = {}

    def mpas_remap_setup(self, grid_id, lat_rad, lon_rad):
        """target_mesh_setup + remap_info_setup (remapper.F:83-286); lat_rad / lon_rad [nlat][nlon] in radians."""
        nlat, nlon = lat_rad.shape
        check(lib().wpsgpu_mpas_remap_setup(self._h, int(grid_id), C.c_void_p(ptr(lat_rad)), C.c_void_p(ptr(lon_rad)), int(nlon), int(nlat)))
        self._mpas_grids[grid_id] = (nlat, nlon)

    def mpas_get_tables(self, grid_id, kind):
        """(nearest [nlat][nlon] or None, sources [nlat][nlon][width], weights [nlat][nlon][width]) as host arrays."""
        assert self.ptr_mode == PTR_HOST
        nlat, nlon = self._mpas_grids[grid_id]
        width = 3 if kind in (MPAS_CELLS, MPAS_MASKED_CELLS) else self._mpas_dims[3]
        near = None if kind == MPAS_MASKED_CELLS else np.empty((nlat, nlon), np.int32)
        src = np.empty((nlat, nlon, width), np.int32)
        w = np.empty((nlat, nlon, width), np.float32)
        check(lib().wpsgpu_mpas_get_tables(self._h, int(grid_id), int(kind), C.c_void_p(ptr(near)), C.c_void_p(ptr(src)), C.c_void_p(ptr(w))))
        return near, src, w

    def mpas_remap_field(self, grid_id, decomp, src, masked=False, nw=0, out=None):
        """remap_field (remapper.F:417-663): src [nNodes][nlev] (or [nNodes]) float32 / float64 / int32 ->
        [nlat][nlon][nlev], level index fastest as in the reference's wrf_field."""
        nlat, nlon = self._mpas_grids[grid_id]
        nlev = 1 if src.ndim == 1 else int(src.shape[1])
        xtype = {"float32": MPAS_REAL, "float64": MPAS_DOUBLE, "int32": MPAS_INTEGER}[str(src.dtype).replace("torch.", "")]
        if out is None:
            if isinstance(src, np.ndarray):
                out = np.empty((nlat, nlon, nlev), src.dtype)
            else:
                import torch
                out = torch.empty((nlat, nlon, nlev), dtype=src.dtype, device=src.device)
        check(lib().wpsgpu_mpas_remap_field(self._h, int(grid_id), int(decomp), xtype, int(bool(masked)), C.c_void_p(ptr(src)), nlev,
                                            int(nw), C.c_void_p(ptr(out))))
        return out

    def mpas_remap_planes(self, grid_id, decomp, src, planes, mode=0, masked=False, landmask=None, fill_missing=0.0):
        """remap_field fused with process_mpas_fields' per-level extraction (process_domain_module.F:1700-1956):
        planes = sequence of nlev arrays [nlat][nlon] (the r_arr of each stored level)."""
        nlev = 1 if src.ndim == 1 else int(src.shape[1])
        assert len(planes) == nlev
        pp = (C.c_void_p * nlev)(*[ptr(a) for a in planes])
        check(lib().wpsgpu_mpas_remap_planes(self._h, int(grid_id), int(decomp), int(bool(masked)), C.c_void_p(ptr(src)), nlev, int(mode),
                                             C.c_void_p(ptr(landmask)), C.c_float(fill_missing), pp))

    def mpas_derive_tt(self, theta_planes, pressure_planes):
        """derive_mpas_fields (process_domain_module.F:1957-2053): theta planes become TT, in place."""
        n = len(theta_planes)
        assert n == len(pressure_planes)
        npts = int(np.prod(theta_planes[0].shape)) if n else 1
        pt = (C.c_void_p * max(n, 1))(*[ptr(a) for a in theta_planes])
        pp = (C.c_void_p * max(n, 1))(*[ptr(a) for a in pressure_planes])
        check(lib().wpsgpu_mpas_derive_tt(self._h, pt, pp, n, C.c_int64(npts)))


def _like(a):
    if isinstance(a, np.ndarray):
        return np.empty_like(a)
    import torch
    return torch.empty_like(a)
