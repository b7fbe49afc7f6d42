"""ctypes binding of the CPU ORACLE (oracle/libwps_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(wps_b200/) must never import this module.  PARITY UNPINNED -- see
oracle/wps_oracle.h and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NAN = np.float32(1.0e20)
SIXTEEN_POINT, FOUR_POINT, N_NEIGHBOR, AVERAGE4, AVERAGE16, W_AVERAGE4, W_AVERAGE16, SEARCH = range(1, 9)
M, U, V, HH, VV, CORNER = range(1, 7)
PROJ_LATLON, PROJ_LC, PROJ_PS, PROJ_MERC, PROJ_GAUSS, PROJ_CYL, PROJ_CASSINI = 0, 1, 2, 3, 4, 5, 6
PROJ_PS_WGS84, PROJ_ALBERS_NAD83, PROJ_ROTLL = 102, 105, 203

_A = dict(lat1=0, lon1=1, lat0=2, lon0=3, knowni=4, knownj=5, dx=6, dy=7, latinc=8, loninc=9, stdlon=10,
          truelat1=11, truelat2=12, nlat=13, nlon=14, ixdim=15, jydim=16, nxmin=17, nxmax=18, stagger=19,
          phi=20, lambda_=21, r_earth=22)


class Proj(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("code", "nlat", "nlon", "nxmin", "nxmax", "ixdim", "jydim", "stagger")] + \
               [(n, C.c_float) for n in ("phi", "lambda_", "lat1", "lon1", "lat0", "lon0", "dx", "dy", "latinc",
                                         "loninc", "dlat", "dlon", "stdlon", "truelat1", "truelat2", "hemi", "cone",
                                         "polei", "polej", "rsw", "rebydx", "knowni", "knownj", "re_m", "rho0",
                                         "nc", "bigc")] + \
               [(n, C.c_int) for n in ("init", "wrap", "comp_ll", "n_gauss")] + \
               [("gauss_lat", C.c_float * 2048)]


class MapArgs(C.Structure):
    _fields_ = [("present", C.c_uint)] + \
               [(n, C.c_float) for n in ("lat1", "lon1", "lat0", "lon0", "knowni", "knownj", "dx", "dy", "latinc",
                                         "loninc", "stdlon", "truelat1", "truelat2", "phi", "lambda_", "r_earth")] + \
               [(n, C.c_int) for n in ("nlat", "nlon", "ixdim", "jydim", "nxmin", "nxmax", "stagger")]


class FieldOpts(C.Structure):
    _fields_ = [("ops", C.c_int * 16), ("opts", C.c_int * 16), ("missing_value", C.c_float),
                ("fill_missing", C.c_float), ("masked", C.c_float), ("mask_val", C.c_float * 3),
                ("mask_rel", C.c_char * 3)]


def build(force=False):
    so = os.path.join(_HERE, "libwps_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_interp_sequence.restype = C.c_float
        _LIB.orc_oned.restype = C.c_float
        _LIB.orc_oned.argtypes = [C.c_float] * 5
        _LIB.orc_interp_to_latlon.restype = C.c_float
    return _LIB


def fptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def set_transc_mode(mode):
    lib().orc_set_transc_mode(int(mode))


def map_set(code, **kw):
    """map_set(module_map_utils.F:243) with keyword OPTIONALs; returns a populated Proj."""
    a = MapArgs()
    present = 0
    for k, v in kw.items():
        key = "lambda_" if k == "lambda" else k
        present |= 1 << _A[key]
        setattr(a, key, v)
    a.present = present
    p = Proj()
    rc = lib().orc_map_set(int(code), C.byref(p), C.byref(a))
    if rc != 0:
        raise ValueError("map_set error %d" % rc)
    return p


def compute_nest_locations(iproj, nests, **coarse_kw):
    """compute_nest_locations (llxy_module.F:331-593); same arguments as wps_b200.capi.compute_nest_locations."""
    a = MapArgs()
    present = 0
    for k, v in coarse_kw.items():
        key = "lambda_" if k == "lambda" else k
        present |= 1 << _A[key]
        setattr(a, key, v)
    a.present = present
    n = len(nests)
    arr = lambda col: (C.c_int * n)(*[int(t[col]) for t in nests])
    out = (Proj * n)()
    rc = lib().orc_compute_nest_locations(int(iproj), C.byref(a), n, arr(0), arr(1), arr(2), arr(3), arr(4), arr(5), out)
    if rc != 0:
        raise ValueError("compute_nest_locations error %d" % rc)
    return [out[k] for k in range(n)]


def push_source_projection(iproj, stand_lon=0., truelat1=0., truelat2=0., dxkm=0., dykm=0., dlat=0., dlon=0.,
                           known_x=1., known_y=1., known_lat=0., known_lon=0., pole_lat=None, pole_lon=None,
                           centerlat=None, centerlon=None, centeri=None, centerj=None, earth_radius=None):
    p = Proj()
    has_cass = pole_lat is not None
    f = C.c_float
    rc = lib().orc_push_source_projection(
        C.byref(p), int(iproj), f(stand_lon), f(truelat1), f(truelat2), f(dxkm), f(dykm), f(dlat), f(dlon),
        f(known_x), f(known_y), f(known_lat), f(known_lon), int(has_cass), f(pole_lat or 0.), f(pole_lon or 0.),
        f(centerlat or 0.), f(centerlon or 0.), f(centeri or 0.), f(centerj or 0.),
        int(earth_radius is not None), f(earth_radius or 0.))
    if rc != 0:
        raise ValueError("push_source_projection error %d" % rc)
    return p


def lltoxy(p, lat, lon, stagger=M):
    lat = np.ascontiguousarray(lat, np.float32)
    lon = np.ascontiguousarray(lon, np.float32)
    x = np.empty_like(lat)
    y = np.empty_like(lat)
    lib().orc_lltoxy_array(C.byref(p), C.c_long(lat.size), fptr(lat), fptr(lon), fptr(x), fptr(y), int(stagger))
    return x, y


def xytoll(p, x, y, stagger=M):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    lat = np.empty_like(x)
    lon = np.empty_like(x)
    lib().orc_xytoll_array(C.byref(p), C.c_long(x.size), fptr(x), fptr(y), fptr(lat), fptr(lon), int(stagger))
    return lat, lon


def parse_interp_option(s, maxn=16):
    codes = (C.c_int * maxn)()
    opts = (C.c_int * maxn)()
    n = lib().orc_interp_array_from_string(s.encode(), codes, maxn)
    n2 = lib().orc_interp_options_from_string(s.encode(), opts, maxn)
    if n < 0 or n2 < 0:
        raise ValueError("bad interp_option %r" % s)
    return list(codes[:n]), list(opts[:n])


def gcell_threshold(s):
    t = C.c_float()
    rc = lib().orc_get_gcell_threshold(s.encode(), C.byref(t))
    if rc != 0:
        raise ValueError("bad average_gcell threshold in %r" % s)
    return t.value


def _chain(chain):
    if isinstance(chain, str):
        codes, opts = parse_interp_option(chain)
    else:
        codes, opts = chain
    codes = np.ascontiguousarray(list(codes) + [0], np.int32)
    opts = np.ascontiguousarray(list(opts) + [0], np.int32)
    return codes, opts


def interp_sequence(xx, yy, array, sx, sy, msgval, chain, izz=1, sz=1, mask=None, rel=' ', maskval=0.0):
    """Vectorised interp_sequence (interp_module.F:304) over arrays xx, yy.
    array is [nz][ny][nx] (x fastest) with lower bounds (sx, sy, sz)."""
    array = np.ascontiguousarray(array, np.float32)
    if array.ndim == 2:
        array = array[None]
    nz, ny, nx = array.shape
    xx = np.ascontiguousarray(xx, np.float32)
    yy = np.ascontiguousarray(yy, np.float32)
    out = np.empty_like(xx)
    codes, opts = _chain(chain)
    m = None if mask is None else np.ascontiguousarray(mask, np.float32)
    lib().orc_interp_sequence_array(
        C.c_long(xx.size), fptr(xx), fptr(yy), int(izz), fptr(array), int(sx), int(sx + nx - 1), int(sy),
        int(sy + ny - 1), int(sz), int(sz + nz - 1), C.c_float(msgval), iptr(codes), iptr(opts),
        int(m is not None), C.c_char(rel.encode()), C.c_float(maskval), fptr(m), fptr(out))
    return out


def field_opts(chain, missing_value=NAN, fill_missing=NAN, masked=-2.0, mask_val=(NAN, NAN, NAN), mask_rel="   "):
    fo = FieldOpts()
    codes, opts = _chain(chain)
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


def interp_met_field(src, fo, ifieldstagger, xlat, xlon, sm1, sm2, sd, slab, minx, miny, bdr,
                     mask=None, land_mask=None, water_mask=None, landmask=None, process_bdy_width=0,
                     r_arr=None, valid_mask=None, sub=None):
    """interp_met_field non-gcell branch (process_domain_module.F:2483-2581).
    xlat/xlon [ny][nx] with lower bounds (sm1, sm2); sd=(sd1,ed1,sd2,ed2); slab [sny][snx] lower bounds (minx,miny).
    Returns (r_arr, new_pts words)."""
    xlat = np.ascontiguousarray(xlat, np.float32)
    xlon = np.ascontiguousarray(xlon, np.float32)
    ny, nx = xlat.shape
    slab = np.ascontiguousarray(slab, np.float32)
    sny, snx = slab.shape
    if r_arr is None:
        r_arr = np.zeros((ny, nx), np.float32)
    if valid_mask is None:
        valid_mask = np.zeros((ny, nwords(nx)), np.int32)
    new_pts = np.zeros((ny, nwords(nx)), np.int32)
    em1, em2 = sm1 + nx - 1, sm2 + ny - 1
    if sub is None:
        sub = (sm1, em1, sm2, em2)
    cm = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    mask, land_mask, water_mask, landmask = cm(mask), cm(land_mask), cm(water_mask), cm(landmask)
    lib().orc_interp_met_field(
        C.byref(src), C.byref(fo), int(ifieldstagger), fptr(xlat), fptr(xlon), sm1, em1, sm2, em2,
        int(sd[0]), int(sd[1]), int(sd[2]), int(sd[3]), fptr(slab), int(minx), int(minx + snx - 1), int(miny),
        int(miny + sny - 1), int(bdr), fptr(mask), fptr(land_mask), fptr(water_mask), fptr(landmask),
        int(process_bdy_width), fptr(r_arr), iptr(valid_mask), iptr(new_pts),
        int(sub[0]), int(sub[1]), int(sub[2]), int(sub[3]))
    return r_arr, new_pts


def interp_met_field_gcell(src, dom, fo, ifieldstagger, xlat, xlon, sm1, sm2, sd, slab, minx, miny, bdr,
                           mask=None, landmask=None, process_bdy_width=0, r_arr=None, valid_mask=None):
    xlat = np.ascontiguousarray(xlat, np.float32)
    xlon = np.ascontiguousarray(xlon, np.float32)
    ny, nx = xlat.shape
    slab = np.ascontiguousarray(slab, np.float32)
    sny, snx = slab.shape
    if r_arr is None:
        r_arr = np.zeros((ny, nx), np.float32)
    if valid_mask is None:
        valid_mask = np.zeros((ny, nwords(nx)), np.int32)
    new_pts = np.zeros((ny, nwords(nx)), np.int32)
    cm = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    mask, landmask = cm(mask), cm(landmask)
    lib().orc_interp_met_field_gcell(
        C.byref(src), C.byref(dom), C.byref(fo), int(ifieldstagger), fptr(xlat), fptr(xlon), sm1, sm1 + nx - 1,
        sm2, sm2 + ny - 1, int(sd[0]), int(sd[1]), int(sd[2]), int(sd[3]), fptr(slab), int(minx),
        int(minx + snx - 1), int(miny), int(miny + sny - 1), int(bdr), fptr(mask), fptr(landmask),
        int(process_bdy_width), fptr(r_arr), iptr(valid_mask), iptr(new_pts))
    return r_arr, new_pts


def metmap_xform(proj, u, u_mask, v, v_mask, xlon_u, xlon_v, idir, us=(1, 1), vs=(1, 1), xlat_u=None, xlat_v=None):
    """metmap_xform (rotate_winds_module.F:85); u [uny][unx], v [vny][vnx]; in place on copies.
    xlat_u / xlat_v are needed for PROJ_CASSINI only (:174-222, :281-329)."""
    u = np.array(u, np.float32, order="C")
    v = np.array(v, np.float32, order="C")
    uny, unx = u.shape
    vny, vnx = v.shape
    if xlat_u is not None:
        lib().orc_metmap_xform_ll(C.byref(proj), fptr(u), iptr(np.ascontiguousarray(u_mask, np.int32)), fptr(v),
                                  iptr(np.ascontiguousarray(v_mask, np.int32)), us[0], us[1], us[0] + unx - 1,
                                  us[1] + uny - 1, vs[0], vs[1], vs[0] + vnx - 1, vs[1] + vny - 1,
                                  fptr(np.ascontiguousarray(xlon_u, np.float32)), fptr(np.ascontiguousarray(xlon_v, np.float32)),
                                  fptr(np.ascontiguousarray(xlat_u, np.float32)), fptr(np.ascontiguousarray(xlat_v, np.float32)),
                                  int(idir))
        return u, v
    lib().orc_metmap_xform(C.byref(proj), fptr(u), iptr(np.ascontiguousarray(u_mask, np.int32)), fptr(v),
                           iptr(np.ascontiguousarray(v_mask, np.int32)), us[0], us[1], us[0] + unx - 1,
                           us[1] + uny - 1, vs[0], vs[1], vs[0] + vnx - 1, vs[1] + vny - 1,
                           fptr(np.ascontiguousarray(xlon_u, np.float32)),
                           fptr(np.ascontiguousarray(xlon_v, np.float32)), int(idir))
    return u, v


def smooth(array, npass, opt):
    """smooth_module.F:20-239 on array [nz][ny][nx]; returns smoothed copy."""
    a = np.array(array, np.float32, order="C")
    if a.ndim == 2:
        a = a[None]
    nz, ny, nx = a.shape
    lib().orc_smooth(fptr(a), 1, nx, 1, ny, 1, nz, int(npass), int(opt))
    return a.reshape(np.shape(array))


# ---- geogrid ----
def accum_categorical(src, dom, istagger, tile, src_min_x, src_min_y, bdr, dst, start_i, start_j, start_k,
                      processed, new_pts, ilevel, msgval, sr_x=1.0, sr_y=1.0):
    """proc_point_module.F:100-206 for one tile; tile [ny][nx]; dst [nk][ny][nx] modified in place."""
    tile = np.ascontiguousarray(tile, np.float32)
    ty, tx = tile.shape[-2:]
    nk, ny, nx = dst.shape
    inv = C.c_long(0)
    lib().orc_accum_categorical(C.byref(src), C.byref(dom), int(istagger), fptr(tile), int(src_min_x),
                                int(src_min_x + tx - 1), int(src_min_y), int(src_min_y + ty - 1), int(bdr), fptr(dst),
                                int(start_i), int(start_i + nx - 1), int(start_j), int(start_j + ny - 1), int(start_k),
                                int(start_k + nk - 1), iptr(processed), iptr(new_pts), int(ilevel), C.c_float(msgval),
                                C.c_float(sr_x), C.c_float(sr_y), C.byref(inv))
    return inv.value


def accum_continuous(src, dom, istagger, tile, src_min_x, src_min_y, src_min_z, bdr, dst, n, start_i, start_j,
                     start_k, processed, new_pts, msgval, sr_x=1.0, sr_y=1.0):
    tile = np.ascontiguousarray(tile, np.float32)
    tz, ty, tx = tile.shape
    nk, ny, nx = dst.shape
    lib().orc_accum_continuous(C.byref(src), C.byref(dom), int(istagger), fptr(tile), int(src_min_x),
                               int(src_min_x + tx - 1), int(src_min_y), int(src_min_y + ty - 1), int(src_min_z),
                               int(src_min_z + tz - 1), int(bdr), fptr(dst), fptr(n), int(start_i),
                               int(start_i + nx - 1), int(start_j), int(start_j + ny - 1), int(start_k),
                               int(start_k + nk - 1), iptr(processed), iptr(new_pts), C.c_float(msgval),
                               C.c_float(sr_x), C.c_float(sr_y))


def tile_point_fallback(src, itype, tile, src_min_x, src_min_y, src_min_z, bdr, xlat, xlon, landmask, use_landmask,
                        mask_val, chain, msg_val, msg_fill_val, field, data_count, start_i, start_j, start_k,
                        processed, level):
    tile = np.ascontiguousarray(tile, np.float32)
    tz, ty, tx = tile.shape
    nk, ny, nx = field.shape
    codes, opts = _chain(chain)
    inv = C.c_long(0)
    lib().orc_tile_point_fallback(C.byref(src), int(itype), fptr(tile), int(src_min_x), int(src_min_x + tx - 1),
                                  int(src_min_y), int(src_min_y + ty - 1), int(src_min_z), int(src_min_z + tz - 1),
                                  int(bdr), fptr(np.ascontiguousarray(xlat, np.float32)),
                                  fptr(np.ascontiguousarray(xlon, np.float32)),
                                  fptr(None if landmask is None else np.ascontiguousarray(landmask, np.float32)),
                                  int(use_landmask), C.c_float(mask_val), iptr(codes), iptr(opts), C.c_float(msg_val),
                                  C.c_float(msg_fill_val), fptr(field), fptr(data_count), int(start_i),
                                  int(start_i + nx - 1), int(start_j), int(start_j + ny - 1), int(start_k),
                                  int(start_k + nk - 1), iptr(processed), iptr(level), C.byref(inv))
    return inv.value


def calc_field_finish(itype, field, data_count, landmask, use_landmask, mask_val, msg_fill_val, scale_factor,
                      start_i, start_j, start_k, processed, level):
    nk, ny, nx = field.shape
    lib().orc_calc_field_finish(int(itype), fptr(field), fptr(data_count),
                                fptr(None if landmask is None else np.ascontiguousarray(landmask, np.float32)),
                                int(use_landmask), C.c_float(mask_val), C.c_float(msg_fill_val),
                                int(scale_factor is not None), C.c_float(scale_factor or 1.0), int(start_i),
                                int(start_i + nx - 1), int(start_j), int(start_j + ny - 1), int(start_k),
                                int(start_k + nk - 1), iptr(processed), iptr(level))


def fractions(field, msg_fill_val):
    f = np.array(field, np.float32, order="C")
    ncat = f.shape[0]
    lib().orc_fractions(fptr(f), C.c_long(int(np.prod(f.shape[1:]))), int(ncat), C.c_float(msg_fill_val))
    return f


def landmask_from_fractions(field, min_cat, lm_vals, is_water_mask, msg_fill_val):
    f = np.ascontiguousarray(field, np.float32)
    ncat = f.shape[0]
    out = np.empty(f.shape[1:], np.float32)
    lm = np.ascontiguousarray(lm_vals, np.int32)
    lib().orc_landmask_from_fractions(fptr(f), C.c_long(out.size), int(min_cat), int(min_cat + ncat - 1), iptr(lm),
                                      int(lm.size), int(is_water_mask), C.c_float(msg_fill_val), fptr(out))
    return out


def dominant_category(field, min_cat, msg_fill_val):
    f = np.ascontiguousarray(field, np.float32)
    out = np.empty(f.shape[1:], np.float32)
    lib().orc_dominant_category(fptr(f), C.c_long(out.size), int(min_cat), int(min_cat + f.shape[0] - 1),
                                C.c_float(msg_fill_val), fptr(out))
    return out


def dominant_category_landmask(field, landmask, min_cat, lm_vals, is_water_mask):
    f = np.ascontiguousarray(field, np.float32)
    out = np.empty(f.shape[1:], np.float32)
    lm = np.ascontiguousarray(lm_vals, np.int32)
    lib().orc_dominant_category_landmask(fptr(f), fptr(np.ascontiguousarray(landmask, np.float32)),
                                         C.c_long(out.size), int(min_cat), int(min_cat + f.shape[0] - 1), iptr(lm),
                                         int(lm.size), int(is_water_mask), fptr(out))
    return out


def get_lat_lon_fields(dom, si, sj, nx, ny, stagger, sub_x=1.0, sub_y=1.0):
    xlat = np.empty((ny, nx), np.float32)
    xlon = np.empty((ny, nx), np.float32)
    lib().orc_get_lat_lon_fields(C.byref(dom), fptr(xlat), fptr(xlon), int(si), int(sj), int(si + nx - 1),
                                 int(sj + ny - 1), int(stagger), C.c_float(sub_x), C.c_float(sub_y))
    return xlat, xlon


def get_map_factor(dom, xlat, xlon, truelat1, truelat2, pole_lat=90.0, pole_lon=0.0, stand_lon=0.0):
    xlat = np.ascontiguousarray(xlat, np.float32)
    xlon = np.ascontiguousarray(xlon, np.float32)
    mx = np.empty_like(xlat)
    my = np.empty_like(xlat)
    f = C.c_float
    lib().orc_get_map_factor(C.byref(dom), fptr(xlat), fptr(xlon), fptr(mx), fptr(my), C.c_long(xlat.size),
                             f(truelat1), f(truelat2), f(pole_lat), f(pole_lon), f(stand_lon))
    return mx, my


def get_coriolis(xlat):
    xlat = np.ascontiguousarray(xlat, np.float32)
    fo = np.empty_like(xlat)
    e = np.empty_like(xlat)
    lib().orc_get_coriolis(fptr(xlat), fptr(fo), fptr(e), C.c_long(xlat.size))
    return fo, e


def get_rotang(iproj, xlat, xlon, truelat1, truelat2, stand_lon):
    xlat = np.ascontiguousarray(xlat, np.float32)
    xlon = np.ascontiguousarray(xlon, np.float32)
    ny, nx = xlat.shape
    cosa = np.empty_like(xlat)
    sina = np.empty_like(xlat)
    f = C.c_float
    lib().orc_get_rotang(int(iproj), fptr(xlat), fptr(xlon), fptr(cosa), fptr(sina), 1, 1, nx, ny, f(truelat1),
                         f(truelat2), f(stand_lon))
    return cosa, sina


def calc_dfd(which, src, sr, mapfac, dom_d):
    src = np.ascontiguousarray(src, np.float32)
    a3 = src if src.ndim == 3 else src[None]
    nz, ny, nx = a3.shape
    dst = np.empty_like(src)
    fn = lib().orc_calc_dfdx if which == "x" else lib().orc_calc_dfdy
    fn(fptr(src), fptr(dst), 1, 1, 1, nx, ny, nz, int(sr), fptr(None if mapfac is None else np.ascontiguousarray(mapfac, np.float32)), C.c_float(dom_d))
    return dst


def search_order(x0, y0, sx, ex, sy, ey, maxdepth, maxn=100000):
    xs = np.zeros(maxn, np.int32)
    ys = np.zeros(maxn, np.int32)
    n = lib().orc_search_order(int(x0), int(y0), int(sx), int(ex), int(sy), int(ey), int(maxdepth), iptr(xs), iptr(ys), maxn)
    return xs[:n], ys[:n]


def fill_missing_levels(levels, r_arr, valid):
    """fill_missing_levels vertical interpolation (process_domain_module.F:2862-2921); r_arr [nlev][ny][nx],
    valid [nlev][ny][nxw]; returns updated copies."""
    r = np.array(r_arr, np.float32, order="C")
    v = np.array(valid, np.int32, order="C")
    nlev, ny, nx = r.shape
    lv = np.ascontiguousarray(levels, np.int32)
    lib().orc_fill_missing_levels(int(nlev), lv.ctypes.data_as(C.c_void_p), fptr(r), iptr(v), int(nx), int(ny))
    return r, v


# ---- MPAS remapping (orc_mpas.c) ----
_MESH_INT = ("landmask", "nEdgesOnCell", "cellsOnCell", "verticesOnCell", "edgesOnCell", "cellsOnVertex", "cellsOnEdge")
_MESH_REAL = ("latCell", "lonCell", "latVertex", "lonVertex", "latEdge", "lonEdge")


class MpasMesh(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nCells", "nVertices", "nEdges", "maxEdges")] + [(n, C.c_void_p) for n in _MESH_INT + _MESH_REAL]


class MpasRemap(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("nearestCell", "nearestVertex", "nearestEdge", "sourceCells", "cellWeights", "sourceVertices",
                                          "vertexWeights", "sourceEdges", "edgeWeights", "sourceMaskedCells", "cellMaskedWeights")]


def _mpas_mesh(mesh):
    d = MpasMesh(nCells=mesh["nCells"], nVertices=mesh["nVertices"], nEdges=mesh["nEdges"], maxEdges=mesh["maxEdges"])
    keep = []
    for n in _MESH_INT + _MESH_REAL:
        a = np.ascontiguousarray(mesh[n], dtype=np.int32 if n in _MESH_INT else np.float32)
        keep.append(a)
        setattr(d, n, a.ctypes.data)
    return d, keep


def mpas_remap_setup(mesh, lat_rad, lon_rad):
    """remap_info_setup (remapper.F:83-286), sequential scan as in the reference; returns a dict of the tables."""
    d, keep = _mpas_mesh(mesh)
    nlat, nlon = lat_rad.shape
    me = mesh["maxEdges"]
    t = dict(nearestCell=np.zeros((nlat, nlon), np.int32), nearestVertex=np.zeros((nlat, nlon), np.int32),
             nearestEdge=np.zeros((nlat, nlon), np.int32),
             sourceCells=np.zeros((nlat, nlon, 3), np.int32), cellWeights=np.zeros((nlat, nlon, 3), np.float32),
             sourceVertices=np.zeros((nlat, nlon, me), np.int32), vertexWeights=np.zeros((nlat, nlon, me), np.float32),
             sourceEdges=np.zeros((nlat, nlon, me), np.int32), edgeWeights=np.zeros((nlat, nlon, me), np.float32),
             sourceMaskedCells=np.zeros((nlat, nlon, 3), np.int32), cellMaskedWeights=np.zeros((nlat, nlon, 3), np.float32))
    r = MpasRemap(**{k: v.ctypes.data for k, v in t.items()})
    lib().orc_mpas_remap_setup(C.byref(d), fptr(np.ascontiguousarray(lat_rad)), fptr(np.ascontiguousarray(lon_rad)), nlon, nlat, C.byref(r))
    return t


def mpas_remap_field(sources, weights, src, nw=0):
    """remap_field for REAL / DOUBLE sources [nNodes][nlev] with the given tables [nlat][nlon][width]."""
    nlat, nlon, tw = sources.shape
    src2 = src.reshape(src.shape[0], -1)
    nlev = src2.shape[1]
    if nw == 0:
        nw = tw if src.ndim == 1 else 3
    out = np.zeros((nlat, nlon, nlev), src.dtype)
    fn = lib().orc_mpas_remap_field_real if src.dtype == np.float32 else lib().orc_mpas_remap_field_double
    fn(iptr(np.ascontiguousarray(sources)), fptr(np.ascontiguousarray(weights)), tw, nw, C.c_long(nlat * nlon),
       C.c_void_p(np.ascontiguousarray(src2).ctypes.data), nlev, C.c_void_p(out.ctypes.data))
    return out


def mpas_remap_field_int(nearest, src):
    nlat, nlon = nearest.shape
    src2 = np.ascontiguousarray(src.reshape(src.shape[0], -1), dtype=np.int32)
    out = np.zeros((nlat, nlon, src2.shape[1]), np.int32)
    lib().orc_mpas_remap_field_int(iptr(np.ascontiguousarray(nearest)), C.c_long(nlat * nlon), iptr(src2), src2.shape[1], iptr(out))
    return out


def mpas_extract_planes(wrf, mode=0, landmask=None, fill=0.0):
    """process_mpas_fields' per-level planes of a remapped field wrf [nlat][nlon][nlev] -> [nlev][nlat][nlon]."""
    nlat, nlon, nlev = wrf.shape
    out = np.zeros((nlev, nlat, nlon), np.float32)
    lib().orc_mpas_extract_planes(fptr(np.ascontiguousarray(wrf)), nlev, C.c_long(nlat * nlon), int(mode),
                                  fptr(None if landmask is None else np.ascontiguousarray(landmask, dtype=np.float32)), C.c_float(fill), fptr(out))
    return out


def metmap_xform_nmm(proj, u, u_mask, v, v_mask, xlat_v, xlon_v, idir):
    """metmap_xform_nmm (rotate_winds_module.F:455-631): one level, U and V on the E-grid's velocity points; returns copies."""
    ny, nx = u.shape
    uo, vo = np.ascontiguousarray(u).copy(), np.ascontiguousarray(v).copy()
    lib().orc_metmap_xform_nmm(C.byref(proj), fptr(uo), iptr(np.ascontiguousarray(u_mask)), fptr(vo), iptr(np.ascontiguousarray(v_mask)),
                               1, 1, nx, ny, fptr(np.ascontiguousarray(xlat_v)), fptr(np.ascontiguousarray(xlon_v)), int(idir))
    return uo, vo


def smooth_egrid(array, npass, opt, msgval=1.0e20, hflag=1.0, sx=1, sy=1):
    """one_two_one_egrid / smth_desmth_egrid (smooth_module.F:252-585) on array [nz][ny][nx] with lower bounds (sx, sy); returns a copy."""
    a = np.ascontiguousarray(array).copy()
    nz, ny, nx = a.shape
    lib().orc_smooth_egrid(fptr(a), int(sx), int(sx + nx - 1), int(sy), int(sy + ny - 1), 1, nz, int(npass), int(opt), C.c_float(msgval), C.c_float(hflag))
    return a


def mpas_derive_tt(theta, pressure):
    """derive_mpas_fields (process_domain_module.F:1957-2053) on one level; returns TT."""
    t = np.ascontiguousarray(theta).copy()
    lib().orc_mpas_derive_tt(fptr(t), fptr(np.ascontiguousarray(pressure)), C.c_long(t.size))
    return t
