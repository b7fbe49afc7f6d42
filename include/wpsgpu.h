/*
 * wpsgpu.h -- C ABI of the B200-native WPS horizontal-interpolation library (libwpsgpu.so).
 *
 * WPS v4.6.0 has no plugin / FFI interface: the path lives behind ordinary Fortran module
 * procedures.  The seam is therefore defined at their call sites; every entry
 * point below names the reference procedure it replaces (file:line in the WPS tree).  The
 * Fortran host binds these through ISO_C_BINDING (fortran/wpsgpu_mod.F90).
 *
 * Conventions (taken from the callers):
 *  - all arrays are caller-owned, Fortran column-major (x fastest) binary32 / int32; explicit
 *    lower/upper bounds are passed exactly as the Fortran dummy-argument bounds;
 *  - bit arrays are bitarray_module.F words: int32 words[ceil(nx/32)][ny], x-word fastest,
 *    bit (i-1) mod 32 of word (i-1)/32 (bitarray_module.F:91-94);
 *  - "no value" is in-band: missing_value / fill_missing (default NAN = 1.E20,
 *    misc_definitions_module.F:12) and the new_pts bits;
 *  - every function returns 0 on success, <0 on error (wpsgpu_last_error() gives the text);
 *    the Fortran wrapper turns non-zero into mprintf(.true.,ERROR,...) (module_debug.F:50);
 *  - one handle per MPI rank / GPU; a handle owns one CUDA stream, the device arena and the
 *    coordinate cache.  No call retains a host pointer after it (or the flush it belongs to) returns;
 *  - pointer mode: WPSGPU_PTR_HOST (default) = arrays are host memory, the library stages
 *    H2D/D2H on its stream; WPSGPU_PTR_DEVICE = arrays are device memory already (used to time
 *    the HBM-resident path and to chain calls without leaving the device).
 *  - there is NO CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef WPSGPU_H
#define WPSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* misc_definitions_module.F:12-47 */
#define WPSGPU_NAN 1.0e20f
#define WPSGPU_NOT_MASKED (-2.0f)
#define WPSGPU_MASKED_BOTH (-1.0f)
#define WPSGPU_MASKED_WATER 0.0f
#define WPSGPU_MASKED_LAND 1.0f
enum { WPSGPU_SIXTEEN_POINT = 1, WPSGPU_FOUR_POINT = 2, WPSGPU_N_NEIGHBOR = 3, WPSGPU_AVERAGE4 = 4,
       WPSGPU_AVERAGE16 = 5, WPSGPU_W_AVERAGE4 = 6, WPSGPU_W_AVERAGE16 = 7, WPSGPU_SEARCH = 8 };
enum { WPSGPU_M = 1, WPSGPU_U = 2, WPSGPU_V = 3, WPSGPU_HH = 4, WPSGPU_VV = 5, WPSGPU_CORNER = 6 };
enum { WPSGPU_ONETWOONE = 1, WPSGPU_SMTHDESMTH = 2, WPSGPU_SMTHDESMTH_SPECIAL = 3 };
enum { WPSGPU_CONTINUOUS = 0, WPSGPU_CATEGORICAL = 1, WPSGPU_SP_CONTINUOUS = 2 };
enum { WPSGPU_PROJ_LATLON = 0, WPSGPU_PROJ_LC = 1, WPSGPU_PROJ_PS = 2, WPSGPU_PROJ_PS_WGS84 = 102,
       WPSGPU_PROJ_MERC = 3, WPSGPU_PROJ_GAUSS = 4, WPSGPU_PROJ_CYL = 5, WPSGPU_PROJ_CASSINI = 6,
       WPSGPU_PROJ_ALBERS_NAD83 = 105, WPSGPU_PROJ_ROTLL = 203 };
enum { WPSGPU_PTR_HOST = 0, WPSGPU_PTR_DEVICE = 1 };
#define WPSGPU_MAX_OPS 16
#define WPSGPU_MAX_GAUSS 2048

typedef struct wpsgpu_ctx *wpsgpu_handle_t;

/* proj_info (module_map_utils.F:140-192), populated by wpsgpu_map_set */
typedef struct wpsgpu_proj {
    int32_t code, nlat, nxmin, nxmax, stagger, comp_ll, init, n_gauss;
    float lat1, lon1, lat0, lon0, dx, dy, latinc, loninc, dlon, stdlon, truelat1, truelat2, hemi, cone,
          polei, polej, rsw, rebydx, knowni, knownj, re_m, rho0, nc, bigc;
    float gauss_lat[WPSGPU_MAX_GAUSS];
    int32_t ixdim, jydim;       /* PROJ_ROTLL (NMM E-grid): mass points per row / number of rows, module_map_utils.F:150-153 */
    float phi, lambda;          /* PROJ_ROTLL: half extent of the rotated grid in latitude / longitude, degrees */
} wpsgpu_proj;

/* OPTIONAL arguments of map_set (module_map_utils.F:243-283); bit set in `present` = argument passed */
enum { WPSGPU_A_LAT1 = 1 << 0, WPSGPU_A_LON1 = 1 << 1, WPSGPU_A_LAT0 = 1 << 2, WPSGPU_A_LON0 = 1 << 3,
       WPSGPU_A_KNOWNI = 1 << 4, WPSGPU_A_KNOWNJ = 1 << 5, WPSGPU_A_DX = 1 << 6, WPSGPU_A_DY = 1 << 7,
       WPSGPU_A_LATINC = 1 << 8, WPSGPU_A_LONINC = 1 << 9, WPSGPU_A_STDLON = 1 << 10,
       WPSGPU_A_TRUELAT1 = 1 << 11, WPSGPU_A_TRUELAT2 = 1 << 12, WPSGPU_A_NLAT = 1 << 13,
       WPSGPU_A_IXDIM = 1 << 15, WPSGPU_A_JYDIM = 1 << 16, WPSGPU_A_PHI = 1 << 20, WPSGPU_A_LAMBDA = 1 << 21,
       WPSGPU_A_NXMIN = 1 << 17, WPSGPU_A_NXMAX = 1 << 18, WPSGPU_A_STAGGER = 1 << 19,
       WPSGPU_A_R_EARTH = 1 << 22 };
typedef struct wpsgpu_map_args {
    uint32_t present;
    float lat1, lon1, lat0, lon0, knowni, knownj, dx, dy, latinc, loninc, stdlon, truelat1, truelat2, r_earth;
    int32_t nlat, nxmin, nxmax, stagger;
    int32_t ixdim, jydim;       /* PROJ_ROTLL */
    float phi, lambda;
} wpsgpu_map_args;

/* ---- lifetime / errors ---- */
int wpsgpu_create(int device_id, wpsgpu_handle_t *out);
int wpsgpu_destroy(wpsgpu_handle_t h);
const char *wpsgpu_last_error(void);
int wpsgpu_set_pointer_mode(wpsgpu_handle_t h, int mode);
int wpsgpu_synchronize(wpsgpu_handle_t h);
/* the handle's CUDA stream (cudaStream_t as void*), so callers can record CUDA events on it */
void *wpsgpu_stream(wpsgpu_handle_t h);
/* number of library kernels launched on this handle since creation (bench.py "gpu_launches") */
int64_t wpsgpu_launch_count(wpsgpu_handle_t h);

/* ---- projections ---- */
/* map_set (module_map_utils.F:243-567) */
int wpsgpu_map_set(int proj_code, const wpsgpu_map_args *args, wpsgpu_proj *out);
/* compute_nest_locations + find_known_latlon (geogrid/src/llxy_module.F:331-593): the projection of every domain of a nested
 * configuration.  `coarse` holds the map_set arguments of domain 1 exactly as :343-438 passes them (known_lat/known_lon as
 * lat1/lon1, known_x/known_y as knowni/knownj, dxkm as dx).  The arrays have n_domains entries
 * (entry 0 = domain 1): parent_id (1-based), parent_grid_ratio, i_parent_start / j_parent_start (gridinfo_module.F:581
 * parent_ll_x/y), ixdim = e_we - s_we + 1 and jydim = e_sn - s_sn + 1 (:461-462).  out[k] receives proj_stack(k + 1): nests
 * are located through their centre point (ixdim/2, jydim/2) mapped up the parent chain, grid distances divided by the ratios.
 * Host-only: no device work.  Lambert conformal, polar stereographic (both), Mercator and Albers; the reference's own
 * PROJ_LATLON branch omits map_set's mandatory knowni/knownj and stops in MAP_INIT, Cassini / rotated lat-lon are not built. */
int wpsgpu_compute_nest_locations(int iproj, const wpsgpu_map_args *coarse, int n_domains, const int32_t *parent_id,
                                  const int32_t *parent_grid_ratio, const int32_t *i_parent_start,
                                  const int32_t *j_parent_start, const int32_t *ixdim, const int32_t *jydim,
                                  wpsgpu_proj *out);
/* push_source_projection (llxy_module.F:39-165); cassini[6] = {pole_lat,pole_lon,centerlat,centerlon,
 * centeri,centerj} or NULL; earth_radius NULL = absent */
int wpsgpu_push_source_projection(int iproj, float stand_lon, float truelat1, float truelat2, float dxkm,
                                  float dykm, float dlat, float dlon, float known_x, float known_y,
                                  float known_lat, float known_lon, const float *cassini,
                                  const float *earth_radius, wpsgpu_proj *out);
/* lltoxy / xytoll (llxy_module.F:788-887) over n points */
int wpsgpu_lltoxy(wpsgpu_handle_t h, const wpsgpu_proj *p, int stagger, const float *lat, const float *lon,
                  int64_t n, float *x, float *y);
int wpsgpu_xytoll(wpsgpu_handle_t h, const wpsgpu_proj *p, int stagger, const float *x, const float *y,
                  int64_t n, float *lat, float *lon);

/* ---- interp_option chains ---- */
/* interp_array_from_string + interp_options_from_string + get_search_depth (interp_module.F:20-295)
 * and get_gcell_threshold (interp_option_module.F:692).  codes/opts receive *n entries (0-terminated,
 * *n = number of '+'-separated tokens + 1).  has_gcell=1 when an average_gcell token is present. */
int wpsgpu_parse_interp_option(const char *interp_string, int32_t *codes, int32_t *opts, int maxn, int *n,
                               int *has_gcell, float *gcell_threshold);

/* ---- metgrid: batched replacement of interp_met_field (process_domain_module.F:2205-2586) ---- */
typedef struct wpsgpu_field_opts {
    int32_t ops[WPSGPU_MAX_OPS];   /* interp_array_from_string(interp_method(idx)) */
    int32_t opts[WPSGPU_MAX_OPS];  /* interp_options_from_string(interp_method(idx)) */
    float missing_value;           /* missing_value(idx) */
    float fill_missing;            /* fill_missing(idx) */
    float masked;                  /* masked(idx): -2 none, -1 both, 0 water, 1 land */
    float mask_val[3];             /* interp_mask_val, interp_land_mask_val, interp_water_mask_val */
    char mask_rel[4];              /* the three *_relational characters ' ', '<', '>' */
} wpsgpu_field_opts;

/* Destination grid `grid_id` (0..7): the xlat/xlon(sm1:em1,sm2:em2) pair interp_met_field receives
 * (xlat, xlat_u, xlat_v ..., process_domain_module.F:1246-1357).  Replaces nothing by itself; it
 * lets the per-(domain,stagger,source projection) coordinates be computed once and cached on device. */
int wpsgpu_metgrid_set_grid(wpsgpu_handle_t h, int grid_id, const float *xlat, const float *xlon,
                            int sm1, int em1, int sm2, int em2);
/* landmask(sm1:em1,sm2:em2) of grid `grid_id` (the optional LANDMASK argument, :1350) */
int wpsgpu_metgrid_set_landmask(wpsgpu_handle_t h, int grid_id, const float *landmask);
/* we_domain_s, we_domain_e, sn_domain_s, sn_domain_e (sd1,ed1,sd2,ed2 of interp_met_field) */
int wpsgpu_metgrid_set_domain(wpsgpu_handle_t h, int sd1, int ed1, int sd2, int ed2);

typedef struct wpsgpu_slab {
    const float *slab;              /* slab(minx:maxx,miny:maxy), halo included as the caller built it (:1121-1135) */
    int32_t minx, maxx, miny, maxy, bdr;
    const float *mask[3];           /* "<NAME>.mask" slabs for interp_mask / interp_land_mask / interp_water_mask,
                                       same bounds as slab; NULL when storage_get_field status /= 0 (:2293-2322) */
    int32_t grid_id;                /* which destination grid (xlat/xlon pair) */
    int32_t field_stagger;          /* ifieldstagger argument (M,U,V,VV,HH) */
    int32_t use_landmask;           /* 1 when the optional landmask argument is passed (:1350,:1357) */
    int32_t process_bdy_width;
    float *r_arr;                   /* field%r_arr(sm1:em1,sm2:em2), inout */
    const int32_t *valid_mask;      /* field%valid_mask words, NULL = freshly created (all clear) */
    int32_t *new_pts;               /* field%modified_mask words, out (caller passes a zeroed array) */
    /* Raw-record ingest (read_met_module.F:375-386 + halo_slab, process_domain_module.F:1119-1140), optional:
     * raw_nx > 0 says `slab` is not the halo'd array but the record's slab(1:raw_nx, miny:maxy) exactly as it sits in
     * the intermediate file -- 4-byte words, big-endian when raw_big_endian != 0.  The library byte-swaps it on the
     * device and builds slab(minx:maxx) with bdr = (maxx-minx+1-raw_nx)/2 wrap-around columns on each side (bdr may
     * be 0 for regional data), so the host neither converts nor copies the slab.  raw_nx == 0: `slab` is ready. */
    int32_t raw_nx;
    int32_t raw_big_endian;
} wpsgpu_slab;

/* Queue one interp_met_field call (non-average_gcell branch, :2483-2581). */
int wpsgpu_metgrid_enqueue_slab(wpsgpu_handle_t h, const wpsgpu_proj *src_proj, const wpsgpu_field_opts *fo,
                                const wpsgpu_slab *slab);
/* Queue n calls that share projection and options -- the levels of one field (the `do k` record loop around
 * :1346-1360 visits them one after another); equivalent to n wpsgpu_metgrid_enqueue_slab calls in array order. */
int wpsgpu_metgrid_enqueue_slabs(wpsgpu_handle_t h, const wpsgpu_proj *src_proj, const wpsgpu_field_opts *fo,
                                 const wpsgpu_slab *slabs, int32_t n);
/* Execute everything queued as ONE batched launch over field x level (x time) and fill r_arr / new_pts. */
int wpsgpu_metgrid_flush(wpsgpu_handle_t h);
/* Same without waiting for the device: returns once everything is queued.  The caller's input and output arrays must
 * stay allocated and untouched until wpsgpu_synchronize(h) returns (then r_arr / new_pts are complete); with pinned
 * host buffers the host is free to read and decode the next file meanwhile.  Further calls on the handle are ordered
 * behind the pending copies automatically. */
int wpsgpu_metgrid_flush_async(wpsgpu_handle_t h);
/* Bytes the host-pointer mode has moved over PCIe for metgrid flushes since the handle was created (slabs, masks,
 * carried-over fields in; r_arr and new_pts out; the small job tables are not counted).  Chains without `search` only
 * need the rectangle of the source slab their destination grid can touch, so h2d is usually far below the slab sizes. */
int wpsgpu_metgrid_copy_bytes(wpsgpu_handle_t h, int64_t *h2d, int64_t *d2h);
/* Measurement aid, off by default: when on, every flush records the CUDA events wpsgpu_metgrid_last_kernel / _last_timings
 * read (a dozen per flush, ~2 % of a config-2 step); those two calls fail after a flush that was not timed. */
int wpsgpu_metgrid_set_timing(wpsgpu_handle_t h, int on);
/* Device time (CUDA events on the handle's stream) of the dominant kernel of the last flush and the number of
 * (slab x destination point) outputs it produced -- for roofline reporting only. */
int wpsgpu_metgrid_last_kernel(wpsgpu_handle_t h, float *ms, double *points);
/* Same, broken down (arrays of 7): [0] packing pre-pass, [1] packed sixteen_pt kernel, [2] generic kernel,
 * [3] four_pt / nearest_neighbor first-method kernels, [4] slow-bitmap pass (points the first-method kernels handed
 * back to the full fallback chain), [5] deferred search passes, [6] first launch to last kernel of the flush. */
int wpsgpu_metgrid_last_timings(wpsgpu_handle_t h, float *ms7, double *points7);
/* Arithmetic of chains that start with sixteen_pt (interp_module.F:1179-1374).
 * WPSGPU_PRECISION_EXACT (default): every operation of sixteen_pt / oned is rounded separately in the reference's
 *   order -- results are bit-identical to a build of the Fortran without FMA contraction.
 * WPSGPU_PRECISION_TOLERANCE: the general branch of oned is evaluated as a weighted sum of the 16 stencil values with
 *   fused multiply-adds (weights are level-independent); results agree with the reference to a few ulp of the largest
 *   stencil value (<= 1e-5 relative to it, the bound the project allows for sixteen_pt).  Points where the reference
 *   leaves the general branch -- missing or masked stencil values, exact zeros among the row interpolants, values below
 *   1e-21 in magnitude -- are detected conservatively and still take the exact code; all other operators, masks,
 *   new_pts bits and fill rules are unchanged and exact.  About 3x the throughput of the exact mode. */
enum { WPSGPU_PRECISION_EXACT = 0, WPSGPU_PRECISION_TOLERANCE = 1 };
int wpsgpu_metgrid_set_precision(wpsgpu_handle_t h, int mode);
/* Testing aid: disable!=0 routes every job through the generic kernel (used to test both kernels against the oracle). */
int wpsgpu_debug_disable_fast(wpsgpu_handle_t h, int disable);
/* The average_gcell branch (:2331-2478) with metgrid's accum_continuous (:3347-3717); dom_proj is the
 * projection set_domain_projection built (llxy_module.F:191).  The caller decides per slab whether the branch
 * applies (do_gcell_interp, :1198-1220).  Cell sums are accumulated in binary64 and rounded once (the reference adds
 * in binary32 in quadtree order), so cell averages agree to ~1e-7 relative rather than bit for bit; everything else
 * (which cells receive data, the fallback chain for the others, new_pts) is exact.
 * enqueue_gcell_slab queues the call for the next wpsgpu_metgrid_flush; gcell_slab queues it and flushes. */
int wpsgpu_metgrid_enqueue_gcell_slab(wpsgpu_handle_t h, const wpsgpu_proj *src_proj, const wpsgpu_proj *dom_proj,
                                      const wpsgpu_field_opts *fo, const wpsgpu_slab *slab);
int wpsgpu_metgrid_gcell_slab(wpsgpu_handle_t h, const wpsgpu_proj *src_proj, const wpsgpu_proj *dom_proj,
                              const wpsgpu_field_opts *fo, const wpsgpu_slab *slab);

/* fill_missing_levels, the vertical-interpolation loop (process_domain_module.F:2862-2921) over the nlev levels of
 * one 3-d field: a point whose valid_mask bit is clear at level k receives the linear interpolation in `levels`
 * between the nearest valid levels below and above in the list (storage_get_levels order), and its bit is set;
 * levels filled earlier count as valid neighbours for later ones, as in the reference's k loop.
 * levels, r_arr and valid_mask are HOST arrays of nlev entries; the arrays r_arr[k] / valid_mask[k] point to follow
 * the handle's pointer mode (nx*ny floats, ny*ceil(nx/32) words). */
int wpsgpu_fill_missing_levels(wpsgpu_handle_t h, int32_t nlev, const int32_t *levels, float *const *r_arr,
                               int32_t *const *valid_mask, int32_t nx, int32_t ny);
/* create_level (process_domain_module.F:3191-3332): a new level filled with a constant (src_* NULL; every valid bit
 * set, :3263-3272) or copied from another field's level together with its valid bits (:3300-3312).  The storage
 * lookups around it ("already exists; leaving it alone", source missing) stay with the caller. */
int wpsgpu_create_level(wpsgpu_handle_t h, int32_t nx, int32_t ny, float fill_const, const float *src_r_arr,
                        const int32_t *src_valid_mask, float *r_arr, int32_t *valid_mask);

/* met_to_map / map_to_met -> metmap_xform (rotate_winds_module.F:23-387), C grid, nlev levels at once.
 * u(us1:ue1,us2:ue2,nlev), v(vs1:ve1,vs2:ve2,nlev) in place; masks are per-level bit arrays laid out
 * [nlev][ny][nwords].  idir=+1 map_to_met (proj = source projection), -1 met_to_map (proj = domain). */
int wpsgpu_rotate_winds(wpsgpu_handle_t h, int idir, const wpsgpu_proj *proj, float *u, const int32_t *u_mask,
                        float *v, const int32_t *v_mask, int us1, int us2, int ue1, int ue2,
                        int vs1, int vs2, int ve1, int ve2, const float *xlon_u, const float *xlon_v, int nlev);
/* Same with the latitudes of the U and V points, which the PROJ_CASSINI branches read (rotate_winds_module.F:174-222,
 * :281-329: idir=+1 differentiates ij_to_latlon of `proj` at y +- 0.1, idir=-1 differentiates xlat/xlon along j).
 * xlat_u / xlat_v may be NULL for Lambert conformal and polar stereographic. */
int wpsgpu_rotate_winds_ll(wpsgpu_handle_t h, int idir, const wpsgpu_proj *proj, float *u, const int32_t *u_mask,
                           float *v, const int32_t *v_mask, int us1, int us2, int ue1, int ue2,
                           int vs1, int vs2, int ve1, int ve2, const float *xlon_u, const float *xlon_v,
                           const float *xlat_u, const float *xlat_v, int nlev);

/* map_to_met_nmm / met_to_map_nmm -> metmap_xform_nmm (rotate_winds_module.F:399-631): NMM E-grid, U and V collocated on the
 * velocity points (vs1:ve1, vs2:ve2), nlev levels in place, masks [nlev][ny][nwords].  PROJ_ROTLL uses lat1 / lon1 and the
 * latitudes / longitudes of the points; PROJ_LC / PS / CASSINI the rotation by stdlon of :571-583. */
int wpsgpu_rotate_winds_nmm(wpsgpu_handle_t h, int idir, const wpsgpu_proj *proj, float *u, const int32_t *u_mask, float *v,
                            const int32_t *v_mask, int vs1, int vs2, int ve1, int ve2, const float *xlat_v, const float *xlon_v, int nlev);

/* ---- geogrid: calc_field (process_tile_module.F:1208-1681) ---- */
typedef struct wpsgpu_geo_field {
    wpsgpu_proj dom_proj;           /* proj_stack(current_nest_number) */
    int32_t istagger;
    const float *xlat, *xlon;       /* xlat_array/xlon_array(start_i:end_i,start_j:end_j) */
    int32_t start_i, end_i, start_j, end_j, start_k, end_k;
    const float *landmask;          /* NULL = optional landmask absent */
    float *field;                   /* field(start_i:end_i,start_j:end_j,start_k:end_k), inout */
} wpsgpu_geo_field;
typedef struct wpsgpu_geo_level {
    wpsgpu_proj src_proj;           /* push_source_projection for this priority level (:1288) */
    int32_t ilevel, itype;          /* itype: CONTINUOUS / CATEGORICAL / SP_CONTINUOUS after the :1306-1327 decision */
    int32_t ops[WPSGPU_MAX_OPS], opts[WPSGPU_MAX_OPS];
    float msg_val, msg_fill_val, mask_val;
    int32_t has_scale; float scale_factor;
    int32_t sr_x, sr_y;
} wpsgpu_geo_level;
int wpsgpu_geogrid_field_begin(wpsgpu_handle_t h, const wpsgpu_geo_field *f);
int wpsgpu_geogrid_level_begin(wpsgpu_handle_t h, const wpsgpu_geo_level *l);
/* One source tile as get_data_tile returns it (source_data_module.F:2830): tile(src_min_x:src_max_x,
 * src_min_y:src_max_y, src_min_z:src_max_z) in global source-index space, bdr = tile_bdr.  Runs
 * accum_categorical / accum_continuous (proc_point_module.F:100,473) and then the per-point fallback of
 * calc_field's inner loop (:1408-1519, get_point proc_point_module.F:788) for the destination points
 * inside the tile (is_point_in_tile, :894).  The call returns once the tile has been read (host-pointer mode: copied
 * to the device; device-pointer mode: the work is queued and the caller keeps the buffer alive until the level ends);
 * the tile is staged on a second stream while the previous one is still being accumulated, so the host can fetch the
 * next tile from disk meanwhile.  Results are complete after wpsgpu_geogrid_field_finish. */
int wpsgpu_geogrid_add_tile(wpsgpu_handle_t h, const float *tile, int src_min_x, int src_max_x, int src_min_y,
                            int src_max_y, int src_min_z, int src_max_z, int bdr);
/* Same, from the raw file bytes of read_geogrid.c:26-140 (decode + row flip on device):
 * wordsize 1..4, is_signed, endian 0 big / 1 little, row_order 1 bottom_top / 2 top_bottom. */
int wpsgpu_geogrid_add_tile_raw(wpsgpu_handle_t h, const void *bytes, int wordsize, int is_signed, int endian,
                                int row_order, int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                                int src_min_z, int src_max_z, int bdr);
/* read_geogrid (geogrid/src/read_geogrid.c:26-140) without the file access, plus the TOP_BOTTOM row flip of
 * source_data_module.F:2887-2898: nx*ny*nz words of `wordsize` bytes -> rarray(nx,ny,nz) binary32, multiplied by
 * scalefactor when it is not 1.  Reproduces the reference's sign rule as written (`ival > 2^(8w-1)`, so the most negative
 * value of a signed word stays positive).  bytes / rarray follow the handle's pointer mode. */
int wpsgpu_geogrid_decode_tile(wpsgpu_handle_t h, const void *bytes, int wordsize, int is_signed, int endian,
                               int row_order, int nx, int ny, int nz, float scalefactor, float *rarray);
/* post-processing of the level (:1580-1659) */
int wpsgpu_geogrid_level_end(wpsgpu_handle_t h);
/* copies field (and optionally the processed_domain words) back to the caller */
int wpsgpu_geogrid_field_finish(wpsgpu_handle_t h, int32_t *processed_domain_out, int64_t *n_invalid_category);

/* fractions + landmask + dominant category (process_tile_module.F:539-622,699-727,1035-1051,1123-1144).
 * field(npts, min_cat:max_cat) in place; outputs may be NULL.  lm_vals/n_lm/is_water_mask describe the
 * landmask categories (get_landmask_field); pass n_lm=0 for a non-landmask field (plain dominant). */
int wpsgpu_fractions_dominant(wpsgpu_handle_t h, float *field, int64_t npts, int min_cat, int max_cat,
                              float msg_fill_val, const int32_t *lm_vals, int n_lm, int is_water_mask,
                              float *landmask_out, float *dominant_out);
/* one_two_one / smth_desmth / smth_desmth_special (smooth_module.F:20,73,154), array(sx:ex,sy:ey,sz:ez) */
int wpsgpu_smooth(wpsgpu_handle_t h, int opt, int npass, float *array, int sx, int ex, int sy, int ey,
                  int sz, int ez);
/* one_two_one_egrid / smth_desmth_egrid (smooth_module.F:252-338, :448-585) for the NMM E-grid: msgval == 1.0 leaves points whose
 * value is 0 alone, hflag == 1.0 selects mass points (else velocity points); smth-desmth_special is a no-op as in
 * process_tile_module.F:675.  Levels other than the first: see oracle/orc_geogrid.c (the reference initialises its scratch
 * for level index 1 only). */
int wpsgpu_smooth_egrid(wpsgpu_handle_t h, int opt, int npass, float *array, int sx, int ex, int sy, int ey,
                        int sz, int ez, float msgval, float hflag);
/* get_lat_lon_fields (process_tile_module.F:1692) */
int wpsgpu_xytoll_grid(wpsgpu_handle_t h, const wpsgpu_proj *dom, int stagger, int si, int sj, int ei, int ej,
                       float sub_x, float sub_y, float *xlat, float *xlon);
/* get_map_factor (:1735); pole_lat/pole_lon/stand_lon are the gridinfo_module values used for Cassini */
int wpsgpu_map_factor(wpsgpu_handle_t h, int iproj, float truelat1, float truelat2, float pole_lat,
                      float pole_lon, float stand_lon, const float *xlat, const float *xlon, int64_t npts,
                      float *mapfac_x, float *mapfac_y);
/* get_coriolis_parameters (:1885) */
int wpsgpu_coriolis(wpsgpu_handle_t h, const float *xlat, int64_t npts, float *f, float *e);
/* get_rotang (:1920) */
int wpsgpu_rotang(wpsgpu_handle_t h, int iproj, float truelat1, float truelat2, float stand_lon,
                  const float *xlat, const float *xlon, int si, int sj, int ei, int ej, float *cosa, float *sina);
/* calc_dfdx / calc_dfdy (:2197,:2132); mapfac may be NULL */
int wpsgpu_dfdx(wpsgpu_handle_t h, const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek,
                int sr_x, const float *mapfac, float dom_dx);
int wpsgpu_dfdy(wpsgpu_handle_t h, const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek,
                int sr_y, const float *mapfac, float dom_dy);

/* ---- metgrid: MPAS source remapping (metgrid/src/remapper.F, process_domain_module.F:1457-1962) ---- */
#define WPSGPU_MPAS_CELLS 0      /* decomposed dimension nCells (remapper.F:445) */
#define WPSGPU_MPAS_VERTICES 1   /* nVertices (:454) */
#define WPSGPU_MPAS_EDGES 2      /* nEdges (:458) */
#define WPSGPU_MPAS_MASKED_CELLS 3   /* table selector only: sourceMaskedCells / cellMaskedWeights (:448-450) */
#define WPSGPU_MPAS_REAL 0       /* FIELD_TYPE_REAL / _DOUBLE / _INTEGER of scan_input.F */
#define WPSGPU_MPAS_DOUBLE 1
#define WPSGPU_MPAS_INTEGER 2
/* mpas_mesh_type (mpas_mesh.F:5-30) as mpas_mesh_setup fills it: connectivity 1-based as in the file, Fortran
 * (maxEdges, nCells) = C [nCells][maxEdges]; latitudes / longitudes in radians.  Always HOST arrays. */
typedef struct wpsgpu_mpas_mesh {
    int32_t nCells, nVertices, nEdges, maxEdges;
    const int32_t *landmask, *nEdgesOnCell;
    const int32_t *cellsOnCell, *verticesOnCell, *edgesOnCell;
    const int32_t *cellsOnVertex;    /* [nVertices][3] */
    const int32_t *cellsOnEdge;      /* [nEdges][2] */
    const float *latCell, *lonCell, *latVertex, *lonVertex, *latEdge, *lonEdge;
} wpsgpu_mpas_mesh;
/* mpas_mesh_setup (mpas_mesh.F:36): the mesh is copied to the device and kept by the handle. */
int wpsgpu_mpas_set_mesh(wpsgpu_handle_t h, const wpsgpu_mpas_mesh *mesh);
/* target_mesh_setup(lat2d, lon2d) + remap_info_setup (remapper.F:83-286) for target grid `grid_id` (0..2: the mass, U
 * and V grids of process_mpas_fields, :1535-1590).  lat_rad / lon_rad(nlon, nlat), x fastest, radians.  Builds
 * nearestCell / nearestVertex / nearestEdge, the Wachspress tables of cells, vertices and edges and the masked cell
 * table (landmask == 1 cells only, search_for_cells where all three are water) on the device; one thread per target
 * point instead of the reference's sequential scan, which hands every point the previous point's result as the start
 * of its greedy walk -- the walks end in the same cell / vertex / edge except where two candidates are equidistant
 * in binary32. */
int wpsgpu_mpas_remap_setup(wpsgpu_handle_t h, int grid_id, const float *lat_rad, const float *lon_rad, int nlon, int nlat);
/* the tables of remap_info_type (remapper.F:19-40) for `kind`: nearest(nlon*nlat) [NULL for MASKED_CELLS],
 * sources / weights(nlon*nlat, width) with width 3 (cells, masked cells) or maxEdges; any output may be NULL */
int wpsgpu_mpas_get_tables(wpsgpu_handle_t h, int grid_id, int kind, int32_t *nearest, int32_t *sources, float *weights);
/* remap_field (remapper.F:417-663): dst(nlev, nlon, nlat) from src(nlev, nNodes), level index fastest in both (nlev =
 * product of the non-decomposed dimensions, 1 for a 1-d field).  REAL / DOUBLE: weighted sum over the first nw table
 * entries in the reference's order (nw = 3 for 2-d and 3-d sources as :522, :535; the table width for 1-d sources as
 * :506; 0 selects by nlev like the reference); INTEGER: nearest node.  masked != 0 selects the masked cell table. */
int wpsgpu_mpas_remap_field(wpsgpu_handle_t h, int grid_id, int decomp, int xtype, int masked, const void *src, int nlev,
                            int nw, void *dst);
/* remap_field of a REAL field fused with process_mpas_fields' per-level extraction (:1700-1956): writes the r_arr
 * planes(nlon, nlat) directly.  mode 0: plane k = level k (nVertLevels, nSoilLevels, 2-d fields); mode 1
 * (nVertLevelsP1, :1733-1776): plane 0 = level 1 (stored at 200100), plane k = 0.5 * (level k + level k+1), k = 1 ..
 * nlev-1.  landmask (may be NULL): `where (landmask == 0) r_arr = fill_missing` (:1854, :1948).  planes: HOST array of
 * nlev pointers that follow the pointer mode. */
int wpsgpu_mpas_remap_planes(wpsgpu_handle_t h, int grid_id, int decomp, int masked, const float *src, int nlev, int mode,
                             const float *landmask, float fill_missing, float *const *planes);
/* derive_mpas_fields (process_domain_module.F:1957-2053): TT = theta * (pressure / P0) ** (RD / CP) in place on the stored
 * planes of every level; theta / pressure: HOST arrays of nlev pointers that follow the pointer mode, npts values each. */
int wpsgpu_mpas_derive_tt(wpsgpu_handle_t h, float *const *theta, const float *const *pressure, int nlev, int64_t npts);

#ifdef __cplusplus
}
#endif
#endif
