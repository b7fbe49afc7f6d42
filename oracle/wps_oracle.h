/*
 * wps_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the horizontal-interpolation hot path of
 * wrf-model/WPS v4.6.0 (geogrid.exe / metgrid.exe).  Every function cites the
 * reference file:line it follows (paths relative to the WPS source tree).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path and cannot be compiled in this image (no Fortran
 * compiler, no MPI, no netCDF, no WRF I/O libs).  The pins are therefore
 * manufactured: hand-derived known-answer tests from the Fortran source
 * (tests/test_oracle_kat.py), an independent second transliteration in Python
 * (oracle/pyref.py) that the C oracle is cross-checked against, and metamorphic
 * properties.  See DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (wps_b200/) never does.
 *
 * Arithmetic: IEEE binary32 everywhere the Fortran uses default REAL, compiled
 * with -O1 -ffp-contract=off -fno-fast-math (mirrors gfortran -O on generic
 * x86-64: SSE scalar float, no FMA contraction, no flush-to-zero).
 * Transcendentals: orc_set_transc_mode(0) calls glibc's float functions exactly
 * like gfortran would (sinf, cosf, tanf, powf, log10f, logf, atan2f, ...);
 * mode 1 evaluates them in binary64 and rounds once to binary32 (the policy the
 * CUDA product uses, see DESIGN.md "Transcendental policy").
 */
#ifndef WPS_ORACLE_H
#define WPS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* misc_definitions_module.F:12-47 */
#define ORC_NAN_F 1.0e20f
#define ORC_NOT_MASKED (-2.0f)
#define ORC_MASKED_BOTH (-1.0f)
#define ORC_MASKED_WATER 0.0f
#define ORC_MASKED_LAND 1.0f
#define ORC_OUTSIDE_DOMAIN 100000000
#define ORC_NOT_PROCESSED 1000000000

enum { ORC_SIXTEEN_POINT = 1, ORC_FOUR_POINT = 2, ORC_N_NEIGHBOR = 3, ORC_AVERAGE4 = 4,
       ORC_AVERAGE16 = 5, ORC_W_AVERAGE4 = 6, ORC_W_AVERAGE16 = 7, ORC_SEARCH = 8 };
enum { ORC_M = 1, ORC_U = 2, ORC_V = 3, ORC_HH = 4, ORC_VV = 5, ORC_CORNER = 6 };
enum { ORC_ONETWOONE = 1, ORC_SMTHDESMTH = 2, ORC_SMTHDESMTH_SPECIAL = 3 };
enum { ORC_PROJ_LATLON = 0, ORC_PROJ_LC = 1, ORC_PROJ_PS = 2, ORC_PROJ_PS_WGS84 = 102,
       ORC_PROJ_MERC = 3, ORC_PROJ_GAUSS = 4, ORC_PROJ_CYL = 5, ORC_PROJ_CASSINI = 6,
       ORC_PROJ_ALBERS_NAD83 = 105, ORC_PROJ_ROTLL = 203 };

/* proj_info, module_map_utils.F:140-192 */
typedef struct orc_proj {
    int code, nlat, nlon, nxmin, nxmax, ixdim, jydim, stagger;
    float phi, lambda, lat1, lon1, lat0, lon0, dx, dy, latinc, loninc, dlat, dlon, stdlon,
          truelat1, truelat2, hemi, cone, polei, polej, rsw, rebydx, knowni, knownj, re_m,
          rho0, nc, bigc;
    int init, wrap, comp_ll;
    int n_gauss;
    float gauss_lat[2048];
} orc_proj;

/* OPTIONAL arguments of map_set (module_map_utils.F:243-283): bit i of `present` set
 * when the argument was passed. */
enum { ORC_A_LAT1 = 1 << 0, ORC_A_LON1 = 1 << 1, ORC_A_LAT0 = 1 << 2, ORC_A_LON0 = 1 << 3,
       ORC_A_KNOWNI = 1 << 4, ORC_A_KNOWNJ = 1 << 5, ORC_A_DX = 1 << 6, ORC_A_DY = 1 << 7,
       ORC_A_LATINC = 1 << 8, ORC_A_LONINC = 1 << 9, ORC_A_STDLON = 1 << 10,
       ORC_A_TRUELAT1 = 1 << 11, ORC_A_TRUELAT2 = 1 << 12, ORC_A_NLAT = 1 << 13,
       ORC_A_NLON = 1 << 14, ORC_A_IXDIM = 1 << 15, ORC_A_JYDIM = 1 << 16,
       ORC_A_NXMIN = 1 << 17, ORC_A_NXMAX = 1 << 18, ORC_A_STAGGER = 1 << 19,
       ORC_A_PHI = 1 << 20, ORC_A_LAMBDA = 1 << 21, ORC_A_R_EARTH = 1 << 22 };
typedef struct orc_map_args {
    unsigned present;
    float lat1, lon1, lat0, lon0, knowni, knownj, dx, dy, latinc, loninc, stdlon, truelat1,
          truelat2, phi, lambda, r_earth;
    int nlat, nlon, ixdim, jydim, nxmin, nxmax, stagger;
} orc_map_args;

void orc_set_transc_mode(int mode);
int  orc_get_transc_mode(void);

/* ---- projections (module_map_utils.F, llxy_module.F) ---- */
void orc_map_init(orc_proj *p);
int  orc_map_set(int code, orc_proj *p, const orc_map_args *a);
void orc_latlon_to_ij(const orc_proj *p, float lat, float lon, float *i, float *j);
void orc_ij_to_latlon(const orc_proj *p, float i, float j, float *lat, float *lon);
void orc_lltoxy(orc_proj *p, float xlat, float xlon, float *x, float *y, int stagger);
void orc_xytoll(orc_proj *p, float x, float y, float *xlat, float *xlon, int stagger);
void orc_lltoxy_array(orc_proj *p, long n, const float *lat, const float *lon, float *x, float *y, int stagger);
void orc_xytoll_array(orc_proj *p, long n, const float *x, const float *y, float *lat, float *lon, int stagger);
void orc_rotate_coords(float ilat, float ilon, float *olat, float *olon, float lat_np, float lon_np,
                       float lon_0, int direction);
void orc_lc_cone(float truelat1, float truelat2, float *cone);
/* compute_nest_locations + find_known_latlon (llxy_module.F:331-593); arrays of n_domains entries, entry 0 = domain 1 */
int orc_compute_nest_locations(int iproj, const orc_map_args *coarse, int n_domains, const int *parent_id, const int *ratio,
                               const int *ips, const int *jps, const int *ixdim, const int *jydim, orc_proj *out);
/* push_source_projection (llxy_module.F:39-165) */
int orc_push_source_projection(orc_proj *p, int iproj, float stand_lon, float truelat1, float truelat2,
                               float dxkm, float dykm, float dlat, float dlon, float known_x, float known_y,
                               float known_lat, float known_lon, int has_cass, float pole_lat, float pole_lon,
                               float centerlat, float centerlon, float centeri, float centerj,
                               int has_re, float earth_radius);

/* ---- interp_module.F ---- */
int   orc_interp_array_from_string(const char *s, int *codes, int maxn);
int   orc_interp_options_from_string(const char *s, int *opts, int maxn);
int   orc_get_gcell_threshold(const char *s, float *threshold);
float orc_oned(float x, float a, float b, float c, float d);
/* mask_mode: 0 = no mask args; 1 = mask_array + maskval + mask_relational present */
float orc_interp_sequence(float xx, float yy, int izz, const float *array, int sx, int ex, int sy, int ey,
                          int sz, int ez, float msgval, const int *list, const int *opts, int idx,
                          int mask_mode, char rel, float maskval, const float *mask);
void  orc_interp_sequence_array(long n, const float *xx, const float *yy, int izz, const float *array,
                                int sx, int ex, int sy, int ey, int sz, int ez, float msgval,
                                const int *list, const int *opts, int mask_mode, char rel, float maskval,
                                const float *mask, float *out);
/* search_extrap BFS visit order (for the closed-form tests): writes visited (x,y) in pop order */
int   orc_search_order(int x0, int y0, int sx, int ex, int sy, int ey, int maxdepth, int *xs, int *ys, int maxn);

/* ---- bitarray_module.F ---- */
static inline int orc_bit_test(const int32_t *w, int nxw, int i, int j)
{ return (w[(long)(j - 1) * nxw + (i - 1) / 32] >> ((i - 1) % 32)) & 1; }
static inline void orc_bit_set(int32_t *w, int nxw, int i, int j)
{ w[(long)(j - 1) * nxw + (i - 1) / 32] |= (int32_t)(1u << ((i - 1) % 32)); }
static inline void orc_bit_clear(int32_t *w, int nxw, int i, int j)
{ w[(long)(j - 1) * nxw + (i - 1) / 32] &= (int32_t)~(1u << ((i - 1) % 32)); }

/* ---- metgrid: process_domain_module.F ---- */
typedef struct orc_field_opts {
    int   ops[16];            /* interp_array_from_string, 0-terminated */
    int   opts[16];           /* interp_options_from_string */
    float missing_value, fill_missing, masked;
    /* index 0: interp_mask, 1: interp_land_mask, 2: interp_water_mask */
    float mask_val[3];
    char  mask_rel[3];
} orc_field_opts;

/* interp_to_latlon, process_domain_module.F:2594-2669 */
float orc_interp_to_latlon(orc_proj *src, float rlat, float rlon, const int *list, const int *opts,
                           const float *slab, int minx, int maxx, int miny, int maxy, int bdr, float msg,
                           int mask_mode, char rel, float maskval, const float *mask);
/* interp_met_field non-gcell branch, process_domain_module.F:2483-2581.
 * mask slabs are NULL when the "<NAME>.mask" field is absent (status /= 0).
 * i0..i1/j0..j1 restrict the loop to a sub-rectangle (patch decomposition for the CPU baseline). */
void orc_interp_met_field(orc_proj *src, const orc_field_opts *fo, int ifieldstagger,
                          const float *xlat, const float *xlon, int sm1, int em1, int sm2, int em2,
                          int sd1, int ed1, int sd2, int ed2,
                          const float *slab, int minx, int maxx, int miny, int maxy, int bdr,
                          const float *mask, const float *land_mask, const float *water_mask,
                          const float *landmask, int process_bdy_width,
                          float *r_arr, const int32_t *valid_mask, int32_t *new_pts,
                          int i0, int i1, int j0, int j1);
/* average_gcell branch, process_domain_module.F:2331-2478 + 3347-3717 */
void orc_interp_met_field_gcell(orc_proj *src, orc_proj *dom, const orc_field_opts *fo, int ifieldstagger,
                                const float *xlat, const float *xlon, int sm1, int em1, int sm2, int em2,
                                int sd1, int ed1, int sd2, int ed2,
                                const float *slab, int minx, int maxx, int miny, int maxy, int bdr,
                                const float *mask, const float *landmask, int process_bdy_width,
                                float *r_arr, const int32_t *valid_mask, int32_t *new_pts);
/* metmap_xform, rotate_winds_module.F:85-387 (C grid; LC / PS / "plain" alpha only, no Cassini) */
void orc_metmap_xform_ll(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                         int us1, int us2, int ue1, int ue2, int vs1, int vs2, int ve1, int ve2,
                         const float *xlon_u, const float *xlon_v, const float *xlat_u, const float *xlat_v, int idir);
void orc_metmap_xform(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                      int us1, int us2, int ue1, int ue2, int vs1, int vs2, int ve1, int ve2,
                      const float *xlon_u, const float *xlon_v, int idir);

/* ---- geogrid ---- */
/* smooth_module.F:20-239; opt = ORC_ONETWOONE | ORC_SMTHDESMTH | ORC_SMTHDESMTH_SPECIAL */
void orc_smooth(float *array, int sx, int ex, int sy, int ey, int sz, int ez, int npass, int opt);
void orc_smooth_egrid(float *array, int sx, int ex, int sy, int ey, int sz, int ez, int npass, int opt, float msgval, float hflag);
/* one tile of accum_categorical / accum_continuous (proc_point_module.F:100-779).
 * tile(src_min_x:src_max_x, src_min_y:src_max_y, src_min_z:src_max_z) in global source index space. */
void orc_accum_categorical(orc_proj *src, orc_proj *dom, int istagger, const float *tile,
                           int src_min_x, int src_max_x, int src_min_y, int src_max_y, int bdr,
                           float *dst, int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                           int32_t *processed_pts, int32_t *new_pts, int ilevel, float msgval,
                           float sr_x, float sr_y, long *n_invalid);
void orc_accum_continuous(orc_proj *src, orc_proj *dom, int istagger, const float *tile,
                          int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                          int src_min_z, int src_max_z, int bdr,
                          float *dst, float *n, int start_i, int end_i, int start_j, int end_j,
                          int start_k, int end_k, int32_t *processed_pts, int32_t *new_pts, float msgval,
                          float sr_x, float sr_y);
/* per-tile point fallback of calc_field's inner loop (process_tile_module.F:1408-1519) for all
 * destination points inside the tile (is_point_in_tile, proc_point_module.F:894-937). */
void orc_tile_point_fallback(orc_proj *src, int itype, const float *tile,
                             int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                             int src_min_z, int src_max_z, int bdr,
                             const float *xlat, const float *xlon, const float *landmask, int use_landmask,
                             float mask_val, const int *list, const int *opts, float msg_val, float msg_fill_val,
                             float *field, float *data_count,
                             int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                             const int32_t *processed_domain, int32_t *level_domain, long *n_invalid);
/* calc_field post-processing, process_tile_module.F:1580-1654 */
void orc_calc_field_finish(int itype, float *field, float *data_count, const float *landmask, int use_landmask,
                           float mask_val, float msg_fill_val, int has_scale, float scale_factor,
                           int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                           int32_t *processed_domain, const int32_t *level_domain);
/* fractions / landmask / dominant category, process_tile_module.F:539-622,699-727,1035-1051,1123-1144 */
void orc_fractions(float *field, long npts, int ncat, float msg_fill_val);
void orc_landmask_from_fractions(const float *field, long npts, int min_cat, int max_cat, const int *lm_vals,
                                 int n_lm, int is_water_mask, float msg_fill_val, float *landmask);
void orc_dominant_category(const float *field, long npts, int min_cat, int max_cat, float msg_fill_val,
                           float *dominant);
void orc_dominant_category_landmask(const float *field, const float *landmask, long npts, int min_cat,
                                    int max_cat, const int *lm_vals, int n_lm, int is_water_mask,
                                    float *dominant);
/* get_lat_lon_fields / get_map_factor / get_coriolis_parameters / get_rotang / calc_dfdx / calc_dfdy
 * (process_tile_module.F:1692-2253) */
void orc_get_lat_lon_fields(orc_proj *dom, float *xlat, float *xlon, int si, int sj, int ei, int ej,
                            int stagger, float sub_x, float sub_y);
void orc_get_map_factor(const orc_proj *dom, const float *xlat, const float *xlon, float *mx, float *my,
                        long npts, float truelat1, float truelat2, float pole_lat, float pole_lon, float stand_lon);
void orc_get_coriolis(const float *xlat, float *f, float *e, long npts);
void orc_get_rotang(int iproj, const float *xlat, const float *xlon, float *cosa, float *sina,
                    int si, int sj, int ei, int ej, float truelat1, float truelat2, float stand_lon);
void orc_calc_dfdx(const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek, int sr_x,
                   const float *mapfac, float dom_dx);
void orc_calc_dfdy(const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek, int sr_y,
                   const float *mapfac, float dom_dy);

#ifdef __cplusplus
}
#endif
/* fill_missing_levels vertical interpolation, process_domain_module.F:2862-2921 */
void orc_fill_missing_levels(int nlev, const int32_t *levels, float *r_arr, int32_t *valid, int nx, int ny);

void orc_metmap_xform_nmm(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                          int vs1, int vs2, int ve1, int ve2, const float *xlat_v, const float *xlon_v, int idir);

/* ---- MPAS remapping (metgrid/src/remapper.F, mpas_mesh.F:5-30); connectivity 1-based, C row = Fortran last index ---- */
typedef struct orc_mpas_mesh {
    int nCells, nVertices, nEdges, maxEdges;
    const int *landmask, *nEdgesOnCell;                       /* [nCells] */
    const int *cellsOnCell, *verticesOnCell, *edgesOnCell;    /* [nCells][maxEdges] */
    const int *cellsOnVertex;                                 /* [nVertices][3] */
    const int *cellsOnEdge;                                   /* [nEdges][2] */
    const float *latCell, *lonCell, *latVertex, *lonVertex, *latEdge, *lonEdge;   /* radians */
} orc_mpas_mesh;
typedef struct orc_mpas_remap {      /* remap_info_type, remapper.F:19-40; point index = ix + nlon * iy */
    int *nearestCell, *nearestVertex, *nearestEdge;           /* [npts] */
    int *sourceCells; float *cellWeights;                     /* [npts][3] */
    int *sourceVertices; float *vertexWeights;                /* [npts][maxEdges] */
    int *sourceEdges; float *edgeWeights;                     /* [npts][maxEdges] */
    int *sourceMaskedCells; float *cellMaskedWeights;         /* [npts][3] */
} orc_mpas_remap;
int  orc_mpas_nearest_cell(const orc_mpas_mesh *m, float tlat, float tlon, int start_cell);
int  orc_mpas_nearest_vertex(const orc_mpas_mesh *m, float tlat, float tlon, int start, int degree);
void orc_mpas_wachspress3(const float vert[3][3], const float p[3], float w[3]);
void orc_mpas_search_for_cells(const orc_mpas_mesh *m, float tlat, float tlon, int start, int cells[3], float w[3]);
void orc_mpas_remap_setup(const orc_mpas_mesh *m, const float *lat, const float *lon, int nlon, int nlat, orc_mpas_remap *r);
void orc_mpas_remap_field_real(const int *sourceNodes, const float *nodeWeights, int tw, int nw, long npts, const float *src,
                               int nlev, float *dst);
void orc_mpas_remap_field_double(const int *sourceNodes, const float *nodeWeights, int tw, int nw, long npts, const double *src,
                                 int nlev, double *dst);
void orc_mpas_remap_field_int(const int *nearest, long npts, const int *src, int nlev, int *dst);
void orc_mpas_derive_tt(float *theta, const float *pressure, long n);
void orc_mpas_extract_planes(const float *wrf, int nlev, long npts, int mode, const float *landmask, float fill, float *planes);

#endif
