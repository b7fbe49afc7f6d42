/*
 * orc_geogrid.c -- CPU ORACLE (test infrastructure): geogrid accumulators, calc_field pieces,
 * fractions / dominant category / landmask, smoothers and the elementwise geometry outputs.
 * Restates geogrid/src/proc_point_module.F:100-937, process_tile_module.F:539-622,699-727,
 * 1035-1051,1123-1144,1408-1654,1692-2253 and smooth_module.F:20-239 of WPS v4.6.0.
 * PARITY UNPINNED (see wps_oracle.h).  Not product code.
 */
#include "orc_internal.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* smooth_module.F:20-239.  array(sx:ex, sy:ey, sz:ez), x fastest. */
void orc_smooth(float *array, int sx, int ex, int sy, int ey, int sz, int ez, int npass, int opt)
{
    int nx = ex - sx + 1, ny = ey - sy + 1, nz = ez - sz + 1;
    long n2 = (long)nx * ny, n3 = n2 * nz;
#define AR(a, ix, iy, iz) (a)[(long)((iz) - sz) * n2 + (long)((iy) - sy) * nx + ((ix) - sx)]
    float *scratch = (float *)malloc(sizeof(float) * (size_t)n3);
    float *orig = NULL;
    if (opt == ORC_SMTHDESMTH_SPECIAL) { orig = (float *)malloc(sizeof(float) * (size_t)n3); memcpy(orig, array, sizeof(float) * (size_t)n3); }
    for (int ipass = 1; ipass <= npass; ++ipass) {
        for (int iy = sy; iy <= ey; ++iy) for (int ix = sx + 1; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            AR(scratch, ix, iy, iz) = 0.5f * AR(array, ix, iy, iz) + 0.25f * (AR(array, ix - 1, iy, iz) + AR(array, ix + 1, iy, iz));
        for (int iy = sy + 1; iy <= ey - 1; ++iy) for (int ix = sx + 1; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            AR(array, ix, iy, iz) = 0.5f * AR(scratch, ix, iy, iz) + 0.25f * (AR(scratch, ix, iy - 1, iz) + AR(scratch, ix, iy + 1, iz));
        if (opt == ORC_ONETWOONE) continue;
        for (int iy = sy; iy <= ey; ++iy) for (int ix = sx + 1; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            AR(scratch, ix, iy, iz) = 1.52f * AR(array, ix, iy, iz) - 0.26f * (AR(array, ix - 1, iy, iz) + AR(array, ix + 1, iy, iz));
        for (int iy = sy + 1; iy <= ey - 1; ++iy) for (int ix = sx + 1; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            AR(array, ix, iy, iz) = 1.52f * AR(scratch, ix, iy, iz) - 0.26f * (AR(scratch, ix, iy - 1, iz) + AR(scratch, ix, iy + 1, iz));
    }
    if (opt == ORC_SMTHDESMTH_SPECIAL) {
        for (long k = 0; k < n3; ++k) if (array[k] < 0.0f && orig[k] >= 0.0f) array[k] = orig[k];
        free(orig);
    }
    free(scratch);
#undef AR
}

/* ------------------------------------------------------------------ */
/* accumulators, proc_point_module.F:100-779 */
typedef struct {
    orc_proj *src, *dom;
    const float *tile; int smx, sMx, smy, sMy, smz, sMz;
    float msgval, sr_x, sr_y; int istagger;
    float *dst, *n; int sx, ex, sy, ey, sz, ez;
    int32_t *processed, *new_pts; int nxw;
    int *wm; long n_invalid; int categorical;
    int min_k, max_k;
} actx;
#define WM(g, i, j, c) (g)->wm[(((long)((j) - (g)->smy)) * ((g)->sMx - (g)->smx + 1) + ((i) - (g)->smx)) * 2 + (c)]
#define TL(g, i, j, k) (g)->tile[((long)((k) - (g)->smz) * ((g)->sMy - (g)->smy + 1) + ((j) - (g)->smy)) * ((g)->sMx - (g)->smx + 1) + ((i) - (g)->smx)]
#define GD(g, a, i, j, k) (a)[((long)((k) - (g)->sz) * ((g)->ey - (g)->sy + 1) + ((j) - (g)->sy)) * ((g)->ex - (g)->sx + 1) + ((i) - (g)->sx)]

static void a_corner(actx *g, int i, int j)
{
    if (WM(g, i, j, 0) != ORC_NOT_PROCESSED) return;
    float lat, lon, rx, ry;
    orc_xytoll(g->src, (float)i, (float)j, &lat, &lon, ORC_M);
    orc_lltoxy(g->dom, lat, lon, &rx, &ry, g->istagger);
    rx = (rx - 0.5f) * g->sr_x + 0.5f;
    ry = (ry - 0.5f) * g->sr_y + 0.5f;
    if ((float)g->sx <= rx && rx <= (float)g->ex && (float)g->sy <= ry && ry <= (float)g->ey) {
        WM(g, i, j, 0) = o_nint(rx); WM(g, i, j, 1) = o_nint(ry);
    } else WM(g, i, j, 0) = ORC_OUTSIDE_DOMAIN;
}
static void a_init_cell(actx *g, int xd, int yd)
{
    if (!orc_bit_test(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1)) {
        if (g->categorical) { for (int k = g->sz; k <= g->ez; ++k) GD(g, g->dst, xd, yd, k) = 0.0f; }
        else { for (int k = g->min_k; k <= g->max_k; ++k) GD(g, g->dst, xd, yd, k) = 0.0f; }
    }
}
static void a_credit(actx *g, int i, int j, int xd, int yd)
{
    if (g->categorical) {
        float v = TL(g, i, j, g->min_k);
        if (v != g->msgval) {
            int cat = (int)v;
            if (cat >= g->sz && cat <= g->ez) {
                GD(g, g->dst, xd, yd, cat) = GD(g, g->dst, xd, yd, cat) + 1.0f;
                orc_bit_set(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1);
            } else g->n_invalid++;
        }
    } else {
        for (int k = g->min_k; k <= g->max_k; ++k) {
            float v = TL(g, i, j, k);
            if (v != g->msgval) {
                GD(g, g->dst, xd, yd, k) = GD(g, g->dst, xd, yd, k) + v;
                GD(g, g->n, xd, yd, k) = GD(g, g->n, xd, yd, k) + 1.0f;
                orc_bit_set(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1);
            }
        }
    }
}
static void a_block(actx *g, int min_i, int min_j, int max_i, int max_j)
{
    a_corner(g, min_i, min_j); a_corner(g, min_i, max_j); a_corner(g, max_i, max_j); a_corner(g, max_i, min_j);
    if (WM(g, min_i, min_j, 0) == WM(g, min_i, max_j, 0) && WM(g, min_i, min_j, 0) == WM(g, max_i, max_j, 0) &&
        WM(g, min_i, min_j, 0) == WM(g, max_i, min_j, 0) && WM(g, min_i, min_j, 1) == WM(g, min_i, max_j, 1) &&
        WM(g, min_i, min_j, 1) == WM(g, max_i, max_j, 1) && WM(g, min_i, min_j, 1) == WM(g, max_i, min_j, 1) &&
        WM(g, min_i, min_j, 0) != ORC_OUTSIDE_DOMAIN) {
        int xd = WM(g, min_i, min_j, 0), yd = WM(g, min_i, min_j, 1);
        if (!orc_bit_test(g->processed, g->nxw, xd - g->sx + 1, yd - g->sy + 1)) {
            a_init_cell(g, xd, yd);
            for (int i = min_i; i <= max_i; ++i) for (int j = min_j; j <= max_j; ++j) a_credit(g, i, j, xd, yd);
        }
    } else if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) {
        for (int i = min_i; i <= max_i; ++i) for (int j = min_j; j <= max_j; ++j) {
            int xd = WM(g, i, j, 0), yd = WM(g, i, j, 1);
            if (xd != ORC_OUTSIDE_DOMAIN) {
                if (!orc_bit_test(g->processed, g->nxw, xd - g->sx + 1, yd - g->sy + 1)) {
                    a_init_cell(g, xd, yd);
                    a_credit(g, i, j, xd, yd);
                }
            }
        }
    } else {
        int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
        a_block(g, min_i, min_j, ci, cj);
        if (ci < max_i) a_block(g, ci + 1, min_j, max_i, cj);
        if (cj < max_j) a_block(g, min_i, cj + 1, ci, max_j);
        if (ci < max_i && cj < max_j) a_block(g, ci + 1, cj + 1, max_i, max_j);
    }
}
static void a_run(actx *g, int bdr)
{
    long nw = (long)(g->sMx - g->smx + 1) * (g->sMy - g->smy + 1) * 2;
    g->wm = (int *)malloc(sizeof(int) * (size_t)nw);
    for (long k = 0; k < nw; k += 2) g->wm[k] = ORC_NOT_PROCESSED;
    a_block(g, g->smx + bdr, g->smy + bdr, g->sMx - bdr, g->sMy - bdr);
    free(g->wm);
}

/* proc_point_module.F:100-206 */
void orc_accum_categorical(orc_proj *src, orc_proj *dom, int istagger, const float *tile,
                           int src_min_x, int src_max_x, int src_min_y, int src_max_y, int bdr,
                           float *dst, int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                           int32_t *processed_pts, int32_t *new_pts, int ilevel, float msgval,
                           float sr_x, float sr_y, long *n_invalid)
{
    actx g; memset(&g, 0, sizeof(g));
    g.src = src; g.dom = dom; g.tile = tile; g.smx = src_min_x; g.sMx = src_max_x; g.smy = src_min_y; g.sMy = src_max_y;
    g.smz = 1; g.sMz = 1; g.min_k = 1; g.max_k = 1;
    g.msgval = msgval; g.sr_x = sr_x; g.sr_y = sr_y; g.istagger = istagger;
    g.dst = dst; g.n = NULL; g.sx = start_i; g.ex = end_i; g.sy = start_j; g.ey = end_j; g.sz = start_k; g.ez = end_k;
    g.processed = processed_pts; g.new_pts = new_pts; g.nxw = (end_i - start_i + 1 + 31) / 32; g.categorical = 1;
    a_run(&g, bdr);
    if (n_invalid) *n_invalid += g.n_invalid;
    /* half-area rule, proc_point_module.F:165-202 */
    if (ilevel > 1) {
        for (int i = start_i; i <= end_i; ++i) for (int j = start_j; j <= end_j; ++j) {
            if (orc_bit_test(new_pts, g.nxw, i - start_i + 1, j - start_j + 1) &&
                !orc_bit_test(processed_pts, g.nxw, i - start_i + 1, j - start_j + 1)) {
                float rarea = 0.0f;
                for (int k = start_k; k <= end_k; ++k) rarea = rarea + GD(&g, dst, i, j, k);
                if (src->dx < 0.0f) { float s = src->latinc * 111000.0f; rarea = rarea * (s * s); }
                else rarea = rarea * (src->dx * src->dx);
                float darea;
                if (dom->dx < 0.0f) { float s = dom->latinc * 111000.0f; darea = s * s; } else darea = dom->dx * dom->dx;
                if (darea > 2.0f * rarea) {
                    for (int k = start_k; k <= end_k; ++k) GD(&g, dst, i, j, k) = 0.0f;
                    orc_bit_clear(new_pts, g.nxw, i - start_i + 1, j - start_j + 1);
                }
            }
        }
    }
}

/* proc_point_module.F:473-541 */
void orc_accum_continuous(orc_proj *src, orc_proj *dom, int istagger, const float *tile,
                          int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                          int src_min_z, int src_max_z, int bdr,
                          float *dst, float *n, int start_i, int end_i, int start_j, int end_j,
                          int start_k, int end_k, int32_t *processed_pts, int32_t *new_pts, float msgval,
                          float sr_x, float sr_y)
{
    actx g; memset(&g, 0, sizeof(g));
    g.src = src; g.dom = dom; g.tile = tile; g.smx = src_min_x; g.sMx = src_max_x; g.smy = src_min_y; g.sMy = src_max_y;
    g.smz = src_min_z; g.sMz = src_max_z; g.min_k = src_min_z; g.max_k = src_max_z;
    g.msgval = msgval; g.sr_x = sr_x; g.sr_y = sr_y; g.istagger = istagger;
    g.dst = dst; g.n = n; g.sx = start_i; g.ex = end_i; g.sy = start_j; g.ey = end_j; g.sz = start_k; g.ez = end_k;
    g.processed = processed_pts; g.new_pts = new_pts; g.nxw = (end_i - start_i + 1 + 31) / 32; g.categorical = 0;
    a_run(&g, bdr);
}

/* is_point_in_tile, proc_point_module.F:894-937 (tile-interior test on source coordinates) */
static int in_tile(float rx, float ry, int smx, int sMx, int smy, int sMy, int bdr)
{
    return (smx + bdr <= o_floor(rx + 0.5f) && o_ceiling(rx - 0.5f) <= sMx - bdr &&
            smy + bdr <= o_floor(ry + 0.5f) && o_ceiling(ry - 0.5f) <= sMy - bdr);
}

/* inner loop of calc_field, process_tile_module.F:1408-1519, applied to every destination point inside
 * the tile.  itype: 0 CONTINUOUS, 1 CATEGORICAL, 2 SP_CONTINUOUS.  get_point: proc_point_module.F:788-855. */
void orc_tile_point_fallback(orc_proj *src, int itype, const float *tile,
                             int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                             int src_min_z, int src_max_z, int bdr,
                             const float *xlat, const float *xlon, const float *landmask, int use_landmask,
                             float mask_val, const int *list, const int *opts, float msg_val, float msg_fill_val,
                             float *field, float *data_count,
                             int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                             const int32_t *processed_domain, int32_t *level_domain, long *n_invalid)
{
    int nx = end_i - start_i + 1, ny = end_j - start_j + 1, nxw = (nx + 31) / 32;
    long n2 = (long)nx * ny;
#define F3(a, i, j, k) (a)[(long)((k) - start_k) * n2 + (long)((j) - start_j) * nx + ((i) - start_i)]
#define P2(a, i, j) (a)[(long)((j) - start_j) * nx + ((i) - start_i)]
    for (int iy = start_j; iy <= end_j; ++iy) for (int ix = start_i; ix <= end_i; ++ix) {
        float rlat = P2(xlat, ix, iy), rlon = P2(xlon, ix, iy), rx, ry;
        if (rlon >= 180.0f) rlon = rlon - 360.0f;
        orc_lltoxy(src, rlat, rlon, &rx, &ry, ORC_M);
        if (!in_tile(rx, ry, src_min_x, src_max_x, src_min_y, src_max_y, bdr)) continue;
        if (orc_bit_test(level_domain, nxw, ix - start_i + 1, iy - start_j + 1)) continue;
        if (orc_bit_test(processed_domain, nxw, ix - start_i + 1, iy - start_j + 1)) {
            orc_bit_set(level_domain, nxw, ix - start_i + 1, iy - start_j + 1);
            continue;
        }
        int lm_ok = 1;
        if (use_landmask && !(P2(landmask, ix, iy) != mask_val)) lm_ok = 0;
        if (itype == 0 || itype == 2) {
            if (lm_ok) {
                for (int k = src_min_z; k <= src_max_z; ++k) {
                    float temp = orc_interp_sequence(rx, ry, k, tile, src_min_x, src_max_x, src_min_y, src_max_y,
                                                     src_min_z, src_max_z, msg_val, list, opts, 1, 0, ' ', 0.0f, NULL);
                    if (temp != msg_val) {
                        F3(field, ix, iy, k) = temp;
                        orc_bit_set(level_domain, nxw, ix - start_i + 1, iy - start_j + 1);
                        if (itype == 2) F3(data_count, ix, iy, k) = 1.0f;
                    } else F3(field, ix, iy, k) = msg_fill_val;
                }
            } else {
                for (int k = start_k; k <= end_k; ++k) F3(field, ix, iy, k) = msg_fill_val;
            }
        } else {
            if (lm_ok) {
                float temp = orc_interp_sequence(rx, ry, 1, tile, src_min_x, src_max_x, src_min_y, src_max_y,
                                                 src_min_z, src_max_z, msg_val, list, opts, 1, 0, ' ', 0.0f, NULL);
                for (int k = start_k; k <= end_k; ++k) F3(field, ix, iy, k) = 0.0f;
                if (temp != msg_val) {
                    int c = (int)temp;
                    if (c >= start_k && c <= end_k) F3(field, ix, iy, c) = F3(field, ix, iy, c) + 1.0f;
                    else if (n_invalid) (*n_invalid)++;
                    orc_bit_set(level_domain, nxw, ix - start_i + 1, iy - start_j + 1);
                }
            } else {
                for (int k = start_k; k <= end_k; ++k) F3(field, ix, iy, k) = 0.0f;
            }
        }
    }
#undef F3
#undef P2
}

/* process_tile_module.F:1580-1659 */
void orc_calc_field_finish(int itype, float *field, float *data_count, const float *landmask, int use_landmask,
                           float mask_val, float msg_fill_val, int has_scale, float scale_factor,
                           int start_i, int end_i, int start_j, int end_j, int start_k, int end_k,
                           int32_t *processed_domain, const int32_t *level_domain)
{
    int nx = end_i - start_i + 1, ny = end_j - start_j + 1, nxw = (nx + 31) / 32;
    long n2 = (long)nx * ny;
#define F3(a, i, j, k) (a)[(long)((k) - start_k) * n2 + (long)((j) - start_j) * nx + ((i) - start_i)]
#define P2(a, i, j) (a)[(long)((j) - start_j) * nx + ((i) - start_i)]
    if (itype == 2) {
        for (int j = start_j; j <= end_j; ++j) for (int i = start_i; i <= end_i; ++i) {
            int proc = orc_bit_test(processed_domain, nxw, i - start_i + 1, j - start_j + 1);
            if (!use_landmask || P2(landmask, i, j) != mask_val) {
                for (int k = start_k; k <= end_k; ++k) {
                    if (F3(data_count, i, j, k) > 0.0f) F3(field, i, j, k) = F3(field, i, j, k) / F3(data_count, i, j, k);
                    else if (!proc) F3(field, i, j, k) = msg_fill_val;
                }
            } else if (!proc) {
                for (int k = start_k; k <= end_k; ++k) F3(field, i, j, k) = msg_fill_val;
            }
        }
    } else if (itype == 1) {
        if (use_landmask)
            for (int j = start_j; j <= end_j; ++j) for (int i = start_i; i <= end_i; ++i)
                if (P2(landmask, i, j) == mask_val) for (int k = start_k; k <= end_k; ++k) F3(field, i, j, k) = 0.0f;
    }
    if (has_scale) {
        for (int i = start_i; i <= end_i; ++i) for (int j = start_j; j <= end_j; ++j)
            if (orc_bit_test(level_domain, nxw, i - start_i + 1, j - start_j + 1) &&
                !orc_bit_test(processed_domain, nxw, i - start_i + 1, j - start_j + 1))
                for (int k = start_k; k <= end_k; ++k)
                    if (F3(field, i, j, k) != msg_fill_val) F3(field, i, j, k) = F3(field, i, j, k) * scale_factor;
    }
    long nw = (long)nxw * ny;
    for (long w = 0; w < nw; ++w) processed_domain[w] |= level_domain[w];
#undef F3
#undef P2
}

/* process_tile_module.F:539-555, 1035-1051.  field(npts, ncat), point index fastest. */
void orc_fractions(float *field, long npts, int ncat, float msg_fill_val)
{
    for (long p = 0; p < npts; ++p) {
        float sum = 0.0f;
        for (int k = 0; k < ncat; ++k) sum = sum + field[(long)k * npts + p];
        if (sum > 0.0f) for (int k = 0; k < ncat; ++k) field[(long)k * npts + p] = field[(long)k * npts + p] / sum;
        else for (int k = 0; k < ncat; ++k) field[(long)k * npts + p] = msg_fill_val;
    }
}

/* process_tile_module.F:569-622 */
void orc_landmask_from_fractions(const float *field, long npts, int min_cat, int max_cat, const int *lm_vals,
                                 int n_lm, int is_water_mask, float msg_fill_val, float *landmask)
{
    for (long p = 0; p < npts; ++p) {
        float total = -1.0f;
        for (int k = 0; k < n_lm; ++k) {
            if (lm_vals[k] >= min_cat && lm_vals[k] <= max_cat) {
                float v = field[(long)(lm_vals[k] - min_cat) * npts + p];
                if (v != msg_fill_val) { if (total < 0.0f) total = 0.0f; total = total + v; }
                else { total = -1.0f; break; }
            }
        }
        if (total >= 0.0f) {
            if (is_water_mask) landmask[p] = (total < 0.50f) ? 1.0f : 0.0f;
            else landmask[p] = (total > 0.50f) ? 1.0f : 0.0f;
        } else landmask[p] = -1.0f;
    }
}

/* process_tile_module.F:1123-1144 */
void orc_dominant_category(const float *field, long npts, int min_cat, int max_cat, float msg_fill_val,
                           float *dominant_field)
{
    for (long p = 0; p < npts; ++p) {
        float dominant = 0.0f, d = (float)(min_cat - 1);
        for (int k = min_cat; k <= max_cat; ++k) {
            float v = field[(long)(k - min_cat) * npts + p];
            if (v > dominant && v != msg_fill_val) { d = (float)k; dominant = v; }
        }
        if (d == (float)(min_cat - 1)) d = msg_fill_val;
        dominant_field[p] = d;
    }
}

/* process_tile_module.F:699-727 (dominant category of the landmask field) */
void orc_dominant_category_landmask(const float *field, const float *landmask, long npts, int min_cat,
                                    int max_cat, const int *lm_vals, int n_lm, int is_water_mask,
                                    float *dominant_field)
{
    for (long p = 0; p < npts; ++p) {
        float dominant = 0.0f, d = (float)(min_cat - 1);
        if ((landmask[p] == 1.0f && is_water_mask) || (landmask[p] == 0.0f && !is_water_mask)) {
            for (int k = min_cat; k <= max_cat; ++k) {
                int kk;
                for (kk = 1; kk <= n_lm; ++kk) if (k == lm_vals[kk - 1]) break;
                float v = field[(long)(k - min_cat) * npts + p];
                if (v > dominant && kk > n_lm) { d = (float)k; dominant = v; }
            }
        } else {
            for (int k = min_cat; k <= max_cat; ++k)
                for (int kk = 1; kk <= n_lm; ++kk) {
                    float v = field[(long)(k - min_cat) * npts + p];
                    if (v > dominant && k == lm_vals[kk - 1]) { d = (float)k; dominant = v; }
                }
        }
        dominant_field[p] = d;
    }
}

/* ------------------------------------------------------------------ */
/* process_tile_module.F:1692-1724 */
void orc_get_lat_lon_fields(orc_proj *dom, float *xlat, float *xlon, int si, int sj, int ei, int ej,
                            int stagger, float sub_x, float sub_y)
{
    int nx = ei - si + 1;
    for (int i = si; i <= ei; ++i) for (int j = sj; j <= ej; ++j) {
        float x = ((float)i - 0.5f) / sub_x + 0.5f, y = ((float)j - 0.5f) / sub_y + 0.5f;
        orc_xytoll(dom, x, y, &xlat[(long)(j - sj) * nx + (i - si)], &xlon[(long)(j - sj) * nx + (i - si)], stagger);
    }
}

/* process_tile_module.F:1735-1876 (PROJ_ROTLL => 1.0) */
void orc_get_map_factor(const orc_proj *dom, const float *xlat, const float *xlon, float *mx, float *my,
                        long npts, float truelat1, float truelat2, float pole_lat, float pole_lon, float stand_lon)
{
    const float R = ORC_RAD_PER_DEG;
    int code = dom->code;
    if (code == ORC_PROJ_LC) {
        if (truelat1 != truelat2) {
            float colat1 = R * (90.0f - truelat1), colat2 = R * (90.0f - truelat2);
            float n = (o_log(o_sin(colat1)) - o_log(o_sin(colat2))) / (o_log(o_tan(colat1 / 2.0f)) - o_log(o_tan(colat2 / 2.0f)));
            for (long p = 0; p < npts; ++p) {
                float colat = R * (90.0f - xlat[p]);
                mx[p] = o_sin(colat2) / o_sin(colat) * o_pow(o_tan(colat / 2.0f) / o_tan(colat2 / 2.0f), n);
                my[p] = mx[p];
            }
        } else {
            float colat0 = R * (90.0f - truelat1);
            for (long p = 0; p < npts; ++p) {
                float colat = R * (90.0f - xlat[p]);
                mx[p] = o_sin(colat0) / o_sin(colat) * o_pow(o_tan(colat / 2.0f) / o_tan(colat0 / 2.0f), o_cos(colat0));
                my[p] = mx[p];
            }
        }
    } else if (code == ORC_PROJ_PS) {
        float sg = (truelat1 >= 0.0f) ? 1.0f : -1.0f; /* sign(1.,truelat1) */
        for (long p = 0; p < npts; ++p) {
            mx[p] = (1.0f + o_sin(R * fabsf(truelat1))) / (1.0f + o_sin(R * sg * xlat[p]));
            my[p] = mx[p];
        }
    } else if (code == ORC_PROJ_MERC) {
        float colat0 = R * (90.0f - truelat1);
        for (long p = 0; p < npts; ++p) { float colat = R * (90.0f - xlat[p]); mx[p] = o_sin(colat0) / o_sin(colat); my[p] = mx[p]; }
    } else if (code == ORC_PROJ_CYL) {
        for (long p = 0; p < npts; ++p) {
            if (fabsf(xlat[p]) == 90.0f) mx[p] = 0.0f; else mx[p] = 1.0f / o_cos(xlat[p] * R);
            my[p] = 1.0f;
        }
    } else if (code == ORC_PROJ_CASSINI) {
        if (fabsf(pole_lat) == 90.0f) {
            for (long p = 0; p < npts; ++p) {
                if (fabsf(xlat[p]) >= 90.0f) mx[p] = 0.0f; else mx[p] = 1.0f / o_cos(xlat[p] * R);
                my[p] = 1.0f;
            }
        } else {
            for (long p = 0; p < npts; ++p) {
                float cl, co;
                orc_rotate_coords(xlat[p], xlon[p], &cl, &co, pole_lat, pole_lon, stand_lon, -1);
                if (fabsf(cl) >= 90.0f) mx[p] = 0.0f; else mx[p] = 1.0f / o_cos(cl * R);
                my[p] = 1.0f;
            }
        }
    } else {
        for (long p = 0; p < npts; ++p) { mx[p] = 1.0f; my[p] = 1.0f; }
    }
}

/* process_tile_module.F:1885-1909; OMEGA_E constants_module.F:8 */
void orc_get_coriolis(const float *xlat, float *f, float *e, long npts)
{
    const float OMEGA_E = 7.292e-5f;
    for (long p = 0; p < npts; ++p) {
        f[p] = 2.0f * OMEGA_E * o_sin(ORC_RAD_PER_DEG * xlat[p]);
        e[p] = 2.0f * OMEGA_E * o_cos(ORC_RAD_PER_DEG * xlat[p]);
    }
}

/* process_tile_module.F:1920-2067 */
void orc_get_rotang(int iproj, const float *xlat, const float *xlon, float *cosa, float *sina,
                    int si, int sj, int ei, int ej, float truelat1, float truelat2, float stand_lon)
{
    int nx = ei - si + 1;
#define Q(a, i, j) (a)[(long)((j) - sj) * nx + ((i) - si)]
    if (iproj == ORC_PROJ_LC || iproj == ORC_PROJ_PS) {
        float cone = 1.0f;
        if (iproj == ORC_PROJ_LC) orc_lc_cone(truelat1, truelat2, &cone);
        for (int i = si; i <= ei; ++i) for (int j = sj; j <= ej; ++j) {
            float d_lon = stand_lon - Q(xlon, i, j);
            if (d_lon > 180.0f) d_lon = d_lon - 360.0f; else if (d_lon < -180.0f) d_lon = d_lon + 360.0f;
            float alpha = (iproj == ORC_PROJ_LC) ? d_lon * cone * ORC_RAD_PER_DEG : d_lon * ORC_RAD_PER_DEG;
            Q(sina, i, j) = o_sin(alpha); Q(cosa, i, j) = o_cos(alpha);
        }
    } else if (iproj == ORC_PROJ_MERC || iproj == ORC_PROJ_CYL) {
        for (int i = si; i <= ei; ++i) for (int j = sj; j <= ej; ++j) { Q(sina, i, j) = 0.0f; Q(cosa, i, j) = 1.0f; }
    } else if (iproj == ORC_PROJ_CASSINI) {
        for (int i = si; i <= ei; ++i) for (int j = sj; j <= ej; ++j) {
            int jm = (j == sj) ? j : j - 1, jp = (j == ej) ? j : j + 1;
            float d_lon = Q(xlon, i, jp) - Q(xlon, i, jm);
            if (d_lon > 180.0f) d_lon = d_lon - 360.0f; else if (d_lon < -180.0f) d_lon = d_lon + 360.0f;
            float alpha = o_atan2(-o_cos(Q(xlat, i, j) * ORC_RAD_PER_DEG) * (d_lon * ORC_RAD_PER_DEG),
                                  ((Q(xlat, i, jp) - Q(xlat, i, jm)) * ORC_RAD_PER_DEG));
            Q(sina, i, j) = o_sin(alpha); Q(cosa, i, j) = o_cos(alpha);
        }
    }
#undef Q
}

/* process_tile_module.F:2197-2253.  Edge columns use mapfac(i,j) with the STALE loop index i == end_mem_i
 * (value of the DO variable after `do i=start_mem_i+1,end_mem_i-1` terminates), :2228,2232. */
void orc_calc_dfdx(const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek, int sr_x,
                   const float *mapfac, float dom_dx)
{
    int nx = ei - si + 1, ny = ej - sj + 1; long n2 = (long)nx * ny;
#define S3(a, i, j, k) (a)[(long)((k) - sk) * n2 + (long)((j) - sj) * nx + ((i) - si)]
#define M2(i, j) mapfac[(long)((j) - sj) * nx + ((i) - si)]
    float sr = (float)sr_x;
    int istale = (si + 1 <= ei - 1) ? ei : si + 1; /* DO variable after loop exit */
    for (int k = sk; k <= ek; ++k) {
        for (int i = si + 1; i <= ei - 1; ++i) for (int j = sj; j <= ej; ++j) {
            if (mapfac) S3(dst, i, j, k) = (S3(src, i + 1, j, k) - S3(src, i - 1, j, k)) / (2.0f * dom_dx * M2(i, j) / sr);
            else S3(dst, i, j, k) = (S3(src, i + 1, j, k) - S3(src, i - 1, j, k)) / (2.0f * dom_dx / sr);
        }
        for (int j = sj; j <= ej; ++j) {
            if (mapfac) {
                S3(dst, si, j, k) = (S3(src, si + 1, j, k) - S3(src, si, j, k)) / (dom_dx * M2(istale, j) / sr);
                S3(dst, ei, j, k) = (S3(src, ei, j, k) - S3(src, ei - 1, j, k)) / (dom_dx * M2(istale, j) / sr);
            } else {
                S3(dst, si, j, k) = (S3(src, si + 1, j, k) - S3(src, si, j, k)) / (dom_dx / sr);
                S3(dst, ei, j, k) = (S3(src, ei, j, k) - S3(src, ei - 1, j, k)) / (dom_dx / sr);
            }
        }
    }
}

/* process_tile_module.F:2132-2188; edge rows use mapfac(i,j) with stale j == end_mem_j (:2163,2167) */
void orc_calc_dfdy(const float *src, float *dst, int si, int sj, int sk, int ei, int ej, int ek, int sr_y,
                   const float *mapfac, float dom_dy)
{
    int nx = ei - si + 1, ny = ej - sj + 1; long n2 = (long)nx * ny;
    float sr = (float)sr_y;
    int jstale = (sj + 1 <= ej - 1) ? ej : sj + 1;
    for (int k = sk; k <= ek; ++k) {
        for (int i = si; i <= ei; ++i) for (int j = sj + 1; j <= ej - 1; ++j) {
            if (mapfac) S3(dst, i, j, k) = (S3(src, i, j + 1, k) - S3(src, i, j - 1, k)) / (2.0f * dom_dy * M2(i, j) / sr);
            else S3(dst, i, j, k) = (S3(src, i, j + 1, k) - S3(src, i, j - 1, k)) / (2.0f * dom_dy / sr);
        }
        for (int i = si; i <= ei; ++i) {
            if (mapfac) {
                S3(dst, i, sj, k) = (S3(src, i, sj + 1, k) - S3(src, i, sj, k)) / (dom_dy * M2(i, jstale) / sr);
                S3(dst, i, ej, k) = (S3(src, i, ej, k) - S3(src, i, ej - 1, k)) / (dom_dy * M2(i, jstale) / sr);
            } else {
                S3(dst, i, sj, k) = (S3(src, i, sj + 1, k) - S3(src, i, sj, k)) / (dom_dy / sr);
                S3(dst, i, ej, k) = (S3(src, i, ej, k) - S3(src, i, ej - 1, k)) / (dom_dy / sr);
            }
        }
    }
#undef S3
#undef M2
}

/* one_two_one_egrid / smth_desmth_egrid, smooth_module.F:252-338, :448-585 (NMM E-grid; hflag 1.0 = mass points, else velocity
 * points).  `scratch(ix,iy,1) = array(ix,iy,1)` (:291, :503) initialises level index 1 only: for the other levels of a multi-level
 * field the reference reads uninitialised scratch at the points the first sweep does not write; here every level starts from
 * the array's values.  smth-desmth_special is 'not currently implemented for NMM' (process_tile_module.F:675): no-op. */
void orc_smooth_egrid(float *array, int sx, int ex, int sy, int ey, int sz, int ez, int npass, int opt, float msgval, float hflag)
{
    if (opt != ORC_ONETWOONE && opt != ORC_SMTHDESMTH) return;
    int nx = ex - sx + 1, ny = ey - sy + 1, nz = ez - sz + 1;
    long n2 = (long)nx * ny, n3 = n2 * nz;
#define EA(a, ix, iy, iz) (a)[(long)((iz) - sz) * n2 + (long)((iy) - sy) * nx + ((ix) - sx)]
    float *scratch = (float *)malloc(sizeof(float) * (size_t)n3);
    int *ihe = (int *)malloc(sizeof(int) * ny), *ihw = (int *)malloc(sizeof(int) * ny), *istart = (int *)malloc(sizeof(int) * ny);
    for (int iy = sy; iy <= ey; ++iy) {
        if (hflag == 1.0f) { ihe[iy - sy] = abs((iy + 1) % 2); istart[iy - sy] = (iy % 2 == 0) ? sx : sx + 1; }
        else { ihe[iy - sy] = abs(iy % 2); istart[iy - sy] = (abs(iy % 2) == 1) ? sx : sx + 1; }
        ihw[iy - sy] = ihe[iy - sy] - 1;
    }
    const float cenwgt = 1.52f, endwgt = 0.26f;
#define ECOND(ix, iy, iz) ((msgval == 1.0f && EA(array, ix, iy, iz) != 0.0f) || msgval != 1.0f)
    for (int ipass = 1; ipass <= npass; ++ipass) {
        memcpy(scratch, array, sizeof(float) * (size_t)n3);
        for (int iy = sy + 1; iy <= ey - 1; ++iy) for (int ix = istart[iy - sy]; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            if (ECOND(ix, iy, iz))
                EA(scratch, ix, iy, iz) = 0.50f * EA(array, ix, iy, iz) + 0.25f * (EA(array, ix + ihw[iy - sy], iy - 1, iz) + EA(array, ix + ihe[iy - sy], iy + 1, iz));
        for (int iy = sy + 1; iy <= ey - 1; ++iy) for (int ix = istart[iy - sy]; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            if (ECOND(ix, iy, iz))
                EA(array, ix, iy, iz) = 0.50f * EA(scratch, ix, iy, iz) + 0.25f * (EA(scratch, ix + ihe[iy - sy], iy - 1, iz) + EA(scratch, ix + ihw[iy - sy], iy + 1, iz));
        if (opt != ORC_SMTHDESMTH) continue;
        for (int iy = sy + 2; iy <= ey - 2; ++iy) for (int ix = istart[iy - sy]; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            if (ECOND(ix, iy, iz))
                EA(scratch, ix, iy, iz) = cenwgt * EA(array, ix, iy, iz) - endwgt * (EA(array, ix + ihw[iy - sy], iy - 1, iz) + EA(array, ix + ihe[iy - sy], iy + 1, iz));
        for (int iy = sy + 2; iy <= ey - 2; ++iy) for (int ix = istart[iy - sy]; ix <= ex - 1; ++ix) for (int iz = sz; iz <= ez; ++iz)
            if (ECOND(ix, iy, iz))
                EA(array, ix, iy, iz) = cenwgt * EA(scratch, ix, iy, iz) - endwgt * (EA(scratch, ix + ihe[iy - sy], iy - 1, iz) + EA(scratch, ix + ihw[iy - sy], iy + 1, iz));
    }
#undef ECOND
#undef EA
    free(scratch); free(ihe); free(ihw); free(istart);
}
