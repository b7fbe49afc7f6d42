/*
 * orc_metgrid.c -- CPU ORACLE (test infrastructure): metgrid per-slab driver and wind rotation.
 * Restates metgrid/src/process_domain_module.F:2205-2669,3347-3717 and
 * metgrid/src/rotate_winds_module.F:85-387 of WPS v4.6.0.
 * PARITY UNPINNED (see wps_oracle.h).  Not product code.
 */
#include "orc_internal.h"
#include <stdlib.h>
#include <string.h>

/* process_domain_module.F:2594-2669 (istagger is always M at every call site, :1246-1357) */
float orc_interp_to_latlon(orc_proj *src, float rlat, float rlon, const int *list, const int *opts,
                           const float *slab, int minx, int maxx, int miny, int maxy, int bdr, float msg,
                           int mask_mode, char rel, float maskval, const float *mask)
{
    float r, rx, ry;
    orc_lltoxy(src, rlat, rlon, &rx, &ry, ORC_M);
    if (rx >= (float)(minx + bdr) - 0.5f && rx <= (float)(maxx - bdr) + 0.5f)
        r = orc_interp_sequence(rx, ry, 1, slab, minx, maxx, miny, maxy, 1, 1, msg, list, opts, 1, mask_mode, rel, maskval, mask);
    else r = msg;
    if (r == msg) {
        if (rlon < 0.0f) {
            orc_lltoxy(src, rlat, rlon + 360.0f, &rx, &ry, ORC_M);
            if (rx >= (float)(minx + bdr) - 0.5f && rx <= (float)(maxx - bdr) + 0.5f)
                r = orc_interp_sequence(rx, ry, 1, slab, minx, maxx, miny, maxy, 1, 1, msg, list, opts, 1, mask_mode, rel, maskval, mask);
            else r = msg;
        }
    }
    return r;
}

#define D2(a, i, j) (a)[(long)((j) - sm2) * (em1 - sm1 + 1) + ((i) - sm1)]

static int process_width_of(int process_bdy_width, int ifieldstagger, int sd1, int ed1, int sd2, int ed2)
{
    /* process_domain_module.F:2270-2278 */
    int pw;
    if (process_bdy_width == 0) {
        pw = ed1 - sd1 + 1; if (ed2 - sd2 + 1 > pw) pw = ed2 - sd2 + 1;
    } else {
        pw = process_bdy_width;
        if (ifieldstagger != ORC_M) pw = pw + 2;
    }
    return pw;
}

/* process_domain_module.F:2483-2581 */
void orc_interp_met_field(orc_proj *src, const orc_field_opts *fo, int ifieldstagger,
                          const float *xlat, const float *xlon, int sm1, int em1, int sm2, int em2,
                          int sd1, int ed1, int sd2, int ed2,
                          const float *slab, int minx, int maxx, int miny, int maxy, int bdr,
                          const float *mask, const float *land_mask, const float *water_mask,
                          const float *landmask, int process_bdy_width,
                          float *r_arr, const int32_t *valid_mask, int32_t *new_pts,
                          int i0, int i1, int j0, int j1)
{
    int nxw = (em1 - sm1 + 1 + 31) / 32;
    int pw = process_width_of(process_bdy_width, ifieldstagger, sd1, ed1, sd2, ed2);
    float msg = fo->missing_value, fill = fo->fill_missing;
    float temp = 0.0f; /* uninitialised in the Fortran; only observable when landmask is neither 0 nor 1 with masked=both */
    for (int j = j0; j <= j1; ++j) {
        for (int i = i0; i <= i1; ++i) {
            if (abs(i - sd1) >= pw && abs(i - ed1) >= pw && abs(j - sd2) >= pw && abs(j - ed2) >= pw) {
                D2(r_arr, i, j) = fill;
                orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
                continue;
            }
            float lat = D2(xlat, i, j), lon = D2(xlon, i, j);
            if (landmask) {
                float lm = D2(landmask, i, j);
                if (fo->masked == ORC_MASKED_BOTH) {
                    if (lm == 0.0f) {
                        temp = orc_interp_to_latlon(src, lat, lon, fo->ops, fo->opts, slab, minx, maxx, miny, maxy, bdr, msg,
                                                    land_mask != NULL, fo->mask_rel[1], fo->mask_val[1], land_mask);
                    } else if (lm == 1.0f) {
                        temp = orc_interp_to_latlon(src, lat, lon, fo->ops, fo->opts, slab, minx, maxx, miny, maxy, bdr, msg,
                                                    water_mask != NULL, fo->mask_rel[2], fo->mask_val[2], water_mask);
                    }
                } else if (lm != fo->masked) {
                    temp = orc_interp_to_latlon(src, lat, lon, fo->ops, fo->opts, slab, minx, maxx, miny, maxy, bdr, msg,
                                                mask != NULL, fo->mask_rel[0], fo->mask_val[0], mask);
                } else temp = msg;
            } else {
                temp = orc_interp_to_latlon(src, lat, lon, fo->ops, fo->opts, slab, minx, maxx, miny, maxy, bdr, msg,
                                            mask != NULL, fo->mask_rel[0], fo->mask_val[0], mask);
            }
            if (temp != msg) {
                D2(r_arr, i, j) = temp;
                orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
            } else if (landmask) {
                if (D2(landmask, i, j) == fo->masked) {
                    D2(r_arr, i, j) = fill;
                    orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
                }
            }
            if (!orc_bit_test(new_pts, nxw, i - sm1 + 1, j - sm2 + 1) &&
                !orc_bit_test(valid_mask, nxw, i - sm1 + 1, j - sm2 + 1)) {
                D2(r_arr, i, j) = fill;
                if (fill != ORC_NAN_F) orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
            }
        }
    }
}

/* ---- metgrid's private accum_continuous, process_domain_module.F:3347-3717 ---- */
typedef struct {
    orc_proj *src, *dom;
    const float *tile; int smx, sMx, smy, sMy;   /* src_min_x.. */
    const float *mask; char rel; float maskval; int mask_mode; /* 0 none, 1 relational */
    float msgval, sr_x, sr_y; int istagger;
    float *dst, *n; int sx, ex, sy, ey;
    int32_t *new_pts; int nxw;
    int *wm; /* where_maps_to(:,:,2) */
} gctx;
#define WM(g, i, j, c) (g)->wm[(((long)((j) - (g)->smy)) * ((g)->sMx - (g)->smx + 1) + ((i) - (g)->smx)) * 2 + (c)]
#define TL(g, i, j) (g)->tile[(long)((j) - (g)->smy) * ((g)->sMx - (g)->smx + 1) + ((i) - (g)->smx)]
#define TM(g, i, j) (g)->mask[(long)((j) - (g)->smy) * ((g)->sMx - (g)->smx + 1) + ((i) - (g)->smx)]
#define GD(g, a, i, j) (a)[(long)((j) - (g)->sy) * ((g)->ex - (g)->sx + 1) + ((i) - (g)->sx)]

static void g_corner(gctx *g, int i, int j)
{
    if (WM(g, i, j, 0) != ORC_NOT_PROCESSED) return;
    float lat, lon, rx, ry;
    orc_xytoll(g->src, (float)i, (float)j, &lat, &lon, g->istagger);
    orc_lltoxy(g->dom, lat, lon, &rx, &ry, g->istagger);
    rx = (rx - 0.5f) * g->sr_x + 0.5f;
    ry = (ry - 0.5f) * g->sr_y + 0.5f;
    if ((float)g->sx <= rx && rx <= (float)g->ex && (float)g->sy <= ry && ry <= (float)g->ey) {
        WM(g, i, j, 0) = o_nint(rx); WM(g, i, j, 1) = o_nint(ry);
    } else WM(g, i, j, 0) = ORC_OUTSIDE_DOMAIN;
}
static inline int g_usable(const gctx *g, int i, int j)
{
    float v = TL(g, i, j);
    if (v == g->msgval) return 0;
    if (g->mask_mode) {
        float m = TM(g, i, j);
        if (g->rel == ' ') return m != g->maskval;
        if (g->rel == '<') return m >= g->maskval;
        if (g->rel == '>') return m <= g->maskval;
        return 0;
    }
    return 1;
}
static void g_credit(gctx *g, int i, int j, int xd, int yd)
{
    if (g_usable(g, i, j)) {
        GD(g, g->dst, xd, yd) = GD(g, g->dst, xd, yd) + TL(g, i, j);
        GD(g, g->n, xd, yd) = GD(g, g->n, xd, yd) + 1.0f;
        orc_bit_set(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1);
    }
}
static void g_block(gctx *g, int min_i, int min_j, int max_i, int max_j)
{
    g_corner(g, min_i, min_j); g_corner(g, min_i, max_j); g_corner(g, max_i, max_j); g_corner(g, max_i, min_j);
    if (WM(g, min_i, min_j, 0) == WM(g, min_i, max_j, 0) && WM(g, min_i, min_j, 0) == WM(g, max_i, max_j, 0) &&
        WM(g, min_i, min_j, 0) == WM(g, max_i, min_j, 0) && WM(g, min_i, min_j, 1) == WM(g, min_i, max_j, 1) &&
        WM(g, min_i, min_j, 1) == WM(g, max_i, max_j, 1) && WM(g, min_i, min_j, 1) == WM(g, max_i, min_j, 1) &&
        WM(g, min_i, min_j, 0) != ORC_OUTSIDE_DOMAIN) {
        int xd = WM(g, min_i, min_j, 0), yd = WM(g, min_i, min_j, 1);
        if (!orc_bit_test(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1)) GD(g, g->dst, xd, yd) = 0.0f;
        for (int i = min_i; i <= max_i; ++i) for (int j = min_j; j <= max_j; ++j) g_credit(g, i, j, xd, yd);
    } else if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) {
        for (int i = min_i; i <= max_i; ++i) for (int j = min_j; j <= max_j; ++j) {
            int xd = WM(g, i, j, 0), yd = WM(g, i, j, 1);
            if (xd != ORC_OUTSIDE_DOMAIN) {
                if (!orc_bit_test(g->new_pts, g->nxw, xd - g->sx + 1, yd - g->sy + 1)) GD(g, g->dst, xd, yd) = 0.0f;
                g_credit(g, i, j, xd, yd);
            }
        }
    } else {
        int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
        g_block(g, min_i, min_j, ci, cj);
        if (ci < max_i) g_block(g, ci + 1, min_j, max_i, cj);
        if (cj < max_j) g_block(g, min_i, cj + 1, ci, max_j);
        if (ci < max_i && cj < max_j) g_block(g, ci + 1, cj + 1, max_i, max_j);
    }
}

/* process_domain_module.F:2331-2478 */
void orc_interp_met_field_gcell(orc_proj *src, orc_proj *dom, const orc_field_opts *fo, int ifieldstagger,
                                const float *xlat, const float *xlon, int sm1, int em1, int sm2, int em2,
                                int sd1, int ed1, int sd2, int ed2,
                                const float *slab, int minx, int maxx, int miny, int maxy, int bdr,
                                const float *mask, const float *landmask, int process_bdy_width,
                                float *r_arr, const int32_t *valid_mask, int32_t *new_pts)
{
    long npts = (long)(em1 - sm1 + 1) * (em2 - sm2 + 1);
    int nxw = (em1 - sm1 + 1 + 31) / 32;
    int pw = process_width_of(process_bdy_width, ifieldstagger, sd1, ed1, sd2, ed2);
    float msg = fo->missing_value, fill = fo->fill_missing;
    float *cur = (float *)malloc(sizeof(float) * (size_t)npts);
    float *cnt = (float *)calloc((size_t)npts, sizeof(float));
    memcpy(cur, r_arr, sizeof(float) * (size_t)npts);
    gctx g;
    g.src = src; g.dom = dom; g.tile = slab; g.smx = minx; g.sMx = maxx; g.smy = miny; g.sMy = maxy;
    g.mask = mask; g.rel = fo->mask_rel[0]; g.maskval = mask ? fo->mask_val[0] : -1.0f; g.mask_mode = mask != NULL;
    g.msgval = msg; g.sr_x = 1.0f; g.sr_y = 1.0f; g.istagger = ORC_M;
    g.dst = cur; g.n = cnt; g.sx = sm1; g.ex = em1; g.sy = sm2; g.ey = em2; g.new_pts = new_pts; g.nxw = nxw;
    long nw = (long)(maxx - minx + 1) * (maxy - miny + 1) * 2;
    g.wm = (int *)malloc(sizeof(int) * (size_t)nw);
    for (long k = 0; k < nw; k += 2) g.wm[k] = ORC_NOT_PROCESSED;
    g_block(&g, minx + bdr, miny, maxx - bdr, maxy);
    free(g.wm);
    for (int j = sm2; j <= em2; ++j) {
        for (int i = sm1; i <= em1; ++i) {
            if (abs(i - sd1) >= pw && abs(i - ed1) >= pw && abs(j - sd2) >= pw && abs(j - ed2) >= pw) {
                D2(r_arr, i, j) = fill; orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1); continue;
            }
            int do_pt = 1;
            if (landmask && !(D2(landmask, i, j) != fo->masked)) do_pt = 0;
            if (do_pt) {
                if (D2(cnt, i, j) > 0.0f) {
                    D2(r_arr, i, j) = D2(cur, i, j) / D2(cnt, i, j);
                    orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
                } else {
                    float temp = orc_interp_to_latlon(src, D2(xlat, i, j), D2(xlon, i, j), fo->ops, fo->opts, slab,
                                                      minx, maxx, miny, maxy, bdr, msg, mask != NULL, fo->mask_rel[0],
                                                      fo->mask_val[0], mask);
                    if (temp != msg) { D2(r_arr, i, j) = temp; orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1); }
                }
            } else {
                D2(r_arr, i, j) = fill; orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
            }
            if (!orc_bit_test(new_pts, nxw, i - sm1 + 1, j - sm2 + 1) &&
                !orc_bit_test(valid_mask, nxw, i - sm1 + 1, j - sm2 + 1)) {
                D2(r_arr, i, j) = fill;
                if (fill != ORC_NAN_F) orc_bit_set(new_pts, nxw, i - sm1 + 1, j - sm2 + 1);
            }
        }
    }
    free(cur); free(cnt);
}

/* Cassini rotation angle, rotate_winds_module.F:174-222 (U) and :281-329 (V): same formulas on the array passed in.
 * idir == 1: finite difference of ij_to_latlon(SOURCE_PROJ) at y +- 0.1; idir == -1: finite difference of the
 * domain's own lat/lon arrays along j (one-sided on the first/last row). */
static float cassini_alpha(const orc_proj *proj, int idir, const float *xlat, const float *xlon,
                           int s1, int s2, int e2, int nx, int i, int j)
{
#define LL(a, ii, jj) (a)[(long)((jj) - s2) * nx + ((ii) - s1)]
    if (idir == 1) {
        float x, y, latp, lonp, latm, lonm;
        orc_latlon_to_ij(proj, LL(xlat, i, j), LL(xlon, i, j), &x, &y);
        orc_ij_to_latlon(proj, x, y + 0.1f, &latp, &lonp);
        orc_ij_to_latlon(proj, x, y - 0.1f, &latm, &lonm);
        float diff_lon = lonp - lonm;
        if (diff_lon > 180.0f) diff_lon = diff_lon - 360.0f; else if (diff_lon < -180.0f) diff_lon = diff_lon + 360.0f;
        float diff_lat = latp - latm;
        return -o_atan2(-o_cos(LL(xlat, i, j) * ORC_RAD_PER_DEG) * diff_lon * ORC_RAD_PER_DEG, diff_lat * ORC_RAD_PER_DEG);
    }
    int ja, jb;   /* :190-222: last row looks back, first row looks ahead, interior rows are centred */
    if (j == e2) { ja = j; jb = j - 1; } else if (j == s2) { ja = j + 1; jb = j; } else { ja = j + 1; jb = j - 1; }
    float diff = LL(xlon, i, ja) - LL(xlon, i, jb);
    if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
    return o_atan2(-o_cos(LL(xlat, i, j) * ORC_RAD_PER_DEG) * diff * ORC_RAD_PER_DEG, (LL(xlat, i, ja) - LL(xlat, i, jb)) * ORC_RAD_PER_DEG);
#undef LL
}

/* rotate_winds_module.F:85-387, C grid.  xlat_u / xlat_v are only read for PROJ_CASSINI (may be NULL otherwise). */
void orc_metmap_xform_ll(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                         int us1, int us2, int ue1, int ue2, int vs1, int vs2, int ve1, int ve2,
                         const float *xlon_u, const float *xlon_v, const float *xlat_u, const float *xlat_v, int idir);
void orc_metmap_xform(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                      int us1, int us2, int ue1, int ue2, int vs1, int vs2, int ve1, int ve2,
                      const float *xlon_u, const float *xlon_v, int idir)
{
    orc_metmap_xform_ll(proj, u, u_mask, v, v_mask, us1, us2, ue1, ue2, vs1, vs2, ve1, ve2, xlon_u, xlon_v, NULL, NULL, idir);
}
void orc_metmap_xform_ll(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                         int us1, int us2, int ue1, int ue2, int vs1, int vs2, int ve1, int ve2,
                         const float *xlon_u, const float *xlon_v, const float *xlat_u, const float *xlat_v, int idir)
{
    if (!(proj->code == ORC_PROJ_LC || proj->code == ORC_PROJ_PS || proj->code == ORC_PROJ_CASSINI)) return;
    if (proj->code == ORC_PROJ_CASSINI && (!xlat_u || !xlat_v)) return;
    int unx = ue1 - us1 + 1, uny = ue2 - us2 + 1, vnx = ve1 - vs1 + 1, vny = ve2 - vs2 + 1;
    int unxw = (unx + 31) / 32, vnxw = (vnx + 31) / 32;
#define UU(a, i, j) (a)[(long)((j) - us2) * unx + ((i) - us1)]
#define VV(a, i, j) (a)[(long)((j) - vs2) * vnx + ((i) - vs1)]
    float *u_mult = (float *)malloc(sizeof(float) * (size_t)unx * uny), *v_mult = (float *)malloc(sizeof(float) * (size_t)vnx * vny);
    float *u_new = (float *)malloc(sizeof(float) * (size_t)unx * uny), *v_new = (float *)malloc(sizeof(float) * (size_t)vnx * vny);
    for (int j = us2; j <= ue2; ++j) for (int i = us1; i <= ue1; ++i)
        UU(u_mult, i, j) = orc_bit_test(u_mask, unxw, i - us1 + 1, j - us2 + 1) ? 1.0f : 0.0f;
    for (int j = vs2; j <= ve2; ++j) for (int i = vs1; i <= ve1; ++i)
        VV(v_mult, i, j) = orc_bit_test(v_mask, vnxw, i - vs1 + 1, j - vs2 + 1) ? 1.0f : 0.0f;
    int do_last_col_u, do_last_col_v, do_last_row_u, do_last_row_v;
    if (ue1 - us1 == ve1 - vs1) { do_last_col_u = 0; do_last_col_v = 1; } else { do_last_col_u = 1; do_last_col_v = 0; }
    if (ue2 - us2 == ve2 - vs2) { do_last_row_u = 1; do_last_row_v = 0; } else { do_last_row_u = 0; do_last_row_v = 1; }
    (void)do_last_row_v;
    for (int j = us2; j <= ue2; ++j) for (int i = us1; i <= ue1; ++i) {
        float diff = (float)idir * (UU(xlon_u, i, j) - proj->stdlon);
        if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
        float alpha;
        if (proj->code == ORC_PROJ_LC) alpha = diff * proj->cone * ORC_RAD_PER_DEG * proj->hemi;
        else if (proj->code == ORC_PROJ_CASSINI) alpha = cassini_alpha(proj, idir, xlat_u, xlon_u, us1, us2, ue2, unx, i, j);
        else alpha = diff * ORC_RAD_PER_DEG * proj->hemi;
        float v_weight = 0.0f, v_map = 0.0f;
        if (orc_bit_test(u_mask, unxw, i - us1 + 1, j - us2 + 1)) {
            float u_map = UU(u, i, j);
            if (i == us1) {
                if (j == ue2 && do_last_row_u) { v_weight = VV(v_mult, i, j); v_map = VV(v, i, j) * VV(v_mult, i, j); }
                else { v_weight = VV(v_mult, i, j) + VV(v_mult, i, j + 1); v_map = VV(v, i, j) * VV(v_mult, i, j) + VV(v, i, j + 1) * VV(v_mult, i, j + 1); }
            } else if (i == ue1 && do_last_col_u) {
                if (j == ue2 && do_last_row_u) { v_weight = VV(v_mult, i - 1, j); v_map = VV(v, i - 1, j); /* :244, not multiplied by v_mult */ }
                else { v_weight = VV(v_mult, i - 1, j) + VV(v_mult, i - 1, j + 1); v_map = VV(v, i - 1, j) * VV(v_mult, i - 1, j) + VV(v, i - 1, j + 1) * VV(v_mult, i - 1, j + 1); }
            } else if (j == ue2 && do_last_row_u) {
                v_weight = VV(v_mult, i - 1, j - 1) + VV(v_mult, i, j - 1);
                v_map = VV(v, i - 1, j - 1) * VV(v_mult, i - 1, j - 1) + VV(v, i, j - 1) * VV(v_mult, i, j - 1);
            } else {
                v_weight = VV(v_mult, i - 1, j) + VV(v_mult, i - 1, j + 1) + VV(v_mult, i, j) + VV(v_mult, i, j + 1);
                v_map = VV(v, i - 1, j) * VV(v_mult, i - 1, j) + VV(v, i - 1, j + 1) * VV(v_mult, i - 1, j + 1) +
                        VV(v, i, j) * VV(v_mult, i, j) + VV(v, i, j + 1) * VV(v_mult, i, j + 1);
            }
            if (v_weight > 0.0f) UU(u_new, i, j) = o_cos(alpha) * u_map + o_sin(alpha) * v_map / v_weight;
            else UU(u_new, i, j) = UU(u, i, j);
        } else UU(u_new, i, j) = UU(u, i, j);
    }
    for (int j = vs2; j <= ve2; ++j) for (int i = vs1; i <= ve1; ++i) {
        float diff = (float)idir * (VV(xlon_v, i, j) - proj->stdlon);
        if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
        float alpha;
        if (proj->code == ORC_PROJ_LC) alpha = diff * proj->cone * ORC_RAD_PER_DEG * proj->hemi;
        else if (proj->code == ORC_PROJ_CASSINI) alpha = cassini_alpha(proj, idir, xlat_v, xlon_v, vs1, vs2, ve2, vnx, i, j);
        else alpha = diff * ORC_RAD_PER_DEG * proj->hemi;
        float u_weight = 0.0f, u_map = 0.0f;
        if (orc_bit_test(v_mask, vnxw, i - vs1 + 1, j - vs2 + 1)) {
            float v_map = VV(v, i, j);
            if (j == vs2) {
                if (i == ve1 && do_last_col_v) { u_weight = UU(u_mult, i, j); u_map = UU(u, i, j) * UU(u_mult, i, j); }
                else { u_weight = UU(u_mult, i, j) + UU(u_mult, i + 1, j); u_map = UU(u, i, j) * UU(u_mult, i, j) + UU(u, i + 1, j) * UU(u_mult, i + 1, j); }
            } else if (j == ve2 && do_last_row_v) {
                if (i == ve1 && do_last_col_v) { u_weight = UU(u_mult, i, j - 1); u_map = UU(u, i, j - 1) * UU(u_mult, i, j - 1); }
                else { u_weight = UU(u_mult, i, j - 1) + UU(u_mult, i + 1, j - 1); u_map = UU(u, i, j - 1) * UU(u_mult, i, j - 1) + UU(u, i + 1, j - 1) * UU(u_mult, i + 1, j - 1); }
            } else if (i == ve1 && do_last_col_v) {
                u_weight = UU(u_mult, i, j) + UU(u_mult, i, j - 1);
                u_map = UU(u, i, j) * UU(u_mult, i, j) + UU(u, i, j - 1) * UU(u_mult, i, j - 1);
            } else {
                u_weight = UU(u_mult, i, j - 1) + UU(u_mult, i, j) + UU(u_mult, i + 1, j - 1) + UU(u_mult, i + 1, j);
                u_map = UU(u, i, j - 1) * UU(u_mult, i, j - 1) + UU(u, i, j) * UU(u_mult, i, j) +
                        UU(u, i + 1, j - 1) * UU(u_mult, i + 1, j - 1) + UU(u, i + 1, j) * UU(u_mult, i + 1, j);
            }
            if (u_weight > 0.0f) VV(v_new, i, j) = -o_sin(alpha) * u_map / u_weight + o_cos(alpha) * v_map;
            else VV(v_new, i, j) = VV(v, i, j);
        } else VV(v_new, i, j) = VV(v, i, j);
    }
    memcpy(u, u_new, sizeof(float) * (size_t)unx * uny);
    memcpy(v, v_new, sizeof(float) * (size_t)vnx * vny);
    free(u_mult); free(v_mult); free(u_new); free(v_new);
#undef UU
#undef VV
}

/* fill_missing_levels, the vertical-interpolation loop -- process_domain_module.F:2862-2921.  levels[] in the order
 * storage_get_levels returns them; r_arr [nlev][ny][nx], valid [nlev][ny][nxw] (bit arrays), both updated in place. */
void orc_fill_missing_levels(int nlev, const int32_t *levels, float *r_arr, int32_t *valid, int nx, int ny)
{
    int nxw = (nx + 31) / 32;
    long np = (long)nx * ny, nw = (long)ny * nxw;
    if (nlev <= 1) return;                                   /* "If this isn't a 3-d array, nothing to do" */
    for (int k = 0; k < nlev; ++k) {
        for (int jx = 1; jx <= ny; ++jx) {
            for (int ix = 1; ix <= nx; ++ix) {
                if (orc_bit_test(valid + k * nw, nxw, ix, jx)) continue;
                int lower, upper;
                for (lower = k - 1; lower >= 0; --lower)
                    if (orc_bit_test(valid + lower * nw, nxw, ix, jx)) break;
                for (upper = k + 1; upper < nlev; ++upper)
                    if (orc_bit_test(valid + upper * nw, nxw, ix, jx)) break;
                if (upper < nlev && lower >= 0) {
                    long o = (long)(jx - 1) * nx + (ix - 1);
                    float wl = (float)abs(levels[upper] - levels[k]) / (float)abs(levels[upper] - levels[lower]);
                    float wu = (float)abs(levels[k] - levels[lower]) / (float)abs(levels[upper] - levels[lower]);
                    r_arr[k * np + o] = wl * r_arr[lower * np + o] + wu * r_arr[upper * np + o];
                    orc_bit_set(valid + k * nw, nxw, ix, jx);
                }
            }
        }
    }
}

/* metmap_xform_nmm, rotate_winds_module.F:455-631: U and V collocated on the E-grid's velocity points */
void orc_metmap_xform_nmm(const orc_proj *proj, float *u, const int32_t *u_mask, float *v, const int32_t *v_mask,
                          int vs1, int vs2, int ve1, int ve2, const float *xlat_v, const float *xlon_v, int idir)
{
    const int nx = ve1 - vs1 + 1, ny = ve2 - vs2 + 1, nxw = (nx + 31) / 32;
    const int rot = proj->code == ORC_PROJ_ROTLL;
    if (!rot && !(proj->code == ORC_PROJ_LC || proj->code == ORC_PROJ_PS || proj->code == ORC_PROJ_CASSINI)) return;
    float *u_new = (float *)malloc(sizeof(float) * (size_t)nx * ny), *v_new = (float *)malloc(sizeof(float) * (size_t)nx * ny);
    float phi0 = proj->lat1 * ORC_RAD_PER_DEG, clontemp = proj->lon1, lmbd0;
    if (clontemp < 0.0f) lmbd0 = (clontemp + 360.0f) * ORC_RAD_PER_DEG; else lmbd0 = clontemp * ORC_RAD_PER_DEG;
    float sin_phi0 = o_sin(phi0), cos_phi0 = o_cos(phi0);
    const float fdir = (float)idir;
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const long k = (long)j * nx + i;
        float sin_alpha, cos_alpha;
        if (rot) {
            float rlat_v = xlat_v[k] * ORC_RAD_PER_DEG, rlon_v = xlon_v[k] * ORC_RAD_PER_DEG;
            float relm = rlon_v - lmbd0;
            float big_denominator = o_cos(o_asin(cos_phi0 * o_sin(rlat_v) - sin_phi0 * o_cos(rlat_v) * o_cos(relm)));
            sin_alpha = sin_phi0 * o_sin(relm) / big_denominator;
            cos_alpha = (cos_phi0 * o_cos(rlat_v) + sin_phi0 * o_sin(rlat_v) * o_cos(relm)) / big_denominator;
        } else {
            float diff = fdir * (xlon_v[k] - proj->stdlon);
            if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
            float alpha;
            if (proj->code == ORC_PROJ_LC) alpha = diff * proj->cone * ORC_RAD_PER_DEG * proj->hemi;
            else alpha = diff * ORC_RAD_PER_DEG * proj->hemi;
            sin_alpha = o_sin(alpha); cos_alpha = o_cos(alpha);
        }
        const int ub = orc_bit_test(u_mask, nxw, i + 1, j + 1), vb = orc_bit_test(v_mask, nxw, i + 1, j + 1);
        if (ub) { float u_map = u[k], v_map = vb ? v[k] : 0.0f; u_new[k] = cos_alpha * u_map + fdir * sin_alpha * v_map; }
        else u_new[k] = u[k];
        if (vb) { float v_map = v[k], u_map = ub ? u[k] : 0.0f; v_new[k] = -fdir * sin_alpha * u_map + cos_alpha * v_map; }
        else v_new[k] = v[k];
    }
    memcpy(u, u_new, sizeof(float) * (size_t)nx * ny);
    memcpy(v, v_new, sizeof(float) * (size_t)nx * ny);
    free(u_new); free(v_new);
}
