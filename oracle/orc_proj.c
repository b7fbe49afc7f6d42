/*
 * orc_proj.c -- CPU ORACLE (test infrastructure): map projections.
 * Restates geogrid/src/module_map_utils.F and geogrid/src/llxy_module.F of WPS v4.6.0.
 * PARITY UNPINNED (see wps_oracle.h).  Not product code.
 */
#include "wps_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* constants_module.F:7-25 -- default-REAL parameters (binary32) */
static const float PI_F = 3.141592653589793f;
#define DEG_PER_RAD (180.0f / PI_F)
#define RAD_PER_DEG (PI_F / 180.0f)
static const float A_WGS84 = 6378137.0f;
static const float E_WGS84 = 0.081819192f;
static const float A_NAD83 = 6378137.0f;
static const float E_NAD83 = 0.0818187034f;
static const float EARTH_RADIUS_M = 6370000.0f;

static int g_transc_mode = 0;
void orc_set_transc_mode(int mode) { g_transc_mode = mode; }
int orc_get_transc_mode(void) { return g_transc_mode; }

/* Fortran intrinsics on default REAL.  mode 0: glibc float functions (what gfortran calls);
 * mode 1: binary64 evaluation rounded once to binary32. */
/* ACOS(-1.0) on default REAL: both evaluations round to 3.1415927 */
float o_acosf_m1(void) { return g_transc_mode ? (float)acos(-1.0) : acosf(-1.0f); }
float o_sin(float x) { return g_transc_mode ? (float)sin((double)x) : sinf(x); }
float o_cos(float x) { return g_transc_mode ? (float)cos((double)x) : cosf(x); }
float o_tan(float x) { return g_transc_mode ? (float)tan((double)x) : tanf(x); }
float o_asin(float x) { return g_transc_mode ? (float)asin((double)x) : asinf(x); }
float o_atan(float x) { return g_transc_mode ? (float)atan((double)x) : atanf(x); }
float o_atan2(float y, float x) { return g_transc_mode ? (float)atan2((double)y, (double)x) : atan2f(y, x); }
float o_log(float x) { return g_transc_mode ? (float)log((double)x) : logf(x); }
float o_log10(float x) { return g_transc_mode ? (float)log10((double)x) : log10f(x); }
float o_exp(float x) { return g_transc_mode ? (float)exp((double)x) : expf(x); }
float o_pow(float x, float y) { return g_transc_mode ? (float)pow((double)x, (double)y) : powf(x, y); }
float o_sqrt(float x) { return sqrtf(x); } /* correctly rounded either way */

/* module_map_utils.F:198-240 */
void orc_map_init(orc_proj *p)
{
    p->lat1 = -999.9f; p->lon1 = -999.9f; p->lat0 = -999.9f; p->lon0 = -999.9f;
    p->dx = -999.9f; p->dy = -999.9f; p->latinc = -999.9f; p->loninc = -999.9f;
    p->stdlon = -999.9f; p->truelat1 = -999.9f; p->truelat2 = -999.9f;
    p->phi = -999.9f; p->lambda = -999.9f; p->ixdim = -999; p->jydim = -999;
    p->stagger = ORC_HH; p->nlat = 0; p->nlon = 0; p->nxmin = 1; p->nxmax = 43200;
    p->hemi = 0.0f; p->cone = -999.9f; p->polei = -999.9f; p->polej = -999.9f;
    p->rsw = -999.9f; p->knowni = -999.9f; p->knownj = -999.9f; p->re_m = EARTH_RADIUS_M;
    p->init = 0; p->wrap = 0; p->rho0 = 0.0f; p->nc = 0.0f; p->bigc = 0.0f; p->comp_ll = 0;
    p->n_gauss = 0;
    p->dlat = 0.0f; p->dlon = 0.0f; p->rebydx = 0.0f; p->code = -1;
}

/* module_map_utils.F:1124-1157 */
void orc_lc_cone(float truelat1, float truelat2, float *cone)
{
    if (fabsf(truelat1 - truelat2) > 0.1f) {
        float c = o_log10(o_cos(truelat1 * RAD_PER_DEG)) - o_log10(o_cos(truelat2 * RAD_PER_DEG));
        c = c / (o_log10(o_tan((45.0f - fabsf(truelat1) / 2.0f) * RAD_PER_DEG)) -
                 o_log10(o_tan((45.0f - fabsf(truelat2) / 2.0f) * RAD_PER_DEG)));
        *cone = c;
    } else {
        *cone = o_sin(fabsf(truelat1) * RAD_PER_DEG);
    }
}

/* module_map_utils.F:682-715 */
static void set_ps(orc_proj *p)
{
    float reflon = p->stdlon + 90.0f;
    float scale_top = 1.0f + p->hemi * o_sin(p->truelat1 * RAD_PER_DEG);
    float ala1 = p->lat1 * RAD_PER_DEG;
    p->rsw = p->rebydx * o_cos(ala1) * scale_top / (1.0f + p->hemi * o_sin(ala1));
    float alo1 = (p->lon1 - reflon) * RAD_PER_DEG;
    p->polei = p->knowni - p->rsw * o_cos(alo1);
    p->polej = p->knownj - p->hemi * p->rsw * o_sin(alo1);
}

/* helper shared by the WGS84 polar-stereographic routines: the "t" term,
 * module_map_utils.F:842-847,878-882 */
static float wgs84_t(float h, float lat)
{
    float s = o_sin(h * lat * RAD_PER_DEG);
    return o_sqrt(((1.0f - s) / (1.0f + s)) * o_pow((1.0f + E_WGS84 * s) / (1.0f - E_WGS84 * s), E_WGS84));
}
static float wgs84_mc(float h, float truelat1)
{
    float es = E_WGS84 * o_sin(h * truelat1 * RAD_PER_DEG);
    return o_cos(h * truelat1 * RAD_PER_DEG) / o_sqrt(1.0f - es * es);
}

/* module_map_utils.F:825-854 */
static void set_ps_wgs84(orc_proj *p)
{
    float h = p->hemi;
    float mc = wgs84_mc(h, p->truelat1);
    float tc = wgs84_t(h, p->truelat1);
    float t = wgs84_t(h, p->lat1);
    float rho = h * (A_WGS84 / p->dx) * mc * t / tc;
    p->polei = rho * o_sin((h * p->lon1 - h * p->stdlon) * RAD_PER_DEG);
    p->polej = -rho * o_cos((h * p->lon1 - h * p->stdlon) * RAD_PER_DEG);
}

static float nad83_q(float sinphi)
{
    float es = E_NAD83 * sinphi;
    return (1.0f - E_NAD83 * E_NAD83) *
           ((sinphi / (1.0f - es * es)) -
            1.0f / (2.0f * E_NAD83) * o_log((1.0f - E_NAD83 * sinphi) / (1.0f + E_NAD83 * sinphi)));
}

/* module_map_utils.F:947-995 */
static void set_albers_nad83(orc_proj *p)
{
    float h = p->hemi;
    float e1 = E_NAD83 * o_sin(h * p->truelat1 * RAD_PER_DEG);
    float e2 = E_NAD83 * o_sin(h * p->truelat2 * RAD_PER_DEG);
    float m1 = o_cos(h * p->truelat1 * RAD_PER_DEG) / o_sqrt(1.0f - e1 * e1);
    float m2 = o_cos(h * p->truelat2 * RAD_PER_DEG) / o_sqrt(1.0f - e2 * e2);
    float q1 = nad83_q(o_sin(p->truelat1 * RAD_PER_DEG));
    float q2 = nad83_q(o_sin(p->truelat2 * RAD_PER_DEG));
    if (p->truelat1 == p->truelat2) p->nc = o_sin(p->truelat1 * RAD_PER_DEG);
    else p->nc = (m1 * m1 - m2 * m2) / (q2 - q1);
    p->bigc = m1 * m1 + p->nc * q1;
    float q = nad83_q(o_sin(p->lat1 * RAD_PER_DEG));
    p->rho0 = h * (A_NAD83 / p->dx) * o_sqrt(p->bigc - p->nc * q) / p->nc;
    float theta = p->nc * (p->lon1 - p->stdlon) * RAD_PER_DEG;
    p->polei = p->rho0 * o_sin(h * theta);
    p->polej = p->rho0 - p->rho0 * o_cos(h * theta);
}

/* module_map_utils.F:1083-1121 */
static void set_lc(orc_proj *p)
{
    orc_lc_cone(p->truelat1, p->truelat2, &p->cone);
    float deltalon1 = p->lon1 - p->stdlon;
    if (deltalon1 > 180.0f) deltalon1 = deltalon1 - 360.0f;
    if (deltalon1 < -180.0f) deltalon1 = deltalon1 + 360.0f;
    float tl1r = p->truelat1 * RAD_PER_DEG;
    float ctl1r = o_cos(tl1r);
    p->rsw = p->rebydx * ctl1r / p->cone *
             o_pow(o_tan((90.0f * p->hemi - p->lat1) * RAD_PER_DEG / 2.0f) /
                       o_tan((90.0f * p->hemi - p->truelat1) * RAD_PER_DEG / 2.0f),
                   p->cone);
    float arg = p->cone * (deltalon1 * RAD_PER_DEG);
    p->polei = p->hemi * p->knowni - p->hemi * p->rsw * o_sin(arg);
    p->polej = p->hemi * p->knownj + p->rsw * o_cos(arg);
}

/* module_map_utils.F:1293-1317 */
static void set_merc(orc_proj *p)
{
    float clain = o_cos(RAD_PER_DEG * p->truelat1);
    p->dlon = p->dx / (p->re_m * clain);
    p->rsw = 0.0f;
    if (p->lat1 != 0.0f) p->rsw = (o_log(o_tan(0.5f * ((p->lat1 + 90.0f) * RAD_PER_DEG)))) / p->dlon;
}

/* module_map_utils.F:1602-1656 */
void orc_rotate_coords(float ilat, float ilon, float *olat, float *olon, float lat_np, float lon_np,
                       float lon_0, int direction)
{
    float phi_np = lat_np * RAD_PER_DEG, lam_np = lon_np * RAD_PER_DEG, lam_0 = lon_0 * RAD_PER_DEG;
    float rlat = ilat * RAD_PER_DEG, rlon = ilon * RAD_PER_DEG, dlam;
    if (direction < 0) dlam = PI_F - lam_0; else dlam = lam_np;
    float sinphi = o_cos(phi_np) * o_cos(rlat) * o_cos(rlon - dlam) + o_sin(phi_np) * o_sin(rlat);
    float cosphi = o_sqrt(1.0f - sinphi * sinphi);
    float coslam = o_sin(phi_np) * o_cos(rlat) * o_cos(rlon - dlam) - o_cos(phi_np) * o_sin(rlat);
    float sinlam = o_cos(rlat) * o_sin(rlon - dlam);
    if (cosphi != 0.0f) { coslam = coslam / cosphi; sinlam = sinlam / cosphi; }
    *olat = DEG_PER_RAD * o_asin(sinphi);
    float ol = DEG_PER_RAD * (o_atan2(sinlam, coslam) - dlam - lam_0 + lam_np);
    /* DO ... IF (olon >= -180.) EXIT loops of the reference; bounded here so a NaN cannot hang the test harness */
    for (int it = 0; it < 64; ++it) { if (ol >= -180.0f) break; ol = ol + 360.0f; }
    for (int it = 0; it < 64; ++it) { if (ol <= 180.0f) break; ol = ol - 360.0f; }
    *olon = ol;
}

/* module_map_utils.F:1512-1540 */
static void set_cassini(orc_proj *p)
{
    int global_domain;
    p->hemi = 1.0f;
    if (fabsf(p->lat1 - p->latinc / 2.0f + 90.0f) < 0.001f &&
        fabsf(fmodf(p->lon1 - p->loninc / 2.0f - p->stdlon, 360.0f)) < 0.001f) global_domain = 1;
    else global_domain = 0;
    if (fabsf(p->lat0) != 90.0f && !global_domain) {
        float comp_lat, comp_lon;
        orc_rotate_coords(p->lat1, p->lon1, &comp_lat, &comp_lon, p->lat0, p->lon0, p->stdlon, -1);
        comp_lon = comp_lon + p->stdlon;
        p->lat1 = comp_lat;
        p->lon1 = comp_lon;
    }
}

/* ---- Gaussian latitudes: module_map_utils.F:1943-2127 (REAL(KIND=8) throughout) ---- */
static void lgord(double *f, double cosc, int n)
{
    double colat = acos(cosc);
    double c1 = sqrt(2.0);
    for (int k = 1; k <= n; ++k) c1 = c1 * sqrt(1.0 - 1.0 / (double)(4 * k * k));
    double fn = n, ang = fn * colat, s1 = 0.0, c4 = 1.0, a = -1.0, b = 0.0;
    for (int k = 0; k <= n; k += 2) {
        if (k == n) c4 = 0.5 * c4;
        s1 = s1 + c4 * cos(ang);
        a = a + 2.0; b = b + 1.0;
        double fk = k;
        ang = colat * (fn - fk - 2.0);
        c4 = (a * (fn - b + 1.0) / (b * (fn + fn - a))) * c4;
    }
    *f = s1 * c1;
}

static void gausll(int nlat, float *lat_sp)
{
    const double pi = 3.141592653589793;
    double *cosc = (double *)calloc((size_t)nlat, sizeof(double));
    double *sinc = (double *)calloc((size_t)nlat, sizeof(double));
    double *colat = (double *)calloc((size_t)nlat, sizeof(double));
    const float xlim = 1.0e-14f; /* REAL parameter, module_map_utils.F:1996 */
    int nzero = nlat / 2;
    for (int i = 1; i <= nzero; ++i) cosc[i - 1] = sin((i - 0.5) * pi / nlat + pi * 0.5);
    double fi = nlat, fi1 = fi + 1.0;
    double a = fi * fi1 / sqrt(4.0 * fi1 * fi1 - 1.0);
    double b = fi1 * fi / sqrt(4.0 * fi * fi - 1.0);
    for (int i = 1; i <= nzero; ++i) {
        for (;;) {
            double g, gm, gp;
            lgord(&g, cosc[i - 1], nlat);
            lgord(&gm, cosc[i - 1], nlat - 1);
            lgord(&gp, cosc[i - 1], nlat + 1);
            double gt = (cosc[i - 1] * cosc[i - 1] - 1.0) / (a * gp - b * gm);
            double delta = g * gt;
            cosc[i - 1] = cosc[i - 1] - delta;
            if (fabs(delta) > (double)xlim) continue;
            break;
        }
    }
    for (int i = 1; i <= nzero; ++i) { colat[i - 1] = acos(cosc[i - 1]); sinc[i - 1] = sin(colat[i - 1]); }
    if (nlat % 2 != 0) { int i = nzero + 1; cosc[i - 1] = 0.0; colat[i - 1] = pi * 0.5; sinc[i - 1] = 1.0; }
    for (int i = nlat - nzero + 1; i <= nlat; ++i) {
        cosc[i - 1] = -cosc[nlat + 1 - i - 1];
        colat[i - 1] = pi - colat[nlat + 1 - i - 1];
        sinc[i - 1] = sinc[nlat + 1 - i - 1];
    }
    /* gausll, module_map_utils.F:1943-1961: lat = ACOS(sinc)*180/pi, negated for the southern half */
    for (int i = 1; i <= nlat; ++i) {
        double lat = acos(sinc[i - 1]) * 180.0 / pi;
        if (i > nlat / 2) lat = -lat;
        lat_sp[i - 1] = (float)lat;
    }
    free(cosc); free(sinc); free(colat);
}

/* module_map_utils.F:1900-1940 */
static int set_gauss(orc_proj *p)
{
    int n = p->nlat * 2;
    if (n > (int)(sizeof(p->gauss_lat) / sizeof(float))) return -1;
    p->n_gauss = n;
    gausll(n, p->gauss_lat);
    if (fabsf(p->gauss_lat[0] - p->lat1) > 0.01f)
        for (int i = 0; i < n; ++i) p->gauss_lat[i] = -1.0f * p->gauss_lat[i];
    if (fabsf(p->gauss_lat[0] - p->lat1) > 0.01f) return -2;
    return 0;
}

static float wrap180(float lon, int *ok)
{
    /* module_map_utils.F:420-435 */
    float d = lon;
    *ok = 1;
    if (fabsf(d) > 180.0f) {
        int iter = 0;
        while (fabsf(d) > 180.0f && iter < 10) {
            if (d < -180.0f) d = d + 360.0f;
            if (d > 180.0f) d = d - 360.0f;
            iter = iter + 1;
        }
        if (fabsf(d) > 180.0f) *ok = 0;
    }
    return d;
}

/* module_map_utils.F:243-567.  Returns 0, or <0 where the reference would mprintf(ERROR). */
int orc_map_set(int code, orc_proj *p, const orc_map_args *a)
{
    unsigned pr = a->present;
#define HAS(b) ((pr & (b)) != 0)
    unsigned need = 0;
    switch (code) {
    case ORC_PROJ_LC: case ORC_PROJ_ALBERS_NAD83:
        need = ORC_A_TRUELAT1 | ORC_A_TRUELAT2 | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_STDLON | ORC_A_DX; break;
    case ORC_PROJ_PS: case ORC_PROJ_PS_WGS84:
        need = ORC_A_TRUELAT1 | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_STDLON | ORC_A_DX; break;
    case ORC_PROJ_MERC:
        need = ORC_A_TRUELAT1 | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_DX; break;
    case ORC_PROJ_LATLON:
        need = ORC_A_LATINC | ORC_A_LONINC | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_LAT1 | ORC_A_LON1; break;
    case ORC_PROJ_CYL:
        need = ORC_A_LATINC | ORC_A_LONINC | ORC_A_STDLON; break;
    case ORC_PROJ_CASSINI:
        need = ORC_A_LATINC | ORC_A_LONINC | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_LAT0 | ORC_A_LON0 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_STDLON; break;
    case ORC_PROJ_GAUSS:
        need = ORC_A_NLAT | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_LONINC; break;
    case ORC_PROJ_ROTLL:
        need = ORC_A_IXDIM | ORC_A_JYDIM | ORC_A_PHI | ORC_A_LAMBDA | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_STAGGER; break;
    default: return -1;
    }
    if ((pr & need) != need) return -2;
    if (HAS(ORC_A_LAT1) && fabsf(a->lat1) > 90.0f) return -3;
    int ok;
    float dummy_lon1 = 0.0f, dummy_lon0 = 0.0f, dummy_stdlon = 0.0f;
    if (HAS(ORC_A_LON1)) { dummy_lon1 = wrap180(a->lon1, &ok); if (!ok) return -4; }
    if (HAS(ORC_A_LON0)) { dummy_lon0 = wrap180(a->lon0, &ok); if (!ok) return -5; }
    if (HAS(ORC_A_DX) && a->dx <= 0.0f && code != ORC_PROJ_LATLON) return -6;
    if (HAS(ORC_A_STDLON)) {
        dummy_stdlon = a->stdlon;
        if (fabsf(dummy_stdlon) > 180.0f && code != ORC_PROJ_MERC) { dummy_stdlon = wrap180(a->stdlon, &ok); if (!ok) return -7; }
    }
    if (HAS(ORC_A_TRUELAT1) && fabsf(a->truelat1) > 90.0f) return -8;

    orc_map_init(p);
    p->code = code;
    if (HAS(ORC_A_LAT1)) p->lat1 = a->lat1;
    if (HAS(ORC_A_LON1)) p->lon1 = dummy_lon1;
    if (HAS(ORC_A_LAT0)) p->lat0 = a->lat0;
    if (HAS(ORC_A_LON0)) p->lon0 = dummy_lon0;
    if (HAS(ORC_A_LATINC)) p->latinc = a->latinc;
    if (HAS(ORC_A_LONINC)) p->loninc = a->loninc;
    if (HAS(ORC_A_KNOWNI)) p->knowni = a->knowni;
    if (HAS(ORC_A_KNOWNJ)) p->knownj = a->knownj;
    if (HAS(ORC_A_NXMIN)) p->nxmin = a->nxmin;
    if (HAS(ORC_A_NXMAX)) p->nxmax = a->nxmax;
    if (HAS(ORC_A_DX)) p->dx = a->dx;
    if (HAS(ORC_A_DY)) p->dy = a->dy; else if (HAS(ORC_A_DX)) p->dy = a->dx;
    if (HAS(ORC_A_STDLON)) p->stdlon = dummy_stdlon;
    if (HAS(ORC_A_TRUELAT1)) p->truelat1 = a->truelat1;
    if (HAS(ORC_A_TRUELAT2)) p->truelat2 = a->truelat2;
    if (HAS(ORC_A_NLAT)) p->nlat = a->nlat;
    if (HAS(ORC_A_NLON)) p->nlon = a->nlon;
    if (HAS(ORC_A_IXDIM)) p->ixdim = a->ixdim;
    if (HAS(ORC_A_JYDIM)) p->jydim = a->jydim;
    if (HAS(ORC_A_STAGGER)) p->stagger = a->stagger;
    if (HAS(ORC_A_PHI)) p->phi = a->phi;
    if (HAS(ORC_A_LAMBDA)) p->lambda = a->lambda;
    if (HAS(ORC_A_R_EARTH)) p->re_m = a->r_earth;
    if (HAS(ORC_A_DX)) {
        if (code == ORC_PROJ_LC || code == ORC_PROJ_PS || code == ORC_PROJ_PS_WGS84 ||
            code == ORC_PROJ_ALBERS_NAD83 || code == ORC_PROJ_MERC) {
            p->dx = a->dx;
            if (a->truelat1 < 0.0f) p->hemi = -1.0f; else p->hemi = 1.0f;
            p->rebydx = p->re_m / a->dx;
        }
    }
    int rc = 0;
    switch (code) {
    case ORC_PROJ_PS: set_ps(p); break;
    case ORC_PROJ_PS_WGS84: set_ps_wgs84(p); break;
    case ORC_PROJ_ALBERS_NAD83: set_albers_nad83(p); break;
    case ORC_PROJ_LC:
        if (fabsf(p->truelat2) > 90.0f) p->truelat2 = p->truelat1;
        set_lc(p); break;
    case ORC_PROJ_MERC: set_merc(p); break;
    case ORC_PROJ_LATLON: break;
    case ORC_PROJ_GAUSS: rc = set_gauss(p); break;
    case ORC_PROJ_CYL: p->hemi = 1.0f; break; /* set_cyl :1431-1440 */
    case ORC_PROJ_CASSINI: set_cassini(p); break;
    default: break;
    }
    p->init = 1;
    return rc;
#undef HAS
}

/* ---------- forward / inverse pairs ---------- */

/* module_map_utils.F:718-760 */
static void llij_ps(float lat, float lon, const orc_proj *p, float *i, float *j)
{
    float reflon = p->stdlon + 90.0f;
    float scale_top = 1.0f + p->hemi * o_sin(p->truelat1 * RAD_PER_DEG);
    float ala = lat * RAD_PER_DEG;
    float rm = p->rebydx * o_cos(ala) * scale_top / (1.0f + p->hemi * o_sin(ala));
    float alo = (lon - reflon) * RAD_PER_DEG;
    *i = p->polei + rm * o_cos(alo);
    *j = p->polej + p->hemi * rm * o_sin(alo);
}

/* module_map_utils.F:763-822 (r2, gi2 and the arccos argument are REAL(KIND=8)) */
static void ijll_ps(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float reflon = p->stdlon + 90.0f;
    float scale_top = 1.0f + p->hemi * o_sin(p->truelat1 * RAD_PER_DEG);
    float xx = i - p->polei;
    float yy = (j - p->polej) * p->hemi;
    double r2 = (double)(xx * xx + yy * yy); /* xx**2+yy**2 is a default-REAL expression */
    float la, lo;
    if (r2 == 0.0) {
        la = p->hemi * 90.0f;
        lo = reflon;
    } else {
        float g = p->rebydx * scale_top;
        double gi2 = (double)(g * g);
        /* lat = deg_per_rad*hemi*ASIN((gi2-r2)/(gi2+r2)): the ASIN argument is REAL(8), so ASIN is
         * evaluated in double, the product in double, then stored to REAL(4) */
        la = (float)((double)(DEG_PER_RAD * p->hemi) * asin((gi2 - r2) / (gi2 + r2)));
        double q = (double)xx / sqrt(r2);
        if (q < -1.0) q = -1.0;
        if (q > 1.0) q = 1.0;
        float arccos = (float)acos(q);
        if (yy > 0.0f) lo = reflon + DEG_PER_RAD * arccos; else lo = reflon - DEG_PER_RAD * arccos;
    }
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    *lat = la; *lon = lo;
}

/* module_map_utils.F:857-895 */
static void llij_ps_wgs84(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float h = p->hemi;
    float mc = wgs84_mc(h, p->truelat1), tc = wgs84_t(h, p->truelat1), t = wgs84_t(h, lat);
    float rho = (A_WGS84 / p->dx) * mc * t / tc;
    float i = h * rho * o_sin((h * lon - h * p->stdlon) * RAD_PER_DEG);
    float j = h * (-rho) * o_cos((h * lon - h * p->stdlon) * RAD_PER_DEG);
    *io = p->knowni + (i - p->polei);
    *jo = p->knownj + (j - p->polej);
}

/* module_map_utils.F:898-944 */
static void ijll_ps_wgs84(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float h = p->hemi;
    float x = (i - p->knowni + p->polei), y = (j - p->knownj + p->polej);
    float mc = wgs84_mc(h, p->truelat1), tc = wgs84_t(h, p->truelat1);
    float xd = x * p->dx, yd = y * p->dx;
    float rho = o_sqrt(xd * xd + yd * yd);
    float t = rho * tc / (A_WGS84 * mc);
    float lo = h * p->stdlon * RAD_PER_DEG + h * o_atan2(h * x, h * (-y));
    float chi = PI_F / 2.0f - 2.0f * o_atan(t);
    float e2 = E_WGS84 * E_WGS84, e4 = e2 * e2, e6 = e4 * e2, e8 = e4 * e4;
    float a = 1.0f / 2.0f * e2 + 5.0f / 24.0f * e4 + 1.0f / 40.0f * e6 + 73.0f / 2016.0f * e8;
    float b = 7.0f / 24.0f * e4 + 29.0f / 120.0f * e6 + 54113.0f / 40320.0f * e8;
    float c = 7.0f / 30.0f * e6 + 81.0f / 280.0f * e8;
    float d = 4279.0f / 20160.0f * e8;
    float c2 = o_cos(2.0f * chi);
    float la = chi + o_sin(2.0f * chi) * (a + c2 * (b + c2 * (c + d * c2)));
    la = h * la;
    *lat = la * DEG_PER_RAD;
    *lon = lo * DEG_PER_RAD;
}

/* module_map_utils.F:998-1036 */
static void llij_albers_nad83(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float h = p->hemi;
    float q = nad83_q(o_sin(h * lat * RAD_PER_DEG));
    float rho = h * (A_NAD83 / p->dx) * o_sqrt(p->bigc - p->nc * q) / p->nc;
    float theta = p->nc * (h * lon - h * p->stdlon) * RAD_PER_DEG;
    float i = h * rho * o_sin(theta);
    float j = h * p->rho0 - h * rho * o_cos(theta);
    *io = p->knowni + (i - p->polei);
    *jo = p->knownj + (j - p->polej);
}

/* module_map_utils.F:1039-1080 */
static void ijll_albers_nad83(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float h = p->hemi;
    float x = (i - p->knowni + p->polei), y = (j - p->knownj + p->polej);
    float ry = p->rho0 - y;
    float rho = o_sqrt(x * x + ry * ry);
    float theta = o_atan2(x, p->rho0 - y);
    float rr = rho * p->nc * p->dx / A_NAD83;
    float q = (p->bigc - rr * rr) / p->nc;
    float beta = o_asin(q / (1.0f - o_log((1.0f - E_NAD83) / (1.0f + E_NAD83)) * (1.0f - E_NAD83 * E_NAD83) / (2.0f * E_NAD83)));
    float e2 = E_NAD83 * E_NAD83, e4 = e2 * e2, e6 = e4 * e2;
    float a = 1.0f / 3.0f * e2 + 31.0f / 180.0f * e4 + 517.0f / 5040.0f * e6;
    float b = 23.0f / 360.0f * e4 + 251.0f / 3780.0f * e6;
    float c = 761.0f / 45360.0f * e6;
    float la = beta + a * o_sin(2.0f * beta) + b * o_sin(4.0f * beta) + c * o_sin(6.0f * beta);
    *lat = h * la * DEG_PER_RAD;
    *lon = p->stdlon + theta * DEG_PER_RAD / p->nc;
}

/* module_map_utils.F:1160-1233 */
static void ijll_lc(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float chi1 = (90.0f - p->hemi * p->truelat1) * RAD_PER_DEG;
    float chi2 = (90.0f - p->hemi * p->truelat2) * RAD_PER_DEG;
    float inew = p->hemi * i, jnew = p->hemi * j;
    float xx = inew - p->polei, yy = p->polej - jnew;
    float r2 = (xx * xx + yy * yy);
    float r = o_sqrt(r2) / p->rebydx;
    float la, lo;
    if (r2 == 0.0f) {
        la = p->hemi * 90.0f;
        lo = p->stdlon;
    } else {
        lo = p->stdlon + DEG_PER_RAD * o_atan2(p->hemi * xx, yy) / p->cone;
        lo = fmodf(lo + 360.0f, 360.0f);
        float chi;
        if (chi1 == chi2) chi = 2.0f * o_atan(o_pow(r / o_tan(chi1), 1.0f / p->cone) * o_tan(chi1 * 0.5f));
        else chi = 2.0f * o_atan(o_pow(r * p->cone / o_sin(chi1), 1.0f / p->cone) * o_tan(chi1 * 0.5f));
        la = (90.0f - chi * DEG_PER_RAD) * p->hemi;
    }
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    *lat = la; *lon = lo;
}

/* module_map_utils.F:1236-1290 */
static void llij_lc(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float deltalon = lon - p->stdlon;
    if (deltalon > 180.0f) deltalon = deltalon - 360.0f;
    if (deltalon < -180.0f) deltalon = deltalon + 360.0f;
    float tl1r = p->truelat1 * RAD_PER_DEG;
    float ctl1r = o_cos(tl1r);
    float rm = p->rebydx * ctl1r / p->cone *
               o_pow(o_tan((90.0f * p->hemi - lat) * RAD_PER_DEG / 2.0f) /
                         o_tan((90.0f * p->hemi - p->truelat1) * RAD_PER_DEG / 2.0f),
                     p->cone);
    float arg = p->cone * (deltalon * RAD_PER_DEG);
    float i = p->polei + p->hemi * rm * o_sin(arg);
    float j = p->polej - rm * o_cos(arg);
    *io = p->hemi * i;
    *jo = p->hemi * j;
}

/* module_map_utils.F:1320-1341 */
static void llij_merc(float lat, float lon, const orc_proj *p, float *i, float *j)
{
    float deltalon = lon - p->lon1;
    if (deltalon < -180.0f) deltalon = deltalon + 360.0f;
    if (deltalon > 180.0f) deltalon = deltalon - 360.0f;
    *i = p->knowni + (deltalon / (p->dlon * DEG_PER_RAD));
    *j = p->knownj + (o_log(o_tan(0.5f * ((lat + 90.0f) * RAD_PER_DEG)))) / p->dlon - p->rsw;
}

/* module_map_utils.F:1344-1362 */
static void ijll_merc(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    *lat = 2.0f * o_atan(o_exp(p->dlon * (p->rsw + j - p->knownj))) * DEG_PER_RAD - 90.0f;
    float lo = (i - p->knowni) * p->dlon * DEG_PER_RAD + p->lon1;
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    *lon = lo;
}

/* module_map_utils.F:1365-1395 */
static void llij_latlon(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float deltalat = lat - p->lat1;
    float deltalon = lon - p->lon1;
    float i = deltalon / p->loninc;
    float j = deltalat / p->latinc;
    i = i + p->knowni;
    j = j + p->knownj;
    if (i < (float)p->nxmin - 0.5f) i = i + (float)(p->nxmax - p->nxmin + 1);
    if (i >= (float)p->nxmax + 0.5f) i = i - (float)(p->nxmax - p->nxmin + 1);
    *io = i; *jo = j;
}

/* module_map_utils.F:1398-1428 */
static void ijll_latlon(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float i_work = i;
    if (i < (float)p->nxmin - 0.5f) i_work = i + (float)(p->nxmax - p->nxmin + 1);
    if (i >= (float)p->nxmax + 0.5f) i_work = i - (float)(p->nxmax - p->nxmin + 1);
    i_work = i_work - p->knowni;
    float j_work = j - p->knownj;
    float deltalat = j_work * p->latinc;
    float deltalon = i_work * p->loninc;
    *lat = p->lat1 + deltalat;
    *lon = p->lon1 + deltalon;
}

/* module_map_utils.F:1443-1475 */
static void llij_cyl(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float deltalat = lat - p->lat1;
    float deltalon = lon - p->lon1;
    if (deltalon < 0.0f) deltalon = deltalon + 360.0f;
    if (deltalon > 360.0f) deltalon = deltalon - 360.0f;
    float i = deltalon / p->loninc;
    float j = deltalat / p->latinc;
    i = i + p->knowni;
    j = j + p->knownj;
    if (i <= 0.0f) i = i + 360.0f / p->loninc;
    if (i > 360.0f / p->loninc) i = i - 360.0f / p->loninc;
    *io = i; *jo = j;
}

/* module_map_utils.F:1478-1509 */
static void ijll_cyl(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float i_work = i - p->knowni, j_work = j - p->knownj;
    if (i_work < 0.0f) i_work = i_work + 360.0f / p->loninc;
    if (i_work >= 360.0f / p->loninc) i_work = i_work - 360.0f / p->loninc;
    float deltalat = j_work * p->latinc, deltalon = i_work * p->loninc;
    *lat = deltalat + p->lat1;
    float lo = deltalon + p->lon1;
    if (lo < -180.0f) lo = lo + 360.0f;
    if (lo > 180.0f) lo = lo - 360.0f;
    *lon = lo;
}

/* module_map_utils.F:1543-1567 */
static void llij_cassini(float lat, float lon, const orc_proj *p, float *i, float *j)
{
    float comp_lat, comp_lon;
    if (fabsf(p->lat0) != 90.0f && !p->comp_ll) {
        orc_rotate_coords(lat, lon, &comp_lat, &comp_lon, p->lat0, p->lon0, p->stdlon, -1);
        comp_lon = comp_lon + p->stdlon;
    } else { comp_lat = lat; comp_lon = lon; }
    llij_cyl(comp_lat, comp_lon, p, i, j);
}

/* module_map_utils.F:1570-1594 */
static void ijll_cassini(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float comp_lat, comp_lon;
    ijll_cyl(i, j, p, &comp_lat, &comp_lon);
    if (fabsf(p->lat0) != 90.0f && !p->comp_ll) {
        comp_lon = comp_lon - p->stdlon;
        orc_rotate_coords(comp_lat, comp_lon, lat, lon, p->lat0, p->lon0, p->stdlon, 1);
    } else { *lat = comp_lat; *lon = comp_lon; }
}

/* module_map_utils.F:2130-2210 */
static void llij_gauss(float lat, float lon, const orc_proj *p, float *io, float *jo)
{
    float i = (lon - p->lon1) / p->loninc + 1.0f, j = 0.0f;
    const float *g = p->gauss_lat; /* 1-based in Fortran */
    int n2 = p->nlat * 2;
    if (fabsf(lat) > fabsf(g[0])) {
        float diff_1 = lat - g[0], diff_nlat = lat - g[n2 - 1];
        if (fabsf(diff_1) < fabsf(diff_nlat)) j = 1.0f; else j = (float)n2;
    } else {
        int n_low = -1;
        for (int n = 1; n <= n2 - 1; ++n)
            if ((g[n - 1] - lat) * (g[n] - lat) <= 0.0f) { n_low = n; break; }
        if (n_low > 0)
            j = ((g[n_low - 1] - lat) * (float)(n_low + 1) + (lat - g[n_low]) * (float)(n_low)) /
                (g[n_low - 1] - g[n_low]);
    }
    if (i < (float)p->nxmin - 0.5f) i = i + (float)(p->nxmax - p->nxmin + 1);
    if (i >= (float)p->nxmax + 0.5f) i = i - (float)(p->nxmax - p->nxmin + 1);
    *io = i; *jo = j;
}

/* llij_rotlatlon, module_map_utils.F:1659-1829.  REAL(KIND=HIGH) = binary64; `pi = ACOS(-1.0)` is the DEFAULT-REAL arc cosine
 * widened, `phi / REAL(..)` a default-real division.  The routine's nearest-mass-point search (:1729-1815) ends in the locals
 * i, j, which are never returned: the outputs are the real-valued position set at :1731-1737 / :1773-1779. */
static double d_sin(double x) { return sin(x); }
static void llij_rotlatlon(float lat, float lon, const orc_proj *p, float *i_real, float *j_real)
{
    double glatd = lat, glond = -lon;
    double dphd = (double)(p->phi / (float)((p->jydim - 1) / 2)), dlmd = (double)(p->lambda / (float)(p->ixdim - 1));
    double pi = (double)o_acosf_m1(), d2r = pi / 180.0, r2d = 1.0 / d2r;
    int jmt = p->jydim / 2 + 1;
    double glat = glatd * d2r, glon = glond * d2r;
    double tph0 = (double)p->lat1 * d2r, tlm0 = -(double)p->lon1 * d2r;
    double x = cos(tph0) * cos(glat) * cos(glon - tlm0) + d_sin(tph0) * d_sin(glat);
    double y = -cos(glat) * d_sin(glon - tlm0);
    double z = cos(tph0) * d_sin(glat) - d_sin(tph0) * cos(glat) * cos(glon - tlm0);
    double tlat = r2d * atan(z / sqrt(x * x + y * y));
    double tlon = r2d * atan(y / x);
    double row = tlat / dphd + jmt, col = tlon / dlmd + p->ixdim;
    if ((row - (int)row) > (double)0.999f) row = row + (double)0.0002f;
    else if ((col - (int)col) > (double)0.999f) col = col + (double)0.0002f;
    int nrow = (int)row;
    if (p->stagger == ORC_HH) {
        *i_real = (nrow % 2 == 0) ? (float)(col / 2.0) : (float)(col / 2.0 + 0.5);
        *j_real = (float)row;
    } else if (p->stagger == ORC_VV) {
        *i_real = (nrow % 2 == 0) ? (float)(col / 2.0 + 0.5) : (float)(col / 2.0);
        *j_real = (float)row;
    } else { *i_real = 0.0f; *j_real = 0.0f; }      /* the reference leaves i_real, j_real unset; map_set is only given HH or VV */
}
/* ijll_rotlatlon, :1832-1909: dphd, dlmd, col, jj and the (jj - midrow) * dphd products are default REAL here */
static void ijll_rotlatlon(float i, float j, const orc_proj *p, float *lat, float *lon)
{
    float jj = j;
    if ((j - (float)(int)j) > 0.999f) jj = j + 0.0002f;
    int jh = (int)jj;
    float dphd = p->phi / (float)((p->jydim - 1) / 2), dlmd = p->lambda / (float)(p->ixdim - 1);
    double pi = (double)o_acosf_m1(), d2r = pi / 180.0, r2d = 1.0 / d2r;
    double tph0 = (double)p->lat1 * d2r, tlm0 = -(double)p->lon1 * d2r;
    int midrow = (p->jydim + 1) / 2, midcol = p->ixdim;
    float col = 2.0f * i - 1.0f + (float)abs((jh + 1) % 2);
    double tlatd = (double)((jj - (float)midrow) * dphd), tlond = (double)((col - (float)midcol) * dlmd);
    if (p->stagger == ORC_VV) {
        if (jh % 2 == 0) tlond = tlond - (double)dlmd; else tlond = tlond + (double)dlmd;
    }
    double tlatr = tlatd * d2r, tlonr = tlond * d2r;
    double arg1 = d_sin(tlatr) * cos(tph0) + cos(tlatr) * d_sin(tph0) * cos(tlonr);
    double glatr = asin(arg1), glatd = glatr * r2d;
    double arg2 = cos(tlatr) * cos(tlonr) / (cos(glatr) * cos(tph0)) - tan(glatr) * tan(tph0);
    if (fabs(arg2) > 1.0) arg2 = fabs(arg2) / arg2;
    double fctr = 1.0;
    if (tlond > 0.0) fctr = -1.0;
    double glond = tlm0 * r2d + fctr * acos(arg2) * r2d;
    *lat = (float)glatd;
    *lon = (float)(-glond);
    if (*lon > 180.0f) *lon = *lon - 360.0f;
    if (*lon < -180.0f) *lon = *lon + 360.0f;
}

/* module_map_utils.F:570-626 */
void orc_latlon_to_ij(const orc_proj *p, float lat, float lon, float *i, float *j)
{
    switch (p->code) {
    case ORC_PROJ_LATLON: llij_latlon(lat, lon, p, i, j); break;
    case ORC_PROJ_MERC: llij_merc(lat, lon, p, i, j); break;
    case ORC_PROJ_PS: llij_ps(lat, lon, p, i, j); break;
    case ORC_PROJ_PS_WGS84: llij_ps_wgs84(lat, lon, p, i, j); break;
    case ORC_PROJ_ALBERS_NAD83: llij_albers_nad83(lat, lon, p, i, j); break;
    case ORC_PROJ_LC: llij_lc(lat, lon, p, i, j); break;
    case ORC_PROJ_GAUSS: llij_gauss(lat, lon, p, i, j); break;
    case ORC_PROJ_CYL: llij_cyl(lat, lon, p, i, j); break;
    case ORC_PROJ_CASSINI: llij_cassini(lat, lon, p, i, j); break;
    case ORC_PROJ_ROTLL: llij_rotlatlon(lat, lon, p, i, j); break;
    default: *i = 0.0f; *j = 0.0f; break;
    }
}

/* module_map_utils.F:629-679 */
void orc_ij_to_latlon(const orc_proj *p, float i, float j, float *lat, float *lon)
{
    switch (p->code) {
    case ORC_PROJ_LATLON: ijll_latlon(i, j, p, lat, lon); break;
    case ORC_PROJ_MERC: ijll_merc(i, j, p, lat, lon); break;
    case ORC_PROJ_PS: ijll_ps(i, j, p, lat, lon); break;
    case ORC_PROJ_PS_WGS84: ijll_ps_wgs84(i, j, p, lat, lon); break;
    case ORC_PROJ_ALBERS_NAD83: ijll_albers_nad83(i, j, p, lat, lon); break;
    case ORC_PROJ_LC: ijll_lc(i, j, p, lat, lon); break;
    case ORC_PROJ_CYL: ijll_cyl(i, j, p, lat, lon); break;
    case ORC_PROJ_CASSINI: ijll_cassini(i, j, p, lat, lon); break;
    case ORC_PROJ_ROTLL: ijll_rotlatlon(i, j, p, lat, lon); break;
    default: *lat = 0.0f; *lon = 0.0f; break;
    }
}

/* llxy_module.F:788-829 */
void orc_lltoxy(orc_proj *p, float xlat, float xlon, float *x, float *y, int stagger)
{
    if (stagger == ORC_HH) p->stagger = ORC_HH; else if (stagger == ORC_VV) p->stagger = ORC_VV;
    orc_latlon_to_ij(p, xlat, xlon, x, y);
    if (stagger == ORC_U) *x = *x + 0.5f;
    else if (stagger == ORC_V) *y = *y + 0.5f;
    else if (stagger == ORC_CORNER) { *x = *x + 0.5f; *y = *y + 0.5f; }
}

/* llxy_module.F:837-887 */
void orc_xytoll(orc_proj *p, float x, float y, float *xlat, float *xlon, int stagger)
{
    float rx, ry;
    if (stagger == ORC_U) { rx = x - 0.5f; ry = y; }
    else if (stagger == ORC_V) { rx = x; ry = y - 0.5f; }
    else if (stagger == ORC_HH) { p->stagger = ORC_HH; rx = x; ry = y; }
    else if (stagger == ORC_VV) { p->stagger = ORC_VV; rx = x; ry = y; }
    else if (stagger == ORC_CORNER) { p->stagger = ORC_CORNER; rx = x - 0.5f; ry = y - 0.5f; }
    else { rx = x; ry = y; }
    orc_ij_to_latlon(p, rx, ry, xlat, xlon);
}

void orc_lltoxy_array(orc_proj *p, long n, const float *lat, const float *lon, float *x, float *y, int stagger)
{ for (long k = 0; k < n; ++k) orc_lltoxy(p, lat[k], lon[k], &x[k], &y[k], stagger); }
void orc_xytoll_array(orc_proj *p, long n, const float *x, const float *y, float *lat, float *lon, int stagger)
{ for (long k = 0; k < n; ++k) orc_xytoll(p, x[k], y[k], &lat[k], &lon[k], stagger); }

/* llxy_module.F:39-165.  has_cass: the optional Cassini arguments are present; has_re: earth_radius present. */
int orc_push_source_projection(orc_proj *p, int iproj, float stand_lon, float truelat1, float truelat2,
                               float dxkm, float dykm, float dlat, float dlon, float known_x, float known_y,
                               float known_lat, float known_lon, int has_cass, float pole_lat, float pole_lon,
                               float centerlat, float centerlon, float centeri, float centerj,
                               int has_re, float earth_radius)
{
    orc_map_args a;
    memset(&a, 0, sizeof(a));
    (void)dykm;
    unsigned re = has_re ? ORC_A_R_EARTH : 0;
    a.r_earth = earth_radius;
    orc_map_init(p);
    switch (iproj) {
    case ORC_PROJ_LATLON:
        a.present = ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_NXMAX | ORC_A_LATINC | ORC_A_LONINC | re;
        a.lat1 = known_lat; a.lon1 = known_lon; a.knowni = known_x; a.knownj = known_y;
        a.nxmax = (int)lroundf(360.0f / dlon); a.latinc = dlat; a.loninc = dlon;
        break;
    case ORC_PROJ_MERC:
        a.present = ORC_A_TRUELAT1 | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_DX | re;
        a.truelat1 = truelat1; a.lat1 = known_lat; a.lon1 = known_lon; a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case ORC_PROJ_CASSINI:
        if (!has_cass) return -20;
        a.present = ORC_A_LATINC | ORC_A_LONINC | ORC_A_STDLON | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_LAT0 | ORC_A_LON0 | ORC_A_KNOWNI | ORC_A_KNOWNJ | re;
        a.latinc = dlat; a.loninc = dlon; a.stdlon = stand_lon; a.lat1 = centerlat; a.lon1 = centerlon;
        a.lat0 = pole_lat; a.lon0 = pole_lon; a.knowni = centeri; a.knownj = centerj;
        break;
    case ORC_PROJ_LC: case ORC_PROJ_ALBERS_NAD83:
        a.present = ORC_A_TRUELAT1 | ORC_A_TRUELAT2 | ORC_A_STDLON | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_DX | re;
        a.truelat1 = truelat1; a.truelat2 = truelat2; a.stdlon = stand_lon; a.lat1 = known_lat; a.lon1 = known_lon;
        a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case ORC_PROJ_PS: case ORC_PROJ_PS_WGS84:
        a.present = ORC_A_TRUELAT1 | ORC_A_STDLON | ORC_A_LAT1 | ORC_A_LON1 | ORC_A_KNOWNI | ORC_A_KNOWNJ | ORC_A_DX | re;
        a.truelat1 = truelat1; a.stdlon = stand_lon; a.lat1 = known_lat; a.lon1 = known_lon;
        a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case ORC_PROJ_GAUSS:
        a.present = ORC_A_LAT1 | ORC_A_LON1 | ORC_A_NXMAX | ORC_A_NLAT | ORC_A_LONINC | re;
        a.lat1 = known_lat; a.lon1 = known_lon; a.nxmax = (int)lroundf(360.0f / dlon);
        a.nlat = (int)lroundf(dlat); a.loninc = dlon;
        break;
    default:
        return -21;
    }
    return orc_map_set(iproj, p, &a);
}


/* find_known_latlon (llxy_module.F:550-593), recursive exactly like the reference: returns the coarse-domain lat/lon of
 * point (rx, ry) of domain n (1-based) and the grid distances divided by the ratios on the way back up. */
static void find_known_latlon(const orc_proj *coarse, const orc_map_args *ca, int n, float rx, float ry, const int *parent_id,
                              const int *ratio, const int *ips, const int *jps, float *rlat, float *rlon, float *dx, float *dy,
                              float *dlat, float *dlon)
{
    if (n == 1) {
        *dx = (ca->present & (1u << 6)) ? ca->dx : 0.0f;
        *dy = (ca->present & (1u << 7)) ? ca->dy : 0.0f;
        *dlat = (ca->present & (1u << 8)) ? ca->latinc : 0.0f;
        *dlon = (ca->present & (1u << 9)) ? ca->loninc : 0.0f;
        orc_ij_to_latlon(coarse, rx, ry, rlat, rlon);
        return;
    }
    float r = (float)ratio[n - 1];
    float xp = (rx - ((r + 1.0f) / 2.0f)) / r + (float)ips[n - 1];
    float yp = (ry - ((r + 1.0f) / 2.0f)) / r + (float)jps[n - 1];
    find_known_latlon(coarse, ca, parent_id[n - 1], xp, yp, parent_id, ratio, ips, jps, rlat, rlon, dx, dy, dlat, dlon);
    *dx = *dx / r; *dy = *dy / r; *dlat = *dlat / r; *dlon = *dlon / r;
}

/* compute_nest_locations (llxy_module.F:331-531) for the projections with a dx (LC, PS, PS_WGS84, MERC, ALBERS) */
int orc_compute_nest_locations(int iproj, const orc_map_args *coarse, int n_domains, const int *parent_id, const int *ratio,
                               const int *ips, const int *jps, const int *ixdim, const int *jydim, orc_proj *out)
{
    orc_map_init(&out[0]);
    int rc = orc_map_set(iproj, &out[0], coarse);
    if (rc) return rc;
    for (int i = 2; i <= n_domains; ++i) {
        float kx = (float)ixdim[i - 1] / 2.0f, ky = (float)jydim[i - 1] / 2.0f, klat, klon, dx, dy, dlat, dlon;
        find_known_latlon(&out[0], coarse, i, kx, ky, parent_id, ratio, ips, jps, &klat, &klon, &dx, &dy, &dlat, &dlon);
        orc_map_args a = *coarse;
        a.present |= (1u << 0) | (1u << 1);
        a.lat1 = klat; a.lon1 = klon;
        a.present |= (1u << 4) | (1u << 5) | (1u << 6); a.knowni = kx; a.knownj = ky; a.dx = dx;
        (void)dy; (void)dlat; (void)dlon;
        orc_map_init(&out[i - 1]);
        rc = orc_map_set(iproj, &out[i - 1], &a);
        if (rc) return rc;
    }
    return 0;
}
