// proj.cuh -- map projections of WPS (geogrid/src/module_map_utils.F, llxy_module.F) for device and host.
//
// Arithmetic policy (DESIGN.md "Transcendental policy"): every operation the Fortran does in default
// REAL is done in binary32 with no FMA contraction (-fmad=false); each transcendental intrinsic is
// evaluated in binary64 and rounded once to binary32, which reproduces a correctly rounded libm
// except for ~2^-29 double-rounding cases.  The same source runs on the host for map_set.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/wpsgpu.h"

#define WPS_HD __host__ __device__ __forceinline__

namespace wps {

// constants_module.F:7-25 (default REAL)
#define WPS_PI 3.141592653589793f
#define WPS_DEG_PER_RAD (180.0f / WPS_PI)
#define WPS_RAD_PER_DEG (WPS_PI / 180.0f)
#define WPS_A_WGS84 6378137.0f
#define WPS_E_WGS84 0.081819192f
#define WPS_A_NAD83 6378137.0f
#define WPS_E_NAD83 0.0818187034f
#define WPS_EARTH_RADIUS_M 6370000.0f

WPS_HD float t_sin(float x) { return (float)sin((double)x); }
WPS_HD float t_cos(float x) { return (float)cos((double)x); }
WPS_HD float t_tan(float x) { return (float)tan((double)x); }
WPS_HD float t_asin(float x) { return (float)asin((double)x); }
WPS_HD float t_atan(float x) { return (float)atan((double)x); }
WPS_HD float t_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
WPS_HD float t_log(float x) { return (float)log((double)x); }
WPS_HD float t_log10(float x) { return (float)log10((double)x); }
WPS_HD float t_exp(float x) { return (float)exp((double)x); }
WPS_HD float t_pow(float x, float y) { return (float)pow((double)x, (double)y); }
WPS_HD float t_sqrt(float x) { return sqrtf(x); }  // IEEE correctly rounded (-prec-sqrt=true)
WPS_HD float t_fmod(float x, float y) { return fmodf(x, y); }  // exact
// Fortran NINT / FLOOR / CEILING / INT on default REAL
WPS_HD int f_nint(float x) { return (int)lroundf(x); }
WPS_HD int f_floor(float x) { return (int)floorf(x); }
WPS_HD int f_ceiling(float x) { return (int)ceilf(x); }

// proj_info without the Gaussian latitude table (passed by pointer)
struct DevProj {
    int code, nlat, nxmin, nxmax, comp_ll;
    float lat1, lon1, lat0, lon0, dx, dy, latinc, loninc, dlon, stdlon, truelat1, truelat2, hemi, cone,
          polei, polej, rsw, rebydx, knowni, knownj, re_m, rho0, nc, bigc;
    const float* gauss_lat;
    int stagger, ixdim, jydim;      // PROJ_ROTLL
    float phi, lambda;
};

inline DevProj to_dev(const wpsgpu_proj& p, const float* gauss)
{
    DevProj d;
    d.code = p.code; d.nlat = p.nlat; d.nxmin = p.nxmin; d.nxmax = p.nxmax; d.comp_ll = p.comp_ll;
    d.lat1 = p.lat1; d.lon1 = p.lon1; d.lat0 = p.lat0; d.lon0 = p.lon0; d.dx = p.dx; d.dy = p.dy;
    d.latinc = p.latinc; d.loninc = p.loninc; d.dlon = p.dlon; d.stdlon = p.stdlon; d.truelat1 = p.truelat1;
    d.truelat2 = p.truelat2; d.hemi = p.hemi; d.cone = p.cone; d.polei = p.polei; d.polej = p.polej;
    d.rsw = p.rsw; d.rebydx = p.rebydx; d.knowni = p.knowni; d.knownj = p.knownj; d.re_m = p.re_m;
    d.rho0 = p.rho0; d.nc = p.nc; d.bigc = p.bigc; d.gauss_lat = gauss;
    d.stagger = p.stagger; d.ixdim = p.ixdim; d.jydim = p.jydim; d.phi = p.phi; d.lambda = p.lambda;
    return d;
}

// rotate_coords, module_map_utils.F:1602-1656
WPS_HD void rotate_coords(float ilat, float ilon, float& olat, float& olon, float lat_np, float lon_np,
                          float lon_0, int direction)
{
    float phi_np = lat_np * WPS_RAD_PER_DEG, lam_np = lon_np * WPS_RAD_PER_DEG, lam_0 = lon_0 * WPS_RAD_PER_DEG;
    float rlat = ilat * WPS_RAD_PER_DEG, rlon = ilon * WPS_RAD_PER_DEG;
    float dlam = (direction < 0) ? (WPS_PI - lam_0) : lam_np;
    float sinphi = t_cos(phi_np) * t_cos(rlat) * t_cos(rlon - dlam) + t_sin(phi_np) * t_sin(rlat);
    float cosphi = t_sqrt(1.0f - sinphi * sinphi);
    float coslam = t_sin(phi_np) * t_cos(rlat) * t_cos(rlon - dlam) - t_cos(phi_np) * t_sin(rlat);
    float sinlam = t_cos(rlat) * t_sin(rlon - dlam);
    if (cosphi != 0.0f) { coslam = coslam / cosphi; sinlam = sinlam / cosphi; }
    olat = WPS_DEG_PER_RAD * t_asin(sinphi);
    float ol = WPS_DEG_PER_RAD * (t_atan2(sinlam, coslam) - dlam - lam_0 + lam_np);
    while (!(ol >= -180.0f)) { if (ol != ol) break; ol = ol + 360.0f; }
    while (!(ol <= 180.0f)) { if (ol != ol) break; ol = ol - 360.0f; }
    olon = ol;
}

// lc_cone, module_map_utils.F:1124-1157
WPS_HD float lc_cone(float truelat1, float truelat2)
{
    if (fabsf(truelat1 - truelat2) > 0.1f) {
        float c = t_log10(t_cos(truelat1 * WPS_RAD_PER_DEG)) - t_log10(t_cos(truelat2 * WPS_RAD_PER_DEG));
        c = c / (t_log10(t_tan((45.0f - fabsf(truelat1) / 2.0f) * WPS_RAD_PER_DEG)) -
                 t_log10(t_tan((45.0f - fabsf(truelat2) / 2.0f) * WPS_RAD_PER_DEG)));
        return c;
    }
    return t_sin(fabsf(truelat1) * WPS_RAD_PER_DEG);
}

WPS_HD float wgs84_t(float h, float lat)
{
    float s = t_sin(h * lat * WPS_RAD_PER_DEG);
    return t_sqrt(((1.0f - s) / (1.0f + s)) * t_pow((1.0f + WPS_E_WGS84 * s) / (1.0f - WPS_E_WGS84 * s), WPS_E_WGS84));
}
WPS_HD float wgs84_mc(float h, float truelat1)
{
    float es = WPS_E_WGS84 * t_sin(h * truelat1 * WPS_RAD_PER_DEG);
    return t_cos(h * truelat1 * WPS_RAD_PER_DEG) / t_sqrt(1.0f - es * es);
}
WPS_HD float nad83_q(float sinphi)
{
    float es = WPS_E_NAD83 * sinphi;
    return (1.0f - WPS_E_NAD83 * WPS_E_NAD83) *
           ((sinphi / (1.0f - es * es)) -
            1.0f / (2.0f * WPS_E_NAD83) * t_log((1.0f - WPS_E_NAD83 * sinphi) / (1.0f + WPS_E_NAD83 * sinphi)));
}

// ---- forward transforms (lat,lon) -> (i,j) ----
// llij_latlon, module_map_utils.F:1365-1395
WPS_HD void llij_latlon(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float deltalat = lat - p.lat1, deltalon = lon - p.lon1;
    float i = deltalon / p.loninc, j = deltalat / p.latinc;
    i = i + p.knowni;
    j = j + p.knownj;
    if (i < (float)p.nxmin - 0.5f) i = i + (float)(p.nxmax - p.nxmin + 1);
    if (i >= (float)p.nxmax + 0.5f) i = i - (float)(p.nxmax - p.nxmin + 1);
    io = i; jo = j;
}
// llij_lc, :1236-1290.  Split into its latitude-only and longitude-only terms so that kernels walking a regular
// lat-lon source grid can evaluate each once per row / column (same operations, same order, same bits).
WPS_HD float llij_lc_rm(float lat, const DevProj& p)
{
    float tl1r = p.truelat1 * WPS_RAD_PER_DEG;
    float ctl1r = t_cos(tl1r);
    return p.rebydx * ctl1r / p.cone *
           t_pow(t_tan((90.0f * p.hemi - lat) * WPS_RAD_PER_DEG / 2.0f) /
                     t_tan((90.0f * p.hemi - p.truelat1) * WPS_RAD_PER_DEG / 2.0f),
                 p.cone);
}
WPS_HD void llij_lc_sc(float lon, const DevProj& p, float& s, float& c)
{
    float deltalon = lon - p.stdlon;
    if (deltalon > 180.0f) deltalon = deltalon - 360.0f;
    if (deltalon < -180.0f) deltalon = deltalon + 360.0f;
    float arg = p.cone * (deltalon * WPS_RAD_PER_DEG);
    s = t_sin(arg); c = t_cos(arg);
}
WPS_HD void llij_lc_combine(float rm, float s, float c, const DevProj& p, float& io, float& jo)
{
    float i = p.polei + p.hemi * rm * s;
    float j = p.polej - rm * c;
    io = p.hemi * i; jo = p.hemi * j;
}
WPS_HD void llij_lc(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float s, c;
    float rm = llij_lc_rm(lat, p);
    llij_lc_sc(lon, p, s, c);
    llij_lc_combine(rm, s, c, p, io, jo);
}
// llij_ps, :718-760 (same split)
WPS_HD float llij_ps_rm(float lat, const DevProj& p)
{
    float scale_top = 1.0f + p.hemi * t_sin(p.truelat1 * WPS_RAD_PER_DEG);
    float ala = lat * WPS_RAD_PER_DEG;
    return p.rebydx * t_cos(ala) * scale_top / (1.0f + p.hemi * t_sin(ala));
}
WPS_HD void llij_ps_sc(float lon, const DevProj& p, float& s, float& c)
{
    float reflon = p.stdlon + 90.0f;
    float alo = (lon - reflon) * WPS_RAD_PER_DEG;
    s = t_sin(alo); c = t_cos(alo);
}
WPS_HD void llij_ps_combine(float rm, float s, float c, const DevProj& p, float& i, float& j)
{
    i = p.polei + rm * c;
    j = p.polej + p.hemi * rm * s;
}
WPS_HD void llij_ps(float lat, float lon, const DevProj& p, float& i, float& j)
{
    float s, c;
    float rm = llij_ps_rm(lat, p);
    llij_ps_sc(lon, p, s, c);
    llij_ps_combine(rm, s, c, p, i, j);
}
// llij_merc, :1320-1341
WPS_HD void llij_merc(float lat, float lon, const DevProj& p, float& i, float& j)
{
    float deltalon = lon - p.lon1;
    if (deltalon < -180.0f) deltalon = deltalon + 360.0f;
    if (deltalon > 180.0f) deltalon = deltalon - 360.0f;
    i = p.knowni + (deltalon / (p.dlon * WPS_DEG_PER_RAD));
    j = p.knownj + (t_log(t_tan(0.5f * ((lat + 90.0f) * WPS_RAD_PER_DEG)))) / p.dlon - p.rsw;
}
// llij_cyl, :1443-1475
WPS_HD void llij_cyl(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float deltalat = lat - p.lat1, deltalon = lon - p.lon1;
    if (deltalon < 0.0f) deltalon = deltalon + 360.0f;
    if (deltalon > 360.0f) deltalon = deltalon - 360.0f;
    float i = deltalon / p.loninc, j = deltalat / p.latinc;
    i = i + p.knowni;
    j = j + p.knownj;
    if (i <= 0.0f) i = i + 360.0f / p.loninc;
    if (i > 360.0f / p.loninc) i = i - 360.0f / p.loninc;
    io = i; jo = j;
}
// llij_cassini, :1543-1567
WPS_HD void llij_cassini(float lat, float lon, const DevProj& p, float& i, float& j)
{
    float comp_lat, comp_lon;
    if (fabsf(p.lat0) != 90.0f && !p.comp_ll) {
        rotate_coords(lat, lon, comp_lat, comp_lon, p.lat0, p.lon0, p.stdlon, -1);
        comp_lon = comp_lon + p.stdlon;
    } else { comp_lat = lat; comp_lon = lon; }
    llij_cyl(comp_lat, comp_lon, p, i, j);
}
// llij_ps_wgs84, :857-895
WPS_HD void llij_ps_wgs84(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float h = p.hemi;
    float mc = wgs84_mc(h, p.truelat1), tc = wgs84_t(h, p.truelat1), t = wgs84_t(h, lat);
    float rho = (WPS_A_WGS84 / p.dx) * mc * t / tc;
    float i = h * rho * t_sin((h * lon - h * p.stdlon) * WPS_RAD_PER_DEG);
    float j = h * (-rho) * t_cos((h * lon - h * p.stdlon) * WPS_RAD_PER_DEG);
    io = p.knowni + (i - p.polei);
    jo = p.knownj + (j - p.polej);
}
// llij_albers_nad83, :998-1036
WPS_HD void llij_albers_nad83(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float h = p.hemi;
    float q = nad83_q(t_sin(h * lat * WPS_RAD_PER_DEG));
    float rho = h * (WPS_A_NAD83 / p.dx) * t_sqrt(p.bigc - p.nc * q) / p.nc;
    float theta = p.nc * (h * lon - h * p.stdlon) * WPS_RAD_PER_DEG;
    float i = h * rho * t_sin(theta);
    float j = h * p.rho0 - h * rho * t_cos(theta);
    io = p.knowni + (i - p.polei);
    jo = p.knownj + (j - p.polej);
}
// llij_gauss, :2130-2210
WPS_HD void llij_gauss(float lat, float lon, const DevProj& p, float& io, float& jo)
{
    float i = (lon - p.lon1) / p.loninc + 1.0f, j = 0.0f;
    const float* g = p.gauss_lat;
    int n2 = p.nlat * 2;
    if (fabsf(lat) > fabsf(g[0])) {
        float diff_1 = lat - g[0], diff_nlat = lat - g[n2 - 1];
        j = (fabsf(diff_1) < fabsf(diff_nlat)) ? 1.0f : (float)n2;
    } else {
        int n_low = -1;
        for (int n = 1; n <= n2 - 1; ++n)
            if ((g[n - 1] - lat) * (g[n] - lat) <= 0.0f) { n_low = n; break; }
        if (n_low > 0)
            j = ((g[n_low - 1] - lat) * (float)(n_low + 1) + (lat - g[n_low]) * (float)(n_low)) / (g[n_low - 1] - g[n_low]);
    }
    if (i < (float)p.nxmin - 0.5f) i = i + (float)(p.nxmax - p.nxmin + 1);
    if (i >= (float)p.nxmax + 0.5f) i = i - (float)(p.nxmax - p.nxmin + 1);
    io = i; jo = j;
}

WPS_HD void llij_rotlatlon(float lat, float lon, const DevProj& p, int stagger, float& i_real, float& j_real);
// latlon_to_ij, :570-626
WPS_HD void latlon_to_ij(const DevProj& p, float lat, float lon, float& i, float& j)
{
    switch (p.code) {
    case WPSGPU_PROJ_LATLON: llij_latlon(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_LC: llij_lc(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_PS: llij_ps(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_MERC: llij_merc(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_GAUSS: llij_gauss(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_CYL: llij_cyl(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_CASSINI: llij_cassini(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_PS_WGS84: llij_ps_wgs84(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_ALBERS_NAD83: llij_albers_nad83(lat, lon, p, i, j); break;
    case WPSGPU_PROJ_ROTLL: llij_rotlatlon(lat, lon, p, p.stagger, i, j); break;
    default: i = 0.0f; j = 0.0f; break;
    }
}

// ---- inverse transforms (i,j) -> (lat,lon) ----
// ijll_latlon, :1398-1428
WPS_HD void ijll_latlon(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float i_work = i;
    if (i < (float)p.nxmin - 0.5f) i_work = i + (float)(p.nxmax - p.nxmin + 1);
    if (i >= (float)p.nxmax + 0.5f) i_work = i - (float)(p.nxmax - p.nxmin + 1);
    i_work = i_work - p.knowni;
    float j_work = j - p.knownj;
    float deltalat = j_work * p.latinc, deltalon = i_work * p.loninc;
    lat = p.lat1 + deltalat;
    lon = p.lon1 + deltalon;
}
// ijll_lc, :1160-1233
WPS_HD void ijll_lc(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float chi1 = (90.0f - p.hemi * p.truelat1) * WPS_RAD_PER_DEG;
    float chi2 = (90.0f - p.hemi * p.truelat2) * WPS_RAD_PER_DEG;
    float inew = p.hemi * i, jnew = p.hemi * j;
    float xx = inew - p.polei, yy = p.polej - jnew;
    float r2 = (xx * xx + yy * yy);
    float r = t_sqrt(r2) / p.rebydx;
    float la, lo;
    if (r2 == 0.0f) {
        la = p.hemi * 90.0f;
        lo = p.stdlon;
    } else {
        lo = p.stdlon + WPS_DEG_PER_RAD * t_atan2(p.hemi * xx, yy) / p.cone;
        lo = t_fmod(lo + 360.0f, 360.0f);
        float chi;
        if (chi1 == chi2) chi = 2.0f * t_atan(t_pow(r / t_tan(chi1), 1.0f / p.cone) * t_tan(chi1 * 0.5f));
        else chi = 2.0f * t_atan(t_pow(r * p.cone / t_sin(chi1), 1.0f / p.cone) * t_tan(chi1 * 0.5f));
        la = (90.0f - chi * WPS_DEG_PER_RAD) * p.hemi;
    }
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    lat = la; lon = lo;
}
// ijll_ps, :763-822 (r2, gi2 and the ACOS argument are REAL(KIND=8) in the reference)
WPS_HD void ijll_ps(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float reflon = p.stdlon + 90.0f;
    float scale_top = 1.0f + p.hemi * t_sin(p.truelat1 * WPS_RAD_PER_DEG);
    float xx = i - p.polei, yy = (j - p.polej) * p.hemi;
    double r2 = (double)(xx * xx + yy * yy);
    float la, lo;
    if (r2 == 0.0) {
        la = p.hemi * 90.0f;
        lo = reflon;
    } else {
        float g = p.rebydx * scale_top;
        double gi2 = (double)(g * g);
        la = (float)((double)(WPS_DEG_PER_RAD * p.hemi) * asin((gi2 - r2) / (gi2 + r2)));
        double q = (double)xx / sqrt(r2);
        if (q < -1.0) q = -1.0;
        if (q > 1.0) q = 1.0;
        float arccos = (float)acos(q);
        lo = (yy > 0.0f) ? (reflon + WPS_DEG_PER_RAD * arccos) : (reflon - WPS_DEG_PER_RAD * arccos);
    }
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    lat = la; lon = lo;
}
// ijll_merc, :1344-1362
WPS_HD void ijll_merc(float i, float j, const DevProj& p, float& lat, float& lon)
{
    lat = 2.0f * t_atan(t_exp(p.dlon * (p.rsw + j - p.knownj))) * WPS_DEG_PER_RAD - 90.0f;
    float lo = (i - p.knowni) * p.dlon * WPS_DEG_PER_RAD + p.lon1;
    if (lo > 180.0f) lo = lo - 360.0f;
    if (lo < -180.0f) lo = lo + 360.0f;
    lon = lo;
}
// ijll_cyl, :1478-1509
WPS_HD void ijll_cyl(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float i_work = i - p.knowni, j_work = j - p.knownj;
    if (i_work < 0.0f) i_work = i_work + 360.0f / p.loninc;
    if (i_work >= 360.0f / p.loninc) i_work = i_work - 360.0f / p.loninc;
    float deltalat = j_work * p.latinc, deltalon = i_work * p.loninc;
    lat = deltalat + p.lat1;
    float lo = deltalon + p.lon1;
    if (lo < -180.0f) lo = lo + 360.0f;
    if (lo > 180.0f) lo = lo - 360.0f;
    lon = lo;
}
// ijll_cassini, :1570-1594
WPS_HD void ijll_cassini(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float comp_lat, comp_lon;
    ijll_cyl(i, j, p, comp_lat, comp_lon);
    if (fabsf(p.lat0) != 90.0f && !p.comp_ll) {
        comp_lon = comp_lon - p.stdlon;
        rotate_coords(comp_lat, comp_lon, lat, lon, p.lat0, p.lon0, p.stdlon, 1);
    } else { lat = comp_lat; lon = comp_lon; }
}
// ijll_ps_wgs84, :898-944
WPS_HD void ijll_ps_wgs84(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float h = p.hemi;
    float x = (i - p.knowni + p.polei), y = (j - p.knownj + p.polej);
    float mc = wgs84_mc(h, p.truelat1), tc = wgs84_t(h, p.truelat1);
    float xd = x * p.dx, yd = y * p.dx;
    float rho = t_sqrt(xd * xd + yd * yd);
    float t = rho * tc / (WPS_A_WGS84 * mc);
    float lo = h * p.stdlon * WPS_RAD_PER_DEG + h * t_atan2(h * x, h * (-y));
    float chi = WPS_PI / 2.0f - 2.0f * t_atan(t);
    float e2 = WPS_E_WGS84 * WPS_E_WGS84, e4 = e2 * e2, e6 = e4 * e2, e8 = e4 * e4;
    float a = 1.0f / 2.0f * e2 + 5.0f / 24.0f * e4 + 1.0f / 40.0f * e6 + 73.0f / 2016.0f * e8;
    float b = 7.0f / 24.0f * e4 + 29.0f / 120.0f * e6 + 54113.0f / 40320.0f * e8;
    float c = 7.0f / 30.0f * e6 + 81.0f / 280.0f * e8;
    float d = 4279.0f / 20160.0f * e8;
    float c2 = t_cos(2.0f * chi);
    float la = chi + t_sin(2.0f * chi) * (a + c2 * (b + c2 * (c + d * c2)));
    la = h * la;
    lat = la * WPS_DEG_PER_RAD;
    lon = lo * WPS_DEG_PER_RAD;
}
// ijll_albers_nad83, :1039-1080
WPS_HD void ijll_albers_nad83(float i, float j, const DevProj& p, float& lat, float& lon)
{
    float h = p.hemi;
    float x = (i - p.knowni + p.polei), y = (j - p.knownj + p.polej);
    float ry = p.rho0 - y;
    float rho = t_sqrt(x * x + ry * ry);
    float theta = t_atan2(x, p.rho0 - y);
    float rr = rho * p.nc * p.dx / WPS_A_NAD83;
    float q = (p.bigc - rr * rr) / p.nc;
    float beta = t_asin(q / (1.0f - t_log((1.0f - WPS_E_NAD83) / (1.0f + WPS_E_NAD83)) * (1.0f - WPS_E_NAD83 * WPS_E_NAD83) / (2.0f * WPS_E_NAD83)));
    float e2 = WPS_E_NAD83 * WPS_E_NAD83, e4 = e2 * e2, e6 = e4 * e2;
    float a = 1.0f / 3.0f * e2 + 31.0f / 180.0f * e4 + 517.0f / 5040.0f * e6;
    float b = 23.0f / 360.0f * e4 + 251.0f / 3780.0f * e6;
    float c = 761.0f / 45360.0f * e6;
    float la = beta + a * t_sin(2.0f * beta) + b * t_sin(4.0f * beta) + c * t_sin(6.0f * beta);
    lat = h * la * WPS_DEG_PER_RAD;
    lon = p.stdlon + theta * WPS_DEG_PER_RAD / p.nc;
}

// llij_rotlatlon, :1659-1829 (NMM E-grid).  REAL(KIND=HIGH) = binary64; `pi = ACOS(-1.0)` is the default-REAL arc cosine widened
// (3.1415927410125732), `phi / REAL(..)` a default-real division.  The routine's nearest-mass-point search (:1729-1815) ends in
// the locals i, j, which are never returned: the outputs are the real-valued position set at :1731-1737 / :1773-1779.
#define WPS_PI_SP_AS_DP 3.1415927410125732
WPS_HD void llij_rotlatlon(float lat, float lon, const DevProj& p, int stagger, float& i_real, float& j_real)
{
    const double glatd = lat, glond = -lon;
    const double dphd = (double)(p.phi / (float)((p.jydim - 1) / 2)), dlmd = (double)(p.lambda / (float)(p.ixdim - 1));
    const double pi = WPS_PI_SP_AS_DP, d2r = pi / 180.0, r2d = 1.0 / d2r;
    const int jmt = p.jydim / 2 + 1;
    const double glat = glatd * d2r, glon = glond * d2r;
    const double tph0 = (double)p.lat1 * d2r, tlm0 = -(double)p.lon1 * d2r;
    const double x = cos(tph0) * cos(glat) * cos(glon - tlm0) + sin(tph0) * sin(glat);
    const double y = -cos(glat) * sin(glon - tlm0);
    const double z = cos(tph0) * sin(glat) - sin(tph0) * cos(glat) * cos(glon - tlm0);
    const double tlat = r2d * atan(z / sqrt(x * x + y * y));
    const double tlon = r2d * atan(y / x);
    double row = tlat / dphd + jmt, col = tlon / dlmd + p.ixdim;
    if ((row - (int)row) > (double)0.999f) row = row + (double)0.0002f;
    else if ((col - (int)col) > (double)0.999f) col = col + (double)0.0002f;
    const int nrow = (int)row;
    if (stagger == WPSGPU_HH) {
        i_real = (nrow % 2 == 0) ? (float)(col / 2.0) : (float)(col / 2.0 + 0.5);
        j_real = (float)row;
    } else if (stagger == WPSGPU_VV) {
        i_real = (nrow % 2 == 0) ? (float)(col / 2.0 + 0.5) : (float)(col / 2.0);
        j_real = (float)row;
    } else { i_real = 0.0f; j_real = 0.0f; }      // unset in the reference; map_set is only given HH or VV
}
// ijll_rotlatlon, :1832-1909: dphd, dlmd, col, jj and the (jj - midrow) * dphd products are default REAL here
WPS_HD void ijll_rotlatlon(float i, float j, const DevProj& p, int stagger, float& lat, float& lon)
{
    float jj = j;
    if ((j - (float)(int)j) > 0.999f) jj = j + 0.0002f;
    const int jh = (int)jj;
    const float dphd = p.phi / (float)((p.jydim - 1) / 2), dlmd = p.lambda / (float)(p.ixdim - 1);
    const double pi = WPS_PI_SP_AS_DP, d2r = pi / 180.0, r2d = 1.0 / d2r;
    const double tph0 = (double)p.lat1 * d2r, tlm0 = -(double)p.lon1 * d2r;
    const int midrow = (p.jydim + 1) / 2, midcol = p.ixdim;
    const int par = (jh + 1) % 2;
    const float col = 2.0f * i - 1.0f + (float)(par < 0 ? -par : par);
    const double tlatd = (double)((jj - (float)midrow) * dphd);
    double tlond = (double)((col - (float)midcol) * dlmd);
    if (stagger == WPSGPU_VV) {
        if (jh % 2 == 0) tlond = tlond - (double)dlmd; else tlond = tlond + (double)dlmd;
    }
    const double tlatr = tlatd * d2r, tlonr = tlond * d2r;
    const double arg1 = sin(tlatr) * cos(tph0) + cos(tlatr) * sin(tph0) * cos(tlonr);
    const double glatr = asin(arg1), glatd = glatr * r2d;
    double arg2 = cos(tlatr) * cos(tlonr) / (cos(glatr) * cos(tph0)) - tan(glatr) * tan(tph0);
    if (fabs(arg2) > 1.0) arg2 = fabs(arg2) / arg2;
    double fctr = 1.0;
    if (tlond > 0.0) fctr = -1.0;
    const double glond = tlm0 * r2d + fctr * acos(arg2) * r2d;
    lat = (float)glatd;
    lon = (float)(-glond);
    if (lon > 180.0f) lon = lon - 360.0f;
    if (lon < -180.0f) lon = lon + 360.0f;
}

// ij_to_latlon, :629-679
WPS_HD void ij_to_latlon(const DevProj& p, float i, float j, float& lat, float& lon)
{
    switch (p.code) {
    case WPSGPU_PROJ_LATLON: ijll_latlon(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_LC: ijll_lc(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_PS: ijll_ps(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_MERC: ijll_merc(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_CYL: ijll_cyl(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_CASSINI: ijll_cassini(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_PS_WGS84: ijll_ps_wgs84(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_ALBERS_NAD83: ijll_albers_nad83(i, j, p, lat, lon); break;
    case WPSGPU_PROJ_ROTLL: ijll_rotlatlon(i, j, p, p.stagger, lat, lon); break;
    default: lat = 0.0f; lon = 0.0f; break;
    }
}

// lltoxy / xytoll, llxy_module.F:788-887
WPS_HD void lltoxy(const DevProj& p, float xlat, float xlon, float& x, float& y, int stagger)
{
    if (p.code == WPSGPU_PROJ_ROTLL) {
        // proj_stack(current_nest_number)%stagger = HH / VV (:800-804): the module state the reference mutates, by value here
        llij_rotlatlon(xlat, xlon, p, (stagger == WPSGPU_HH || stagger == WPSGPU_VV) ? stagger : p.stagger, x, y);
    } else
    latlon_to_ij(p, xlat, xlon, x, y);
    if (stagger == WPSGPU_U) x = x + 0.5f;
    else if (stagger == WPSGPU_V) y = y + 0.5f;
    else if (stagger == WPSGPU_CORNER) { x = x + 0.5f; y = y + 0.5f; }
}
WPS_HD void xytoll(const DevProj& p, float x, float y, float& xlat, float& xlon, int stagger)
{
    float rx = x, ry = y;
    if (stagger == WPSGPU_U) rx = x - 0.5f;
    else if (stagger == WPSGPU_V) ry = y - 0.5f;
    else if (stagger == WPSGPU_CORNER) { rx = x - 0.5f; ry = y - 0.5f; }
    if (p.code == WPSGPU_PROJ_ROTLL) {
        // :851-862: HH, VV and CORNER overwrite proj%stagger before ij_to_latlon (CORNER makes ijll_rotlatlon take its H branch)
        const int st = (stagger == WPSGPU_HH || stagger == WPSGPU_VV || stagger == WPSGPU_CORNER) ? stagger : p.stagger;
        ijll_rotlatlon(rx, ry, p, st, xlat, xlon);
        return;
    }
    ij_to_latlon(p, rx, ry, xlat, xlon);
}

}  // namespace wps
