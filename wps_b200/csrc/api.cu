// api.cu -- handle lifetime, errors, interp_option parsing, map_set and the point-wise lltoxy/xytoll entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include "ctx.h"

namespace wps {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int Arena::reserve(size_t bytes, cudaStream_t st)
{
    if (bytes <= cap) return 0;
    WPS_CUDA(cudaStreamSynchronize(st));
    if (base) WPS_CUDA(cudaFree(base));
    base = nullptr; cap = 0; off = 0;
    size_t want = bytes + (bytes >> 3) + (1 << 20);
    WPS_CUDA(cudaMalloc((void**)&base, want));
    cap = want;
    return 0;
}
void Arena::release()
{
    if (base) cudaFree(base);
    base = nullptr; cap = 0; off = 0;
}

int stage_in(wpsgpu_ctx* h, const void* p, size_t bytes, void** dptr)
{
    if (!p) { *dptr = nullptr; return 0; }
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) { *dptr = const_cast<void*>(p); return 0; }
    void* d = h->arena.take(bytes);
    WPS_CHECK(d != nullptr, -3, "device arena exhausted (%zu bytes requested)", bytes);
    WPS_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, h->stream));
    *dptr = d;
    return 0;
}
int alloc_out(wpsgpu_ctx* h, void* user_ptr, size_t bytes, void** dptr)
{
    if (!user_ptr) { *dptr = nullptr; return 0; }
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) { *dptr = user_ptr; return 0; }
    void* d = h->arena.take(bytes);
    WPS_CHECK(d != nullptr, -3, "device arena exhausted (%zu bytes requested)", bytes);
    *dptr = d;
    return 0;
}
int copy_out(wpsgpu_ctx* h, void* user_ptr, const void* dptr, size_t bytes)
{
    if (!user_ptr || h->ptr_mode == WPSGPU_PTR_DEVICE) return 0;
    WPS_CUDA(cudaMemcpyAsync(user_ptr, dptr, bytes, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}
int arena_begin(wpsgpu_ctx* h)
{
    if (h->async_pending) {
        WPS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_out_done, 0));
        h->async_pending = false;
    }
    return 0;
}

int upload_gauss(wpsgpu_ctx* h, const wpsgpu_proj& p, const float** d_gauss)
{
    *d_gauss = nullptr;
    if (p.code != WPSGPU_PROJ_GAUSS) return 0;
    void* d = h->arena.take(sizeof(float) * WPSGPU_MAX_GAUSS);
    WPS_CHECK(d != nullptr, -3, "device arena exhausted (gauss table)");
    WPS_CUDA(cudaMemcpyAsync(d, p.gauss_lat, sizeof(float) * WPSGPU_MAX_GAUSS, cudaMemcpyHostToDevice, h->stream));
    *d_gauss = (const float*)d;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// map_set and its set_* helpers (module_map_utils.F:243-567, 682-715, 825-854, 947-995, 1083-1121,
// 1293-1317, 1512-1540, 1900-2127), evaluated on the host with the same arithmetic policy as the device.
static void h_map_init(wpsgpu_proj& p)
{
    memset(&p, 0, sizeof(p));
    p.lat1 = -999.9f; p.lon1 = -999.9f; p.lat0 = -999.9f; p.lon0 = -999.9f; p.dx = -999.9f; p.dy = -999.9f;
    p.latinc = -999.9f; p.loninc = -999.9f; p.stdlon = -999.9f; p.truelat1 = -999.9f; p.truelat2 = -999.9f;
    p.stagger = WPSGPU_HH; p.nlat = 0; p.nxmin = 1; p.nxmax = 43200; p.hemi = 0.0f; p.cone = -999.9f;
    p.polei = -999.9f; p.polej = -999.9f; p.rsw = -999.9f; p.knowni = -999.9f; p.knownj = -999.9f;
    p.re_m = WPS_EARTH_RADIUS_M; p.init = 0; p.rho0 = 0.0f; p.nc = 0.0f; p.bigc = 0.0f; p.comp_ll = 0;
    p.code = -1;
    p.ixdim = -999; p.jydim = -999; p.phi = -999.9f; p.lambda = -999.9f;    // map_init :215-218
}

static void h_lgord(double* f, double cosc, int n)
{
    double colat = acos(cosc), c1 = sqrt(2.0);
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
static void h_gausll(int nlat, float* lat_sp)
{
    const double pi = 3.141592653589793;
    std::vector<double> cosc(nlat, 0.0), sinc(nlat, 0.0), colat(nlat, 0.0);
    const float xlim = 1.0e-14f;
    int nzero = nlat / 2;
    for (int i = 1; i <= nzero; ++i) cosc[i - 1] = sin((i - 0.5) * pi / nlat + pi * 0.5);
    double fi = nlat, fi1 = fi + 1.0;
    double a = fi * fi1 / sqrt(4.0 * fi1 * fi1 - 1.0), b = fi1 * fi / sqrt(4.0 * fi * fi - 1.0);
    for (int i = 1; i <= nzero; ++i) {
        for (int it = 0; it < 1000; ++it) {
            double g, gm, gp;
            h_lgord(&g, cosc[i - 1], nlat);
            h_lgord(&gm, cosc[i - 1], nlat - 1);
            h_lgord(&gp, cosc[i - 1], nlat + 1);
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
    for (int i = 1; i <= nlat; ++i) {
        double lat = acos(sinc[i - 1]) * 180.0 / pi;
        if (i > nlat / 2) lat = -lat;
        lat_sp[i - 1] = (float)lat;
    }
}

static float h_wrap180(float lon, bool& ok)
{
    float d = lon;
    ok = true;
    if (fabsf(d) > 180.0f) {
        int iter = 0;
        while (fabsf(d) > 180.0f && iter < 10) {
            if (d < -180.0f) d = d + 360.0f;
            if (d > 180.0f) d = d - 360.0f;
            ++iter;
        }
        if (fabsf(d) > 180.0f) ok = false;
    }
    return d;
}

static int h_map_set(int code, const wpsgpu_map_args& a, wpsgpu_proj& p)
{
    uint32_t pr = a.present;
    auto has = [&](uint32_t b) { return (pr & b) != 0; };
    uint32_t need = 0;
    switch (code) {
    case WPSGPU_PROJ_LC: case WPSGPU_PROJ_ALBERS_NAD83:
        need = WPSGPU_A_TRUELAT1 | WPSGPU_A_TRUELAT2 | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_STDLON | WPSGPU_A_DX; break;
    case WPSGPU_PROJ_PS: case WPSGPU_PROJ_PS_WGS84:
        need = WPSGPU_A_TRUELAT1 | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_STDLON | WPSGPU_A_DX; break;
    case WPSGPU_PROJ_MERC:
        need = WPSGPU_A_TRUELAT1 | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX; break;
    case WPSGPU_PROJ_LATLON:
        need = WPSGPU_A_LATINC | WPSGPU_A_LONINC | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_LAT1 | WPSGPU_A_LON1; break;
    case WPSGPU_PROJ_CYL:
        need = WPSGPU_A_LATINC | WPSGPU_A_LONINC | WPSGPU_A_STDLON; break;
    case WPSGPU_PROJ_CASSINI:
        need = WPSGPU_A_LATINC | WPSGPU_A_LONINC | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_LAT0 | WPSGPU_A_LON0 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_STDLON; break;
    case WPSGPU_PROJ_GAUSS:
        need = WPSGPU_A_NLAT | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_LONINC; break;
    case WPSGPU_PROJ_ROTLL:
        need = WPSGPU_A_IXDIM | WPSGPU_A_JYDIM | WPSGPU_A_PHI | WPSGPU_A_LAMBDA | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_STAGGER; break;
    default:
        WPS_CHECK(false, -10, "map_set: Unknown projection code: %d", code);
    }
    WPS_CHECK((pr & need) == need, -11, "map_set: mandatory parameters missing for projection code %d", code);
    if (has(WPSGPU_A_LAT1)) WPS_CHECK(fabsf(a.lat1) <= 90.0f, -12, "map_set: -90N <= lat1 <= 90N required");
    bool ok;
    float dummy_lon1 = 0.0f, dummy_lon0 = 0.0f, dummy_stdlon = 0.0f;
    if (has(WPSGPU_A_LON1)) { dummy_lon1 = h_wrap180(a.lon1, ok); WPS_CHECK(ok, -13, "map_set: bad lon1"); }
    if (has(WPSGPU_A_LON0)) { dummy_lon0 = h_wrap180(a.lon0, ok); WPS_CHECK(ok, -14, "map_set: bad lon0"); }
    if (has(WPSGPU_A_DX)) WPS_CHECK(!(a.dx <= 0.0f && code != WPSGPU_PROJ_LATLON), -15, "map_set: dx must be positive");
    if (has(WPSGPU_A_STDLON)) {
        dummy_stdlon = a.stdlon;
        if (fabsf(dummy_stdlon) > 180.0f && code != WPSGPU_PROJ_MERC) { dummy_stdlon = h_wrap180(a.stdlon, ok); WPS_CHECK(ok, -16, "map_set: bad stdlon"); }
    }
    if (has(WPSGPU_A_TRUELAT1)) WPS_CHECK(fabsf(a.truelat1) <= 90.0f, -17, "map_set: bad truelat1");

    h_map_init(p);
    p.code = code;
    if (has(WPSGPU_A_LAT1)) p.lat1 = a.lat1;
    if (has(WPSGPU_A_LON1)) p.lon1 = dummy_lon1;
    if (has(WPSGPU_A_LAT0)) p.lat0 = a.lat0;
    if (has(WPSGPU_A_LON0)) p.lon0 = dummy_lon0;
    if (has(WPSGPU_A_LATINC)) p.latinc = a.latinc;
    if (has(WPSGPU_A_LONINC)) p.loninc = a.loninc;
    if (has(WPSGPU_A_KNOWNI)) p.knowni = a.knowni;
    if (has(WPSGPU_A_KNOWNJ)) p.knownj = a.knownj;
    if (has(WPSGPU_A_NXMIN)) p.nxmin = a.nxmin;
    if (has(WPSGPU_A_NXMAX)) p.nxmax = a.nxmax;
    if (has(WPSGPU_A_DX)) p.dx = a.dx;
    if (has(WPSGPU_A_DY)) p.dy = a.dy; else if (has(WPSGPU_A_DX)) p.dy = a.dx;
    if (has(WPSGPU_A_STDLON)) p.stdlon = dummy_stdlon;
    if (has(WPSGPU_A_TRUELAT1)) p.truelat1 = a.truelat1;
    if (has(WPSGPU_A_TRUELAT2)) p.truelat2 = a.truelat2;
    if (has(WPSGPU_A_NLAT)) p.nlat = a.nlat;
    if (has(WPSGPU_A_STAGGER)) p.stagger = a.stagger;
    if (has(WPSGPU_A_IXDIM)) p.ixdim = a.ixdim;
    if (has(WPSGPU_A_JYDIM)) p.jydim = a.jydim;
    if (has(WPSGPU_A_PHI)) p.phi = a.phi;
    if (has(WPSGPU_A_LAMBDA)) p.lambda = a.lambda;
    if (has(WPSGPU_A_R_EARTH)) p.re_m = a.r_earth;
    if (has(WPSGPU_A_DX)) {
        if (code == WPSGPU_PROJ_LC || code == WPSGPU_PROJ_PS || code == WPSGPU_PROJ_PS_WGS84 ||
            code == WPSGPU_PROJ_ALBERS_NAD83 || code == WPSGPU_PROJ_MERC) {
            p.dx = a.dx;
            p.hemi = (a.truelat1 < 0.0f) ? -1.0f : 1.0f;
            p.rebydx = p.re_m / a.dx;
        }
    }
    const float R = WPS_RAD_PER_DEG;
    switch (code) {
    case WPSGPU_PROJ_PS: {  // set_ps
        float reflon = p.stdlon + 90.0f;
        float scale_top = 1.0f + p.hemi * t_sin(p.truelat1 * R);
        float ala1 = p.lat1 * R;
        p.rsw = p.rebydx * t_cos(ala1) * scale_top / (1.0f + p.hemi * t_sin(ala1));
        float alo1 = (p.lon1 - reflon) * R;
        p.polei = p.knowni - p.rsw * t_cos(alo1);
        p.polej = p.knownj - p.hemi * p.rsw * t_sin(alo1);
        break;
    }
    case WPSGPU_PROJ_PS_WGS84: {  // set_ps_wgs84
        float h = p.hemi;
        float mc = wgs84_mc(h, p.truelat1), tc = wgs84_t(h, p.truelat1), t = wgs84_t(h, p.lat1);
        float rho = h * (WPS_A_WGS84 / p.dx) * mc * t / tc;
        p.polei = rho * t_sin((h * p.lon1 - h * p.stdlon) * R);
        p.polej = -rho * t_cos((h * p.lon1 - h * p.stdlon) * R);
        break;
    }
    case WPSGPU_PROJ_ALBERS_NAD83: {  // set_albers_nad83
        float h = p.hemi;
        float e1 = WPS_E_NAD83 * t_sin(h * p.truelat1 * R), e2 = WPS_E_NAD83 * t_sin(h * p.truelat2 * R);
        float m1 = t_cos(h * p.truelat1 * R) / t_sqrt(1.0f - e1 * e1);
        float m2 = t_cos(h * p.truelat2 * R) / t_sqrt(1.0f - e2 * e2);
        float q1 = nad83_q(t_sin(p.truelat1 * R)), q2 = nad83_q(t_sin(p.truelat2 * R));
        if (p.truelat1 == p.truelat2) p.nc = t_sin(p.truelat1 * R); else p.nc = (m1 * m1 - m2 * m2) / (q2 - q1);
        p.bigc = m1 * m1 + p.nc * q1;
        float q = nad83_q(t_sin(p.lat1 * R));
        p.rho0 = h * (WPS_A_NAD83 / p.dx) * t_sqrt(p.bigc - p.nc * q) / p.nc;
        float theta = p.nc * (p.lon1 - p.stdlon) * R;
        p.polei = p.rho0 * t_sin(h * theta);
        p.polej = p.rho0 - p.rho0 * t_cos(h * theta);
        break;
    }
    case WPSGPU_PROJ_LC: {  // set_lc
        if (fabsf(p.truelat2) > 90.0f) p.truelat2 = p.truelat1;
        p.cone = lc_cone(p.truelat1, p.truelat2);
        float deltalon1 = p.lon1 - p.stdlon;
        if (deltalon1 > 180.0f) deltalon1 = deltalon1 - 360.0f;
        if (deltalon1 < -180.0f) deltalon1 = deltalon1 + 360.0f;
        float tl1r = p.truelat1 * R;
        float ctl1r = t_cos(tl1r);
        p.rsw = p.rebydx * ctl1r / p.cone *
                t_pow(t_tan((90.0f * p.hemi - p.lat1) * R / 2.0f) / t_tan((90.0f * p.hemi - p.truelat1) * R / 2.0f), p.cone);
        float arg = p.cone * (deltalon1 * R);
        p.polei = p.hemi * p.knowni - p.hemi * p.rsw * t_sin(arg);
        p.polej = p.hemi * p.knownj + p.rsw * t_cos(arg);
        break;
    }
    case WPSGPU_PROJ_MERC: {  // set_merc
        float clain = t_cos(R * p.truelat1);
        p.dlon = p.dx / (p.re_m * clain);
        p.rsw = 0.0f;
        if (p.lat1 != 0.0f) p.rsw = (t_log(t_tan(0.5f * ((p.lat1 + 90.0f) * R)))) / p.dlon;
        break;
    }
    case WPSGPU_PROJ_GAUSS: {  // set_gauss
        int n = p.nlat * 2;
        WPS_CHECK(n > 0 && n <= WPSGPU_MAX_GAUSS, -18, "map_set: Gaussian grid with %d latitudes exceeds WPSGPU_MAX_GAUSS", n);
        p.n_gauss = n;
        h_gausll(n, p.gauss_lat);
        if (fabsf(p.gauss_lat[0] - p.lat1) > 0.01f) for (int i = 0; i < n; ++i) p.gauss_lat[i] = -1.0f * p.gauss_lat[i];
        WPS_CHECK(!(fabsf(p.gauss_lat[0] - p.lat1) > 0.01f), -19, "map_set: Gaussian latitude computation does not match lat1");
        break;
    }
    case WPSGPU_PROJ_CYL: p.hemi = 1.0f; break;
    case WPSGPU_PROJ_CASSINI: {  // set_cassini
        p.hemi = 1.0f;
        bool global_domain = fabsf(p.lat1 - p.latinc / 2.0f + 90.0f) < 0.001f &&
                             fabsf(t_fmod(p.lon1 - p.loninc / 2.0f - p.stdlon, 360.0f)) < 0.001f;
        if (fabsf(p.lat0) != 90.0f && !global_domain) {
            float comp_lat, comp_lon;
            rotate_coords(p.lat1, p.lon1, comp_lat, comp_lon, p.lat0, p.lon0, p.stdlon, -1);
            comp_lon = comp_lon + p.stdlon;
            p.lat1 = comp_lat;
            p.lon1 = comp_lon;
        }
        break;
    }
    default: break;
    }
    p.init = 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// interp_array_from_string / interp_options_from_string / get_search_depth (interp_module.F:20-295)
static bool tok_is(const std::string& tok, const char* name)
{
    size_t len = tok.size();
    while (len > 0 && tok[len - 1] == ' ') --len;
    return len == strlen(name) && tok.compare(0, len, name) == 0;
}
static int search_depth(const std::string& tok, int& depth)
{
    depth = 1200;
    size_t i = tok.find("search");
    if (i == std::string::npos) return 0;
    size_t p = tok.find('+', i);
    std::string sub = tok.substr(i, p == std::string::npos ? std::string::npos : p - i + 1);
    size_t p1 = sub.find('('), p2 = sub.find(')');
    if (p1 != std::string::npos && p2 != std::string::npos) {
        // the reference uses p1,p2 (relative to the keyword) as absolute positions in the token
        if (p2 <= p1 + 1 || p2 > tok.size()) return -1;
        std::string num = tok.substr(p1 + 1, p2 - p1 - 1);
        char* end;
        long v = strtol(num.c_str(), &end, 10);
        while (*end == ' ') ++end;
        if (end == num.c_str() || *end != 0) return -1;
        depth = (int)v;
    } else if (p1 == std::string::npos && p2 == std::string::npos) {
    } else depth = 1200;
    return 0;
}

}  // namespace wps

using namespace wps;

extern "C" {

const char* wpsgpu_last_error(void) { return g_err; }

int wpsgpu_create(int device_id, wpsgpu_handle_t* out)
{
    WPS_CHECK(out != nullptr, -1, "wpsgpu_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    WPS_CHECK(e == cudaSuccess && n > 0, -2, "wpsgpu_create: no CUDA device available (%s); there is no CPU fallback",
              cudaGetErrorString(e));
    WPS_CHECK(device_id >= 0 && device_id < n, -2, "wpsgpu_create: device %d out of range (0..%d)", device_id, n - 1);
    WPS_CUDA(cudaSetDevice(device_id));
    wpsgpu_ctx* h = new wpsgpu_ctx();
    h->device = device_id;
    WPS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    WPS_CUDA(cudaGetDeviceProperties(&prop, device_id));
    h->sm_count = prop.multiProcessorCount;
    WPS_CUDA(cudaMalloc((void**)&h->d_counters, 64 * sizeof(int)));
    WPS_CUDA(cudaMemsetAsync(h->d_counters, 0, 64 * sizeof(int), h->stream));
    for (int k = 0; k < 12; ++k) WPS_CUDA(cudaEventCreate(&h->ev_t[k]));
    h->ev_k0 = h->ev_t[4]; h->ev_k1 = h->ev_t[5];
    for (int k = 0; k < 8; ++k) WPS_CUDA(cudaEventRecord(h->ev_t[k], h->stream));
    *out = h;
    return 0;
}

void wpsgpu_geo_free(wpsgpu_ctx* h);  // geogrid.cu
void wpsgpu_mpas_free(wpsgpu_ctx* h); // mpas.cu

int wpsgpu_destroy(wpsgpu_handle_t h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto& kv : h->coord_cache) wps::free_coord_entry(kv.second);
    for (auto& g : h->grids)
        if (g.own) { cudaFree(g.d_xlat); cudaFree(g.d_xlon); if (g.d_landmask) cudaFree(g.d_landmask); }
    wpsgpu_geo_free(h);
    wpsgpu_mpas_free(h);
    h->arena.release();
    if (h->d_counters) cudaFree(h->d_counters);
    if (h->d_search_scratch) cudaFree(h->d_search_scratch);
    for (int k = 0; k < 12; ++k) if (h->ev_t[k]) cudaEventDestroy(h->ev_t[k]);
    if (h->s_side) cudaStreamDestroy(h->s_side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->ev_out_done) cudaEventDestroy(h->ev_out_done);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int wpsgpu_set_pointer_mode(wpsgpu_handle_t h, int mode)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    WPS_CHECK(mode == WPSGPU_PTR_HOST || mode == WPSGPU_PTR_DEVICE, -1, "bad pointer mode %d", mode);
    h->ptr_mode = mode;
    return 0;
}

int wpsgpu_synchronize(wpsgpu_handle_t h)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    WPS_CUDA(cudaSetDevice(h->device));
    if (h->s_in) WPS_CUDA(cudaStreamSynchronize(h->s_in));
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    if (h->s_side) WPS_CUDA(cudaStreamSynchronize(h->s_side));
    if (h->s_out) WPS_CUDA(cudaStreamSynchronize(h->s_out));
    h->async_pending = false;
    return 0;
}

void* wpsgpu_stream(wpsgpu_handle_t h) { return h ? (void*)h->stream : nullptr; }
int64_t wpsgpu_launch_count(wpsgpu_handle_t h) { return h ? h->launches : 0; }

int wpsgpu_map_set(int proj_code, const wpsgpu_map_args* args, wpsgpu_proj* out)
{
    WPS_CHECK(args && out, -1, "wpsgpu_map_set: NULL argument");
    return h_map_set(proj_code, *args, *out);
}

// compute_nest_locations + find_known_latlon, llxy_module.F:331-593.  Host arithmetic only (binary32, the reference's order).
int wpsgpu_compute_nest_locations(int iproj, const wpsgpu_map_args* coarse, int n_domains, const int32_t* parent_id,
                                  const int32_t* parent_grid_ratio, const int32_t* i_parent_start, const int32_t* j_parent_start,
                                  const int32_t* ixdim, const int32_t* jydim, wpsgpu_proj* out)
{
    WPS_CHECK(coarse && parent_id && parent_grid_ratio && i_parent_start && j_parent_start && ixdim && jydim && out, -1,
              "wpsgpu_compute_nest_locations: NULL argument");
    WPS_CHECK(n_domains >= 1 && n_domains <= 21, -1, "n_domains %d out of range 1..21 (MAX_DOMAINS)", n_domains);
    WPS_CHECK(iproj == WPSGPU_PROJ_LC || iproj == WPSGPU_PROJ_PS || iproj == WPSGPU_PROJ_PS_WGS84 || iproj == WPSGPU_PROJ_MERC ||
              iproj == WPSGPU_PROJ_ALBERS_NAD83, -1,
              "compute_nest_locations: projection %d is not supported for nesting here", iproj);
    WPS_TRY(h_map_set(iproj, *coarse, out[0]));          // proj_stack(1), :343-438
    const wps::DevProj d1 = wps::to_dev(out[0], out[0].gauss_lat);
    for (int n = 2; n <= n_domains; ++n) {
        WPS_CHECK(parent_id[n - 1] >= 1 && parent_id[n - 1] < n && parent_grid_ratio[n - 1] >= 1, -1,
                  "domain %d: parent_id %d / parent_grid_ratio %d invalid", n, parent_id[n - 1], parent_grid_ratio[n - 1]);
        // find_known_latlon (:550-593) for the nest's centre point, walked up to the coarse domain
        float rx = (float)ixdim[n - 1] / 2.0f, ry = (float)jydim[n - 1] / 2.0f;
        const float known_x = rx, known_y = ry;
        float dx = (coarse->present & WPSGPU_A_DX) ? coarse->dx : 0.0f, dy = (coarse->present & WPSGPU_A_DY) ? coarse->dy : 0.0f;
        float dlat = (coarse->present & WPSGPU_A_LATINC) ? coarse->latinc : 0.0f, dlon = (coarse->present & WPSGPU_A_LONINC) ? coarse->loninc : 0.0f;
        // the recursion divides on the way back up, innermost ratio last: dx = ((dx / r_outer) ... / r_n)
        int chain[32], nc = 0;
        for (int k = n; k != 1; k = parent_id[k - 1]) {
            const float r = (float)parent_grid_ratio[k - 1];
            rx = (rx - ((r + 1.0f) / 2.0f)) / r + (float)i_parent_start[k - 1];
            ry = (ry - ((r + 1.0f) / 2.0f)) / r + (float)j_parent_start[k - 1];
            chain[nc++] = k;
        }
        for (int c = nc - 1; c >= 0; --c) {
            const float r = (float)parent_grid_ratio[chain[c] - 1];
            dx = dx / r; dy = dy / r; dlat = dlat / r; dlon = dlon / r;
        }
        float rlat, rlon;
        wps::ij_to_latlon(d1, rx, ry, rlat, rlon);
        wpsgpu_map_args a = *coarse;                     // truelat1/2, stdlon, ... as for the coarse domain (:451-532)
        a.present |= WPSGPU_A_LAT1 | WPSGPU_A_LON1;
        a.lat1 = rlat; a.lon1 = rlon;
        a.present |= WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX;
        a.knowni = known_x; a.knownj = known_y; a.dx = dx;
        (void)dy; (void)dlat; (void)dlon;
        WPS_TRY(h_map_set(iproj, a, out[n - 1]));
    }
    return 0;
}

int wpsgpu_push_source_projection(int iproj, float stand_lon, float truelat1, float truelat2, float dxkm,
                                  float dykm, float dlat, float dlon, float known_x, float known_y,
                                  float known_lat, float known_lon, const float* cassini,
                                  const float* earth_radius, wpsgpu_proj* out)
{
    WPS_CHECK(out != nullptr, -1, "wpsgpu_push_source_projection: out is NULL");
    wpsgpu_map_args a;
    memset(&a, 0, sizeof(a));
    (void)dykm;
    uint32_t re = earth_radius ? WPSGPU_A_R_EARTH : 0;
    if (earth_radius) a.r_earth = *earth_radius;
    switch (iproj) {
    case WPSGPU_PROJ_LATLON:
        a.present = WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_NXMAX | WPSGPU_A_LATINC | WPSGPU_A_LONINC | re;
        a.lat1 = known_lat; a.lon1 = known_lon; a.knowni = known_x; a.knownj = known_y;
        a.nxmax = f_nint(360.0f / dlon); a.latinc = dlat; a.loninc = dlon;
        break;
    case WPSGPU_PROJ_MERC:
        a.present = WPSGPU_A_TRUELAT1 | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX | re;
        a.truelat1 = truelat1; a.lat1 = known_lat; a.lon1 = known_lon; a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case WPSGPU_PROJ_CASSINI:
        WPS_CHECK(cassini != nullptr, -1, "push_source_projection: Cassini needs pole/center arguments");
        a.present = WPSGPU_A_LATINC | WPSGPU_A_LONINC | WPSGPU_A_STDLON | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_LAT0 | WPSGPU_A_LON0 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | re;
        a.latinc = dlat; a.loninc = dlon; a.stdlon = stand_lon; a.lat0 = cassini[0]; a.lon0 = cassini[1];
        a.lat1 = cassini[2]; a.lon1 = cassini[3]; a.knowni = cassini[4]; a.knownj = cassini[5];
        break;
    case WPSGPU_PROJ_LC: case WPSGPU_PROJ_ALBERS_NAD83:
        a.present = WPSGPU_A_TRUELAT1 | WPSGPU_A_TRUELAT2 | WPSGPU_A_STDLON | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX | re;
        a.truelat1 = truelat1; a.truelat2 = truelat2; a.stdlon = stand_lon; a.lat1 = known_lat; a.lon1 = known_lon;
        a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case WPSGPU_PROJ_PS: case WPSGPU_PROJ_PS_WGS84:
        a.present = WPSGPU_A_TRUELAT1 | WPSGPU_A_STDLON | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX | re;
        a.truelat1 = truelat1; a.stdlon = stand_lon; a.lat1 = known_lat; a.lon1 = known_lon;
        a.knowni = known_x; a.knownj = known_y; a.dx = dxkm;
        break;
    case WPSGPU_PROJ_GAUSS:
        a.present = WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_NXMAX | WPSGPU_A_NLAT | WPSGPU_A_LONINC | re;
        a.lat1 = known_lat; a.lon1 = known_lon; a.nxmax = f_nint(360.0f / dlon); a.nlat = f_nint(dlat); a.loninc = dlon;
        break;
    default:
        WPS_CHECK(false, -10, "push_source_projection: projection %d not supported as a source projection", iproj);
    }
    return h_map_set(iproj, a, *out);
}

int wpsgpu_parse_interp_option(const char* interp_string, int32_t* codes, int32_t* opts, int maxn, int* n,
                               int* has_gcell, float* gcell_threshold)
{
    WPS_CHECK(interp_string && codes && opts && n, -1, "wpsgpu_parse_interp_option: NULL argument");
    std::string s(interp_string);
    while (!s.empty() && s.back() == ' ') s.pop_back();
    int iend = (int)s.size();
    int num_methods = 1;
    for (char c : s) if (c == '+') ++num_methods;
    WPS_CHECK(num_methods + 1 <= maxn, -20, "interp_option has %d methods; at most %d supported", num_methods, maxn - 1);
    for (int k = 0; k < num_methods + 1; ++k) { codes[k] = 0; opts[k] = 0; }
    int j = 0;
    // tokens separated by '+'; the reference only parses the trailing token when it is longer than one
    // character (`if (p1 < iend)`, interp_module.F:89)
    size_t p1 = 0;
    for (;;) {
        size_t p2 = s.find('+', p1);
        bool last = (p2 == std::string::npos);
        if (last && !((int)p1 + 1 < iend)) break;
        std::string tok = s.substr(p1, last ? std::string::npos : p2 - p1);
        int code = 0, depth = 0;
        if (tok_is(tok, "nearest_neighbor")) code = WPSGPU_N_NEIGHBOR;
        else if (tok_is(tok, "average_4pt")) code = WPSGPU_AVERAGE4;
        else if (tok_is(tok, "average_16pt")) code = WPSGPU_AVERAGE16;
        else if (tok_is(tok, "wt_average_4pt")) code = WPSGPU_W_AVERAGE4;
        else if (tok_is(tok, "wt_average_16pt")) code = WPSGPU_W_AVERAGE16;
        else if (tok_is(tok, "four_pt")) code = WPSGPU_FOUR_POINT;
        else if (tok_is(tok, "sixteen_pt")) code = WPSGPU_SIXTEEN_POINT;
        else if (tok.find("search") != std::string::npos) {
            code = WPSGPU_SEARCH;
            WPS_CHECK(search_depth(tok, depth) == 0, -21,
                      "Search depth option to search interpolator must be an integer value, enclosed in parentheses "
                      "immediately after keyword \"search\"");
        }
        // else: average_gcell is skipped silently, anything else is a WARN in the reference
        if (code) { codes[j] = code; opts[j] = depth; ++j; }
        if (last) break;
        p1 = p2 + 1;
    }
    *n = num_methods + 1;
    if (has_gcell) *has_gcell = 0;
    if (gcell_threshold) *gcell_threshold = 1.0f;
    size_t g = s.find("average_gcell");
    if (g != std::string::npos) {
        if (has_gcell) *has_gcell = 1;
        // get_gcell_threshold (interp_option_module.F:692-727): '(' / ')' positions relative to the keyword
        // are used as absolute positions in the string
        size_t q1 = s.find('(', g), q2 = s.find(')', g);
        if (q1 != std::string::npos && q2 != std::string::npos) {
            size_t r1 = q1 - g, r2 = q2 - g;  // 0-based relative == Fortran (p-1)
            WPS_CHECK(r2 > r1 + 1 && r2 <= s.size(), -22, "Threshold option to average_gcell interpolator must be a real number");
            std::string num = s.substr(r1 + 1, r2 - r1 - 1);
            char* end;
            float v = strtof(num.c_str(), &end);
            WPS_CHECK(end != num.c_str(), -22, "Threshold option to average_gcell interpolator must be a real number, "
                      "enclosed in parentheses immediately after keyword \"average_gcell\"");
            if (gcell_threshold) *gcell_threshold = v;
        }
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// lltoxy / xytoll over arrays
namespace wps {
__global__ void k_lltoxy(DevProj p, int stagger, const float* __restrict__ lat, const float* __restrict__ lon,
                         int64_t n, float* __restrict__ x, float* __restrict__ y)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float xx, yy;
    lltoxy(p, lat[k], lon[k], xx, yy, stagger);
    x[k] = xx; y[k] = yy;
}
__global__ void k_xytoll(DevProj p, int stagger, const float* __restrict__ x, const float* __restrict__ y,
                         int64_t n, float* __restrict__ lat, float* __restrict__ lon)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float la, lo;
    xytoll(p, x[k], y[k], la, lo, stagger);
    lat[k] = la; lon[k] = lo;
}
}  // namespace wps

extern "C" {

static int pointwise(wpsgpu_handle_t h, const wpsgpu_proj* p, int stagger, const float* a, const float* b,
                     int64_t n, float* o1, float* o2, bool fwd)
{
    WPS_CHECK(h && p && a && b && o1 && o2, -1, "NULL argument");
    WPS_CHECK(p->init, -1, "projection not initialised (call wpsgpu_map_set)");
    if (n <= 0) return 0;
    WPS_CUDA(cudaSetDevice(h->device));
    size_t bytes = sizeof(float) * (size_t)n;
    h->arena.reset();
    WPS_TRY(h->arena.reserve(h->ptr_mode == WPSGPU_PTR_HOST ? 4 * bytes + (1 << 16) : (1 << 16), h->stream));
    const float* dg;
    WPS_TRY(upload_gauss(h, *p, &dg));
    void *da, *db, *d1, *d2;
    WPS_TRY(stage_in(h, a, bytes, &da));
    WPS_TRY(stage_in(h, b, bytes, &db));
    WPS_TRY(alloc_out(h, o1, bytes, &d1));
    WPS_TRY(alloc_out(h, o2, bytes, &d2));
    DevProj dp = to_dev(*p, dg);
    int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    if (fwd) k_lltoxy<<<(unsigned)blocks, threads, 0, h->stream>>>(dp, stagger, (const float*)da, (const float*)db, n, (float*)d1, (float*)d2);
    else k_xytoll<<<(unsigned)blocks, threads, 0, h->stream>>>(dp, stagger, (const float*)da, (const float*)db, n, (float*)d1, (float*)d2);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    WPS_TRY(copy_out(h, o1, d1, bytes));
    WPS_TRY(copy_out(h, o2, d2, bytes));
    if (h->ptr_mode == WPSGPU_PTR_HOST) WPS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wpsgpu_lltoxy(wpsgpu_handle_t h, const wpsgpu_proj* p, int stagger, const float* lat, const float* lon,
                  int64_t n, float* x, float* y)
{ return pointwise(h, p, stagger, lat, lon, n, x, y, true); }

int wpsgpu_xytoll(wpsgpu_handle_t h, const wpsgpu_proj* p, int stagger, const float* x, const float* y,
                  int64_t n, float* lat, float* lon)
{ return pointwise(h, p, stagger, x, y, n, lat, lon, false); }

}  // extern "C"
