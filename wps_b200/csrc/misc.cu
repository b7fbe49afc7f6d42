// misc.cu -- wind rotation, smoothers, fractions/dominant/landmask and the elementwise geometry outputs.
#include <vector>
#include "ctx.h"

namespace wps {

// ---- metmap_xform (rotate_winds_module.F:85-387), C grid, LC / PS (Cassini branches not ported) ----
struct RotArgs {
    const float *u, *v, *xlon_u, *xlon_v, *xlat_u, *xlat_v;
    int is_cassini;
    DevProj proj;
    const int *u_mask, *v_mask;
    float *u_new, *v_new;
    const float2 *cs_u, *cs_v;     // (cos alpha, sin alpha) per U / V point
    int us1, us2, ue1, ue2, vs1, vs2, ve1, ve2, unx, uny, vnx, vny, unxw, vnxw;
    int idir, is_lc;
    float stdlon, cone, hemi;
    int do_last_col_u, do_last_row_u, do_last_col_v, do_last_row_v;
};

// rotation angle of one point; (ii, jj) 0-based in the nx x ny array the point belongs to (rotate_winds_module.F:161-226)
__device__ __forceinline__ float rot_alpha(const RotArgs& a, const float* __restrict__ xlat, const float* __restrict__ xlon,
                                           int nx, int ny, int ii, int jj)
{
    const size_t o = (size_t)jj * nx + ii;
    float diff = (float)a.idir * (xlon[o] - a.stdlon);
    if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
    if (a.is_lc) return diff * a.cone * WPS_RAD_PER_DEG * a.hemi;
    if (a.is_cassini) {
        if (a.idir == 1) {   // :174-189: finite difference of the source projection along its own y axis
            float x, y, latp, lonp, latm, lonm;
            latlon_to_ij(a.proj, xlat[o], xlon[o], x, y);
            ij_to_latlon(a.proj, x, y + 0.1f, latp, lonp);
            ij_to_latlon(a.proj, x, y - 0.1f, latm, lonm);
            float diff_lon = lonp - lonm;
            if (diff_lon > 180.0f) diff_lon = diff_lon - 360.0f; else if (diff_lon < -180.0f) diff_lon = diff_lon + 360.0f;
            float diff_lat = latp - latm;
            return -t_atan2(-t_cos(xlat[o] * WPS_RAD_PER_DEG) * diff_lon * WPS_RAD_PER_DEG, diff_lat * WPS_RAD_PER_DEG);
        }
        // :190-222: difference of the domain's lat/lon along j, one-sided on the first and last rows
        int ja, jb;
        if (jj == ny - 1) { ja = jj; jb = jj - 1; } else if (jj == 0) { ja = jj + 1; jb = jj; } else { ja = jj + 1; jb = jj - 1; }
        const size_t oa = (size_t)ja * nx + ii, ob = (size_t)jb * nx + ii;
        float d = xlon[oa] - xlon[ob];
        if (d > 180.0f) d = d - 360.0f; else if (d < -180.0f) d = d + 360.0f;
        return t_atan2(-t_cos(xlat[o] * WPS_RAD_PER_DEG) * d * WPS_RAD_PER_DEG, (xlat[oa] - xlat[ob]) * WPS_RAD_PER_DEG);
    }
    return diff * WPS_RAD_PER_DEG * a.hemi;
}

// The rotation angle and its cosine / sine (binary64, rounded once) depend on the point only: evaluated once per call for
// every U and V point (k_rot_angles) and shared by all levels instead of once per point and level.
__global__ void k_rot_angles(RotArgs a, int is_v, float2* __restrict__ cs)
{
    const int nx = is_v ? a.vnx : a.unx, ny = is_v ? a.vny : a.uny;
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ii >= nx || jj >= ny) return;
    const float alpha = is_v ? rot_alpha(a, a.xlat_v, a.xlon_v, nx, ny, ii, jj) : rot_alpha(a, a.xlat_u, a.xlon_u, nx, ny, ii, jj);
    cs[(size_t)jj * nx + ii] = make_float2(t_cos(alpha), t_sin(alpha));
}

__global__ void k_rotate_u(RotArgs a, int nlev)
{
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y, lev = blockIdx.z;
    if (ii >= a.unx || jj >= a.uny) return;
    int i = a.us1 + ii, j = a.us2 + jj;
    size_t un = (size_t)a.unx * a.uny, vn = (size_t)a.vnx * a.vny;
    {
    const float *u = a.u + lev * un, *v = a.v + lev * vn;
    const int *um = a.u_mask + (size_t)lev * a.uny * a.unxw, *vm = a.v_mask + (size_t)lev * a.vny * a.vnxw;
    auto ubit = [&](int x, int y) { int xi = x - a.us1, yi = y - a.us2; return (um[(size_t)yi * a.unxw + (xi >> 5)] >> (xi & 31)) & 1; };
    auto vmul = [&](int x, int y) { int xi = x - a.vs1, yi = y - a.vs2; return ((vm[(size_t)yi * a.vnxw + (xi >> 5)] >> (xi & 31)) & 1) ? 1.0f : 0.0f; };
    auto V = [&](int x, int y) { return v[(size_t)(y - a.vs2) * a.vnx + (x - a.vs1)]; };
    size_t o = (size_t)jj * a.unx + ii;
    float uold = u[o], res = uold;
    if (ubit(i, j)) {
        const float2 cs = __ldg(a.cs_u + o);
        const float ca = cs.x, sa = cs.y;
        float v_weight, v_map;
        if (i == a.us1) {
            if (j == a.ue2 && a.do_last_row_u) { v_weight = vmul(i, j); v_map = V(i, j) * vmul(i, j); }
            else { v_weight = vmul(i, j) + vmul(i, j + 1); v_map = V(i, j) * vmul(i, j) + V(i, j + 1) * vmul(i, j + 1); }
        } else if (i == a.ue1 && a.do_last_col_u) {
            if (j == a.ue2 && a.do_last_row_u) { v_weight = vmul(i - 1, j); v_map = V(i - 1, j); /* :244 */ }
            else { v_weight = vmul(i - 1, j) + vmul(i - 1, j + 1); v_map = V(i - 1, j) * vmul(i - 1, j) + V(i - 1, j + 1) * vmul(i - 1, j + 1); }
        } else if (j == a.ue2 && a.do_last_row_u) {
            v_weight = vmul(i - 1, j - 1) + vmul(i, j - 1);
            v_map = V(i - 1, j - 1) * vmul(i - 1, j - 1) + V(i, j - 1) * vmul(i, j - 1);
        } else {
            // the interior case, with the index arithmetic done once in 32 bits and every mask bit read once
            const int xv = i - a.vs1, yv = j - a.vs2, ov = yv * a.vnx + xv, om = yv * a.vnxw + (xv >> 5), bi = xv & 31;
            const unsigned w0 = (unsigned)vm[om], w1 = (unsigned)vm[om + a.vnxw];
            unsigned wl0 = w0, wl1 = w1;
            int bl = bi - 1;
            if (bi == 0) { wl0 = (unsigned)vm[om - 1]; wl1 = (unsigned)vm[om + a.vnxw - 1]; bl = 31; }
            const float m00 = ((wl0 >> bl) & 1u) ? 1.0f : 0.0f, m01 = ((wl1 >> bl) & 1u) ? 1.0f : 0.0f;     // (i - 1, j), (i - 1, j + 1)
            const float m10 = ((w0 >> bi) & 1u) ? 1.0f : 0.0f, m11 = ((w1 >> bi) & 1u) ? 1.0f : 0.0f;       // (i, j), (i, j + 1)
            v_weight = m00 + m01 + m10 + m11;
            v_map = v[ov - 1] * m00 + v[ov + a.vnx - 1] * m01 + v[ov] * m10 + v[ov + a.vnx] * m11;
        }
        if (v_weight > 0.0f) res = ca * uold + sa * v_map / v_weight;
    }
    a.u_new[lev * un + o] = res;
    }
}

__global__ void k_rotate_v(RotArgs a, int nlev)
{
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y, lev = blockIdx.z;
    if (ii >= a.vnx || jj >= a.vny) return;
    int i = a.vs1 + ii, j = a.vs2 + jj;
    size_t un = (size_t)a.unx * a.uny, vn = (size_t)a.vnx * a.vny;
    {
    const float *u = a.u + lev * un, *v = a.v + lev * vn;
    const int *um = a.u_mask + (size_t)lev * a.uny * a.unxw, *vm = a.v_mask + (size_t)lev * a.vny * a.vnxw;
    auto vbit = [&](int x, int y) { int xi = x - a.vs1, yi = y - a.vs2; return (vm[(size_t)yi * a.vnxw + (xi >> 5)] >> (xi & 31)) & 1; };
    auto umul = [&](int x, int y) { int xi = x - a.us1, yi = y - a.us2; return ((um[(size_t)yi * a.unxw + (xi >> 5)] >> (xi & 31)) & 1) ? 1.0f : 0.0f; };
    auto U = [&](int x, int y) { return u[(size_t)(y - a.us2) * a.unx + (x - a.us1)]; };
    size_t o = (size_t)jj * a.vnx + ii;
    float vold = v[o], res = vold;
    if (vbit(i, j)) {
        const float2 cs = __ldg(a.cs_v + o);
        const float ca = cs.x, sa = cs.y;
        float u_weight, u_map;
        if (j == a.vs2) {
            if (i == a.ve1 && a.do_last_col_v) { u_weight = umul(i, j); u_map = U(i, j) * umul(i, j); }
            else { u_weight = umul(i, j) + umul(i + 1, j); u_map = U(i, j) * umul(i, j) + U(i + 1, j) * umul(i + 1, j); }
        } else if (j == a.ve2 && a.do_last_row_v) {
            if (i == a.ve1 && a.do_last_col_v) { u_weight = umul(i, j - 1); u_map = U(i, j - 1) * umul(i, j - 1); }
            else { u_weight = umul(i, j - 1) + umul(i + 1, j - 1); u_map = U(i, j - 1) * umul(i, j - 1) + U(i + 1, j - 1) * umul(i + 1, j - 1); }
        } else if (i == a.ve1 && a.do_last_col_v) {
            u_weight = umul(i, j) + umul(i, j - 1);
            u_map = U(i, j) * umul(i, j) + U(i, j - 1) * umul(i, j - 1);
        } else {
            const int xu = i - a.us1, yu = j - a.us2, ou = yu * a.unx + xu, om = yu * a.unxw + (xu >> 5), bi = xu & 31;
            const unsigned w0 = (unsigned)um[om - a.unxw], w1 = (unsigned)um[om];                                // rows j - 1, j
            unsigned wr0 = w0, wr1 = w1;
            int br = bi + 1;
            if (bi == 31) { wr0 = (unsigned)um[om - a.unxw + 1]; wr1 = (unsigned)um[om + 1]; br = 0; }
            const float m00 = ((w0 >> bi) & 1u) ? 1.0f : 0.0f, m01 = ((w1 >> bi) & 1u) ? 1.0f : 0.0f;           // (i, j - 1), (i, j)
            const float m10 = ((wr0 >> br) & 1u) ? 1.0f : 0.0f, m11 = ((wr1 >> br) & 1u) ? 1.0f : 0.0f;         // (i + 1, j - 1), (i + 1, j)
            u_weight = m00 + m01 + m10 + m11;
            u_map = u[ou - a.unx] * m00 + u[ou] * m01 + u[ou - a.unx + 1] * m10 + u[ou + 1] * m11;
        }
        if (u_weight > 0.0f) res = -sa * u_map / u_weight + ca * vold;
    }
    a.v_new[lev * vn + o] = res;
    }
}

// ---- smoothers (smooth_module.F:20-239) -----------------------------------------------------------------
// One launch = one PASS of the smoother: NP sweep pairs (1: one_two_one; 2: smth_desmth's smoothing + desmoothing pair),
// each pair = x sweep into a scratch then y sweep (:38-55, :99-143).  A block produces a 32 x 32 tile from the tile plus an
// apron of NP cells held in shared memory, so a pass reads every value once (1.27x with the apron) and writes it once:
// 8.9-10 B per point against 16 B per sweep PAIR for separate x and y kernels through a global scratch array.  Domain-edge
// rows and columns are never modified (the reference's loops run over start+1 .. end-1); they pass through every stage.
// The `_special` rule (:226-234) is applied by the launch of the last pass.  Arithmetic order and rounding are the
// reference's (this file is compiled without FMA contraction), so results stay bit-identical.
constexpr int SM_T = 32;
template <int NP>
__global__ void __launch_bounds__(256) k_smooth_pass(const float* __restrict__ in, float* out, const float* orig,
                                                     int nx, int ny, int restore)
{
    constexpr int W = SM_T + 2 * NP;
    __shared__ float s_a[W][W + 1], s_b[W][W + 1], s_s[W][W + 1];
    const size_t lev = (size_t)blockIdx.z * nx * ny;
    const int x0 = blockIdx.x * SM_T - NP, y0 = blockIdx.y * SM_T - NP;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int e = tid; e < W * W; e += 256) {
        const int lx = e % W, ly = e / W, gx = x0 + lx, gy = y0 + ly;
        s_a[ly][lx] = (gx >= 0 && gx < nx && gy >= 0 && gy < ny) ? __ldg(in + lev + (size_t)gy * nx + gx) : 0.0f;
    }
    __syncthreads();
    float (*A)[W + 1] = s_a;
    float (*B)[W + 1] = s_b;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const float c0 = p == 0 ? 0.5f : 1.52f, c1 = p == 0 ? 0.25f : -0.26f;
        const int m = p;                       // cells within m of the tile edge no longer hold valid data
        for (int e = tid; e < W * W; e += 256) {
            const int lx = e % W, ly = e / W;
            if (lx > m && lx < W - 1 - m && ly >= m && ly < W - m)
                s_s[ly][lx] = c0 * A[ly][lx] + c1 * (A[ly][lx - 1] + A[ly][lx + 1]);
        }
        __syncthreads();
        for (int e = tid; e < W * W; e += 256) {
            const int lx = e % W, ly = e / W, gx = x0 + lx, gy = y0 + ly;
            if (lx > m && lx < W - 1 - m && ly > m && ly < W - 1 - m) {
                const bool interior = gx >= 1 && gx <= nx - 2 && gy >= 1 && gy <= ny - 2;
                B[ly][lx] = interior ? c0 * s_s[ly][lx] + c1 * (s_s[ly - 1][lx] + s_s[ly + 1][lx]) : A[ly][lx];
            }
        }
        __syncthreads();
        float (*t)[W + 1] = A; A = B; B = t;
    }
    for (int e = tid; e < SM_T * SM_T; e += 256) {
        const int lx = NP + e % SM_T, ly = NP + e / SM_T, gx = x0 + lx, gy = y0 + ly;
        if (gx < nx && gy < ny) {
            const size_t o = lev + (size_t)gy * nx + gx;
            float v = A[ly][lx];
            if (restore) {                       // "remove artificially negative values", smooth_module.F:226-234
                const float og = orig ? orig[o] : __ldg(in + o);            // single pass: the pass's input is the original
                if (v < 0.0f && og >= 0.0f) v = og;
            }
            out[o] = v;
        }
    }
}

// ---- fractions / landmask / dominant (process_tile_module.F:539-622,699-727,1035-1051,1123-1144) ---------
struct FracArgs {
    float* field; size_t npts; int min_cat, max_cat; float msg_fill;
    int lm_vals[32]; int n_lm; int is_water_mask;
    float* landmask; float* dominant;
};
__global__ void k_fractions_dominant(FracArgs a)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npts) return;
    int ncat = a.max_cat - a.min_cat + 1;
    float sum = 0.0f;
    for (int k = 0; k < ncat; ++k) sum = sum + a.field[(size_t)k * a.npts + p];
    if (sum > 0.0f) { for (int k = 0; k < ncat; ++k) a.field[(size_t)k * a.npts + p] = a.field[(size_t)k * a.npts + p] / sum; }
    else { for (int k = 0; k < ncat; ++k) a.field[(size_t)k * a.npts + p] = a.msg_fill; }
    float lm = 0.0f;
    if (a.n_lm > 0) {
        float total = -1.0f;
        for (int k = 0; k < a.n_lm; ++k) {
            int c = a.lm_vals[k];
            if (c >= a.min_cat && c <= a.max_cat) {
                float v = a.field[(size_t)(c - a.min_cat) * a.npts + p];
                if (v != a.msg_fill) { if (total < 0.0f) total = 0.0f; total = total + v; }
                else { total = -1.0f; break; }
            }
        }
        if (total >= 0.0f) lm = a.is_water_mask ? ((total < 0.50f) ? 1.0f : 0.0f) : ((total > 0.50f) ? 1.0f : 0.0f);
        else lm = -1.0f;
        if (a.landmask) a.landmask[p] = lm;
    }
    if (a.dominant) {
        float dominant = 0.0f, d = (float)(a.min_cat - 1);
        if (a.n_lm == 0) {
            for (int k = a.min_cat; k <= a.max_cat; ++k) {
                float v = a.field[(size_t)(k - a.min_cat) * a.npts + p];
                if (v > dominant && v != a.msg_fill) { d = (float)k; dominant = v; }
            }
            if (d == (float)(a.min_cat - 1)) d = a.msg_fill;
        } else if ((lm == 1.0f && a.is_water_mask) || (lm == 0.0f && !a.is_water_mask)) {
            for (int k = a.min_cat; k <= a.max_cat; ++k) {
                int kk;
                for (kk = 1; kk <= a.n_lm; ++kk) if (k == a.lm_vals[kk - 1]) break;
                float v = a.field[(size_t)(k - a.min_cat) * a.npts + p];
                if (v > dominant && kk > a.n_lm) { d = (float)k; dominant = v; }
            }
        } else {
            for (int k = a.min_cat; k <= a.max_cat; ++k)
                for (int kk = 1; kk <= a.n_lm; ++kk) {
                    float v = a.field[(size_t)(k - a.min_cat) * a.npts + p];
                    if (v > dominant && k == a.lm_vals[kk - 1]) { d = (float)k; dominant = v; }
                }
        }
        a.dominant[p] = d;
    }
}

// ---- elementwise geometry (process_tile_module.F:1692-2253) ----------------------------------------------
__global__ void k_xytoll_grid(DevProj p, int stagger, int si, int sj, int nx, int ny, float sub_x, float sub_y,
                              float* __restrict__ xlat, float* __restrict__ xlon)
{
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ii >= nx || jj >= ny) return;
    float x = ((float)(si + ii) - 0.5f) / sub_x + 0.5f, y = ((float)(sj + jj) - 0.5f) / sub_y + 0.5f;
    float la, lo;
    xytoll(p, x, y, la, lo, stagger);
    xlat[(size_t)jj * nx + ii] = la; xlon[(size_t)jj * nx + ii] = lo;
}

struct MapfacArgs { int iproj; float truelat1, truelat2, pole_lat, pole_lon, stand_lon; float n, colat0, colat2, s0; };
__global__ void k_map_factor(MapfacArgs a, const float* __restrict__ xlat, const float* __restrict__ xlon, size_t npts,
                             float* __restrict__ mx, float* __restrict__ my)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float R = WPS_RAD_PER_DEG;
    float la = xlat[p], fx, fy;
    if (a.iproj == WPSGPU_PROJ_LC) {
        float colat = R * (90.0f - la);
        if (a.truelat1 != a.truelat2) fx = t_sin(a.colat2) / t_sin(colat) * t_pow(t_tan(colat / 2.0f) / t_tan(a.colat2 / 2.0f), a.n);
        else fx = t_sin(a.colat0) / t_sin(colat) * t_pow(t_tan(colat / 2.0f) / t_tan(a.colat0 / 2.0f), t_cos(a.colat0));
        fy = fx;
    } else if (a.iproj == WPSGPU_PROJ_PS) {
        float sg = (a.truelat1 >= 0.0f) ? 1.0f : -1.0f;
        fx = (1.0f + t_sin(R * fabsf(a.truelat1))) / (1.0f + t_sin(R * sg * la));
        fy = fx;
    } else if (a.iproj == WPSGPU_PROJ_MERC) {
        float colat = R * (90.0f - la);
        fx = t_sin(a.colat0) / t_sin(colat);
        fy = fx;
    } else if (a.iproj == WPSGPU_PROJ_CYL) {
        fx = (fabsf(la) == 90.0f) ? 0.0f : 1.0f / t_cos(la * R);
        fy = 1.0f;
    } else if (a.iproj == WPSGPU_PROJ_CASSINI) {
        if (fabsf(a.pole_lat) == 90.0f) fx = (fabsf(la) >= 90.0f) ? 0.0f : 1.0f / t_cos(la * R);
        else {
            float cl, co;
            rotate_coords(la, xlon[p], cl, co, a.pole_lat, a.pole_lon, a.stand_lon, -1);
            fx = (fabsf(cl) >= 90.0f) ? 0.0f : 1.0f / t_cos(cl * R);
        }
        fy = 1.0f;
    } else { fx = 1.0f; fy = 1.0f; }
    mx[p] = fx; my[p] = fy;
}

__global__ void k_coriolis(const float* __restrict__ xlat, size_t npts, float* __restrict__ f, float* __restrict__ e)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float OMEGA_E = 7.292e-5f;
    f[p] = 2.0f * OMEGA_E * t_sin(WPS_RAD_PER_DEG * xlat[p]);
    e[p] = 2.0f * OMEGA_E * t_cos(WPS_RAD_PER_DEG * xlat[p]);
}

__global__ void k_rotang(int iproj, float cone, float stand_lon, const float* __restrict__ xlat, const float* __restrict__ xlon,
                         int nx, int ny, float* __restrict__ cosa, float* __restrict__ sina)
{
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ii >= nx || jj >= ny) return;
    size_t o = (size_t)jj * nx + ii;
    if (iproj == WPSGPU_PROJ_LC || iproj == WPSGPU_PROJ_PS) {
        float d_lon = stand_lon - xlon[o];
        if (d_lon > 180.0f) d_lon = d_lon - 360.0f; else if (d_lon < -180.0f) d_lon = d_lon + 360.0f;
        float alpha = (iproj == WPSGPU_PROJ_LC) ? d_lon * cone * WPS_RAD_PER_DEG : d_lon * WPS_RAD_PER_DEG;
        sina[o] = t_sin(alpha); cosa[o] = t_cos(alpha);
    } else if (iproj == WPSGPU_PROJ_CASSINI) {
        int jm = (jj == 0) ? jj : jj - 1, jp = (jj == ny - 1) ? jj : jj + 1;
        float d_lon = xlon[(size_t)jp * nx + ii] - xlon[(size_t)jm * nx + ii];
        if (d_lon > 180.0f) d_lon = d_lon - 360.0f; else if (d_lon < -180.0f) d_lon = d_lon + 360.0f;
        float alpha = t_atan2(-t_cos(xlat[o] * WPS_RAD_PER_DEG) * (d_lon * WPS_RAD_PER_DEG),
                              ((xlat[(size_t)jp * nx + ii] - xlat[(size_t)jm * nx + ii]) * WPS_RAD_PER_DEG));
        sina[o] = t_sin(alpha); cosa[o] = t_cos(alpha);
    } else { sina[o] = 0.0f; cosa[o] = 1.0f; }
}

// calc_dfdx / calc_dfdy with the stale-index map factor on the edge rows/columns (:2163,2167,2228,2232)
__global__ void k_dfd(int dir, const float* __restrict__ src, float* __restrict__ dst, int nx, int ny, float sr,
                      const float* __restrict__ mapfac, float dom_d, int stale)
{
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ii >= nx || jj >= ny) return;
    size_t lev = (size_t)blockIdx.z * nx * ny, o2 = (size_t)jj * nx + ii, o = lev + o2;
    if (dir == 0) {
        if (ii >= 1 && ii <= nx - 2) {
            float den = mapfac ? (2.0f * dom_d * mapfac[o2] / sr) : (2.0f * dom_d / sr);
            dst[o] = (src[o + 1] - src[o - 1]) / den;
        } else {
            float den = mapfac ? (dom_d * mapfac[(size_t)jj * nx + stale] / sr) : (dom_d / sr);
            dst[o] = (ii == 0) ? (src[o + 1] - src[o]) / den : (src[o] - src[o - 1]) / den;
        }
    } else {
        if (jj >= 1 && jj <= ny - 2) {
            float den = mapfac ? (2.0f * dom_d * mapfac[o2] / sr) : (2.0f * dom_d / sr);
            dst[o] = (src[o + nx] - src[o - nx]) / den;
        } else {
            float den = mapfac ? (dom_d * mapfac[(size_t)stale * nx + ii] / sr) : (dom_d / sr);
            dst[o] = (jj == 0) ? (src[o + nx] - src[o]) / den : (src[o] - src[o - nx]) / den;
        }
    }
}

}  // namespace wps

using namespace wps;


// ---- fill_missing_levels / create_level (process_domain_module.F:2748-2933, 3191-3332) -------------------------
// One thread per (ix, jx) column, levels walked in storage order exactly like the reference's k loop: a level filled
// at step k is a valid "lower" neighbour for the levels after it.  A warp owns one word of each level's valid_mask,
// so the updated word is assembled with a ballot and stored by lane 0.
struct VFillArgs { float* const* r; int* const* v; const int* lev; int nlev, nx, ny, nxw; };

__global__ void __launch_bounds__(256) k_fill_missing_levels(VFillArgs a)
{
    const int ix = blockIdx.x * 32 + threadIdx.x, jx = blockIdx.y * 8 + threadIdx.y;
    if (jx >= a.ny) return;                       // whole warps only: threadIdx.y selects the warp
    const bool in = ix < a.nx;
    const size_t o = (size_t)jx * a.nx + ix, w = (size_t)jx * a.nxw + blockIdx.x;
    const unsigned lane_bit = 1u << threadIdx.x;
    int lower = -1;                               // last level below k whose bit is set for this column
    int upper = -1;                               // first level above k whose bit is set (searched lazily)
    for (int k = 0; k < a.nlev; ++k) {
        unsigned word = (unsigned)a.v[k][w];
        bool valid = (word & lane_bit) != 0;
        bool fill = false;
        if (in && !valid) {
            if (upper <= k) {
                for (upper = k + 1; upper < a.nlev; ++upper)
                    if (((unsigned)a.v[upper][w]) & lane_bit) break;
            }
            if (upper < a.nlev && lower >= 0) {
                const int lu = a.lev[upper], lk = a.lev[k], ll = a.lev[lower];
                float wl = (float)abs(lu - lk) / (float)abs(lu - ll);
                float wu = (float)abs(lk - ll) / (float)abs(lu - ll);
                a.r[k][o] = wl * a.r[lower][o] + wu * a.r[upper][o];
                fill = true;
            }
        }
        unsigned add = __ballot_sync(0xffffffffu, fill);
        if (threadIdx.x == 0 && add) a.v[k][w] = (int)(word | add);
        if (valid || fill) lower = k;
    }
}

// The same rule for up to 64 levels without walking memory level by level: a warp fetches the valid_mask words of ALL levels
// of its 32 columns at once (lane l holds levels l and l + 32), every lane assembles the validity bits of its own column into
// one 64-bit word, and the level loop then runs on registers -- a gap between two valid levels costs two loads (its lower and
// upper values; a filled level hands its value on in a register) and its stores, instead of a dependent load chain per level.
__global__ void __launch_bounds__(256) k_fill_missing_levels64(VFillArgs a)
{
    const int lane = threadIdx.x;
    const int ix = blockIdx.x * 32 + lane, jx = blockIdx.y * 8 + threadIdx.y;
    if (jx >= a.ny) return;                       // whole warps only
    const bool in = ix < a.nx;
    const size_t o = (size_t)jx * a.nx + ix, w = (size_t)jx * a.nxw + blockIdx.x;
    const unsigned w0 = lane < a.nlev ? (unsigned)a.v[lane][w] : 0u;
    const unsigned w1 = lane + 32 < a.nlev ? (unsigned)a.v[lane + 32][w] : 0u;
    unsigned long long mask = 0ull;               // bit k: level k is valid for this lane's column
    for (int k = 0; k < a.nlev; ++k) {
        const unsigned wk = __shfl_sync(0xffffffffu, k < 32 ? w0 : w1, k & 31);
        mask |= (unsigned long long)((wk >> lane) & 1u) << k;
    }
    int lower = -1, upper = -1;
    float rl = 0.0f, ru = 0.0f;
    bool rl_have = false;                         // rl holds r[lower] (a value this thread has just filled in)
    for (int k = 0; k < a.nlev; ++k) {
        const bool valid = (mask >> k) & 1ull;
        bool fill = false;
        if (in && !valid) {
            if (upper <= k) {
                const unsigned long long above = mask & ~((2ull << k) - 1ull);
                upper = above ? __ffsll((long long)above) - 1 : a.nlev;
                if (upper < a.nlev) ru = a.r[upper][o];
            }
            if (upper < a.nlev && lower >= 0) {
                if (!rl_have) { rl = a.r[lower][o]; rl_have = true; }
                const int lu = a.lev[upper], lk = a.lev[k], ll = a.lev[lower];
                const float wl = (float)abs(lu - lk) / (float)abs(lu - ll);
                const float wu = (float)abs(lk - ll) / (float)abs(lu - ll);
                const float val = wl * rl + wu * ru;
                a.r[k][o] = val;
                rl = val;                         // level k is the lower neighbour of what follows
                fill = true;
            }
        }
        const unsigned add = __ballot_sync(0xffffffffu, fill);
        if (add) {
            const unsigned word = __shfl_sync(0xffffffffu, k < 32 ? w0 : w1, k & 31);
            if (lane == 0) a.v[k][w] = (int)(word | add);
        }
        if (valid) { lower = k; rl_have = false; }
        else if (fill) lower = k;
    }
}

__global__ void k_create_level(float value, const float* __restrict__ src, const int* __restrict__ src_valid,
                               float* __restrict__ r, int* __restrict__ valid, int nx, int ny, int nxw)
{
    const int ix = blockIdx.x * 32 + threadIdx.x, jx = blockIdx.y * blockDim.y + threadIdx.y;
    if (jx >= ny) return;
    if (ix < nx) r[(size_t)jx * nx + ix] = src ? src[(size_t)jx * nx + ix] : value;
    if (threadIdx.x == 0) {
        size_t w = (size_t)jx * nxw + blockIdx.x;
        int rest = nx - blockIdx.x * 32;
        unsigned full = rest >= 32 ? 0xffffffffu : ((1u << rest) - 1u);   // bitarray_set for i = 1..nx only
        valid[w] = src_valid ? (int)((unsigned)src_valid[w] & full) : (int)full;
    }
}


// one sweep of one_two_one_egrid / smth_desmth_egrid (smooth_module.F:252-338, :448-585): the two diagonal neighbours of an
// E-grid point in the rows below and above; `second` swaps the east / west offsets as the reference's second sweep of a pair
// does; rows [sy + ylo, ey - ylo]; cond = the array whose value gates the update when msgval == 1.0 (:296, :306)
__global__ void __launch_bounds__(256) k_egrid_sweep(const float* __restrict__ src, float* dst, const float* cond, int nx, int ny, int sy,
                                                     int hflag1, int ylo, float cw, float ew, int second, int msg1)
{
    const int ix = blockIdx.x * 32 + threadIdx.x, iy0 = blockIdx.y * 8 + threadIdx.y;
    if (ix >= nx || iy0 < ylo || iy0 > ny - 1 - ylo) return;
    const int iy = sy + iy0;
    const int par = hflag1 ? (iy + 1) % 2 : iy % 2;
    const int ihe = par < 0 ? -par : par, ihw = ihe - 1;
    const int istart = hflag1 ? (iy % 2 == 0 ? 0 : 1) : (iy % 2 != 0 ? 0 : 1);
    if (ix < istart || ix > nx - 2) return;
    const size_t pl = (size_t)blockIdx.z * nx * ny, k = pl + (size_t)iy0 * nx + ix;
    if (msg1 && cond[k] == 0.0f) return;
    const float a = src[k - nx + (second ? ihe : ihw)], b = src[k + nx + (second ? ihw : ihe)];
    dst[k] = cw * src[k] + ew * (a + b);
}

// metmap_xform_nmm (rotate_winds_module.F:455-631): U and V collocated on the velocity points of the E-grid; the rotation
// angle of a point once, then every level
__global__ void __launch_bounds__(256) k_rotate_nmm(int code, float lat1, float lon1, float stdlon, float cone, float hemi, int idir,
                                                    const float* __restrict__ xlat_v, const float* __restrict__ xlon_v, int nx, int ny, int nlev,
                                                    float* __restrict__ u, const int* __restrict__ u_mask, float* __restrict__ v, const int* __restrict__ v_mask)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const size_t k = (size_t)j * nx + i;
    const float fdir = (float)idir;
    float sin_alpha, cos_alpha;
    if (code == WPSGPU_PROJ_ROTLL) {
        const float phi0 = lat1 * WPS_RAD_PER_DEG;
        const float lmbd0 = lon1 < 0.0f ? (lon1 + 360.0f) * WPS_RAD_PER_DEG : lon1 * WPS_RAD_PER_DEG;
        const float sin_phi0 = t_sin(phi0), cos_phi0 = t_cos(phi0);
        const float rlat_v = xlat_v[k] * WPS_RAD_PER_DEG, rlon_v = xlon_v[k] * WPS_RAD_PER_DEG;
        const float relm = rlon_v - lmbd0;
        const float big_denominator = t_cos(t_asin(cos_phi0 * t_sin(rlat_v) - sin_phi0 * t_cos(rlat_v) * t_cos(relm)));
        sin_alpha = sin_phi0 * t_sin(relm) / big_denominator;
        cos_alpha = (cos_phi0 * t_cos(rlat_v) + sin_phi0 * t_sin(rlat_v) * t_cos(relm)) / big_denominator;
    } else {
        float diff = fdir * (xlon_v[k] - stdlon);
        if (diff > 180.0f) diff = diff - 360.0f; else if (diff < -180.0f) diff = diff + 360.0f;
        const float alpha = code == WPSGPU_PROJ_LC ? diff * cone * WPS_RAD_PER_DEG * hemi : diff * WPS_RAD_PER_DEG * hemi;
        sin_alpha = t_sin(alpha); cos_alpha = t_cos(alpha);
    }
    const int nxw = (nx + 31) / 32;
    const size_t npl = (size_t)nx * ny, nwl = (size_t)ny * nxw;
    for (int l = 0; l < nlev; ++l) {
        const bool ub = (u_mask[l * nwl + (size_t)j * nxw + (i >> 5)] >> (i & 31)) & 1, vb = (v_mask[l * nwl + (size_t)j * nxw + (i >> 5)] >> (i & 31)) & 1;
        const float uo = u[l * npl + k], vo = v[l * npl + k];
        if (ub) u[l * npl + k] = cos_alpha * uo + fdir * sin_alpha * (vb ? vo : 0.0f);
        if (vb) v[l * npl + k] = -fdir * sin_alpha * (ub ? uo : 0.0f) + cos_alpha * vo;
    }
}

static int begin_call(wpsgpu_ctx* h, size_t host_bytes)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    WPS_CUDA(cudaSetDevice(h->device));
    WPS_TRY(arena_begin(h));
    h->arena.reset();
    return h->arena.reserve((h->ptr_mode == WPSGPU_PTR_HOST ? host_bytes : 0) + (1 << 20), h->stream);
}
static int end_call(wpsgpu_ctx* h)
{
    WPS_CUDA(cudaGetLastError());
    if (h->ptr_mode == WPSGPU_PTR_HOST) WPS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
// inout array: staged in (host mode) and copied back
static int stage_inout(wpsgpu_ctx* h, void* p, size_t bytes, void** d) { return stage_in(h, p, bytes, d); }

extern "C" {

int wpsgpu_fill_missing_levels(wpsgpu_handle_t h, int32_t nlev, const int32_t* levels, float* const* r_arr,
                               int32_t* const* valid_mask, int32_t nx, int32_t ny)
{
    WPS_CHECK(h && levels && r_arr && valid_mask, -1, "wpsgpu_fill_missing_levels: NULL argument");
    WPS_CHECK(nlev >= 0 && nx > 0 && ny > 0, -1, "wpsgpu_fill_missing_levels: bad dimensions");
    if (nlev <= 1) return 0;   // "If this isn't a 3-d array, nothing to do" (:2870)
    const int nxw = (nx + 31) / 32;
    const size_t rb = sizeof(float) * (size_t)nx * ny, vb = sizeof(int) * (size_t)ny * nxw;
    WPS_TRY(begin_call(h, (rb + vb + 512) * nlev + 65536));
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) WPS_TRY(h->arena.reserve((size_t)nlev * 32 + (1 << 20), h->stream));
    std::vector<float*> dr(nlev);
    std::vector<int*> dv(nlev);
    for (int k = 0; k < nlev; ++k) {
        WPS_CHECK(r_arr[k] && valid_mask[k], -1, "wpsgpu_fill_missing_levels: NULL level array");
        void *a, *b;
        WPS_TRY(stage_inout(h, r_arr[k], rb, &a));
        WPS_TRY(stage_inout(h, valid_mask[k], vb, &b));
        dr[k] = (float*)a; dv[k] = (int*)b;
    }
    // pointer and level tables (always host arrays of the ABI) go up through the arena
    float** d_r = (float**)h->arena.take(sizeof(float*) * nlev);
    int** d_v = (int**)h->arena.take(sizeof(int*) * nlev);
    int* d_lev = (int*)h->arena.take(sizeof(int) * nlev);
    WPS_CHECK(d_r && d_v && d_lev, -3, "device arena exhausted");
    WPS_CUDA(cudaMemcpyAsync(d_r, dr.data(), sizeof(float*) * nlev, cudaMemcpyHostToDevice, h->stream));
    WPS_CUDA(cudaMemcpyAsync(d_v, dv.data(), sizeof(int*) * nlev, cudaMemcpyHostToDevice, h->stream));
    WPS_CUDA(cudaMemcpyAsync(d_lev, levels, sizeof(int) * nlev, cudaMemcpyHostToDevice, h->stream));
    VFillArgs a;
    a.r = d_r; a.v = d_v; a.lev = d_lev; a.nlev = nlev; a.nx = nx; a.ny = ny; a.nxw = nxw;
    if (nlev <= 64) k_fill_missing_levels64<<<dim3(nxw, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(a);
    else k_fill_missing_levels<<<dim3(nxw, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(a);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    if (h->ptr_mode == WPSGPU_PTR_HOST)
        for (int k = 0; k < nlev; ++k) {
            WPS_CUDA(cudaMemcpyAsync(r_arr[k], dr[k], rb, cudaMemcpyDeviceToHost, h->stream));
            WPS_CUDA(cudaMemcpyAsync(valid_mask[k], dv[k], vb, cudaMemcpyDeviceToHost, h->stream));
        }
    return end_call(h);
}

int wpsgpu_create_level(wpsgpu_handle_t h, int32_t nx, int32_t ny, float fill_const, const float* src_r_arr,
                        const int32_t* src_valid_mask, float* r_arr, int32_t* valid_mask)
{
    WPS_CHECK(h && r_arr && valid_mask, -1, "wpsgpu_create_level: NULL argument");
    WPS_CHECK((src_r_arr == nullptr) == (src_valid_mask == nullptr), -1,
              "wpsgpu_create_level: source array and source valid_mask must be given together");
    WPS_CHECK(nx > 0 && ny > 0, -1, "wpsgpu_create_level: bad dimensions");
    const int nxw = (nx + 31) / 32;
    const size_t rb = sizeof(float) * (size_t)nx * ny, vb = sizeof(int) * (size_t)ny * nxw;
    WPS_TRY(begin_call(h, 2 * (rb + vb) + 8192));
    void *ds = nullptr, *dsv = nullptr, *dr = r_arr, *dv = valid_mask;
    if (src_r_arr) { WPS_TRY(stage_in(h, src_r_arr, rb, &ds)); WPS_TRY(stage_in(h, src_valid_mask, vb, &dsv)); }
    if (h->ptr_mode == WPSGPU_PTR_HOST) {
        dr = h->arena.take(rb); dv = h->arena.take(vb);
        WPS_CHECK(dr && dv, -3, "device arena exhausted");
    }
    k_create_level<<<dim3(nxw, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(fill_const, (const float*)ds, (const int*)dsv, (float*)dr, (int*)dv, nx, ny, nxw);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    if (h->ptr_mode == WPSGPU_PTR_HOST) {
        WPS_CUDA(cudaMemcpyAsync(r_arr, dr, rb, cudaMemcpyDeviceToHost, h->stream));
        WPS_CUDA(cudaMemcpyAsync(valid_mask, dv, vb, cudaMemcpyDeviceToHost, h->stream));
    }
    return end_call(h);
}

int wpsgpu_rotate_winds(wpsgpu_handle_t h, int idir, const wpsgpu_proj* proj, float* u, const int32_t* u_mask,
                        float* v, const int32_t* v_mask, int us1, int us2, int ue1, int ue2,
                        int vs1, int vs2, int ve1, int ve2, const float* xlon_u, const float* xlon_v, int nlev)
{
    return wpsgpu_rotate_winds_ll(h, idir, proj, u, u_mask, v, v_mask, us1, us2, ue1, ue2, vs1, vs2, ve1, ve2, xlon_u, xlon_v,
                                  nullptr, nullptr, nlev);
}

int wpsgpu_rotate_winds_ll(wpsgpu_handle_t h, int idir, const wpsgpu_proj* proj, float* u, const int32_t* u_mask,
                           float* v, const int32_t* v_mask, int us1, int us2, int ue1, int ue2,
                           int vs1, int vs2, int ve1, int ve2, const float* xlon_u, const float* xlon_v,
                           const float* xlat_u, const float* xlat_v, int nlev)
{
    WPS_CHECK(h && proj && u && v && u_mask && v_mask && xlon_u && xlon_v, -1, "wpsgpu_rotate_winds: NULL argument");
    WPS_CHECK(proj->init, -1, "In metmap_xform(), uninitialized proj_info structure.");
    WPS_CHECK(proj->code != WPSGPU_PROJ_CASSINI || (xlat_u && xlat_v), -30,
              "wpsgpu_rotate_winds: Cassini wind rotation needs xlat_u / xlat_v (use wpsgpu_rotate_winds_ll)");
    if (!(proj->code == WPSGPU_PROJ_LC || proj->code == WPSGPU_PROJ_PS || proj->code == WPSGPU_PROJ_CASSINI)) return 0;  // :112-114
    WPS_CHECK(proj->code != WPSGPU_PROJ_CASSINI || idir == 1 || (ue2 > us2 && ve2 > vs2), -1,
              "wpsgpu_rotate_winds: Cassini met_to_map needs at least two rows");
    RotArgs a;
    a.unx = ue1 - us1 + 1; a.uny = ue2 - us2 + 1; a.vnx = ve1 - vs1 + 1; a.vny = ve2 - vs2 + 1;
    a.unxw = (a.unx + 31) / 32; a.vnxw = (a.vnx + 31) / 32;
    size_t ub = sizeof(float) * a.unx * a.uny, vb = sizeof(float) * a.vnx * a.vny;
    size_t umb = sizeof(int) * a.uny * a.unxw, vmb = sizeof(int) * a.vny * a.vnxw;
    WPS_TRY(begin_call(h, 2 * (ub + vb) * nlev + (umb + vmb) * nlev + 4 * (ub + vb) + 8192));
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) WPS_TRY(h->arena.reserve((ub + vb) * nlev + 2 * (ub + vb) + (1 << 20), h->stream));
    void *du, *dv, *dum, *dvm, *dxu, *dxv;
    WPS_TRY(stage_inout(h, u, ub * nlev, &du));
    WPS_TRY(stage_inout(h, v, vb * nlev, &dv));
    WPS_TRY(stage_in(h, u_mask, umb * nlev, &dum));
    WPS_TRY(stage_in(h, v_mask, vmb * nlev, &dvm));
    WPS_TRY(stage_in(h, xlon_u, ub, &dxu));
    WPS_TRY(stage_in(h, xlon_v, vb, &dxv));
    void *dlu = nullptr, *dlv = nullptr;
    if (proj->code == WPSGPU_PROJ_CASSINI) { WPS_TRY(stage_in(h, xlat_u, ub, &dlu)); WPS_TRY(stage_in(h, xlat_v, vb, &dlv)); }
    a.xlat_u = (const float*)dlu; a.xlat_v = (const float*)dlv;
    a.is_cassini = proj->code == WPSGPU_PROJ_CASSINI;
    a.proj = to_dev(*proj, nullptr);
    // U is rotated into a scratch array (it reads the old V); V is then rotated IN PLACE -- it reads the old U, which is
    // still untouched, and every V point is read only by the thread that overwrites it -- and U is copied back
    a.u_new = (float*)h->arena.take(ub * nlev);
    a.v_new = (float*)dv;
    WPS_CHECK(a.u_new != nullptr, -3, "device arena exhausted");
    a.u = (const float*)du; a.v = (const float*)dv; a.u_mask = (const int*)dum; a.v_mask = (const int*)dvm;
    a.xlon_u = (const float*)dxu; a.xlon_v = (const float*)dxv;
    a.us1 = us1; a.us2 = us2; a.ue1 = ue1; a.ue2 = ue2; a.vs1 = vs1; a.vs2 = vs2; a.ve1 = ve1; a.ve2 = ve2;
    a.idir = idir; a.is_lc = proj->code == WPSGPU_PROJ_LC; a.stdlon = proj->stdlon; a.cone = proj->cone; a.hemi = proj->hemi;
    if (ue1 - us1 == ve1 - vs1) { a.do_last_col_u = 0; a.do_last_col_v = 1; } else { a.do_last_col_u = 1; a.do_last_col_v = 0; }
    if (ue2 - us2 == ve2 - vs2) { a.do_last_row_u = 1; a.do_last_row_v = 0; } else { a.do_last_row_u = 0; a.do_last_row_v = 1; }
    dim3 b(32, 8);
    float2* cs_u = (float2*)h->arena.take(2 * ub);
    float2* cs_v = (float2*)h->arena.take(2 * vb);
    WPS_CHECK(cs_u && cs_v, -3, "device arena exhausted");
    a.cs_u = cs_u; a.cs_v = cs_v;
    k_rot_angles<<<dim3((a.unx + 31) / 32, (a.uny + 7) / 8), b, 0, h->stream>>>(a, 0, cs_u);
    k_rot_angles<<<dim3((a.vnx + 31) / 32, (a.vny + 7) / 8), b, 0, h->stream>>>(a, 1, cs_v);
    k_rotate_u<<<dim3((a.unx + 31) / 32, (a.uny + 7) / 8, nlev), b, 0, h->stream>>>(a, nlev);
    k_rotate_v<<<dim3((a.vnx + 31) / 32, (a.vny + 7) / 8, nlev), b, 0, h->stream>>>(a, nlev);
    h->launches += 4;
    WPS_CUDA(cudaGetLastError());
    // "Copy rotated winds back into argument arrays" (:374-375)
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) {
        WPS_CUDA(cudaMemcpyAsync(u, a.u_new, ub * nlev, cudaMemcpyDeviceToDevice, h->stream));
    } else {
        WPS_CUDA(cudaMemcpyAsync(u, a.u_new, ub * nlev, cudaMemcpyDeviceToHost, h->stream));
        WPS_CUDA(cudaMemcpyAsync(v, a.v_new, vb * nlev, cudaMemcpyDeviceToHost, h->stream));
    }
    return end_call(h);
}

int wpsgpu_smooth(wpsgpu_handle_t h, int opt, int npass, float* array, int sx, int ex, int sy, int ey, int sz, int ez)
{
    WPS_CHECK(h && array, -1, "wpsgpu_smooth: NULL argument");
    WPS_CHECK(opt >= WPSGPU_ONETWOONE && opt <= WPSGPU_SMTHDESMTH_SPECIAL, -1, "unknown smooth option %d", opt);
    int nx = ex - sx + 1, ny = ey - sy + 1, nz = ez - sz + 1;
    size_t n = (size_t)nx * ny * nz, bytes = n * sizeof(float);
    if (npass <= 0) return 0;
    WPS_TRY(begin_call(h, 3 * bytes + 4096));
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) WPS_TRY(h->arena.reserve(2 * bytes + (1 << 20), h->stream));
    void* da;
    WPS_TRY(stage_inout(h, array, bytes, &da));
    float* a = (float*)da;
    float* s1 = (float*)h->arena.take(bytes);
    float* s2 = npass > 1 ? (float*)h->arena.take(bytes) : nullptr;
    WPS_CHECK(s1 != nullptr && (npass == 1 || s2 != nullptr), -3, "device arena exhausted");
    const bool special = opt == WPSGPU_SMTHDESMTH_SPECIAL;
    dim3 b(32, 8), g((nx + SM_T - 1) / SM_T, (ny + SM_T - 1) / SM_T, nz);
    // passes ping-pong a -> s1 -> s2 -> s1 ...; `a` itself stays untouched until the end, so it is the `_special` original.
    // The last pass of a multi-pass call writes straight back into `a` (each point's original is read by the thread that
    // overwrites it, no other thread reads `a` in that launch).
    const float* src = a;
    float* result = nullptr;
    for (int ipass = 0; ipass < npass; ++ipass) {
        const bool last = ipass == npass - 1;
        float* dst = (last && npass > 1) ? a : (src == s1 ? s2 : s1);
        const float* orig = (special && last && npass > 1) ? a : nullptr;
        if (opt == WPSGPU_ONETWOONE) k_smooth_pass<1><<<g, b, 0, h->stream>>>(src, dst, orig, nx, ny, 0);
        else k_smooth_pass<2><<<g, b, 0, h->stream>>>(src, dst, orig, nx, ny, special && last ? 1 : 0);
        h->launches++;
        src = dst;
        result = dst;
    }
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) {
        if (result != a) WPS_CUDA(cudaMemcpyAsync(a, result, bytes, cudaMemcpyDeviceToDevice, h->stream));   // single pass: s1 -> array
    } else WPS_TRY(copy_out(h, array, result, bytes));
    return end_call(h);
}

int wpsgpu_smooth_egrid(wpsgpu_handle_t h, int opt, int npass, float* array, int sx, int ex, int sy, int ey, int sz, int ez, float msgval, float hflag)
{
    WPS_CHECK(h && array, -1, "wpsgpu_smooth_egrid: NULL argument");
    WPS_CHECK(opt >= WPSGPU_ONETWOONE && opt <= WPSGPU_SMTHDESMTH_SPECIAL, -1, "unknown smooth option %d", opt);
    if (opt == WPSGPU_SMTHDESMTH_SPECIAL) return 0;     // 'smth-desmth_special is not currently implemented for NMM. No smoothing will be done.'
    const int nx = ex - sx + 1, ny = ey - sy + 1, nz = ez - sz + 1;
    WPS_CHECK(nx > 0 && ny > 0 && nz > 0, -1, "wpsgpu_smooth_egrid: bad dimensions");
    const size_t bytes = sizeof(float) * (size_t)nx * ny * nz;
    if (npass <= 0) return 0;
    WPS_TRY(begin_call(h, 2 * bytes + 4096));
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) WPS_TRY(h->arena.reserve(bytes + (1 << 20), h->stream));
    void* da;
    WPS_TRY(stage_inout(h, array, bytes, &da));
    float* a = (float*)da;
    float* sc = (float*)h->arena.take(bytes);
    WPS_CHECK(sc != nullptr, -3, "device arena exhausted");
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8, nz);
    const int h1 = hflag == 1.0f ? 1 : 0, m1 = msgval == 1.0f ? 1 : 0;
    for (int ipass = 0; ipass < npass; ++ipass) {
        // scratch starts from the array (the reference initialises level index 1 only, :291, :503; see orc_smooth_egrid)
        WPS_CUDA(cudaMemcpyAsync(sc, a, bytes, cudaMemcpyDeviceToDevice, h->stream));
        k_egrid_sweep<<<g, b, 0, h->stream>>>(a, sc, a, nx, ny, sy, h1, 1, 0.50f, 0.25f, 0, m1);
        k_egrid_sweep<<<g, b, 0, h->stream>>>(sc, a, a, nx, ny, sy, h1, 1, 0.50f, 0.25f, 1, m1);
        h->launches += 2;
        if (opt == WPSGPU_SMTHDESMTH) {
            k_egrid_sweep<<<g, b, 0, h->stream>>>(a, sc, a, nx, ny, sy, h1, 2, 1.52f, -0.26f, 0, m1);
            k_egrid_sweep<<<g, b, 0, h->stream>>>(sc, a, a, nx, ny, sy, h1, 2, 1.52f, -0.26f, 1, m1);
            h->launches += 2;
        }
    }
    WPS_TRY(copy_out(h, array, a, bytes));
    return end_call(h);
}

int wpsgpu_rotate_winds_nmm(wpsgpu_handle_t h, int idir, const wpsgpu_proj* proj, float* u, const int32_t* u_mask, float* v,
                            const int32_t* v_mask, int vs1, int vs2, int ve1, int ve2, const float* xlat_v, const float* xlon_v, int nlev)
{
    WPS_CHECK(h && proj && u && v && u_mask && v_mask && xlat_v && xlon_v, -1, "wpsgpu_rotate_winds_nmm: NULL argument");
    WPS_CHECK(proj->init, -1, "In metmap_xform_nmm(), uninitialized proj_info structure.");
    WPS_CHECK(idir == 1 || idir == -1, -1, "wpsgpu_rotate_winds_nmm: idir must be +1 or -1");
    if (!(proj->code == WPSGPU_PROJ_ROTLL || proj->code == WPSGPU_PROJ_LC || proj->code == WPSGPU_PROJ_PS || proj->code == WPSGPU_PROJ_CASSINI)) return 0;
    const int nx = ve1 - vs1 + 1, ny = ve2 - vs2 + 1, nxw = (nx + 31) / 32;
    WPS_CHECK(nx > 0 && ny > 0 && nlev > 0, -1, "wpsgpu_rotate_winds_nmm: bad dimensions");
    const size_t pb = sizeof(float) * (size_t)nx * ny, mb = sizeof(int) * (size_t)ny * nxw;
    WPS_TRY(begin_call(h, 2 * (pb + mb) * nlev + 2 * pb + 8192));
    void *du, *dv, *dum, *dvm, *dla, *dlo;
    WPS_TRY(stage_inout(h, u, pb * nlev, &du));
    WPS_TRY(stage_inout(h, v, pb * nlev, &dv));
    WPS_TRY(stage_in(h, u_mask, mb * nlev, &dum));
    WPS_TRY(stage_in(h, v_mask, mb * nlev, &dvm));
    WPS_TRY(stage_in(h, xlat_v, pb, &dla));
    WPS_TRY(stage_in(h, xlon_v, pb, &dlo));
    // every point reads and writes only its own U and V: in place
    k_rotate_nmm<<<dim3((nx + 31) / 32, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(proj->code, proj->lat1, proj->lon1, proj->stdlon, proj->cone, proj->hemi, idir,
                                                                                  (const float*)dla, (const float*)dlo, nx, ny, nlev, (float*)du,
                                                                                  (const int*)dum, (float*)dv, (const int*)dvm);
    h->launches++;
    WPS_TRY(copy_out(h, u, du, pb * nlev));
    WPS_TRY(copy_out(h, v, dv, pb * nlev));
    return end_call(h);
}

int wpsgpu_fractions_dominant(wpsgpu_handle_t h, float* field, int64_t npts, int min_cat, int max_cat,
                              float msg_fill_val, const int32_t* lm_vals, int n_lm, int is_water_mask,
                              float* landmask_out, float* dominant_out)
{
    WPS_CHECK(h && field, -1, "wpsgpu_fractions_dominant: NULL argument");
    WPS_CHECK(n_lm >= 0 && n_lm <= 32, -1, "too many landmask categories (%d)", n_lm);
    int ncat = max_cat - min_cat + 1;
    size_t fb = sizeof(float) * (size_t)npts * ncat, pb = sizeof(float) * (size_t)npts;
    WPS_TRY(begin_call(h, fb + 2 * pb + 4096));
    FracArgs a;
    void *df, *dl, *dd;
    WPS_TRY(stage_inout(h, field, fb, &df));
    WPS_TRY(alloc_out(h, landmask_out, pb, &dl));
    WPS_TRY(alloc_out(h, dominant_out, pb, &dd));
    a.field = (float*)df; a.npts = (size_t)npts; a.min_cat = min_cat; a.max_cat = max_cat; a.msg_fill = msg_fill_val;
    a.n_lm = n_lm; a.is_water_mask = is_water_mask; a.landmask = (float*)dl; a.dominant = (float*)dd;
    for (int k = 0; k < n_lm; ++k) a.lm_vals[k] = lm_vals[k];
    k_fractions_dominant<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>(a);
    h->launches++;
    WPS_TRY(copy_out(h, field, df, fb));
    WPS_TRY(copy_out(h, landmask_out, dl, pb));
    WPS_TRY(copy_out(h, dominant_out, dd, pb));
    return end_call(h);
}

int wpsgpu_xytoll_grid(wpsgpu_handle_t h, const wpsgpu_proj* dom, int stagger, int si, int sj, int ei, int ej,
                       float sub_x, float sub_y, float* xlat, float* xlon)
{
    WPS_CHECK(h && dom && xlat && xlon, -1, "wpsgpu_xytoll_grid: NULL argument");
    WPS_CHECK(dom->init, -1, "projection not initialised");
    int nx = ei - si + 1, ny = ej - sj + 1;
    size_t bytes = sizeof(float) * (size_t)nx * ny;
    WPS_TRY(begin_call(h, 2 * bytes + 4096 + sizeof(float) * WPSGPU_MAX_GAUSS));
    const float* dg;
    WPS_TRY(upload_gauss(h, *dom, &dg));
    void *d1, *d2;
    WPS_TRY(alloc_out(h, xlat, bytes, &d1));
    WPS_TRY(alloc_out(h, xlon, bytes, &d2));
    k_xytoll_grid<<<dim3((nx + 31) / 32, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(to_dev(*dom, dg), stagger, si, sj, nx, ny, sub_x, sub_y, (float*)d1, (float*)d2);
    h->launches++;
    WPS_TRY(copy_out(h, xlat, d1, bytes));
    WPS_TRY(copy_out(h, xlon, d2, bytes));
    return end_call(h);
}

int wpsgpu_map_factor(wpsgpu_handle_t h, int iproj, float truelat1, float truelat2, float pole_lat,
                      float pole_lon, float stand_lon, const float* xlat, const float* xlon, int64_t npts,
                      float* mapfac_x, float* mapfac_y)
{
    WPS_CHECK(h && xlat && xlon && mapfac_x && mapfac_y, -1, "wpsgpu_map_factor: NULL argument");
    size_t bytes = sizeof(float) * (size_t)npts;
    WPS_TRY(begin_call(h, 4 * bytes + 4096));
    MapfacArgs a;
    a.iproj = iproj; a.truelat1 = truelat1; a.truelat2 = truelat2; a.pole_lat = pole_lat; a.pole_lon = pole_lon; a.stand_lon = stand_lon;
    const float R = WPS_RAD_PER_DEG;
    a.n = 0.0f; a.colat0 = R * (90.0f - truelat1); a.colat2 = R * (90.0f - truelat2); a.s0 = 0.0f;
    if (iproj == WPSGPU_PROJ_LC && truelat1 != truelat2) {
        float colat1 = R * (90.0f - truelat1), colat2 = a.colat2;
        a.n = (t_log(t_sin(colat1)) - t_log(t_sin(colat2))) / (t_log(t_tan(colat1 / 2.0f)) - t_log(t_tan(colat2 / 2.0f)));
    }
    void *d1, *d2, *o1, *o2;
    WPS_TRY(stage_in(h, xlat, bytes, &d1));
    WPS_TRY(stage_in(h, xlon, bytes, &d2));
    WPS_TRY(alloc_out(h, mapfac_x, bytes, &o1));
    WPS_TRY(alloc_out(h, mapfac_y, bytes, &o2));
    k_map_factor<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>(a, (const float*)d1, (const float*)d2, (size_t)npts, (float*)o1, (float*)o2);
    h->launches++;
    WPS_TRY(copy_out(h, mapfac_x, o1, bytes));
    WPS_TRY(copy_out(h, mapfac_y, o2, bytes));
    return end_call(h);
}

int wpsgpu_coriolis(wpsgpu_handle_t h, const float* xlat, int64_t npts, float* f, float* e)
{
    WPS_CHECK(h && xlat && f && e, -1, "wpsgpu_coriolis: NULL argument");
    size_t bytes = sizeof(float) * (size_t)npts;
    WPS_TRY(begin_call(h, 3 * bytes + 4096));
    void *d1, *o1, *o2;
    WPS_TRY(stage_in(h, xlat, bytes, &d1));
    WPS_TRY(alloc_out(h, f, bytes, &o1));
    WPS_TRY(alloc_out(h, e, bytes, &o2));
    k_coriolis<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>((const float*)d1, (size_t)npts, (float*)o1, (float*)o2);
    h->launches++;
    WPS_TRY(copy_out(h, f, o1, bytes));
    WPS_TRY(copy_out(h, e, o2, bytes));
    return end_call(h);
}

int wpsgpu_rotang(wpsgpu_handle_t h, int iproj, float truelat1, float truelat2, float stand_lon,
                  const float* xlat, const float* xlon, int si, int sj, int ei, int ej, float* cosa, float* sina)
{
    WPS_CHECK(h && xlat && xlon && cosa && sina, -1, "wpsgpu_rotang: NULL argument");
    WPS_CHECK(iproj != WPSGPU_PROJ_ROTLL, -1, "The NMM rotated lat-lon grid is no longer supported. Rotation angles cannot be computed.");
    int nx = ei - si + 1, ny = ej - sj + 1;
    size_t bytes = sizeof(float) * (size_t)nx * ny;
    WPS_TRY(begin_call(h, 4 * bytes + 4096));
    float cone = (iproj == WPSGPU_PROJ_LC) ? lc_cone(truelat1, truelat2) : 1.0f;
    void *d1, *d2, *o1, *o2;
    WPS_TRY(stage_in(h, xlat, bytes, &d1));
    WPS_TRY(stage_in(h, xlon, bytes, &d2));
    WPS_TRY(alloc_out(h, cosa, bytes, &o1));
    WPS_TRY(alloc_out(h, sina, bytes, &o2));
    k_rotang<<<dim3((nx + 31) / 32, (ny + 7) / 8), dim3(32, 8), 0, h->stream>>>(iproj, cone, stand_lon, (const float*)d1, (const float*)d2, nx, ny, (float*)o1, (float*)o2);
    h->launches++;
    WPS_TRY(copy_out(h, cosa, o1, bytes));
    WPS_TRY(copy_out(h, sina, o2, bytes));
    return end_call(h);
}

static int dfd(wpsgpu_handle_t h, int dir, const float* src, float* dst, int si, int sj, int sk, int ei, int ej, int ek,
               int sr, const float* mapfac, float dom_d)
{
    WPS_CHECK(h && src && dst, -1, "wpsgpu_dfdx/dfdy: NULL argument");
    int nx = ei - si + 1, ny = ej - sj + 1, nz = ek - sk + 1;
    size_t b2 = sizeof(float) * (size_t)nx * ny, b3 = b2 * nz;
    WPS_TRY(begin_call(h, 2 * b3 + b2 + 4096));
    void *d1, *dm, *o1;
    WPS_TRY(stage_in(h, src, b3, &d1));
    WPS_TRY(stage_in(h, mapfac, b2, &dm));
    WPS_TRY(alloc_out(h, dst, b3, &o1));
    // DO-variable value after the interior loop terminates (0-based here)
    int stale = (dir == 0) ? ((nx - 2 >= 1) ? nx - 1 : 1) : ((ny - 2 >= 1) ? ny - 1 : 1);
    k_dfd<<<dim3((nx + 31) / 32, (ny + 7) / 8, nz), dim3(32, 8), 0, h->stream>>>(dir, (const float*)d1, (float*)o1, nx, ny, (float)sr, (const float*)dm, dom_d, stale);
    h->launches++;
    WPS_TRY(copy_out(h, dst, o1, b3));
    return end_call(h);
}
int wpsgpu_dfdx(wpsgpu_handle_t h, const float* src, float* dst, int si, int sj, int sk, int ei, int ej, int ek,
                int sr_x, const float* mapfac, float dom_dx)
{ return dfd(h, 0, src, dst, si, sj, sk, ei, ej, ek, sr_x, mapfac, dom_dx); }
int wpsgpu_dfdy(wpsgpu_handle_t h, const float* src, float* dst, int si, int sj, int sk, int ei, int ej, int ek,
                int sr_y, const float* mapfac, float dom_dy)
{ return dfd(h, 1, src, dst, si, sj, sk, ei, ej, ek, sr_y, mapfac, dom_dy); }

}  // extern "C"
