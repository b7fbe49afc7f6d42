// strip16.cuh -- the tolerance-mode sixteen_pt path (WPSGPU_PRECISION_TOLERANCE): window records + strip kernel.
// Included by metgrid.cu after Job / SlabPtrs.
//
// sixteen_pt (interp_module.F:1179-1374) is linear in the 16 stencil values whenever `oned` takes its general branch
// in all five calls: oned(x,a,b,c,d) = wa(x)*a + wb(x)*b + wc(x)*c + wd(x)*d with
//     wa = -u*u*x/2,  wb = u*(1-x*x) + x*u*(1+u)/2,  wc = u*x*(1+x)/2 + x*(1-u*u),  wd = -x*x*u/2,   u = 1-x
// (expand :1370).  The weights are level-independent, so a destination point costs 4 x 4 + 4 fused multiply-adds per
// level instead of the 67 separately rounded operations of the bit-exact kernel.  north_star allows 1e-5 relative
// for sixteen_pt; results agree with the reference's nested evaluation to a few ulp of the largest stencil value.
//
// A thread owns a vertical STRIP of R destination points (same x, R consecutive y): with a source grid much coarser
// than the domain their stencils overlap almost completely, so the thread loads ONE 5 x 5 window of source cells per
// pack of four levels (float4 = 4 levels) and evaluates all R points from registers -- 25 vector loads serve 4*R
// point-levels (k_metgrid_fast16: 24 per 4), which takes the kernel off the L1 return path.  A point whose cell is
// (i0+di, j0+dj), di, dj in {0, 1}, relative to the window origin gets its four weights shifted by (di, dj) inside
// five window weights (one of them 0).
//
// What the linear form cannot reproduce is decided per window by the pre-pass and per point-level by a guard:
//  * a missing or interp-masked value, a NaN/Inf, or a value so small that `b*c`, `a*d` could underflow to 0 in the
//    x direction (|v| < 1e-21) anywhere in the window  -> threshold +Inf, the points go to the exact slow path;
//  * oned's guards in the y direction act on the four row interpolants: a row result within 2^-18 of the window's
//    largest magnitude of zero (the reference's own value may then be exactly 0, or the products may underflow)
//    -> slow path.  Everything else provably takes the general branch in the reference as well.
// Zeros are replaced by 1.E-20 in the window records exactly as :1255-1257 does, and a result equal to 1.E-20 becomes 0
// (:1317).
#pragma once
#include <type_traits>

namespace wps {

constexpr int STRIP_MAX_JOBS = 24;
constexpr int STRIP_MAX_SLABS = 320;
constexpr int STRIP_MAX_SEGS = 128;
constexpr int STRIP_MAX_WORK = 96;

struct StripJob {
    const float2* xy;
    const float* landmask;
    const float4* P;             // window records [npack][pny][pnx], float4 = 4 levels of source cell (px0 + c, py0 + r), clamped like :1227-1241
    const float4* TA;            // guard thresholds per window ORIGIN cell, same index space
    const unsigned* cf;          // per (pack, cell): flag bytes of the four stencils a strip with this window origin can use (k_pack_win)
    int nx, ny, nxw, sm1, sm2, sd1, ed1, sd2, ed2, pw;
    int minx, maxx, miny, maxy, bdr;
    int px0, py0, pnx, pny;
    int nslab, slab0, gslab0;    // slabs [slab0, slab0 + nslab) of the pointer tables below; gslab0 = first slab in the pass's slab table
    int tiles_x, ntiles;
    int both, has_lm;
    float msg, fill, masked;
};

struct StripSeg { short job, pack0, npk, pad; };       // consecutive packs of one job
struct StripWork { short seg0, nseg; };                // what one block does: segments [seg0, seg0 + nseg), same geometry

struct StripParams {
    StripJob job[STRIP_MAX_JOBS];
    StripSeg seg[STRIP_MAX_SEGS];
    StripWork work[STRIP_MAX_WORK];
    float* r_arr[STRIP_MAX_SLABS];
    int* new_pts[STRIP_MAX_SLABS];
    SlowQ sq;
};

// ---- pre-pass: window records, guard thresholds and the slow pass's hints ------------------------------------------------
// One block = 16 x 16 cells of the window array of one (job, pack).  The 20 x 20 halo tile of the four levels goes
// through shared memory once; every thread then writes its cell's record and derives the threshold of the 5 x 5 window
// whose ORIGIN is that cell (cells c-1..c+3, r-1..r+3) and the hint bits of the 4 x 4 stencil anchored there.
struct PackWinTask { int job, pack; };

__global__ void __launch_bounds__(256) k_pack_win(const Job* __restrict__ jobs, const PackWinTask* __restrict__ tasks,
                                                  const SlabPtrs* __restrict__ slabs)
{
    const PackWinTask t = tasks[blockIdx.y];
    const Job& jb = jobs[t.job];
    const int pnx = jb.bnx, pny = jb.bny, ntx = (pnx + 15) >> 4;
    const int bx = blockIdx.x % ntx, by = blockIdx.x / ntx;
    if (by * 16 >= pny) return;
    // tile of 20 x 20 cells: the block's 16 x 16 cells with one cell of halo to the left / top and three to the right / bottom
    __shared__ float s_v[4][20][21];           // record values (zeros replaced, unusable values 0)
    __shared__ float s_a[4][20][21];           // usable magnitude per level (0 where missing / non-finite); later its 5-wide row maximum
    __shared__ unsigned short s_f[20][20];     // per cell: bit lv = missing value, bit 4+lv = unusable for sixteen_pt, bit 8 / 9 = masked under A / B
    __shared__ unsigned short s_h[20][20];     // OR of s_f over the four cells c .. c+3 of the row
    const int tid = threadIdx.x;
    const int sx = jb.minx, ex = jb.maxx, sy = jb.miny, ey = jb.maxy, snx = ex - sx + 1;
    const float msg = jb.msg;
    const bool both = jb.landmask != nullptr && jb.masked == WPSGPU_MASKED_BOTH;
    const int selA = both ? 1 : 0;
    const float* src[4];
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
        const int sl = t.pack * 4 + lv;
        src[lv] = sl < jb.nslab ? slabs[jb.slab0 + sl].src : nullptr;
    }
    for (int e = tid; e < 400; e += 256) {
        const int c = e % 20, r = e / 20;
        const int gx = min(max(jb.bx0 + bx * 16 - 1 + c, sx), ex), gy = min(max(jb.by0 + by * 16 - 1 + r, sy), ey);   // clamped like :1227-1241
        const size_t off = (size_t)(gy - sy) * snx + (gx - sx);
        unsigned f = 0;
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            float v = 0.0f, a = 0.0f;
            if (src[lv]) {
                v = __ldg(src[lv] + off);
                if (v == 0.0f && msg != 0.0f) v = 1.0e-20f;     // interp_module.F:1255-1257
                a = fabsf(v);
                const bool ismsg = v == msg, fin = a <= 3.0e38f;
                if (ismsg) f |= 1u << lv;
                if (ismsg || !fin || !(a >= 1.0e-21f)) f |= 16u << lv;
                // a non-finite or missing value meets zero weights in the neighbouring stencils of a window: 0 keeps those
                // sums finite (its own stencils are flagged and never use the record)
                if (ismsg || !fin) { v = 0.0f; a = 0.0f; }
            }
            s_v[lv][r][c] = v;
            s_a[lv][r][c] = a;
        }
        auto unusable = [&](int sel) -> bool {
            if (!jb.mask[sel]) return false;
            const float mv = __ldg(jb.mask[sel] + off);
            return jb.mask_rel[sel] == REL_GT ? (mv > jb.mask_val[sel]) : (jb.mask_rel[sel] == REL_LT ? (mv < jb.mask_val[sel]) : (mv == jb.mask_val[sel]));
        };
        if (unusable(selA)) f |= 256u;
        if (both && unusable(2)) f |= 512u;
        s_f[r][c] = (unsigned short)f;
    }
    __syncthreads();
    // row pass: flags of the 4 cells c .. c+3 and the largest usable magnitude of the 5 cells c .. c+4 (clipped to the tile)
    float hm[2][4];
    unsigned hf[2];
    for (int e = tid, k = 0; e < 400; e += 256, ++k) {
        const int c = e % 20, r = e / 20;
        unsigned f = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) if (c + d < 20) f |= s_f[r][c + d];
        hf[k] = f;
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            float m = 0.0f;
#pragma unroll
            for (int d = 0; d < 5; ++d) if (c + d < 20) m = fmaxf(m, s_a[lv][r][c + d]);
            hm[k][lv] = m;
        }
    }
    __syncthreads();
    for (int e = tid, k = 0; e < 400; e += 256, ++k) {
        const int c = e % 20, r = e / 20;
        s_h[r][c] = (unsigned short)hf[k];
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) s_a[lv][r][c] = hm[k][lv];
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    const int cx = bx * 16 + tx, cy = by * 16 + ty;
    if (cx >= pnx || cy >= pny) return;
    const int c0 = tx + 1, r0 = ty + 1;        // the thread's cell in the tile
    const size_t o = ((size_t)t.pack * pny + cy) * pnx + cx;
    const_cast<float4*>(jb.rec)[o] = make_float4(s_v[0][r0][c0], s_v[1][r0][c0], s_v[2][r0][c0], s_v[3][r0][c0]);
    // Flag byte of the 4 x 4 stencil anchored at a cell (cells c-1..c+2, r-1..r+2), per level lv:
    //   bit lv      stencil is clean under mask A: no missing / masked / non-finite value, every |v| >= 1e-21
    //   bit 4 + lv  single-mask jobs: the stencil holds a missing or masked value, so sixteen_pt is certain to hand over
    //               to the next method (:1244-1270) and the slow pass may start the chain there;
    //               masked=both jobs: stencil is clean under mask B (interp_water_mask, land points)
    // The cell's flag WORD holds the bytes of the stencils anchored at (c, r), (c+1, r), (c, r+1), (c+1, r+1): the four
    // cells the points of a strip whose window origin is (c, r) can fall into -- one load per thread and pack.
    unsigned present = 0;
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) if (src[lv]) present |= 1u << lv;
    auto stencil_byte = [&](int ax, int ay) -> unsigned {
        unsigned f = 0;
#pragma unroll
        for (int r = -1; r <= 2; ++r) f |= s_h[r0 + ay + r][c0 + ax - 1];
        const unsigned anymsg = f & 15u, bad = (f >> 4) & 15u, mA = (f >> 8) & 1u, mB = (f >> 9) & 1u;
        unsigned bits = mA ? 0u : (~bad & 15u);
        if (both) bits |= mB ? 0u : ((~bad & 15u) << 4);
        else bits |= (mA ? 15u : anymsg) << 4;
        // levels beyond the job's last: records are 0, never flagged
        return (bits & (present | (present << 4))) | ((~present & 15u) * 0x11u);
    };
    // threshold of the 5 x 5 window whose origin is the cell: 2^-18 of the largest usable magnitude in it (an upper bound
    // for every stencil inside the window), at least 1e-21
    float thr[4];
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
        float mx = 0.0f;
#pragma unroll
        for (int r = -1; r <= 3; ++r) mx = fmaxf(mx, s_a[lv][r0 + r][c0 - 1]);
        thr[lv] = src[lv] ? fmaxf(mx * 3.814697265625e-6f, 1.0e-21f) : -1.0f;
        // a weighted sum stays below 1.3 x the largest magnitude: if that cannot reach |missing_value| the kernel need not
        // compare its results with missing_value (it does when missing_value is 0); otherwise the window goes to the slow path
        if (src[lv] && msg != 0.0f && !(mx * 2.0f < fabsf(msg))) thr[lv] = __int_as_float(0x7f800000);
    }
    const_cast<float4*>(jb.thr)[o] = make_float4(thr[0], thr[1], thr[2], thr[3]);
    const_cast<unsigned*>(jb.cfw)[o] = stencil_byte(0, 0) | (stencil_byte(1, 0) << 8) | (stencil_byte(0, 1) << 16) | (stencil_byte(1, 1) << 24);
}

// ---- strip kernel --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// predicated stores: the compiler otherwise wraps every `if (ok) store` around a ballot in a branch + reconvergence pair
__device__ __forceinline__ void st_f32_if(bool p, void* a, float v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.f32 [%1], %2; }" ::"r"((unsigned)p), "l"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void st_u32_if(bool p, void* a, unsigned v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.u32 [%1], %2; }" ::"r"((unsigned)p), "l"(a), "r"(v) : "memory");
}

// mbarrier + bulk-copy (TMA engine, SASS UBLKCP) primitives of the window ring
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Window ring of the staged variant: each warp (= block) copies the bounding box of its lanes' windows for pack p+2 into
// shared memory with one bulk copy per box row while pack p is computed -- the north-star design ("each thread block
// TMA-stages the compact source-grid bounding box its destination tile maps onto").  Box: up to 16 x 8 cells of 16 B.
constexpr int SBW = 16, SBH = 8, SROW = SBW + 1, SNS = 3;     // row stride 17 cells: rows start in different banks

#ifndef STRIP_MINB
#define STRIP_MINB 16      // resident warps (= blocks) per SM the R = 4 kernel is compiled for: 128 registers (12: 0.921 ms, 20: 0.996 ms, 16: 0.907 ms)
#endif

// R points per thread, WIN = window edge (5 when R > 1: cells of a strip may differ by one in each direction; 4 for R = 1).
// One warp per block.  STAGE: windows come from the shared-memory ring (bulk copies), else straight from L1/L2.
// A block works through a list of SEGMENTS (consecutive packs of one job); all jobs of the list share the destination grid,
// the cached coordinates and the record-array geometry, so the level-independent set-up -- cells, weights, window box --
// is done once per block and only the landmask / fill logic is redone per job.
template <int R, bool STAGE>
__global__ void __launch_bounds__(32, R == 4 ? STRIP_MINB : 20)
k_metgrid_strip16(const __grid_constant__ StripParams P)
{
    constexpr int WIN = R > 1 ? 5 : 4;
    const StripWork wk = P.work[blockIdx.y];
    const StripJob& gj = P.job[P.seg[wk.seg0].job];      // geometry: identical for every job of the work item
    if ((int)blockIdx.x >= gj.ntiles) return;
    const int lane = threadIdx.x;
    const int tx = blockIdx.x % gj.tiles_x, ty = blockIdx.x / gj.tiles_x;
    const int nx = gj.nx, ny = gj.ny;
    const int ii = tx * 32 + lane, jj0 = ty * R;

    // ---- geometry: sixteen_pt's cell arithmetic (interp_module.F:1211-1225) for the strip's points
    float2 xyv[R];
#pragma unroll
    for (int p = 0; p < R; ++p) xyv[p] = __ldg(gj.xy + (size_t)min(jj0 + p, ny - 1) * nx + min(ii, nx - 1));   // rows / columns past the grid re-read its last point
    bool geo[R];                                 // the point can be evaluated from a window (if the job's point logic allows)
    int ci[R], cj[R];
    float fx[R], fy[R];
    int i0 = INT_MAX, j0 = INT_MAX;
    const float lo = (float)(gj.minx + gj.bdr) - 0.5f, hi = (float)(gj.maxx - gj.bdr) + 0.5f;
#pragma unroll
    for (int p = 0; p < R; ++p) {
        geo[p] = false; ci[p] = 0; cj[p] = 0; fx[p] = 0.0f; fy[p] = 0.0f;
        const float xx = xyv[p].x, yy = xyv[p].y;
        if (ii < nx && jj0 + p < ny && xx >= lo && xx <= hi && fabsf(xx) < 1.0e9f && fabsf(yy) < 1.0e9f &&
            (int)xx >= gj.minx && (int)xx <= gj.maxx && (int)yy >= gj.miny && (int)yy <= gj.maxy) {
            const int i = (int)(xx + 0.00001f), j = (int)(yy + 0.00001f);
            const float x = xx - (float)i, y = yy - (float)j;
            // a window with origin (i, j) must lie inside the record array, whatever the other points of the strip do
            if ((fabsf(x) > 0.0001f || fabsf(y) > 0.0001f) && i - 1 >= gj.px0 && i + WIN - 2 < gj.px0 + gj.pnx &&
                j - 1 >= gj.py0 && j + WIN - 2 < gj.py0 + gj.pny) {
                geo[p] = true; ci[p] = i; cj[p] = j; fx[p] = x; fy[p] = y;
                i0 = min(i0, i); j0 = min(j0, j);
            }
        }
    }
    const bool anygeo = i0 != INT_MAX;
    if (!anygeo) { i0 = gj.px0 + 1; j0 = gj.py0 + 1; }
    // window weights: the point's four oned weights at offset (di, dj) inside the window (x weights are used as the
    // scalar operand of the packed level-pair arithmetic); R = 4 keeps the y weights in shared memory (20 registers less)
    float2 wx[R][WIN];
    float wy[R][WIN];
    __shared__ float4 s_wy[R == 4 ? WIN * 32 : 1];
    unsigned geom = 0;                           // bit p: geo[p]
    unsigned cfsh0 = 0;                          // 5 bits per point: shift of the point's byte in the flag word, 8 * (dx + 2 dy)
#pragma unroll
    for (int p = 0; p < R; ++p) {
        const int di = ci[p] - i0, dj = cj[p] - j0;
        if (geo[p] && (di > WIN - 4 || dj > WIN - 4)) geo[p] = false;     // strip spans more than two cells: slow path
        if (geo[p]) { geom |= 1u << p; cfsh0 |= (8u * (unsigned)(di + 2 * dj)) << (5 * p); }
        float w4x[4], w4y[4];
        {
            const float x = fx[p], u = 1.0f - x;
            w4x[0] = -0.5f * u * u * x; w4x[3] = -0.5f * x * x * u;
            w4x[1] = u * (1.0f - x * x) + 0.5f * x * u * (1.0f + u);
            w4x[2] = 0.5f * u * x * (1.0f + x) + x * (1.0f - u * u);
            const float y = fy[p], v = 1.0f - y;
            w4y[0] = -0.5f * v * v * y; w4y[3] = -0.5f * y * y * v;
            w4y[1] = v * (1.0f - y * y) + 0.5f * y * v * (1.0f + v);
            w4y[2] = 0.5f * v * y * (1.0f + y) + y * (1.0f - v * v);
        }
        if (!geo[p]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { w4x[k] = 0.0f; w4y[k] = 0.0f; }
        }
        if (WIN == 5) {
            const bool sx_ = geo[p] && di == 1, sy_ = geo[p] && dj == 1;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float a = k < 4 ? w4x[k] : 0.0f, b = k > 0 ? w4x[k - 1] : 0.0f;
                const float c = k < 4 ? w4y[k] : 0.0f, d = k > 0 ? w4y[k - 1] : 0.0f;
                const float wxv = sx_ ? b : a;
                wx[p][k] = make_float2(wxv, wxv);
                wy[p][k] = sy_ ? d : c;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { wx[p][k] = make_float2(w4x[k], w4x[k]); wy[p][k] = w4y[k]; }
        }
    }
    if (R == 4) {
#pragma unroll
        for (int l = 0; l < WIN; ++l) s_wy[l * 32 + lane] = make_float4(wy[0][l], wy[1 % R][l], wy[2 % R][l], wy[3 % R][l]);
        __syncwarp();
    }
    const int pnx = gj.pnx;
    const size_t pstride = (size_t)gj.pny * pnx;
    const size_t toff = (size_t)(j0 - gj.py0) * pnx + (i0 - gj.px0);                   // window origin cell: thresholds, flag words
    const size_t woff = (size_t)(j0 - 1 - gj.py0) * pnx + (i0 - 1 - gj.px0);           // window's first cell: records

    // ---- window source: shared-memory ring fed by bulk copies, or global memory directly
    __shared__ alignas(128) float4 s_win[STAGE ? SNS * SBH * SROW : 1];
    __shared__ alignas(8) unsigned long long s_bar[STAGE ? SNS : 1];
    bool staged = false;
    int bw = 0, bh = 0;
    size_t boff = 0;
    const float4* sw_ = s_win;
    if (STAGE) {
        // bounding box of the warp's windows (lanes without a usable point do not count)
        const int bi0 = __reduce_min_sync(0xffffffffu, anygeo ? i0 - 1 : INT_MAX);
        const int bj0 = __reduce_min_sync(0xffffffffu, anygeo ? j0 - 1 : INT_MAX);
        const int bi1 = __reduce_max_sync(0xffffffffu, anygeo ? i0 + WIN - 2 : INT_MIN);
        const int bj1 = __reduce_max_sync(0xffffffffu, anygeo ? j0 + WIN - 2 : INT_MIN);
        bw = bi1 - bi0 + 1; bh = bj1 - bj0 + 1;
        staged = bi0 != INT_MAX && bw <= SBW && bh <= SBH;
        if (staged) {
            boff = (size_t)(bj0 - gj.py0 + (lane < bh ? lane : 0)) * pnx + (bi0 - gj.px0);    // lane r copies box row r
            if (anygeo) sw_ = s_win + (j0 - 1 - bj0) * SROW + (i0 - 1 - bi0);
        }
    }
    const unsigned bar0 = smem_u32(s_bar), win0 = smem_u32(s_win);
    if (STAGE && staged) {
        if (lane == 0) {
#pragma unroll
            for (int st = 0; st < SNS; ++st) mbar_init(bar0 + 8 * st, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // Output addressing: byte offsets of the strip's first point / bitmap word; the per-level base pointers are warp-uniform
    // (constant bank).  Levels past a job's last one have their pointers aimed at a dump area by the host, so the pack loop
    // has no level test at all; stores are predicated, not branched around.
    const unsigned colw = __ballot_sync(0xffffffffu, ii < nx);   // lanes inside the grid
    const unsigned rowsm = (1u << min(R, ny - jj0)) - 1u;     // rows of the strip that exist
    const unsigned poff0 = (unsigned)(((size_t)jj0 * nx + ii) * sizeof(float)), prow = (unsigned)(nx * sizeof(float));
    const unsigned woff0 = (unsigned)(((size_t)jj0 * gj.nxw + tx) * sizeof(int)), wrow = (unsigned)(gj.nxw * sizeof(int));
    int it0 = 0;                                 // packs this block has taken through the ring so far (stage / phase bookkeeping)

    for (int sg = 0; sg < wk.nseg; ++sg) {
        const StripSeg seg = P.seg[wk.seg0 + sg];
        const StripJob& jb = P.job[seg.job];
        const bool msg0 = jb.msg == 0.0f;
        // ---- the job's point logic (interp_met_field :2490-2541): process_bdy_width interior and landmask == masked
        // receive fill_missing; masked=both selects the interp mask (flag nibble) by the landmask
        unsigned fillm = 0, fastm = geom, cfsh = cfsh0;
        {
            const int pw = jb.pw, di = jb.sm1 + ii;
            const bool fill_i = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw;
#pragma unroll
            for (int p = 0; p < R; ++p) {
                const int jj = jj0 + p, dj = jb.sm2 + jj;
                const bool act = ii < nx && jj < ny;
                const float lm = jb.has_lm ? __ldg(jb.landmask + (size_t)min(jj, ny - 1) * nx + min(ii, nx - 1)) : 0.0f;
                bool fill = fill_i && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw, lm_ok = true;
                if (jb.both) {
                    if (lm == 1.0f) cfsh += 4u << (5 * p);      // land point: interp_water_mask, flag bits 4-7
                    lm_ok = lm == 0.0f || lm == 1.0f;
                } else fill = fill || (jb.has_lm && lm == jb.masked);
                if (fill && act) fillm |= 1u << p;
                if (fill || !lm_ok) fastm &= ~(1u << p);
            }
        }
        const bool anyfill = __any_sync(0xffffffffu, fillm != 0u);
        const float fillv = jb.fill;
        const float4* __restrict__ Ta = jb.TA + (size_t)seg.pack0 * pstride + toff;
        const unsigned* __restrict__ cfp = jb.cf + (size_t)seg.pack0 * pstride + toff;
        const float4* __restrict__ Pw = jb.P + (size_t)seg.pack0 * pstride + woff;
        const float4* Pbox = jb.P + (size_t)seg.pack0 * pstride + boff;
        const int npk = seg.npk;
        const int slab00 = jb.slab0 + seg.pack0 * 4;
        auto issue = [&](int pk) {      // bulk copies of the segment's pack pk box rows into the next ring stage
            const int st = (it0 + pk) % SNS;
            if (lane == 0) mbar_expect_tx(bar0 + 8 * st, (unsigned)(bw * bh * 16));
            __syncwarp();
            if (lane < bh) bulk_g2s(win0 + (unsigned)((st * SBH + lane) * SROW * 16), Pbox + (size_t)pk * pstride, (unsigned)(bw * 16), bar0 + 8 * st);
        };
        if (STAGE && staged) {
            __syncwarp();                            // the previous segment's last stages have been read by every lane
            for (int pk = 0; pk < SNS - 1 && pk < npk; ++pk) issue(pk);
        }
        float4 ta = __ldg(Ta);
        unsigned cfv = __ldg(cfp);                   // flag word of the window origin for the pack

        // the pack loop exists twice (windows from the ring / from global memory) so that neither carries the other's loads
        auto packs = [&](auto direct_tag) {
            constexpr bool DIRECT = decltype(direct_tag)::value;
            for (int pk = 0; pk < npk; ++pk) {
                const float4 cta = ta;
                const unsigned ccf = cfv;
                if (pk + 1 < npk) {                      // thresholds and stencil flags of the next pack, one iteration ahead
                    Ta += pstride; cfp += pstride;
                    ta = __ldg(Ta);
                    cfv = __ldg(cfp);
                }
                const int stg = (it0 + pk) % SNS;
                if (!DIRECT) {
                    __syncwarp();                        // every lane has finished reading the stage that is refilled now
                    if (pk + SNS - 1 < npk) issue(pk + SNS - 1);
                    mbar_wait(bar0 + 8 * stg, (unsigned)(((it0 + pk) / SNS) & 1));
                }
                const float4* __restrict__ sw = sw_ + stg * SBH * SROW;
                float2 a01[R], a23[R];
                bool g[R];                               // the guard holds for every row and level of point p so far
#pragma unroll
                for (int p = 0; p < R; ++p) {
                    a01[p] = make_float2(0.0f, 0.0f); a23[p] = a01[p];
                    // the point's stencil must be clean for all four levels (under the mask the landmask selects, masked=both)
                    g[p] = ((fastm >> p) & 1u) && ((ccf >> ((cfsh >> (5 * p)) & 31u)) & 0xfu) == 0xfu;
                }
#pragma unroll
                for (int l = 0; l < WIN; ++l) {
                    float4 v[WIN];
#pragma unroll
                    for (int k = 0; k < WIN; ++k) v[k] = DIRECT ? __ldg(Pw + l * pnx + k) : sw[l * SROW + k];
                    float wyr[R];
                    if (R == 4) {
                        const float4 w4 = s_wy[l * 32 + lane];
                        wyr[0] = w4.x; wyr[1 % R] = w4.y; wyr[2 % R] = w4.z; wyr[3 % R] = w4.w;
                    } else {
#pragma unroll
                        for (int p = 0; p < R; ++p) wyr[p] = wy[p][l];
                    }
#pragma unroll
                    for (int p = 0; p < R; ++p) {
                        float2 r01 = __fmul2_rn(wx[p][0], lo2(v[0])), r23 = __fmul2_rn(wx[p][0], hi2(v[0]));
#pragma unroll
                        for (int k = 1; k < WIN; ++k) {
                            r01 = __ffma2_rn(wx[p][k], lo2(v[k]), r01);
                            r23 = __ffma2_rn(wx[p][k], hi2(v[k]), r23);
                        }
                        g[p] = g[p] && fabsf(r01.x) > cta.x && fabsf(r01.y) > cta.y && fabsf(r23.x) > cta.z && fabsf(r23.y) > cta.w;
                        const float2 wyl = make_float2(wyr[p], wyr[p]);
                        a01[p] = __ffma2_rn(wyl, r01, a01[p]);
                        a23[p] = __ffma2_rn(wyl, r23, a23[p]);
                    }
                }
                if (DIRECT) Pw += pstride;
                // ---- outputs.  The guard is joint over the pack's four levels: a point either delivers all of them or hands
                // all of them to the slow pass, so one ballot per point serves the pack.
                const int sl0 = slab00 + pk * 4;
                const bool standin = __any_sync(0xffffffffu, cta.x == 1.0e-21f || cta.y == 1.0e-21f || cta.z == 1.0e-21f || cta.w == 1.0e-21f);
                unsigned word[R];
#pragma unroll
                for (int p = 0; p < R; ++p) {
                    float r0 = a01[p].x, r1 = a01[p].y, r2 = a23[p].x, r3 = a23[p].y;
                    bool ok = g[p];
                    // a result equal to the zero stand-in becomes 0 (:1317): possible only where a whole window consists of
                    // stand-ins, i.e. its threshold sits at the 1e-21 floor (warp-uniform test)
                    if (standin) {
                        r0 = r0 == 1.0e-20f ? 0.0f : r0; r1 = r1 == 1.0e-20f ? 0.0f : r1;
                        r2 = r2 == 1.0e-20f ? 0.0f : r2; r3 = r3 == 1.0e-20f ? 0.0f : r3;
                    }
                    // "value /= missing_value" (:2558): only a missing_value of 0 can be met by a weighted sum of usable
                    // values (the pre-pass sends windows whose magnitudes come near |missing_value| to the slow path)
                    if (msg0) ok = ok && r0 != 0.0f && r1 != 0.0f && r2 != 0.0f && r3 != 0.0f;
                    if (anyfill) {                       // warp-uniform: fill_missing points exist in this strip row group
                        const bool f = (fillm >> p) & 1u;
                        r0 = f ? fillv : r0; r1 = f ? fillv : r1; r2 = f ? fillv : r2; r3 = f ? fillv : r3;
                        ok = ok || f;
                    }
                    // (level base + p rows) is warp-uniform arithmetic; the lane adds its own offset
                    st_f32_if(ok, reinterpret_cast<char*>(P.r_arr[sl0 + 0]) + (size_t)p * prow + poff0, r0);
                    st_f32_if(ok, reinterpret_cast<char*>(P.r_arr[sl0 + 1]) + (size_t)p * prow + poff0, r1);
                    st_f32_if(ok, reinterpret_cast<char*>(P.r_arr[sl0 + 2]) + (size_t)p * prow + poff0, r2);
                    st_f32_if(ok, reinterpret_cast<char*>(P.r_arr[sl0 + 3]) + (size_t)p * prow + poff0, r3);
                    word[p] = __ballot_sync(0xffffffffu, ok);
                }
                {
                    // flagged points (rare): one warp-uniform test per pack, the queue code stays out of the way
                    unsigned any = 0;
#pragma unroll
                    for (int p = 0; p < R; ++p) any |= (((rowsm >> p) & 1u) ? colw : 0u) & ~word[p];
                    if (any) {
                        const int nl = min(4, jb.nslab - (seg.pack0 + pk) * 4);
#pragma unroll 1
                        for (int p = 0; p < R; ++p) {
                            const unsigned sword = (((rowsm >> p) & 1u) ? colw : 0u) & ~word[p];
                            if (sword) slowq_push(P.sq, sword, lane, jb.gslab0 + (seg.pack0 + pk) * 4, nl, (int)((poff0 >> 2) + p * nx));
                        }
                    }
                }
                if (lane < R) {
                    // lane p writes the new_pts words of strip row p for the four levels
                    unsigned w = word[0];
#pragma unroll
                    for (int p = 1; p < R; ++p) if (lane == p) w = word[p];
                    const unsigned wo = woff0 + (unsigned)lane * wrow;
                    const bool rowok = (rowsm >> lane) & 1u;
                    st_u32_if(rowok, reinterpret_cast<char*>(P.new_pts[sl0 + 0]) + wo, w);
                    st_u32_if(rowok, reinterpret_cast<char*>(P.new_pts[sl0 + 1]) + wo, w);
                    st_u32_if(rowok, reinterpret_cast<char*>(P.new_pts[sl0 + 2]) + wo, w);
                    st_u32_if(rowok, reinterpret_cast<char*>(P.new_pts[sl0 + 3]) + wo, w);
                }
            }
        };
        if (STAGE && staged) packs(std::false_type{});
        else packs(std::true_type{});
        it0 += npk;
    }
}

}  // namespace wps
