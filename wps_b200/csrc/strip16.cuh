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
#include <cuda.h>
#include <type_traits>

namespace wps {

constexpr int STRIP_MAX_JOBS = 24;
constexpr int STRIP_MAX_SLABS = 320;
constexpr int STRIP_MAX_SEGS = 128;
constexpr int STRIP_MAX_WORK = 96;

struct StripJob {
    const float2* xy;
    const float* landmask;
    const float* landmask_cls;   // the landmask of the grid if ANY job of the geometry class uses one (loaded once per block)
    const float4* P;             // window records [npack][pny][pnx], float4 = 4 levels of source cell (px0 + c, py0 + r), clamped like :1227-1241
    const uint2* TF;             // per (pack, window ORIGIN cell): .x = flag bytes of the four stencils a strip with this window origin can use,
                                 // .y = guard threshold of the 5 x 5 window, sign bit = a level of the window holds zero stand-ins only (k_pack_win)
    int nx, ny, nxw, sm1, sm2, sd1, ed1, sd2, ed2, pw;
    int minx, maxx, miny, maxy, bdr;
    int px0, py0, pnx, pny;
    int nslab, slab0, gslab0;    // slabs [slab0, slab0 + nslab) of the pointer tables below; gslab0 = first slab in the pass's slab table
    int tiles_x, ntiles;
    int both, has_lm;
    float msg, fill, masked;
};

struct StripOut { float* r; int* w; };                 // output plane and new_pts bitmap of one slab (level)
struct StripSeg { short job, pack0, npk, same; };      // consecutive packs of one job; same: point logic (fill, landmask use) as the segment before
// what one block does: segments [seg0, seg0 + nseg) of one geometry class = packs [q0, q0 + npk) of the class arrays
struct StripWork { short seg0, nseg, cls, npk; int q0; };
// window records / thresholds / flag words of a geometry class: the arrays of its jobs back to back, [class packs][pny][pnx]
struct StripClass { const float4* P; const uint2* TF; };
constexpr int STRIP_MAX_CLS = 8;

struct StripParams {
    StripJob job[STRIP_MAX_JOBS];
    StripSeg seg[STRIP_MAX_SEGS];
    StripWork work[STRIP_MAX_WORK];
    StripOut out[STRIP_MAX_SLABS];
    SlowQ sq;
    StripClass cls[STRIP_MAX_CLS];
    alignas(64) CUtensorMap tmap[STRIP_MAX_CLS];       // STAGE 2: a class's record array as a 3-d tensor (4 * pnx floats, pny, class packs)
};

// ---- pre-pass: window records, guard thresholds and the slow pass's hints ------------------------------------------------
// One block = 16 x 16 cells of the window array of one (job, pack).  The 20 x 20 halo tile of the four levels goes
// through shared memory once; every thread then writes its cell's record and derives the threshold of the 5 x 5 window
// whose ORIGIN is that cell (cells c-1..c+3, r-1..r+3) and the hint bits of the 4 x 4 stencil anchored there.
struct PackWinTask { int job, pack; };

__global__ void __launch_bounds__(256, 8) k_pack_win(const Job* __restrict__ jobs, const PackWinTask* __restrict__ tasks,
                                                  const SlabPtrs* __restrict__ slabs)
{
    const PackWinTask t = tasks[blockIdx.y];
    const Job& jb = jobs[t.job];
    const int pnx = jb.bnx, pny = jb.bny, ntx = (pnx + 15) >> 4;
    const int bx = blockIdx.x % ntx, by = blockIdx.x / ntx;
    if (by * 16 >= pny) return;
    // tile of 20 x 20 cells: the block's 16 x 16 cells with one cell of halo to the left / top and three to the right / bottom
    __shared__ float4 s_v[20][21];             // record values of the four levels (zeros replaced, unusable values 0)
    __shared__ float4 s_a[20][21];             // usable magnitude per level (0 where missing / non-finite); later its 5-wide row maximum
    // (the four levels of a cell side by side: one 16-byte shared-memory access where four 4-byte ones throttled the MIO queue)
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
        float v0 = 0.0f;
        float vv[4], aa[4];
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            float v = v0, a = 0.0f;             // levels past the job's last one repeat the pack's first level: their results go to the
                                                // dump area, and their row interpolants must not trip the window's (joint) guard
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
            if (lv == 0) v0 = v;
            vv[lv] = v; aa[lv] = a;
        }
        s_v[r][c] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        s_a[r][c] = make_float4(aa[0], aa[1], aa[2], aa[3]);
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
    float4 hm[2];
    unsigned hf[2];
    for (int e = tid, k = 0; e < 400; e += 256, ++k) {
        const int c = e % 20, r = e / 20;
        unsigned f = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) if (c + d < 20) f |= s_f[r][c + d];
        hf[k] = f;
        float4 m = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int d = 0; d < 5; ++d)
            if (c + d < 20) {
                const float4 q = s_a[r][c + d];
                m.x = fmaxf(m.x, q.x); m.y = fmaxf(m.y, q.y); m.z = fmaxf(m.z, q.z); m.w = fmaxf(m.w, q.w);
            }
        hm[k] = m;
    }
    __syncthreads();
    for (int e = tid, k = 0; e < 400; e += 256, ++k) {
        const int c = e % 20, r = e / 20;
        s_h[r][c] = (unsigned short)hf[k];
        s_a[r][c] = hm[k];
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    const int cx = bx * 16 + tx, cy = by * 16 + ty;
    if (cx >= pnx || cy >= pny) return;
    const int c0 = tx + 1, r0 = ty + 1;        // the thread's cell in the tile
    const size_t o = ((size_t)t.pack * pny + cy) * pnx + cx;
    const_cast<float4*>(jb.rec)[o] = s_v[r0][c0];
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
    // Guard threshold of the 5 x 5 window whose origin is the cell: 2^-18 of the largest usable magnitude in it (an upper bound
    // for every stencil inside the window), at least 1e-21 -- ONE value for the pack's levels, their maximum: the strip kernel
    // tests the smallest |row interpolant| over rows and levels against it (a level whose magnitudes are far below another's of
    // the same pack sends a few more points to the exact path, nothing else changes).  Sign bit: some level sits at the floor,
    // i.e. its window may consist of zero stand-ins only and a result can equal 1.E-20 (:1317).
    float T = 0.0f;
    bool at_floor = false;
    float4 cm = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int r = -1; r <= 3; ++r) {
        const float4 q = s_a[r0 + r][c0 - 1];
        cm.x = fmaxf(cm.x, q.x); cm.y = fmaxf(cm.y, q.y); cm.z = fmaxf(cm.z, q.z); cm.w = fmaxf(cm.w, q.w);
    }
    const float cmx[4] = { cm.x, cm.y, cm.z, cm.w };
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
        if (!src[lv]) continue;
        const float mx = cmx[lv];
        float thr = fmaxf(mx * 3.814697265625e-6f, 1.0e-21f);
        at_floor = at_floor || thr == 1.0e-21f;
        // a weighted sum stays below 1.3 x the largest magnitude: if that cannot reach |missing_value| the kernel need not
        // compare its results with missing_value (it does when missing_value is 0); otherwise the window goes to the slow path
        if (msg != 0.0f && !(mx * 2.0f < fabsf(msg))) thr = __int_as_float(0x7f800000);
        T = fmaxf(T, thr);
    }
    const unsigned fw = stencil_byte(0, 0) | (stencil_byte(1, 0) << 8) | (stencil_byte(0, 1) << 16) | (stencil_byte(1, 1) << 24);
    const_cast<uint2*>(jb.tf)[o] = make_uint2(fw, __float_as_uint(at_floor ? -T : T));
}

// ---- strip kernel --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

// predicated stores: the compiler otherwise wraps every `if (ok) store` around a ballot in a branch + reconvergence pair.
// One predicate serves the four levels of a point.
__device__ __forceinline__ void st4_f32_if(bool p, float* a0, float* a1, float* a2, float* a3, float v0, float v1, float v2, float v3)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0;\n\t@q st.global.f32 [%1], %5;\n\t@q st.global.f32 [%2], %6;\n\t"
                 "@q st.global.f32 [%3], %7;\n\t@q st.global.f32 [%4], %8; }"
                 ::"r"((unsigned)p), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
}
__device__ __forceinline__ void st4_u32_if(bool p, int* a0, int* a1, int* a2, int* a3, unsigned v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0;\n\t@q st.global.u32 [%1], %5;\n\t@q st.global.u32 [%2], %5;\n\t"
                 "@q st.global.u32 [%3], %5;\n\t@q st.global.u32 [%4], %5; }"
                 ::"r"((unsigned)p), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "r"(v) : "memory");
}

// mbarrier + bulk-copy (TMA engine; SASS UBLKCP / UTMALDG) primitives of the window ring
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
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "WAIT_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one lane of the (converged) warp, known to the compiler as exactly one: the uniform-datapath copy below needs no lane loop
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0u;
}
// one tiled tensor copy: box (SBW cells x SBH rows x 1 pack) of the job's record array, out-of-range cells arrive as 0
__device__ __forceinline__ void tma_g2s_3d(unsigned dst, const void* tmap, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// Window ring of the staged variants: each warp (= block) copies the bounding box of its lanes' windows for pack p+2 into
// shared memory while pack p is computed -- the north-star design ("each thread block TMA-stages the compact source-grid
// bounding box its destination tile maps onto").  Box: up to 16 x 8 cells of 16 B.
//   STAGE 1: one bulk copy per box row, issued by lanes 0..bh-1; row stride 17 cells (rows start in different banks)
//   STAGE 2: ONE tiled tensor copy (cp.async.bulk.tensor.3d through a CUtensorMap of the job's record array), dense 16-cell rows
constexpr int SBW = 16, SBH = 8, SNS = 3;

#ifndef STRIP_MINB
#define STRIP_MINB 16      // resident warps (= blocks) per SM the R = 4 kernel is compiled for: 128 registers (12: 0.921 ms, 20: 0.996 ms, 16: 0.907 ms)
#endif

// flagged points of a strip (rare): kept out of line so that the pack loop carries neither the queue code nor its registers
__device__ __noinline__ void strip_slow_push(const SlowQ Q, unsigned s0, unsigned s1, unsigned s2, unsigned s3, int gslab, int nl, int idx0, int nx)
{
    const int lane = threadIdx.x;
    if (s0) slowq_push(Q, s0, lane, gslab, nl, idx0);
    if (s1) slowq_push(Q, s1, lane, gslab, nl, idx0 + nx);
    if (s2) slowq_push(Q, s2, lane, gslab, nl, idx0 + 2 * nx);
    if (s3) slowq_push(Q, s3, lane, gslab, nl, idx0 + 3 * nx);
}

// R points per thread, WIN = window edge (5 when R > 1: cells of a strip may differ by one in each direction; 4 for R = 1).
// One warp per block.  STAGE: 0 = windows straight from L1/L2, 1 / 2 = from the shared-memory ring (see above).
// A block works through a list of SEGMENTS (consecutive packs of one job); all jobs of the list share the destination grid,
// the cached coordinates and the record-array geometry, so the level-independent set-up -- cells, weights, window box --
// is done once per block and only the landmask / fill logic is redone per job.
template <int R, int STAGE>
#ifdef STRIP_MAXREG
__global__ void __launch_bounds__(32) __maxnreg__(STRIP_MAXREG)
#else
__global__ void __launch_bounds__(32, R == 4 ? STRIP_MINB : 20)
#endif
k_metgrid_strip16(const __grid_constant__ StripParams P)
{
    constexpr int WIN = R > 1 ? 5 : 4;
    constexpr int SROW = STAGE == 2 ? SBW : SBW + 1;     // cells per ring row
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
    // landmask of the strip's points: read where a job first needs it (the 3-d fields that come first do not; a select on an
    // early load would make every block wait for it at its first segment), brought towards the SM now
    if (gj.landmask_cls) {
#pragma unroll
        for (int p = 0; p < R; ++p)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gj.landmask_cls + (size_t)min(jj0 + p, ny - 1) * nx + min(ii, nx - 1)));
    }
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
    unsigned csel0 = 0;                          // nibble p: byte of the flag word that belongs to point p's cell, dx + 2 dy
#pragma unroll
    for (int p = 0; p < R; ++p) {
        const int di = ci[p] - i0, dj = cj[p] - j0;
        if (geo[p] && (di > WIN - 4 || dj > WIN - 4)) geo[p] = false;     // strip spans more than two cells: slow path
        if (geo[p]) { geom |= 1u << p; csel0 |= (unsigned)(di + 2 * dj) << (4 * p); }
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

    // ---- window source: shared-memory ring fed by the TMA engine, or global memory directly
    __shared__ alignas(128) float4 s_win[STAGE ? SNS * SBH * SROW : 1];
    __shared__ alignas(8) unsigned long long s_bar[STAGE ? SNS : 1];
    constexpr unsigned STG_BYTES = SBH * SROW * 16;
    bool staged = false;
    int bw = 0, bh = 0, bc0 = 0, bc1 = 0;
    size_t boff = 0;
    unsigned sw_off = 0;                         // byte offset of the lane's window inside a ring stage
    if (STAGE) {
        // bounding box of the warp's windows (lanes without a usable point do not count)
        const int bi0 = __reduce_min_sync(0xffffffffu, anygeo ? i0 - 1 : INT_MAX);
        const int bj0 = __reduce_min_sync(0xffffffffu, anygeo ? j0 - 1 : INT_MAX);
        const int bi1 = __reduce_max_sync(0xffffffffu, anygeo ? i0 + WIN - 2 : INT_MIN);
        const int bj1 = __reduce_max_sync(0xffffffffu, anygeo ? j0 + WIN - 2 : INT_MIN);
        bw = bi1 - bi0 + 1; bh = bj1 - bj0 + 1;
        staged = bi0 != INT_MAX && bw <= SBW && bh <= SBH;
        if (staged) {
            boff = (size_t)(bj0 - gj.py0 + (lane < bh ? lane : 0)) * pnx + (bi0 - gj.px0);    // STAGE 1: lane r copies box row r
            bc0 = 4 * (bi0 - gj.px0); bc1 = bj0 - gj.py0;                                     // STAGE 2: box origin in tensor coordinates
            if (anygeo) sw_off = (unsigned)(((j0 - 1 - bj0) * SROW + (i0 - 1 - bi0)) * 16);
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
    // Output addressing: element index of the strip's points / bitmap words (32-bit: a plane has fewer than 2^31 points), so
    // that an address is one IMAD.WIDE on a level's base pointer.  Levels past a job's last one have their pointers aimed at
    // a dump area by the host, so the pack loop has no level test at all; stores are predicated, not branched around.
    const unsigned colw = __ballot_sync(0xffffffffu, ii < nx);   // lanes inside the grid
    const unsigned rowsm = (1u << min(R, ny - jj0)) - 1u;     // rows of the strip that exist
    const int idx0 = jj0 * nx + ii;
    const int widx = (jj0 + (lane < R ? lane : 0)) * gj.nxw + tx;     // lane p < R writes the bitmap words of strip row p
    const bool wrow = lane < R && ((rowsm >> lane) & 1u);
    const bool fulltile = colw == 0xffffffffu && jj0 + R <= ny;      // all 32 x R points of the tile exist
    unsigned stg = 0, par = 0;                   // ring stage of the next pack to be consumed and its barrier phase
    // The window records, thresholds and flag words of all jobs of a geometry class lie back to back (host), so the block's
    // packs are ONE contiguous range [q0, q0 + nq) of the class arrays, across segment (job) boundaries: the ring copy of the
    // pack two ahead and the threshold / flag loads of the next pack never look at the segment table.
    const StripClass& cl = P.cls[wk.cls];
    const int nq = wk.npk;
    const void* tmap = &P.tmap[wk.cls];
    const float4* Pbox = cl.P + (size_t)wk.q0 * pstride + boff;
    // (ileft, zq) = packs still to be issued to the ring, class index of the next one.  Kept opaque: the compiler otherwise
    // re-derives them from the parameter block (blockIdx -> work item -> ...) through a chain of dependent constant loads
    // in every iteration.
    int ileft = nq, zq = wk.q0;
    asm volatile("" : "+r"(ileft), "+r"(zq), "+l"(tmap));
    auto issue_next = [&](unsigned st) {
        if (ileft > 0) {
            const unsigned bar = bar0 + 8 * st;
            if (STAGE == 2) {
                if (elect_one()) {
                    mbar_expect_tx(bar, (unsigned)(SBW * SBH * 16));
                    tma_g2s_3d(win0 + st * STG_BYTES, tmap, bc0, bc1, zq, bar);
                }
            } else {
                if (lane == 0) mbar_expect_tx(bar, (unsigned)(bw * bh * 16));
                __syncwarp();
                if (lane < bh) bulk_g2s(win0 + st * STG_BYTES + (unsigned)(lane * SROW * 16), Pbox + (size_t)(zq - wk.q0) * pstride, (unsigned)(bw * 16), bar);
            }
        }
        --ileft; ++zq;
    };
    if (STAGE && staged) {
#pragma unroll
        for (int st = 0; st < SNS - 1; ++st) issue_next((unsigned)st);
    }
    const uint2* __restrict__ TFp = cl.TF + (size_t)wk.q0 * pstride + toff;
    const float4* __restrict__ Pw = cl.P + (size_t)wk.q0 * pstride + woff;
    uint2 tfv = __ldg(TFp);                      // flag word and threshold of the first pack; later ones one iteration ahead
    int qleft = nq;                              // packs of the block still to be computed, the current one included
    asm volatile("" : "+r"(qleft));

    unsigned fillm = 0, nibm = 0, csel = 0;
    bool anyfill = false;
    for (int sg = 0; sg < wk.nseg; ++sg) {
        const StripSeg seg = P.seg[wk.seg0 + sg];
        const StripJob& jb = P.job[seg.job];
        // ---- the job's point logic (interp_met_field :2490-2541): process_bdy_width interior and landmask == masked
        // receive fill_missing; masked=both selects the interp mask (flag nibble) by the landmask.  Skipped when the host
        // found the previous segment's job to have the same settings (the levels of TT, RH, GHT ...).
        if (sg == 0 || !seg.same) {
            unsigned fastm = geom;
            fillm = 0; nibm = 0;
            const int pw = jb.pw, di = jb.sm1 + ii;
            const bool fill_i = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw;
#pragma unroll
            for (int p = 0; p < R; ++p) {
                const int jj = jj0 + p, dj = jb.sm2 + jj;
                const bool act = ii < nx && jj < ny;
                float lm = 0.0f;
                if (jb.has_lm) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(lm) : "l"(jb.landmask + (size_t)min(jj, ny - 1) * nx + min(ii, nx - 1)));
                bool fill = fill_i && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw, lm_ok = true;
                bool hi_nib = false;
                if (jb.both) {
                    hi_nib = lm == 1.0f;                        // land point: interp_water_mask, flag bits 4-7
                    lm_ok = lm == 0.0f || lm == 1.0f;
                } else fill = fill || (jb.has_lm && lm == jb.masked);
                if (fill && act) fillm |= 1u << p;
                if (fill || !lm_ok) fastm &= ~(1u << p);
                nibm |= (hi_nib ? 0xf0u : 0x0fu) << (8 * p);
            }
            // flag test of a pack in two instructions: byte p of prmt(flag word, 0, csel) is the flag byte of point p's stencil
            // (a zero byte for points that cannot take the window path); the point is clean iff its nibble is all ones
            csel = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p) csel |= (p < R && ((fastm >> p) & 1u) ? ((csel0 >> (4 * p)) & 3u) : 4u) << (4 * p);
            anyfill = __any_sync(0xffffffffu, fillm != 0u);
        }
        const float fillv = jb.fill;
        const int nslab = jb.nslab, gslab0 = jb.gslab0;
        const int npk = seg.npk;
        const StripOut* outp = P.out + jb.slab0 + seg.pack0 * 4;

        // the pack loop exists twice (windows from the ring / from global memory) so that neither carries the other's loads
        auto packs = [&](auto direct_tag) {
            constexpr bool DIRECT = decltype(direct_tag)::value;
#pragma unroll 1
            for (int pk = 0; pk < npk; ++pk) {
                const unsigned ccf = tfv.x;
                const float cT = __uint_as_float(tfv.y);
                if (--qleft > 0) {                       // stencil flags and threshold of the next pack, one iteration ahead
                    TFp += pstride;
                    tfv = __ldg(TFp);
                }
                unsigned swa = 0;
                if (!DIRECT) {
                    const unsigned prev = stg == 0 ? SNS - 1 : stg - 1;
                    __syncwarp();                        // every lane has finished reading the stage that is refilled now
                    issue_next(prev);
                    mbar_wait(bar0 + 8 * stg, par);
                    swa = win0 + stg * STG_BYTES + sw_off;
                    stg = stg + 1 == SNS ? 0 : stg + 1;
                    par ^= stg == 0 ? 1u : 0u;
                }
                float2 a01[R], a23[R];
                bool g[R];                               // point p's stencil is clean; after the rows: and its guard holds
                float gm[R];                             // smallest |row interpolant| of point p over the window's rows and the four levels
                // the point's stencil must be clean for all four levels (under the mask the landmask selects, masked=both)
                const unsigned dirty = ~__byte_perm(ccf, 0u, csel) & nibm;
#pragma unroll
                for (int p = 0; p < R; ++p) {
                    a01[p] = make_float2(0.0f, 0.0f); a23[p] = a01[p];
                    g[p] = (dirty & (0xffu << (8 * p))) == 0u;
                    gm[p] = 3.0e38f;
                }
#pragma unroll
                for (int l = 0; l < WIN; ++l) {
                    float4 v[WIN];
#pragma unroll
                    for (int k = 0; k < WIN; ++k) {
                        if (DIRECT) v[k] = __ldg(Pw + l * pnx + k);
                        else asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w)
                                          : "r"(swa + (unsigned)((l * SROW + k) * 16)));
                    }
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
                        asm("min.abs.f32 %0, %0, %1, %2;" : "+f"(gm[p]) : "f"(r01.x), "f"(r01.y));      // FMNMX3 |a|, |b|, |c|
                        asm("min.abs.f32 %0, %0, %1, %2;" : "+f"(gm[p]) : "f"(r23.x), "f"(r23.y));
                        const float2 wyl = make_float2(wyr[p], wyr[p]);
                        a01[p] = __ffma2_rn(wyl, r01, a01[p]);
                        a23[p] = __ffma2_rn(wyl, r23, a23[p]);
                    }
                }
                if (DIRECT) Pw += pstride;
                {
                    const float aT = fabsf(cT);
#pragma unroll
                    for (int p = 0; p < R; ++p) g[p] = g[p] && gm[p] > aT;
                }
                // ---- outputs.  The guard is joint over the pack's four levels: a point either delivers all of them or hands
                // all of them to the slow pass, so one ballot per point serves the pack.
                // a result equal to the zero stand-in becomes 0 (:1317): possible only where a whole window consists of
                // stand-ins, i.e. its threshold sits at the 1e-21 floor: the sign bit of the window's threshold word (warp-uniform test, one branch per pack)
                if (__any_sync(0xffffffffu, cT < 0.0f)) {
#pragma unroll
                    for (int p = 0; p < R; ++p) {
                        a01[p].x = a01[p].x == 1.0e-20f ? 0.0f : a01[p].x; a01[p].y = a01[p].y == 1.0e-20f ? 0.0f : a01[p].y;
                        a23[p].x = a23[p].x == 1.0e-20f ? 0.0f : a23[p].x; a23[p].y = a23[p].y == 1.0e-20f ? 0.0f : a23[p].y;
                    }
                }
                // "value /= missing_value" (:2558) needs no test: jobs with missing_value == 0 never come here (host), and the
                // pre-pass sends windows whose magnitudes come near |missing_value| to the slow path
                if (anyfill) {                           // warp-uniform: fill_missing points exist in this strip row group
#pragma unroll
                    for (int p = 0; p < R; ++p) {
                        const bool f = (fillm >> p) & 1u;
                        a01[p].x = f ? fillv : a01[p].x; a01[p].y = f ? fillv : a01[p].y;
                        a23[p].x = f ? fillv : a23[p].x; a23[p].y = f ? fillv : a23[p].y;
                        g[p] = g[p] || f;
                    }
                }
                const StripOut o0 = outp[0], o1 = outp[1], o2 = outp[2], o3 = outp[3];
                outp += 4;
                bool gall = true;
#pragma unroll
                for (int p = 0; p < R; ++p) gall = gall && g[p];
                if (fulltile && __all_sync(0xffffffffu, gall)) {
                    // the common case, warp-uniform: every point of the 32 x R tile delivers -- unconditional stores, full words
#pragma unroll
                    for (int p = 0; p < R; ++p) {
                        const int idx = idx0 + p * nx;
                        o0.r[idx] = a01[p].x; o1.r[idx] = a01[p].y; o2.r[idx] = a23[p].x; o3.r[idx] = a23[p].y;
                    }
                    if (lane < R) { o0.w[widx] = -1; o1.w[widx] = -1; o2.w[widx] = -1; o3.w[widx] = -1; }
                } else {
                    unsigned word[R];
#pragma unroll
                    for (int p = 0; p < R; ++p) word[p] = __ballot_sync(0xffffffffu, g[p]);
#pragma unroll
                    for (int p = 0; p < R; ++p) {
                        const int idx = idx0 + p * nx;
                        if (g[p]) { o0.r[idx] = a01[p].x; o1.r[idx] = a01[p].y; o2.r[idx] = a23[p].x; o3.r[idx] = a23[p].y; }
                    }
                    {
                        // flagged points (rare): one warp-uniform test per pack, the queue code stays out of the way
                        unsigned sw[4] = { 0u, 0u, 0u, 0u };
                        unsigned any = 0;
#pragma unroll
                        for (int p = 0; p < R; ++p) { sw[p] = (((rowsm >> p) & 1u) ? colw : 0u) & ~word[p]; any |= sw[p]; }
                        if (any) {
                            const int pk0 = seg.pack0 + pk;
                            strip_slow_push(P.sq, sw[0], sw[1], sw[2], sw[3], gslab0 + pk0 * 4, min(4, nslab - pk0 * 4), idx0, nx);
                        }
                    }
                    if (wrow) {
                        // lane p writes the new_pts words of strip row p for the four levels
                        unsigned w = word[0];
#pragma unroll
                        for (int p = 1; p < R; ++p) w = lane == p ? word[p] : w;
                        o0.w[widx] = (int)w; o1.w[widx] = (int)w; o2.w[widx] = (int)w; o3.w[widx] = (int)w;
                    }
                }
            }
        };
        if (STAGE && staged) packs(std::false_type{});
        else packs(std::true_type{});
    }
}

}  // namespace wps
