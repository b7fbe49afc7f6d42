// metgrid.cu -- batched replacement of metgrid's interp_met_field (process_domain_module.F:2205-2669):
// coordinate cache, one launch over (destination tile x slab group), deferred search passes.
#include <chrono>
#include <algorithm>
#include <climits>
#include "ctx.h"
#include "interp.cuh"

namespace wps {

constexpr int TILE_X = 32;
constexpr int TILE_Y = 8;

// One group of slabs that share everything but their data pointers.
struct Job {
    const float2* xy;        // cached (rx, ry) = lltoxy(xlat, xlon) w.r.t. the source projection
    const float* xlat;
    const float* xlon;
    const float* landmask;   // nullptr when the optional argument is absent
    int sm1, em1, sm2, em2, nx, ny, nxw;
    int sd1, ed1, sd2, ed2, pw;
    DevProj proj;
    int minx, maxx, miny, maxy, bdr;
    const float* mask[3];
    int ops[WPSGPU_MAX_OPS], opts[WPSGPU_MAX_OPS];
    float msg, fill, masked;
    float mask_val[3];
    int mask_rel[3];
    int nslab, slab0;        // slabs [slab0, slab0+nslab) of the per-slab pointer tables
    int tile0, tiles_x, tiles_y;
    unsigned long long defer_base;  // first slot of this job's deferred-entry region (jobs with SEARCH only)
    // packed sixteen_pt fast path (fast != 0): records over the source-cell bounding box of the destination grid
    int fast, bx0, by0, bnx, bny, npack;
    const float4* rec;           // [npack][bny][ceil(bnx/8)][6][8]: terms {b, c, t1, t2, t3, t4} of 8 cells, float4 = 4 packed levels
    const unsigned char* cf;     // [npack][bny][bnx] bit lv = 4x4 stencil anchored here is "clean" for level lv
    float f_negzero, f_one;      // -0.0f and 1.0f as run-time values (see mul2/add2 in the packed kernel)
    int gcell;                   // average_gcell branch (:2331-2478): masked=both is not special there
    // tolerance mode (strip != 0, strip16.cuh): rec holds the window records [npack][bny][bnx] float4 over cells
    // (bx0 + c, by0 + r) -- not clipped to the source array, columns / rows clamped instead --, thr the guard
    // thresholds (two sets for masked=both), cf the slow pass's hint bits; strip = destination points per thread
    int strip;
    const uint2* tf;             // strip jobs, per (pack, cell): .x = flag word (byte 0 = the cell's own stencil, like cf), .y = guard threshold of the window (k_pack_win)
};

struct SlabPtrs {
    const float* src;
    float* r_arr;
    const int* valid;
    int* new_pts;
    const int* decided;  // average_gcell: bit set = r_arr already holds the cell average (:2393-2400)
    int job, pad_;       // the job the slab belongs to (the queue-driven kernels start from a slab index)
};


// Points a first-method kernel cannot finish (missing values in the stencil, coast lines, nodes, array edges) are
// appended to a compact queue as (slab, point) and redone by k_metgrid_slowlist with the full chain.  cnt[0] = entries
// requested, cnt[1] = overflow flag: when the queue is too small the whole pass is redone by the generic kernel instead.
struct SlowQ { int2* q; unsigned* cnt; unsigned cap; };

// sword: ballot of the lanes whose point is flagged (non-zero, warp-uniform); the point stands for nl consecutive slabs
__device__ __forceinline__ void slowq_push(const SlowQ& Q, unsigned sword, int lane, int gslab, int nl, int idx)
{
    const unsigned n = __popc(sword);
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(Q.cnt, n * (unsigned)nl);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((sword >> lane) & 1u) {
        const unsigned at = base + __popc(sword & ((1u << lane) - 1u)) * (unsigned)nl;
        if (at + (unsigned)nl <= Q.cap) {
            for (int lv = 0; lv < nl; ++lv) Q.q[at + lv] = make_int2(gslab + lv, idx);
        } else Q.cnt[1] = 1u;
    }
}

// ---- coordinate cache -------------------------------------------------------------------------
__global__ void k_coord_cache(DevProj p, const float* __restrict__ xlat, const float* __restrict__ xlon, int64_t n,
                              float2* __restrict__ xy)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float x, y;
    lltoxy(p, xlat[k], xlon[k], x, y, WPSGPU_M);  // istagger is M at every interp_met_field call site
    xy[k] = make_float2(x, y);
}

// ---- raw intermediate-file records -> halo'd source slabs (read_met_module.F:375-386, process_domain_module.F:1119-1140)
struct DecodeTask { const unsigned* raw; float* out; int raw_nx, ny, bdr, big_endian, row0, nrows; };   // rows [row0, row0+nrows) only

__global__ void __launch_bounds__(256) k_decode_records(const DecodeTask* __restrict__ tasks)
{
    const DecodeTask t = tasks[blockIdx.y];
    const int snx = t.raw_nx + 2 * t.bdr;
    const size_t n = (size_t)snx * t.nrows, e0 = (size_t)snx * t.row0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const size_t e = e0 + k;
        int i = (int)(e % snx) - t.bdr, j = (int)(e / snx);          // i: 0-based column of the record, halo columns outside
        if (i < 0) i += t.raw_nx; else if (i >= t.raw_nx) i -= t.raw_nx;   // halo_slab(1-bdr:0) = slab(nx-bdr+1:nx), (nx+1:nx+bdr) = slab(1:bdr)
        unsigned w = __ldg(t.raw + (size_t)j * t.raw_nx + i);
        if (t.big_endian) w = __byte_perm(w, 0, 0x0123);
        t.out[e] = __uint_as_float(w);
    }
}

// ---- per-point logic of interp_met_field (non-gcell branch, :2487-2579) ------------------------
// out flags
enum { PF_WRITE = 1, PF_BIT = 2, PF_DEFER = 4 };

template <int MODE>
__device__ __forceinline__ int interp_to_latlon(const Job& jb, const float* __restrict__ src, int sel, bool use_mask,
                                                float2 xy, float rlat, float rlon, float& r, const LitScratch* ls, int first = 0)
{
    // interp_to_latlon, process_domain_module.F:2594-2669
    Src s;
    s.a = src; s.m = use_mask ? jb.mask[sel] : nullptr;
    s.sx = jb.minx; s.ex = jb.maxx; s.sy = jb.miny; s.ey = jb.maxy; s.nx = jb.maxx - jb.minx + 1;
    s.msg = jb.msg; s.maskval = jb.mask_val[sel]; s.rel = jb.mask_rel[sel];
    float lo = (float)(jb.minx + jb.bdr) - 0.5f, hi = (float)(jb.maxx - jb.bdr) + 0.5f;
    int st = ST_FINAL_MSG;
    r = jb.msg;
    if (xy.x >= lo && xy.x <= hi) {
        st = interp_sequence<MODE>(s, jb.ops, jb.opts, xy.x, xy.y, r, ls, first);   // `first` applies to this position only
        if (st == ST_DEFER) return ST_DEFER;
    }
    if (r == jb.msg) {
        if (rlon < 0.0f) {
            float rx, ry;
            lltoxy(jb.proj, rlat, rlon + 360.0f, rx, ry, WPSGPU_M);
            if (rx >= lo && rx <= hi) {
                st = interp_sequence<MODE>(s, jb.ops, jb.opts, rx, ry, r, ls);
                if (st == ST_DEFER) return ST_DEFER;
            } else r = jb.msg;
        }
    }
    return ST_VAL;
}

template <int MODE>
__device__ __forceinline__ int process_point(const Job& jb, const float* __restrict__ src, int i, int j, size_t idx,
                                             float2 xy, bool valid_bit, float& out, const LitScratch* ls = nullptr, int first = 0)
{
    const int pw = jb.pw;
    if (abs(i - jb.sd1) >= pw && abs(i - jb.ed1) >= pw && abs(j - jb.sd2) >= pw && abs(j - jb.ed2) >= pw) {
        out = jb.fill;
        return PF_WRITE | PF_BIT;
    }
    float temp = jb.msg, lm = 0.0f;
    int sel = 0;
    bool do_interp = true;
    const bool has_lm = jb.landmask != nullptr;
    if (has_lm) {
        lm = __ldg(jb.landmask + idx);
        if (jb.masked == WPSGPU_MASKED_BOTH && !jb.gcell) {
            if (lm == 0.0f) sel = 1;        // water point: interp_land_mask
            else if (lm == 1.0f) sel = 2;   // land point: interp_water_mask
            else do_interp = false;         // reference leaves `temp` stale here; we define it as missing_value
        } else if (lm != jb.masked) sel = 0;
        else do_interp = false;             // temp = missing_value
    }
    if (do_interp) {
        float rlat = 0.0f, rlon = 0.0f;
        rlon = __ldg(jb.xlon + idx);
        if (rlon < 0.0f) rlat = __ldg(jb.xlat + idx);
        int st = interp_to_latlon<MODE>(jb, src, sel, jb.mask[sel] != nullptr, xy, rlat, rlon, temp, ls, first);
        if (st == ST_DEFER) return PF_DEFER;
    }
    int flags = 0;
    if (temp != jb.msg) { out = temp; flags = PF_WRITE | PF_BIT; }
    else if (has_lm && lm == jb.masked) { out = jb.fill; flags = PF_WRITE | PF_BIT; }
    if (!(flags & PF_BIT) && !valid_bit) {
        out = jb.fill;
        flags |= PF_WRITE;
        if (jb.fill != WPSGPU_NAN) flags |= PF_BIT;
    }
    return flags;
}

// ---- main batched kernel -------------------------------------------------------------------------
// grid.x = sum over jobs of tiles; block = 32x8 destination points; each thread walks the job's slabs.
__global__ void __launch_bounds__(TILE_X * TILE_Y, 4)
k_metgrid_interp(const Job* __restrict__ jobs, const int* __restrict__ cls_jobs, const int* __restrict__ cls_tile0, int ncls, int ntiles,
                 const SlabPtrs* __restrict__ slabs, Deferred* __restrict__ deferred, unsigned long long* __restrict__ defer_count,
                 const unsigned* __restrict__ only_if)
{
    // only_if: the launch is the fallback for an overflowed slow queue and has nothing to do otherwise
    if (only_if && *only_if == 0u) return;
    __shared__ Job jb;
    __shared__ int s_job, s_tile0;
    for (int b = blockIdx.x; b < ntiles; b += gridDim.x) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int lo = 0, hi = ncls - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (cls_tile0[mid] <= b) lo = mid; else hi = mid - 1;
        }
        s_job = cls_jobs[lo]; s_tile0 = cls_tile0[lo];
    }
    __syncthreads();
    {
        const int* gsrc = reinterpret_cast<const int*>(jobs + s_job);
        int* sdst = reinterpret_cast<int*>(&jb);
        int tid = threadIdx.y * TILE_X + threadIdx.x;
        for (int w = tid; w < (int)(sizeof(Job) / sizeof(int)); w += TILE_X * TILE_Y) sdst[w] = gsrc[w];
    }
    __syncthreads();
    int t = b - s_tile0;
    int tx = t % jb.tiles_x, ty = t / jb.tiles_x;
    int ii = tx * TILE_X + threadIdx.x, jj = ty * TILE_Y + threadIdx.y;  // 0-based within the grid
    bool active = ii < jb.nx && jj < jb.ny;
    size_t idx = (size_t)jj * jb.nx + ii;
    int i = jb.sm1 + ii, j = jb.sm2 + jj;
    float2 xy = make_float2(0.0f, 0.0f);
    if (active) xy = __ldg(jb.xy + idx);
    size_t widx = (size_t)jj * jb.nxw + tx;
    for (int sl = 0; sl < jb.nslab; ++sl) {
        const SlabPtrs sp = slabs[jb.slab0 + sl];
        int flags = 0;
        float out = 0.0f;
        if (active) {
            bool vbit = false;
            if (sp.valid) vbit = (__ldg(sp.valid + widx) >> threadIdx.x) & 1;
            if (sp.decided && ((__ldg(sp.decided + widx) >> threadIdx.x) & 1)) flags = PF_BIT;   // cell average already stored
            else flags = process_point<0>(jb, sp.src, i, j, idx, xy, vbit, out);
            if (flags & PF_DEFER) {
                unsigned long long slot = atomicAdd(defer_count, 1ull);
                Deferred d; d.slab = jb.slab0 + sl; d.idx = (int)idx;
                deferred[slot] = d;
            } else if (flags & PF_WRITE) sp.r_arr[idx] = out;
        }
        unsigned word = __ballot_sync(0xffffffffu, (flags & PF_BIT) != 0);
        if (threadIdx.x == 0 && jj < jb.ny) sp.new_pts[widx] = (int)word;
    }
    __syncthreads();   // jb is overwritten by the next tile
    }
}


// ---- average_gcell (process_domain_module.F:2331-2478, accum_continuous :3347-3717) ------------------
// where_maps_to of every source cell: xytoll with the source projection, lltoxy with the domain's, the (unit)
// sub-grid scaling of :3448-3449, the inside test and nint (:3441-3461).
#define GC_OUTSIDE 100000000  // OUTSIDE_DOMAIN, misc_definitions_module.F:18
__global__ void k_gc_wm(DevProj src, DevProj dom, int minx, int miny, int snx, int sny, int sm1, int em1, int sm2, int em2,
                        int2* __restrict__ wm)
{
    int ti = blockIdx.x * blockDim.x + threadIdx.x, tj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ti >= snx || tj >= sny) return;
    float lat, lon, rx, ry;
    xytoll(src, (float)(minx + ti), (float)(miny + tj), lat, lon, WPSGPU_M);
    lltoxy(dom, lat, lon, rx, ry, WPSGPU_M);
    rx = (rx - 0.5f) * 1.0f + 0.5f;
    ry = (ry - 0.5f) * 1.0f + 0.5f;
    int2 w;
    if ((float)sm1 <= rx && rx <= (float)em1 && (float)sm2 <= ry && ry <= (float)em2) { w.x = f_nint(rx); w.y = f_nint(ry); }
    else { w.x = GC_OUTSIDE; w.y = 0; }
    wm[(size_t)tj * snx + ti] = w;
}

// One thread per source cell of the block [minx+bdr, maxx-bdr] x [miny, maxy] (:2351-2375): walk the quadtree of
// process_continuous_block top-down (:3519-3525 four-corner shortcut, :3597 2x2 leaf) to find the domain cell this
// source cell is credited to, then add its value unless it is missing or masked (:3542-3592).  The reference sums in
// binary32 in traversal order; here the sum is accumulated in binary64 (order-independent to ~1e-16) and rounded once.
__global__ void k_gc_accum(const int2* __restrict__ wm, const float* __restrict__ slab, const float* __restrict__ mask,
                           int rel, float maskval, float msg, int minx, int maxx, int miny, int maxy, int bdr,
                           int sm1, int sm2, int nx, double* __restrict__ sum, int* __restrict__ cnt)
{
    const int I0 = minx + bdr, I1 = maxx - bdr, J0 = miny, J1 = maxy, snx = maxx - minx + 1;
    int pi = I0 + blockIdx.x * blockDim.x + threadIdx.x, pj = J0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (pi > I1 || pj > J1) return;
    auto W = [&](int i, int j) { return wm[(size_t)(j - miny) * snx + (i - minx)]; };
    int min_i = I0, max_i = I1, min_j = J0, max_j = J1;
    int2 dest;
    for (;;) {
        int2 a = W(min_i, min_j), b = W(min_i, max_j), c = W(max_i, max_j), d = W(max_i, min_j);
        if (a.x == b.x && a.x == c.x && a.x == d.x && a.y == b.y && a.y == c.y && a.y == d.y && a.x != GC_OUTSIDE) { dest = a; break; }
        if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) { dest = W(pi, pj); break; }
        int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
        if (pi <= ci) max_i = ci; else min_i = ci + 1;
        if (pj <= cj) max_j = cj; else min_j = cj + 1;
    }
    if (dest.x == GC_OUTSIDE) return;
    size_t o = (size_t)(pj - miny) * snx + (pi - minx);
    float v = __ldg(slab + o);
    if (v == msg) return;
    if (mask) {
        float m = __ldg(mask + o);
        bool usable = rel == REL_LT ? (m >= maskval) : (rel == REL_GT ? (m <= maskval) : (m != maskval));
        if (!usable) return;
    }
    size_t cell = (size_t)(dest.y - sm2) * nx + (dest.x - sm1);
    atomicAdd(sum + cell, (double)v);
    atomicAdd(cnt + cell, 1);
}

// Cells that received data: r_arr = sum / count (:2393-2400, :2443-2450) unless the point is in the
// process_bdy_width interior or masked by the landmask -- those and the empty cells are left to the chain.
__global__ void k_gc_apply(const Job* __restrict__ jobs, int job, const SlabPtrs* __restrict__ slabs,
                           const double* __restrict__ sum, const int* __restrict__ cnt, int* __restrict__ decided)
{
    const Job& jb = jobs[job];
    const SlabPtrs sp = slabs[jb.slab0];
    int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y;
    bool dec = false;
    if (ii < jb.nx) {
        size_t idx = (size_t)jj * jb.nx + ii;
        int di = jb.sm1 + ii, dj = jb.sm2 + jj, pw = jb.pw;
        bool in_fill = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw;
        bool lm_masked = jb.landmask != nullptr && __ldg(jb.landmask + idx) == jb.masked;
        int c = cnt[idx];
        if (!in_fill && !lm_masked && c > 0) {
            sp.r_arr[idx] = (float)sum[idx] / (float)c;
            dec = true;
        }
    }
    unsigned word = __ballot_sync(0xffffffffu, dec);
    if ((threadIdx.x & 31) == 0 && ii < jb.nxw * 32) decided[(size_t)jj * jb.nxw + (ii >> 5)] = (int)word;
}

// ---- source-cell bounding box of a destination grid (once per cached coordinate set) -----------------
__global__ void k_xy_bbox(const float2* __restrict__ xy, int64_t n, int* __restrict__ bbox)
{
    int imin = INT_MAX, imax = INT_MIN, jmin = INT_MAX, jmax = INT_MIN;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float2 p = xy[k];
        if (fabsf(p.x) < 1.0e9f && fabsf(p.y) < 1.0e9f) {
            int fi = (int)floorf(p.x), fj = (int)floorf(p.y);
            imin = min(imin, fi); imax = max(imax, fi); jmin = min(jmin, fj); jmax = max(jmax, fj);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, off)); imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, off));
        jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, off)); jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, off));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(bbox + 0, imin); atomicMax(bbox + 1, imax); atomicMin(bbox + 2, jmin); atomicMax(bbox + 3, jmax); }
}

// interp_to_latlon tries (lat, lon + 360) when the first position gave nothing and lon < 0 (:2640-2660): the source
// cells that second position can touch, so that partial source transfers cover them too.
__global__ void k_xy_bbox_retry(DevProj p, const float* __restrict__ xlat, const float* __restrict__ xlon, int64_t n, int* __restrict__ bbox)
{
    int imin = INT_MAX, imax = INT_MIN, jmin = INT_MAX, jmax = INT_MIN;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float lon = xlon[k];
        if (lon < 0.0f) {
            float rx, ry;
            lltoxy(p, xlat[k], lon + 360.0f, rx, ry, WPSGPU_M);
            if (fabsf(rx) < 1.0e9f && fabsf(ry) < 1.0e9f) {
                int fi = (int)floorf(rx), fj = (int)floorf(ry);
                imin = min(imin, fi); imax = max(imax, fi); jmin = min(jmin, fj); jmax = max(jmax, fj);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, off)); imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, off));
        jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, off)); jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, off));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(bbox + 0, imin); atomicMax(bbox + 1, imax); atomicMin(bbox + 2, jmin); atomicMax(bbox + 3, jmax); }
}

// ---- pre-pass of the packed sixteen_pt path ------------------------------------------------------------
// For every source cell (i,jj) of the bounding box and every pack of 4 levels: the row segment
// a=s(i-1,jj), b=s(i,jj), c=s(i+1,jj), d=s(i+2,jj) of sixteen_pt's stencil (columns clamped to the array like
// interp_module.F:1227-1241, zeros replaced by 1.E-20 when msgval /= 0, :1255-1257) and the x-independent
// sub-terms of oned (:1370): t1=0.5*(c-a), t2=0.5*(c+a)-b, t3=0.5*(b-d), t4=0.5*(b+d)-c, evaluated in the
// reference's operation order.  cf bit lv says the whole 4x4 stencil anchored at (i,jj) has no missing value
// and takes oned's general branch in the x direction (b*c /= 0 and a*d /= 0 in all four rows).
// One launch covers every packed job of the flush: blockIdx.z picks the job, blockIdx.y the pack of 4 levels.
__global__ void __launch_bounds__(256) k_pack16(const Job* __restrict__ jobs, const int* __restrict__ pack_jobs, const SlabPtrs* __restrict__ slabs)
{
    const Job& jb = jobs[pack_jobs[blockIdx.z]];
    int cell = blockIdx.x * blockDim.x + threadIdx.x, pack = blockIdx.y;
    if (pack >= jb.npack || cell >= jb.bnx * jb.bny) return;
    int ci = cell % jb.bnx, cj = cell / jb.bnx;
    int i = jb.bx0 + ci, jj = jb.by0 + cj;
    const int sx = jb.minx, ex = jb.maxx, sy = jb.miny, ey = jb.maxy, snx = ex - sx + 1;
    int col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) col[k] = min(max(i + k - 1, sx), ex) - sx;
    float out[6][4];
    unsigned cfbits = 0;
    const float msg = jb.msg;
    // masked=both: the landmask picks interp_land_mask (water points, bits 0-3) or interp_water_mask (land points,
    // bits 4-7) per destination point (:2495-2513), so both variants of the flag are kept
    const bool both = jb.landmask != nullptr && jb.masked == WPSGPU_MASKED_BOTH;
    const int selA = both ? 1 : 0;
    // interp_mask: any unusable cell in the stencil sends sixteen_pt to the next method (:1244-1270); level-independent
    bool cleanA = true, cleanB = true;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        int r = min(max(jj + l - 1, sy), ey);
        auto mask_row = [&](int sel, bool& c) {
            if (!jb.mask[sel]) return;
            const float* mp = jb.mask[sel] + (size_t)(r - sy) * snx;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float m = __ldg(mp + col[k]);
                bool out_ = jb.mask_rel[sel] == REL_GT ? (m > jb.mask_val[sel]) : (jb.mask_rel[sel] == REL_LT ? (m < jb.mask_val[sel]) : (m == jb.mask_val[sel]));
                c = c && !out_;
            }
        };
        mask_row(selA, cleanA);
        if (both) mask_row(2, cleanB);
    }
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
        int sl = pack * 4 + lv;
#pragma unroll
        for (int f = 0; f < 6; ++f) out[f][lv] = 0.0f;
        if (sl >= jb.nslab) continue;
        const float* __restrict__ src = slabs[jb.slab0 + sl].src;
        bool clean = true, nomsg = true;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            int r = min(max(jj + l - 1, sy), ey);
            const float* p = src + (size_t)(r - sy) * snx;
            float w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = __ldg(p + col[k]);
                if (t == 0.0f && msg != 0.0f) t = 1.0e-20f;
                w[k] = t;
            }
            nomsg = nomsg && w[0] != msg && w[1] != msg && w[2] != msg && w[3] != msg;
            clean = clean && (w[1] * w[2] != 0.0f) && (w[0] * w[3] != 0.0f);
            if (l == 1) {   // the cell's own row: jj lies inside the array, so row 1 of the stencil is never clamped
                out[0][lv] = w[1];
                out[1][lv] = w[2];
                out[2][lv] = 0.5f * (w[2] - w[0]);
                out[3][lv] = 0.5f * (w[2] + w[0]) - w[1];
                out[4][lv] = 0.5f * (w[1] - w[3]);
                out[5][lv] = 0.5f * (w[1] + w[3]) - w[2];
            }
        }
        clean = clean && nomsg;
        if (clean && cleanA) cfbits |= 1u << lv;
        if (both && clean && cleanB) cfbits |= 16u << lv;
        // single-mask jobs: bits 4-7 say "the stencil holds a missing or masked value", i.e. sixteen_pt is certain to
        // hand over to the next method (:1244-1270) -- the slow-bitmap pass then starts the chain at the second method
        if (!both && !(nomsg && cleanA)) cfbits |= 16u << lv;
    }
    size_t o = ((size_t)pack * jb.bny + cj) * jb.bnx + ci;
    // layout [pack][row][group of 8 cells][term f][cell in group]: one 128-byte line holds one term of 8 neighbouring
    // cells, so the two cells a quarter-warp of destination points typically spans are served by one L1 wavefront,
    // and the six terms of a group are six consecutive lines
    const int bnx8 = (jb.bnx + 7) >> 3;
    float4* rec = const_cast<float4*>(jb.rec) + ((((size_t)pack * jb.bny + cj) * bnx8 + (ci >> 3)) * 6) * 8 + (ci & 7);
#pragma unroll
    for (int f = 0; f < 6; ++f) rec[f * 8] = make_float4(out[f][0], out[f][1], out[f][2], out[f][3]);
    const_cast<unsigned char*>(jb.cf)[o] = (unsigned char)cfbits;
}

// ---- packed sixteen_pt kernel -----------------------------------------------------------------------------
// One thread per destination point, walking the job's levels four at a time.  The level-independent part of
// sixteen_pt (cell, fractional offsets, clamped stencil rows) is computed once per point; per pack of 4 levels
// the work is 24 float4 loads and the reference's arithmetic on level PAIRS with the sm_100 packed-f32x2
// instructions (__fmul2_rn / __fadd2_rn: two IEEE round-to-nearest products / sums per instruction, no
// contraction -- measured 2x the scalar FMUL/FADD rate, tools/ubench_f32x2.cu).  x pass: oned's general branch
// with the pre-computed sub-terms; y pass: oned's general branch, guarded by its own b*c /= 0, a*d /= 0 tests.
// Anything irregular (missing values, products that are exactly zero, points near the array edge or on a node,
// process_bdy_width interior) is flagged in the slow bitmap and redone by k_metgrid_slowbits with the full chain.
constexpr int FAST_MAX_SLABS = 64;   // slabs per job in this kernel (host splits longer level lists)

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// Packed IEEE product / sum of two level values.  ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 even with --fmad=false, which would break bit-parity with the reference; so both are issued as
// fma.rn.f32x2 with a RUN-TIME addend of -0.0 (product: RN(a*b + -0) == RN(a*b), zero signs included) or a
// run-time multiplier of 1.0 (sum: RN(a*1 + c) == RN(a+c)) that the assembler cannot fold.
struct F2K { float2 nz, one; };
__device__ __forceinline__ float2 mul2(float2 a, float2 b, const F2K& k) { return __ffma2_rn(a, b, k.nz); }
__device__ __forceinline__ float2 add2(float2 a, float2 b, const F2K& k) { return __ffma2_rn(a, k.one, b); }
// oned general branch (interp_module.F:1370) on two levels at once, x-independent terms pre-computed
__device__ __forceinline__ float2 xpass2(float2 b, float2 c, float2 t1, float2 t2, float2 t3, float2 t4, float2 x, float2 omx, const F2K& k)
{
    float2 p = add2(b, mul2(x, add2(t1, mul2(x, t2, k), k), k), k);
    float2 q = add2(c, mul2(omx, add2(t3, mul2(omx, t4, k), k), k), k);
    return add2(mul2(omx, p, k), mul2(x, q, k), k);
}
// full general branch in the y direction: (1-y)*(b+y*(0.5*(c-a)+y*(0.5*(c+a)-b))) + y*(c+(1-y)*(0.5*(b-d)+(1-y)*(0.5*(b+d)-c)))
__device__ __forceinline__ float2 ypass2(float2 a, float2 b, float2 c, float2 d, float2 y, float2 omy, const F2K& k)
{
    const float2 h = f2(0.5f);
    float2 t1 = mul2(h, add2(c, neg2(a), k), k);
    float2 t2 = add2(mul2(h, add2(c, a, k), k), neg2(b), k);
    float2 t3 = mul2(h, add2(b, neg2(d), k), k);
    float2 t4 = add2(mul2(h, add2(b, d, k), k), neg2(c), k);
    float2 p = add2(b, mul2(y, add2(t1, mul2(y, t2, k), k), k), k);
    float2 q = add2(c, mul2(omy, add2(t3, mul2(omy, t4, k), k), k), k);
    return add2(mul2(omy, p, k), mul2(y, q, k), k);
}

// Launch parameters live in the constant bank: the job descriptor and the per-level output pointers are
// warp-uniform constant loads that do not consume LSU write-back bandwidth (the limiter of this kernel).
struct FastParams {
    Job jb;
    float* r_arr[FAST_MAX_SLABS];
    int* new_pts[FAST_MAX_SLABS];
    SlowQ sq;
};

// One warp per CTA (32 x 1 destination points): a CTA's registers are released as soon as its single warp is done, which
// keeps the SM's 32 warp slots full; measured 1.59 / 1.53 / 1.51 / 1.48 ms for 8 / 4 / 2 / 1 warps per CTA.
constexpr int F16_Y = 1;
__global__ void __launch_bounds__(TILE_X * F16_Y, 1024 / (TILE_X * F16_Y))
k_metgrid_fast16(const __grid_constant__ FastParams P)
{
    const Job& jb = P.jb;
    const int nslab = jb.nslab, npack = jb.npack, nx = jb.nx, ny = jb.ny;
    const float msg = jb.msg;
    const int tx = blockIdx.x % jb.tiles_x, ty = blockIdx.x / jb.tiles_x;
    const int ii = tx * TILE_X + threadIdx.x, jj = ty * F16_Y + threadIdx.y;
    const bool active = ii < nx && jj < ny;
    const size_t idx = (size_t)jj * nx + ii;
    const size_t widx = (size_t)jj * jb.nxw + tx;
    const bool wlead = threadIdx.x == 0 && jj < ny;

    // level-independent part of interp_met_field's point logic (:2490-2541) and of sixteen_pt (interp_module.F:1211-1241)
    bool fastpt = false, fillpt = false;
    float x = 0.0f, y = 0.0f;
    int roff[4] = {0, 0, 0, 0}, cfoff = 0, cfshift = 0;
    const int bnx8 = (jb.bnx + 7) >> 3;
    const size_t rec_stride = (size_t)jb.bny * bnx8 * 48;   // float4 per pack
    if (active) {
        float2 xy = __ldg(jb.xy + idx);
        int di = jb.sm1 + ii, dj = jb.sm2 + jj;
        const int pw = jb.pw;
        bool in_fill = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw;
        // process_bdy_width interior, or landmask == masked: the point receives fill_missing and counts as new
        float lm = jb.landmask != nullptr ? __ldg(jb.landmask + idx) : 0.0f;
        bool lm_ok = true;
        if (jb.landmask != nullptr && jb.masked == WPSGPU_MASKED_BOTH) {
            // water points use interp_land_mask (flag bits 0-3), land points interp_water_mask (bits 4-7), :2495-2513
            cfshift = lm == 1.0f ? 4 : 0;
            lm_ok = lm == 0.0f || lm == 1.0f;
        } else fillpt = jb.landmask != nullptr && lm == jb.masked;
        fillpt = fillpt || in_fill;
        float lo = (float)(jb.minx + jb.bdr) - 0.5f, hi = (float)(jb.maxx - jb.bdr) + 0.5f;
        float xx = xy.x, yy = xy.y;
        if (!fillpt && lm_ok && xx >= lo && xx <= hi && fabsf(xx) < 1.0e9f && fabsf(yy) < 1.0e9f &&
            (int)xx >= jb.minx && (int)xx <= jb.maxx && (int)yy >= jb.miny && (int)yy <= jb.maxy) {
            int i = (int)(xx + 0.00001f), j = (int)(yy + 0.00001f);
            x = xx - (float)i; y = yy - (float)j;
            if ((fabsf(x) > 0.0001f || fabsf(y) > 0.0001f) && i >= jb.bx0 && i < jb.bx0 + jb.bnx && j >= jb.by0 && j < jb.by0 + jb.bny) {
                fastpt = true;
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    int r = min(max(j + l - 1, jb.miny), jb.maxy) - jb.by0;
                    if (r < 0 || r >= jb.bny) fastpt = false;
                    roff[l] = ((r * bnx8 + ((i - jb.bx0) >> 3)) * 6) * 8 + ((i - jb.bx0) & 7);
                }
                cfoff = (j - jb.by0) * jb.bnx + (i - jb.bx0);
            }
        }
        if (!fastpt) { roff[0] = roff[1] = roff[2] = roff[3] = 0; cfoff = 0; }
    }
    const float2 x2 = f2(x), omx2 = f2(1.0f - x), y2 = f2(y), omy2 = f2(1.0f - y);
    F2K k2;
    k2.nz = f2(jb.f_negzero); k2.one = f2(jb.f_one);   // run-time constants, see mul2/add2
    const size_t pack_stride = (size_t)jb.bny * jb.bnx;
    const float4* __restrict__ rec = jb.rec;
    const unsigned char* __restrict__ cf = jb.cf;
    for (int pack = 0; pack < npack; ++pack) {
        unsigned cfbits = fastpt ? ((unsigned)__ldg(cf + cfoff) >> cfshift) : 0u;
        float2 r01[4], r23[4];  // x-pass results per stencil row, levels (0,1) and (2,3)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const float4* r = rec + roff[l];
            float4 b = __ldg(r), c = __ldg(r + 8), t1 = __ldg(r + 16), t2 = __ldg(r + 24), t3 = __ldg(r + 32), t4 = __ldg(r + 40);
            r01[l] = xpass2(make_float2(b.x, b.y), make_float2(c.x, c.y), make_float2(t1.x, t1.y), make_float2(t2.x, t2.y),
                            make_float2(t3.x, t3.y), make_float2(t4.x, t4.y), x2, omx2, k2);
            r23[l] = xpass2(make_float2(b.z, b.w), make_float2(c.z, c.w), make_float2(t1.z, t1.w), make_float2(t2.z, t2.w),
                            make_float2(t3.z, t3.w), make_float2(t4.z, t4.w), x2, omx2, k2);
        }
        rec += rec_stride;
        cf += pack_stride;
        float2 v01 = ypass2(r01[0], r01[1], r01[2], r01[3], y2, omy2, k2);
        float2 v23 = ypass2(r23[0], r23[1], r23[2], r23[3], y2, omy2, k2);
        // oned's own guards in the y direction: general branch only when b*c /= 0 and a*d /= 0
        float2 bc01 = mul2(r01[1], r01[2], k2), ad01 = mul2(r01[0], r01[3], k2);
        float2 bc23 = mul2(r23[1], r23[2], k2), ad23 = mul2(r23[0], r23[3], k2);
        const float val[4] = {v01.x, v01.y, v23.x, v23.y};
        const bool gen[4] = {bc01.x != 0.0f && ad01.x != 0.0f, bc01.y != 0.0f && ad01.y != 0.0f,
                             bc23.x != 0.0f && ad23.x != 0.0f, bc23.y != 0.0f && ad23.y != 0.0f};
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            int sl = pack * 4 + lv;
            if (sl >= nslab) break;
            float r = val[lv];
            if (r == 1.0e-20f) r = 0.0f;
            bool ok = ((cfbits >> lv) & 1u) && gen[lv] && r != msg;
            if (fillpt) { r = jb.fill; ok = true; }
            if (ok) P.r_arr[sl][idx] = r;
            unsigned word = __ballot_sync(0xffffffffu, ok), sword = __ballot_sync(0xffffffffu, active && !ok);
            if (wlead) P.new_pts[sl][widx] = (int)word;
            if (sword) slowq_push(P.sq, sword, threadIdx.x, jb.slab0 + sl, 1, (int)idx);
        }
    }
}

// ---- four_pt / nearest_neighbor first-method kernel ------------------------------------------------------------
// Same structure for chains that start with four_pt (interp_module.F:1055-1171) or nearest_neighbor (:376-442):
// indices, weights and the interp-mask test are level-independent; per level 4 (or 1) loads and the reference's
// arithmetic.  Missing corners, degenerate cells (integer coordinates) and out-of-range points go to the slow bitmap.
struct Fast4Params {
    Job jb;
    const float* src[FAST_MAX_SLABS];
    float* r_arr[FAST_MAX_SLABS];
    int* new_pts[FAST_MAX_SLABS];
    SlowQ sq;
};

template <int OP>
__global__ void __launch_bounds__(TILE_X * TILE_Y, 4)
k_metgrid_fast4(const __grid_constant__ Fast4Params P)
{
    const Job& jb = P.jb;
    const int nslab = jb.nslab, nx = jb.nx, ny = jb.ny;
    const float msg = jb.msg;
    const int tx = blockIdx.x % jb.tiles_x, ty = blockIdx.x / jb.tiles_x;
    const int ii = tx * TILE_X + threadIdx.x, jj = ty * TILE_Y + threadIdx.y;
    const bool active = ii < nx && jj < ny;
    const size_t idx = (size_t)jj * nx + ii;
    const size_t widx = (size_t)jj * jb.nxw + tx;
    const bool wlead = threadIdx.x == 0 && jj < ny;
    bool fastpt = false, fillpt = false;
    float wx0 = 0.0f, wx1 = 0.0f, wy0 = 0.0f, wy1 = 0.0f;   // (max_x-xx), (xx-min_x), (max_y-yy), (yy-min_y)
    int o00 = 0, o10 = 0, o01 = 0, o11 = 0;
    if (active) {
        float2 xy = __ldg(jb.xy + idx);
        int di = jb.sm1 + ii, dj = jb.sm2 + jj;
        const int pw = jb.pw;
        bool in_fill = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw;
        fillpt = in_fill || (jb.landmask != nullptr && __ldg(jb.landmask + idx) == jb.masked);
        float lo = (float)(jb.minx + jb.bdr) - 0.5f, hi = (float)(jb.maxx - jb.bdr) + 0.5f;
        float xx = xy.x, yy = xy.y;
        const int snx = jb.maxx - jb.minx + 1;
        if (!fillpt && xx >= lo && xx <= hi && fabsf(xx) < 1.0e9f && fabsf(yy) < 1.0e9f) {
            Src ms;   // mask view for the level-independent mask tests
            ms.a = nullptr; ms.m = jb.mask[0]; ms.sx = jb.minx; ms.ex = jb.maxx; ms.sy = jb.miny; ms.ey = jb.maxy; ms.nx = snx;
            ms.msg = msg; ms.maskval = jb.mask_val[0]; ms.rel = jb.mask_rel[0];
            if (OP == WPSGPU_FOUR_POINT) {
                int min_x = f_floor(xx), min_y = f_floor(yy), max_x = f_ceiling(xx), max_y = f_ceiling(yy);
                if (min_x >= jb.minx && max_x <= jb.maxx && min_y >= jb.miny && max_y <= jb.maxy && min_x != max_x && min_y != max_y) {
                    bool m = ms.m && (masked_out(ms, min_x, min_y) || masked_out(ms, max_x, min_y) || masked_out(ms, min_x, max_y) || masked_out(ms, max_x, max_y));
                    if (!m) {
                        fastpt = true;
                        wx0 = (float)max_x - xx; wx1 = xx - (float)min_x; wy0 = (float)max_y - yy; wy1 = yy - (float)min_y;
                        o00 = (min_y - jb.miny) * snx + (min_x - jb.minx); o10 = o00 + 1; o01 = o00 + snx; o11 = o01 + 1;
                    }
                }
            } else {
                int ix = f_nint(xx), iy = f_nint(yy);
                if (ix >= jb.minx && ix <= jb.maxx && iy >= jb.miny && iy <= jb.maxy && !(ms.m && masked_out(ms, ix, iy))) {
                    fastpt = true;
                    o00 = (iy - jb.miny) * snx + (ix - jb.minx);
                }
            }
        }
    }
#pragma unroll 4
    for (int sl = 0; sl < nslab; ++sl) {
        const float* __restrict__ src = P.src[sl];
        bool ok = false;
        float r = 0.0f;
        if (fastpt) {
            if (OP == WPSGPU_FOUR_POINT) {
                float a00 = __ldg(src + o00), a10 = __ldg(src + o10), a01 = __ldg(src + o01), a11 = __ldg(src + o11);
                // four_pt general case, interp_module.F:1165-1168
                r = wy1 * (a01 * wx0 + a11 * wx1) + wy0 * (a00 * wx0 + a10 * wx1);
                ok = a00 != msg && a10 != msg && a01 != msg && a11 != msg && r != msg;
            } else {
                r = __ldg(src + o00);
                ok = r != msg;
            }
        }
        if (fillpt) { r = jb.fill; ok = true; }
        if (ok) P.r_arr[sl][idx] = r;
        unsigned word = __ballot_sync(0xffffffffu, ok), sword = __ballot_sync(0xffffffffu, active && !ok);
        if (wlead) P.new_pts[sl][widx] = (int)word;
        if (sword) slowq_push(P.sq, sword, threadIdx.x, jb.slab0 + sl, 1, (int)idx);
    }
}

// Points the first-method kernels could not take: full generic operator chain, one thread per queue entry.
#ifndef SLOW_MINBLOCKS
#define SLOW_MINBLOCKS 8   // 64 registers
#endif
__global__ void __launch_bounds__(128, SLOW_MINBLOCKS)
k_metgrid_slowlist(const Job* __restrict__ jobs, int njobs, const SlabPtrs* __restrict__ slabs, SlowQ Q,
                   Deferred* __restrict__ deferred, unsigned long long* __restrict__ defer_count)
{
    if (Q.cnt[1] != 0u) return;                  // overflowed: the generic kernel redoes the pass
    const unsigned n = min(Q.cnt[0], Q.cap);
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int2 en = Q.q[e];
        const int slab = en.x, idx = en.y;
        const SlabPtrs sp = slabs[slab];
        const Job& jb = jobs[sp.job];
        const int sl = slab - jb.slab0;
        const int jj = idx / jb.nx, ii = idx - jj * jb.nx;
        const int wwidx = jj * jb.nxw + (ii >> 5), bit = ii & 31;
        const float2 pxy = __ldg(jb.xy + idx);
        const bool hint_ok = jb.fast == 1 && !(jb.landmask != nullptr && jb.masked == WPSGPU_MASKED_BOTH);
        int first = 0;
        if (hint_ok) {
            // same cell arithmetic as sixteen_pt (interp_module.F:1211-1225); off a node and inside the record box
            // the pre-pass has recorded whether the 4x4 stencil holds a missing / masked value
            float xx = pxy.x, yy = pxy.y;
            if (fabsf(xx) < 1.0e9f && fabsf(yy) < 1.0e9f && (int)xx >= jb.minx && (int)xx <= jb.maxx && (int)yy >= jb.miny && (int)yy <= jb.maxy) {
                int ci = (int)(xx + 0.00001f), cj = (int)(yy + 0.00001f);
                float fx = xx - (float)ci, fy = yy - (float)cj;
                ci -= jb.bx0; cj -= jb.by0;
                if ((fabsf(fx) > 0.0001f || fabsf(fy) > 0.0001f) && ci >= 0 && ci < jb.bnx && cj >= 0 && cj < jb.bny) {
                    const size_t co = ((size_t)(sl >> 2) * jb.bny + cj) * jb.bnx + ci;
                    unsigned cfb = jb.strip > 0 ? (__ldg(&jb.tf[co].x) & 0xffu) : (unsigned)__ldg(jb.cf + co);
                    if ((cfb >> (4 + (sl & 3))) & 1u) first = 1;
                }
            }
        }
        bool done = false;
        if (hint_ok && first == 1 && jb.ops[1] == WPSGPU_FOUR_POINT) {
            // what process_point / interp_to_latlon do for such a point (no fill rule applies to a flagged point of
            // a job without masked=both; interp_sequence starts at method 1): the x-range test of :2620, four_pt,
            // and "value /= missing_value => store + new bit" (:2558-2563)
            const int di = jb.sm1 + ii, dj = jb.sm2 + jj, pw = jb.pw;
            const bool in_fill = abs(di - jb.sd1) >= pw && abs(di - jb.ed1) >= pw && abs(dj - jb.sd2) >= pw && abs(dj - jb.ed2) >= pw;
            const bool lm_fill = jb.landmask != nullptr && __ldg(jb.landmask + idx) == jb.masked;
            const float lo = (float)(jb.minx + jb.bdr) - 0.5f, hi = (float)(jb.maxx - jb.bdr) + 0.5f;
            if (!in_fill && !lm_fill && pxy.x >= lo && pxy.x <= hi) {
                Src s4;
                s4.a = sp.src; s4.m = jb.mask[0];
                s4.sx = jb.minx; s4.ex = jb.maxx; s4.sy = jb.miny; s4.ey = jb.maxy; s4.nx = jb.maxx - jb.minx + 1;
                s4.msg = jb.msg; s4.maskval = jb.mask_val[0]; s4.rel = jb.mask_rel[0];
                float v4 = jb.msg;
                if (op_four_pt(s4, pxy.x, pxy.y, v4) == ST_VAL && v4 != jb.msg) {
                    sp.r_arr[idx] = v4;
                    atomicOr(sp.new_pts + wwidx, (int)(1u << bit));
                    done = true;
                }
            }
        }
        if (done) continue;
        bool vbit = sp.valid ? ((((unsigned)__ldg(sp.valid + wwidx)) >> bit) & 1u) : false;
        float out = 0.0f;
        int flags = process_point<0>(jb, sp.src, jb.sm1 + ii, jb.sm2 + jj, (size_t)idx, pxy, vbit, out, nullptr, first);
        if (flags & PF_DEFER) {
            unsigned long long slot = atomicAdd(defer_count, 1ull);
            Deferred d; d.slab = slab; d.idx = idx;
            deferred[slot] = d;
        } else {
            if (flags & PF_WRITE) sp.r_arr[idx] = out;
            if (flags & PF_BIT) atomicOr(sp.new_pts + wwidx, (int)(1u << bit));
        }
    }
}

// ---- deferred passes: warp per point (closed-form search), then thread per point (literal) ---------
__device__ __forceinline__ void finish_point(const Job& jb, const SlabPtrs& sp, int idx, int flags, float out)
{
    if (flags & PF_WRITE) sp.r_arr[idx] = out;
    if (flags & PF_BIT) {
        int ii = idx % jb.nx, jj = idx / jb.nx;
        atomicOr(sp.new_pts + (size_t)jj * jb.nxw + (ii >> 5), (int)(1u << (ii & 31)));
    }
}

__global__ void __launch_bounds__(256, 2)
k_metgrid_search(const Job* __restrict__ jobs, int njobs, const SlabPtrs* __restrict__ slabs,
                 const Deferred* __restrict__ deferred, const unsigned long long* __restrict__ defer_count,
                 Deferred* __restrict__ deferred2, unsigned long long* __restrict__ defer2_count)
{
    unsigned long long n = *defer_count;
    int lane = threadIdx.x & 31;
    unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long e = warp; e < n; e += nwarps) {
        Deferred d = deferred[e];
        const SlabPtrs sp = slabs[d.slab];
        const Job& jb = jobs[sp.job];
        int ii = d.idx % jb.nx, jj = d.idx / jb.nx;
        bool vbit = false;
        if (sp.valid) vbit = (__ldg(sp.valid + (size_t)jj * jb.nxw + (ii >> 5)) >> (ii & 31)) & 1;
        float out = 0.0f;
        int flags = process_point<1>(jb, sp.src, jb.sm1 + ii, jb.sm2 + jj, (size_t)d.idx, __ldg(jb.xy + d.idx), vbit, out);
        if (lane == 0) {
            if (flags & PF_DEFER) { unsigned long long slot = atomicAdd(defer2_count, 1ull); deferred2[slot] = d; }
            else finish_point(jb, sp, d.idx, flags, out);
        }
    }
}

__global__ void __launch_bounds__(64)
k_metgrid_search_literal(const Job* __restrict__ jobs, int njobs, const SlabPtrs* __restrict__ slabs,
                         const Deferred* __restrict__ deferred2, const unsigned long long* __restrict__ defer2_count,
                         unsigned* __restrict__ scratch, size_t words_per_thread, int R, int qcap)
{
    unsigned long long n = *defer2_count;
    unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    unsigned* mine = scratch + tid * words_per_thread;
    LitScratch ls;
    size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
    ls.visited = mine; ls.R = R; ls.q = reinterpret_cast<QNode*>(mine + ((vis_words + 1) & ~size_t(1))); ls.qcap = qcap;
    for (unsigned long long e = tid; e < n; e += nth) {
        Deferred d = deferred2[e];
        const SlabPtrs sp = slabs[d.slab];
        const Job& jb = jobs[sp.job];
        int ii = d.idx % jb.nx, jj = d.idx / jb.nx;
        bool vbit = false;
        if (sp.valid) vbit = (__ldg(sp.valid + (size_t)jj * jb.nxw + (ii >> 5)) >> (ii & 31)) & 1;
        float out = 0.0f;
        int flags = process_point<2>(jb, sp.src, jb.sm1 + ii, jb.sm2 + jj, (size_t)d.idx, __ldg(jb.xy + d.idx), vbit, out, &ls);
        finish_point(jb, sp, d.idx, flags, out);
    }
}

}  // namespace wps
#include "strip16.cuh"
namespace wps {

// ---- host side ---------------------------------------------------------------------------------------
static std::string cache_key(int grid_id, const wpsgpu_proj& p)
{
    std::string k((const char*)&grid_id, sizeof(int));
    size_t nbytes = offsetof(wpsgpu_proj, gauss_lat);
    k.append((const char*)&p, nbytes);
    return k;
}

// tuning knobs of the host-pointer pipeline (slabs per job group / per pipeline stage)
static int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}

// Projection descriptors are interned per flush: callers pass the same one for every slab of a file.
static const wpsgpu_proj* intern_proj(wpsgpu_ctx* h, const wpsgpu_proj* p)
{
    const size_t head = offsetof(wpsgpu_proj, gauss_lat);
    for (auto it = h->proj_pool.rbegin(); it != h->proj_pool.rend(); ++it)
        if (memcmp(&*it, p, head) == 0 && (p->code != WPSGPU_PROJ_GAUSS || memcmp(it->gauss_lat, p->gauss_lat, sizeof(p->gauss_lat)) == 0))
            return &*it;
    h->proj_pool.push_back(*p);
    return &h->proj_pool.back();
}

static int rel_code(char c) { return c == '<' ? REL_LT : (c == '>' ? REL_GT : REL_EQ); }

static bool same_group(const QueuedSlab& a, const QueuedSlab& b)
{
    if (a.gcell || b.gcell) return false;   // average_gcell slabs carry their own accumulation buffers
    if (a.proj != b.proj && memcmp(a.proj, b.proj, offsetof(wpsgpu_proj, gauss_lat)) != 0) return false;
    if (memcmp(a.fo.ops, b.fo.ops, sizeof(a.fo.ops)) || memcmp(a.fo.opts, b.fo.opts, sizeof(a.fo.opts))) return false;
    if (memcmp(&a.fo.missing_value, &b.fo.missing_value, sizeof(float)) || memcmp(&a.fo.fill_missing, &b.fo.fill_missing, sizeof(float)) ||
        memcmp(&a.fo.masked, &b.fo.masked, sizeof(float)) || memcmp(a.fo.mask_val, b.fo.mask_val, sizeof(a.fo.mask_val)) ||
        memcmp(a.fo.mask_rel, b.fo.mask_rel, 3))
        return false;
    const wpsgpu_slab &x = a.s, &y = b.s;
    return x.minx == y.minx && x.maxx == y.maxx && x.miny == y.miny && x.maxy == y.maxy && x.bdr == y.bdr &&
           x.mask[0] == y.mask[0] && x.mask[1] == y.mask[1] && x.mask[2] == y.mask[2] && x.grid_id == y.grid_id &&
           x.field_stagger == y.field_stagger && x.use_landmask == y.use_landmask && x.process_bdy_width == y.process_bdy_width;
}

// strip jobs that share the destination grid, the cached coordinates and the record-array geometry form a class: a block
// does the level-independent set-up once for all of them, and their window records lie back to back in the arena
static bool strip_same_geometry(const Job& a, const Job& b)
{
    return a.strip == b.strip && a.xy == b.xy && a.nx == b.nx && a.ny == b.ny && a.sm1 == b.sm1 && a.sm2 == b.sm2 && a.bx0 == b.bx0 && a.by0 == b.by0 &&
           a.bnx == b.bnx && a.bny == b.bny && a.minx == b.minx && a.maxx == b.maxx && a.miny == b.miny && a.maxy == b.maxy && a.bdr == b.bdr;
}

// Window records, thresholds and flag words of the strip jobs: ONE block of each per geometry class, the jobs' arrays back to
// back in job order, so that a kernel block walks a contiguous range of packs whatever job they belong to.
static int alloc_strip_records(wpsgpu_ctx* h, std::vector<Job>& jobs)
{
    std::vector<char> done(jobs.size(), 0);
    for (size_t a = 0; a < jobs.size(); ++a) {
        if (done[a] || jobs[a].fast != 1 || jobs[a].strip <= 0) continue;
        size_t packs = 0;
        std::vector<size_t> cls;
        for (size_t b = a; b < jobs.size(); ++b)
            if (!done[b] && jobs[b].fast == 1 && jobs[b].strip > 0 && strip_same_geometry(jobs[a], jobs[b])) { cls.push_back(b); packs += jobs[b].npack; }
        const size_t cells = (size_t)jobs[a].bnx * jobs[a].bny;
        float4* rec = (float4*)h->arena.take(packs * cells * sizeof(float4));
        uint2* tf = (uint2*)h->arena.take(packs * cells * sizeof(uint2));
        WPS_CHECK(rec && tf, -3, "device arena exhausted (window records)");
        size_t q = 0;
        for (size_t b : cls) {
            done[b] = 1;
            jobs[b].rec = rec + q * cells; jobs[b].tf = tf + q * cells;
            q += jobs[b].npack;
        }
    }
    return 0;
}

// Tolerance-mode sixteen_pt jobs of one pass -> as few strip launches as the parameter block allows.
template <int R>
static void launch_strip_kernel(const StripParams& sp, unsigned tiles, unsigned nwork, int stage, cudaStream_t st)
{
    if (stage == 2 && R > 1) k_metgrid_strip16<R, (R > 1 ? 2 : 0)><<<dim3(tiles, nwork), 32, 0, st>>>(sp);
    else if (stage == 1 && R > 1) k_metgrid_strip16<R, (R > 1 ? 1 : 0)><<<dim3(tiles, nwork), 32, 0, st>>>(sp);
    else k_metgrid_strip16<R, 0><<<dim3(tiles, nwork), 32, 0, st>>>(sp);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p)
            fn = (EncodeTiledFn)p;
        else (void)cudaGetLastError();
    }
    return fn;
}

// The window records of a strip job, [npack][pny][pnx] float4, as the tensor the ring's tiled copies read: box = SBW cells x
// SBH rows x 1 pack, cells past the array's edge arrive as zeros (they belong to no lane's window).
static bool encode_window_tensor(CUtensorMap* tm, const Job& jb, int packs)
{
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)jb.bnx * 4, (cuuint64_t)jb.bny, (cuuint64_t)packs };
    const cuuint64_t strides[2] = { (cuuint64_t)jb.bnx * 16, (cuuint64_t)jb.bnx * jb.bny * 16 };
    const cuuint32_t box[3] = { SBW * 4, SBH, 1 }, estr[3] = { 1, 1, 1 };
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float4*>(jb.rec), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int launch_strips(wpsgpu_ctx* h, const std::vector<Job>& jobs, const std::vector<int>& sel, const std::vector<SlabPtrs>& sp,
                         float* dump_r, int* dump_w, const SlowQ& sq)
{
    // 2: ring fed by one tiled tensor copy per pack, 1: by one bulk copy per box row, 0: windows straight from global memory
    int stage = env_int("WPSGPU_STRIP_STAGE", 2, 0, 2);
    if (stage == 2 && !tensor_map_encoder()) stage = 1;
    const int work_packs = env_int("WPSGPU_STRIP_CHUNK", 32, 1, 64);    // packs per block
    auto same_logic = [](const Job& a, const Job& b) {      // what the per-segment point logic of the kernel reads
        return a.pw == b.pw && (a.landmask != nullptr) == (b.landmask != nullptr) && memcmp(&a.masked, &b.masked, sizeof(float)) == 0;
    };
    for (int R : { 4, 2, 1 }) {
        std::vector<int> todo;
        for (int j : sel) if (jobs[j].strip == R) todo.push_back(j);
        while (!todo.empty()) {
            // one launch: classes are taken whole as long as the parameter block has room
            StripParams P;
            const Job* seg_job[STRIP_MAX_JOBS] = {};
            P.sq = sq;
            int nj = 0, ns = 0, nseg = 0, nwork = 0, ncls = 0;
            unsigned tiles = 0;
            std::vector<int> rest;
            std::vector<char> taken(todo.size(), 0);
            bool full = false;
            for (size_t a = 0; a < todo.size() && !full; ++a) {
                if (taken[a]) continue;
                // the class of todo[a]
                std::vector<int> cls;
                for (size_t b = a; b < todo.size(); ++b)
                    if (!taken[b] && strip_same_geometry(jobs[todo[a]], jobs[todo[b]])) cls.push_back((int)b);
                int c_slots = 0, c_packs = 0;
                for (int b : cls) { c_slots += jobs[todo[b]].npack * 4; c_packs += jobs[todo[b]].npack; }
                const int c_work = (c_packs + work_packs - 1) / work_packs;
                const int c_seg = (int)cls.size() + c_work;       // upper bound: every cut adds a segment
                WPS_CHECK((int)cls.size() <= STRIP_MAX_JOBS && c_slots <= STRIP_MAX_SLABS && c_seg <= STRIP_MAX_SEGS && c_work <= STRIP_MAX_WORK, -3,
                          "sixteen_pt jobs of one grid too large for one strip launch");
                if (nj + (int)cls.size() > STRIP_MAX_JOBS || ns + c_slots > STRIP_MAX_SLABS || nseg + c_seg > STRIP_MAX_SEGS || nwork + c_work > STRIP_MAX_WORK ||
                    ncls + 1 > STRIP_MAX_CLS) {
                    full = true;
                    break;
                }
                // work items of equal size (e.g. 29 packs -> 15 + 14), cut across the class's jobs
                const int per = (c_packs + c_work - 1) / c_work;
                int in_work = 0, cq = 0;                       // cq: class-relative index of the next pack
                const Job& j0 = jobs[todo[cls[0]]];
                const size_t cells = (size_t)j0.bnx * j0.bny;
                P.cls[ncls].P = j0.rec; P.cls[ncls].TF = j0.tf;
                if (stage == 2 && R > 1) WPS_CHECK(encode_window_tensor(&P.tmap[ncls], j0, c_packs), -3, "cuTensorMapEncodeTiled failed for a window-record array");
                P.work[nwork].seg0 = (short)nseg; P.work[nwork].nseg = 0; P.work[nwork].cls = (short)ncls; P.work[nwork].npk = 0; P.work[nwork].q0 = 0;
                const float* cls_lm = nullptr;                // one grid, one landmask
                for (int b : cls) if (jobs[todo[b]].landmask) cls_lm = jobs[todo[b]].landmask;
                for (int b : cls) {
                    taken[b] = 1;
                    const Job& jb = jobs[todo[b]];
                    const bool jboth = jb.landmask != nullptr && jb.masked == WPSGPU_MASKED_BOTH;
                    const int nslot = jb.npack * 4;          // levels past the last one write to the dump area
                    StripJob& q = P.job[nj];
                    seg_job[nj] = &jb;
                    q.xy = jb.xy; q.landmask = jb.landmask; q.landmask_cls = cls_lm; q.P = jb.rec; q.TF = jb.tf;
                    q.nx = jb.nx; q.ny = jb.ny; q.nxw = jb.nxw; q.sm1 = jb.sm1; q.sm2 = jb.sm2;
                    q.sd1 = jb.sd1; q.ed1 = jb.ed1; q.sd2 = jb.sd2; q.ed2 = jb.ed2; q.pw = jb.pw;
                    q.minx = jb.minx; q.maxx = jb.maxx; q.miny = jb.miny; q.maxy = jb.maxy; q.bdr = jb.bdr;
                    q.px0 = jb.bx0; q.py0 = jb.by0; q.pnx = jb.bnx; q.pny = jb.bny;
                    q.nslab = jb.nslab; q.slab0 = ns; q.gslab0 = jb.slab0;
                    q.tiles_x = jb.tiles_x; q.ntiles = jb.tiles_x * ((jb.ny + R - 1) / R);
                    q.both = jboth ? 1 : 0; q.has_lm = jb.landmask != nullptr ? 1 : 0;
                    q.msg = jb.msg; q.fill = jb.fill; q.masked = jb.masked;
                    for (int k = 0; k < jb.nslab; ++k) {
                        const SlabPtrs& p = sp[jb.slab0 + k];
                        P.out[ns + k].r = p.r_arr; P.out[ns + k].w = p.new_pts;
                    }
                    for (int k = jb.nslab; k < nslot; ++k) { P.out[ns + k].r = dump_r; P.out[ns + k].w = dump_w; }
                    WPS_CHECK(jb.rec == j0.rec + (size_t)cq * cells && jb.tf == j0.tf + (size_t)cq * cells, -3,
                              "window records of a geometry class are not contiguous");
                    for (int pk = 0; pk < jb.npack;) {
                        if (in_work == per) {                 // next block
                            ++nwork;
                            P.work[nwork].seg0 = (short)nseg; P.work[nwork].nseg = 0; P.work[nwork].cls = (short)ncls; P.work[nwork].npk = 0;
                            P.work[nwork].q0 = cq;
                            in_work = 0;
                        }
                        const int n = std::min(per - in_work, jb.npack - pk);
                        StripSeg& sgm = P.seg[nseg++];
                        sgm.job = (short)nj; sgm.pack0 = (short)pk; sgm.npk = (short)n;
                        sgm.same = 0;
                        if (P.work[nwork].nseg > 0) {
                            const StripSeg& prev = P.seg[nseg - 2];
                            sgm.same = (short)(prev.job == sgm.job || same_logic(jb, *seg_job[prev.job]) ? 1 : 0);
                        }
                        P.work[nwork].nseg++;
                        P.work[nwork].npk = (short)(P.work[nwork].npk + n);
                        in_work += n; pk += n; cq += n;
                    }
                    tiles = std::max(tiles, (unsigned)q.ntiles);
                    ns += nslot; nj++;
                }
                ++nwork; ++ncls;
            }
            for (size_t a = 0; a < todo.size(); ++a) if (!taken[a]) rest.push_back(todo[a]);
            if (nwork > 0) {
                if (R == 4) launch_strip_kernel<4>(P, tiles, nwork, stage, h->stream);
                else if (R == 2) launch_strip_kernel<2>(P, tiles, nwork, stage, h->stream);
                else launch_strip_kernel<1>(P, tiles, nwork, stage, h->stream);
                h->launches++;
            }
            WPS_CHECK(rest.size() < todo.size(), -3, "strip launch made no progress");
            todo.swap(rest);
        }
    }
    WPS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace wps

using namespace wps;

extern "C" {

int wpsgpu_metgrid_set_grid(wpsgpu_handle_t h, int grid_id, const float* xlat, const float* xlon,
                            int sm1, int em1, int sm2, int em2)
{
    WPS_CHECK(h && xlat && xlon, -1, "wpsgpu_metgrid_set_grid: NULL argument");
    WPS_CHECK(grid_id >= 0 && grid_id < 8, -1, "grid_id %d out of range 0..7", grid_id);
    WPS_CHECK(em1 >= sm1 && em2 >= sm2, -1, "empty destination grid");
    WPS_CUDA(cudaSetDevice(h->device));
    DstGrid& g = h->grids[grid_id];
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    if (g.own) { cudaFree(g.d_xlat); cudaFree(g.d_xlon); if (g.d_landmask) cudaFree(g.d_landmask); }
    g = DstGrid();
    g.set = true; g.sm1 = sm1; g.em1 = em1; g.sm2 = sm2; g.em2 = em2;
    g.version = ++h->grid_version_counter;
    size_t bytes = g.npts() * sizeof(float);
    if (h->ptr_mode == WPSGPU_PTR_DEVICE) {
        g.d_xlat = const_cast<float*>(xlat); g.d_xlon = const_cast<float*>(xlon); g.own = false;
    } else {
        WPS_CUDA(cudaMalloc((void**)&g.d_xlat, bytes));
        WPS_CUDA(cudaMalloc((void**)&g.d_xlon, bytes));
        g.own = true;
        WPS_CUDA(cudaMemcpyAsync(g.d_xlat, xlat, bytes, cudaMemcpyHostToDevice, h->stream));
        WPS_CUDA(cudaMemcpyAsync(g.d_xlon, xlon, bytes, cudaMemcpyHostToDevice, h->stream));
        WPS_CUDA(cudaStreamSynchronize(h->stream));
    }
    // drop cached coordinates of this grid
    for (auto it = h->coord_cache.begin(); it != h->coord_cache.end();) {
        int gid;
        memcpy(&gid, it->first.data(), sizeof(int));
        if (gid == grid_id) { free_coord_entry(it->second); it = h->coord_cache.erase(it); } else ++it;
    }
    return 0;
}

int wpsgpu_metgrid_set_landmask(wpsgpu_handle_t h, int grid_id, const float* landmask)
{
    WPS_CHECK(h && landmask, -1, "wpsgpu_metgrid_set_landmask: NULL argument");
    WPS_CHECK(grid_id >= 0 && grid_id < 8 && h->grids[grid_id].set, -1, "grid %d not set", grid_id);
    WPS_CUDA(cudaSetDevice(h->device));
    DstGrid& g = h->grids[grid_id];
    size_t bytes = g.npts() * sizeof(float);
    if (h->ptr_mode == WPSGPU_PTR_DEVICE && !g.own) {
        g.d_landmask = const_cast<float*>(landmask);
    } else {
        WPS_CUDA(cudaStreamSynchronize(h->stream));
        if (!g.d_landmask) WPS_CUDA(cudaMalloc((void**)&g.d_landmask, bytes));
        WPS_CUDA(cudaMemcpyAsync(g.d_landmask, landmask, bytes,
                                 h->ptr_mode == WPSGPU_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        WPS_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int wpsgpu_metgrid_set_domain(wpsgpu_handle_t h, int sd1, int ed1, int sd2, int ed2)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    h->sd1 = sd1; h->ed1 = ed1; h->sd2 = sd2; h->ed2 = ed2; h->domain_set = true;
    return 0;
}

int wpsgpu_metgrid_enqueue_slab(wpsgpu_handle_t h, const wpsgpu_proj* src_proj, const wpsgpu_field_opts* fo,
                                const wpsgpu_slab* slab)
{
    WPS_CHECK(h && src_proj && fo && slab, -1, "wpsgpu_metgrid_enqueue_slab: NULL argument");
    WPS_CHECK(src_proj->init, -1, "source projection not initialised");
    WPS_CHECK(slab->grid_id >= 0 && slab->grid_id < 8 && h->grids[slab->grid_id].set, -1, "destination grid %d not set", slab->grid_id);
    WPS_CHECK(h->domain_set, -1, "wpsgpu_metgrid_set_domain has not been called");
    WPS_CHECK(slab->slab && slab->r_arr && slab->new_pts, -1, "slab, r_arr and new_pts must be non-NULL");
    WPS_CHECK(slab->maxx >= slab->minx && slab->maxy >= slab->miny, -1, "empty source slab");
    if (slab->raw_nx != 0) {
        const int halo2 = slab->maxx - slab->minx + 1 - slab->raw_nx;
        WPS_CHECK(slab->raw_nx > 0 && halo2 >= 0 && halo2 % 2 == 0 && halo2 / 2 <= slab->raw_nx, -1,
                  "raw record of %d columns does not fit slab bounds %d..%d", slab->raw_nx, slab->minx, slab->maxx);
    }
    WPS_CHECK(!slab->use_landmask || h->grids[slab->grid_id].d_landmask, -1, "use_landmask set but no landmask given for grid %d", slab->grid_id);
    QueuedSlab q;
    q.proj = intern_proj(h, src_proj); q.fo = *fo; q.s = *slab;
    h->queue.push_back(q);
    return 0;
}

int wpsgpu_metgrid_enqueue_gcell_slab(wpsgpu_handle_t h, const wpsgpu_proj* src_proj, const wpsgpu_proj* dom_proj,
                                      const wpsgpu_field_opts* fo, const wpsgpu_slab* slab)
{
    WPS_CHECK(h && dom_proj, -1, "wpsgpu_metgrid_enqueue_gcell_slab: NULL argument");
    WPS_CHECK(dom_proj->init, -1, "domain projection not initialised");
    WPS_TRY(wpsgpu_metgrid_enqueue_slab(h, src_proj, fo, slab));
    h->queue.back().gcell = true;
    h->queue.back().dom_proj = intern_proj(h, dom_proj);
    return 0;
}

int wpsgpu_metgrid_gcell_slab(wpsgpu_handle_t h, const wpsgpu_proj* src_proj, const wpsgpu_proj* dom_proj,
                              const wpsgpu_field_opts* fo, const wpsgpu_slab* slab)
{
    WPS_TRY(wpsgpu_metgrid_enqueue_gcell_slab(h, src_proj, dom_proj, fo, slab));
    return wpsgpu_metgrid_flush(h);
}

int wpsgpu_metgrid_enqueue_slabs(wpsgpu_handle_t h, const wpsgpu_proj* src_proj, const wpsgpu_field_opts* fo,
                                 const wpsgpu_slab* slabs, int32_t n)
{
    WPS_CHECK(h && src_proj && fo && (slabs || n == 0), -1, "wpsgpu_metgrid_enqueue_slabs: NULL argument");
    WPS_CHECK(n >= 0, -1, "negative slab count");
    const size_t mark = h->queue.size();
    for (int32_t k = 0; k < n; ++k) {
        int rc = wpsgpu_metgrid_enqueue_slab(h, src_proj, fo, slabs + k);
        if (rc) { h->queue.resize(mark); return rc; }   // all or nothing
    }
    return 0;
}

int wpsgpu_metgrid_copy_bytes(wpsgpu_handle_t h, int64_t* h2d, int64_t* d2h)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    if (h2d) *h2d = h->h2d_bytes;
    if (d2h) *d2h = h->d2h_bytes;
    return 0;
}

int wpsgpu_metgrid_last_kernel(wpsgpu_handle_t h, float* ms, double* points)
{
    WPS_CHECK(h && ms && points, -1, "NULL argument");
    WPS_CHECK(h->timed_flush, -2, "wpsgpu_metgrid_last_kernel: the last flush was not timed (wpsgpu_metgrid_set_timing)");
    WPS_CUDA(cudaEventSynchronize(h->ev_k1));
    WPS_CUDA(cudaEventElapsedTime(ms, h->ev_k0, h->ev_k1));
    *points = h->last_points;
    return 0;
}

int wpsgpu_metgrid_last_timings(wpsgpu_handle_t h, float* ms3, double* points3)
{
    WPS_CHECK(h && ms3 && points3, -1, "NULL argument");
    WPS_CHECK(h->timed_flush, -2, "wpsgpu_metgrid_last_timings: the last flush was not timed (wpsgpu_metgrid_set_timing)");
    WPS_CUDA(cudaEventSynchronize(h->ev_t[11]));
    for (int k = 0; k < 6; ++k) {
        WPS_CUDA(cudaEventElapsedTime(&ms3[k], h->ev_t[2 * k], h->ev_t[2 * k + 1]));
        points3[k] = h->last_pts3[k];
    }
    WPS_CUDA(cudaEventElapsedTime(&ms3[6], h->ev_t[0], h->ev_t[11]));   // first launch to last kernel of the flush
    points3[6] = h->last_pts3[4] + h->last_pts3[2];
    return 0;
}

// One pass over a subset of the job groups: stage inputs on s_in, compute on the handle's stream, copy results back
// on s_out.  In device-pointer mode all three are the handle's stream and nothing is copied.
struct Box { int fast, bx0, by0, bnx, bny, npack, strip; };   // strip > 0: tolerance-mode strip kernel with that many points per thread
struct FlushShared {
    std::vector<QueuedSlab>* queue;
    std::vector<std::vector<int>>* groups;
    std::vector<const float2*> job_xy;
    std::vector<const float*> job_gauss;
    std::vector<Box> box;
    std::map<const void*, void*> staged;     // host pointer -> staged device copy (masks are shared by many slabs)
    std::vector<char> tab;                   // host image of a pass's tables
    struct Sub { bool on; int x0, x1, y0, y1; };   // source columns / rows a group can touch (host mode, chains without search)
    std::vector<Sub> sub;
    std::map<const void*, Sub> staged_sub;   // rectangle (rows for raw records) present in a partially staged slab
    // a slab staged before can be shared if it is complete or holds at least the rectangle wanted now
    bool reusable(const void* p, const Sub& want) const
    {
        if (staged.find(p) == staged.end()) return false;
        auto it = staged_sub.find(p);
        if (it == staged_sub.end()) return true;
        const Sub& have = it->second;
        return want.on && have.x0 <= want.x0 && have.x1 >= want.x1 && have.y0 <= want.y0 && have.y1 >= want.y1;
    }
    int max_depth = 0, max_span = 0;
};

static int run_groups(wpsgpu_ctx* h, FlushShared& fs, const std::vector<int>& gsel, cudaStream_t s_in, cudaStream_t s_out, bool last)
{
    std::vector<QueuedSlab>& queue = *fs.queue;
    std::vector<std::vector<int>>& groups = *fs.groups;
    const bool host = h->ptr_mode == WPSGPU_PTR_HOST;
    const bool poison = env_int("WPSGPU_POISON", 0, 0, 1) != 0;
    const int njobs = (int)gsel.size();
    int nq = 0;
    for (int g : gsel) nq += (int)groups[g].size();
    // Host-pointer mode moves data in as few, as large copies as possible: arrays that follow one another in host memory
    // (the levels of a 3-d field) are placed back to back on the device and travel as one piece -- with copies running in
    // both directions, slab-sized pieces reach 37 GB/s per direction on this platform and 60+ MB pieces 49 GB/s
    // (tools/pcie_bw.py).
    struct CopyRun { const char* h; char* d; size_t bytes; };
    std::vector<CopyRun> in_runs, out_r_runs, out_np_runs;
    auto take_run = [&](std::vector<CopyRun>& runs, const void* hp, size_t bytes) -> void* {
        if (!runs.empty()) {
            CopyRun& r = runs.back();
            if (r.h + r.bytes == (const char*)hp && r.d + r.bytes == h->arena.tail()) {
                void* d = h->arena.take_tail(bytes);
                if (d) r.bytes += bytes;
                return d;
            }
        }
        void* d = h->arena.take(bytes);
        if (d) { CopyRun r; r.h = (const char*)hp; r.d = (char*)d; r.bytes = bytes; runs.push_back(r); }
        return d;
    };
    auto stage_shared = [&](const float* p, size_t bytes, const float** out) -> int {
        if (!p) { *out = nullptr; return 0; }
        if (!host) { *out = p; return 0; }
        auto it = fs.staged.find(p);
        if (it != fs.staged.end()) {
            // The same host array was staged before as a source slab.  A mask is read anywhere (search, average_gcell, the
            // pre-passes), so a copy that only holds the rectangle one chain could touch is not good enough: stage it whole.
            WPS_CHECK(it->second != nullptr, -1, "an interp mask aliases a raw (undecoded) record queued in the same flush");
            if (fs.staged_sub.find(p) == fs.staged_sub.end()) { *out = (const float*)it->second; return 0; }
            fs.staged.erase(it);
            fs.staged_sub.erase(p);
        }
        void* d = take_run(in_runs, p, bytes);
        WPS_CHECK(d != nullptr, -3, "device arena exhausted (%zu bytes requested)", bytes);
        fs.staged[p] = d;
        *out = (const float*)d;
        return 0;
    };
    std::vector<DecodeTask> decode_tasks;     // raw records of this pass
    std::vector<int> decode_slots;            // their slab-table slots
    std::vector<const void*> decode_hosts;
    std::vector<std::pair<int, const void*>> decode_pending;
    std::vector<Job> jobs(njobs);
    std::vector<int*> gc_decided(njobs, nullptr);
    std::vector<SlabPtrs> sp(nq);
    std::vector<int> cls_jobs[2], cls_tile0[2];  // 0 = generic kernel, 1 = first-method kernels
    double fast_pairs = 0.0;          // (slab, point) pairs of the first-method jobs: sizes the slow queue
    int tile_cursor[2] = { 0, 0 };
    double cls_points[2] = { 0.0, 0.0 };
    double fast16_points = 0.0, fast4_points = 0.0;
    int slab_cursor = 0;
    unsigned long long defer_cursor = 0;
    bool any_search = false;
    for (int j = 0; j < njobs; ++j) {
        const int g = gsel[j];
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        Job& jb = jobs[j];
        memset(&jb, 0, sizeof(Job));
        jb.xy = fs.job_xy[g]; jb.xlat = dg.d_xlat; jb.xlon = dg.d_xlon;
        jb.landmask = q0.s.use_landmask ? dg.d_landmask : nullptr;
        jb.sm1 = dg.sm1; jb.em1 = dg.em1; jb.sm2 = dg.sm2; jb.em2 = dg.em2;
        jb.nx = dg.nx(); jb.ny = dg.ny(); jb.nxw = (jb.nx + 31) / 32;
        jb.sd1 = h->sd1; jb.ed1 = h->ed1; jb.sd2 = h->sd2; jb.ed2 = h->ed2;
        // process_width, process_domain_module.F:2270-2278
        if (q0.s.process_bdy_width == 0) jb.pw = std::max(h->ed1 - h->sd1 + 1, h->ed2 - h->sd2 + 1);
        else { jb.pw = q0.s.process_bdy_width; if (q0.s.field_stagger != WPSGPU_M) jb.pw += 2; }
        jb.proj = to_dev(*q0.proj, fs.job_gauss[g]);
        jb.minx = q0.s.minx; jb.maxx = q0.s.maxx; jb.miny = q0.s.miny; jb.maxy = q0.s.maxy; jb.bdr = q0.s.bdr;
        size_t sbytes = (size_t)(jb.maxx - jb.minx + 1) * (jb.maxy - jb.miny + 1) * sizeof(float);
        for (int m = 0; m < 3; ++m) {
            WPS_TRY(stage_shared(q0.s.mask[m], sbytes, &jb.mask[m]));
            jb.mask_val[m] = q0.fo.mask_val[m];
            jb.mask_rel[m] = rel_code(q0.fo.mask_rel[m]);
        }
        memcpy(jb.ops, q0.fo.ops, sizeof(jb.ops));
        memcpy(jb.opts, q0.fo.opts, sizeof(jb.opts));
        jb.ops[WPSGPU_MAX_OPS - 1] = 0;
        jb.msg = q0.fo.missing_value; jb.fill = q0.fo.fill_missing; jb.masked = q0.fo.masked;
        jb.nslab = (int)groups[g].size(); jb.slab0 = slab_cursor;
        jb.tiles_x = (jb.nx + TILE_X - 1) / TILE_X; jb.tiles_y = (jb.ny + TILE_Y - 1) / TILE_Y;
        const Box& bx = fs.box[g];
        const int cls = bx.fast ? 1 : 0;
        jb.tile0 = tile_cursor[cls];
        cls_jobs[cls].push_back(j);
        cls_tile0[cls].push_back(tile_cursor[cls]);
        tile_cursor[cls] += jb.tiles_x * jb.tiles_y;
        cls_points[cls] += (double)dg.npts() * groups[g].size();
        if (bx.fast == 1) fast16_points += (double)dg.npts() * groups[g].size();
        if (bx.fast >= 2) fast4_points += (double)dg.npts() * groups[g].size();
        if (cls == 1) fast_pairs += (double)dg.npts() * groups[g].size();
        jb.f_negzero = -0.0f; jb.f_one = 1.0f;
        jb.gcell = q0.gcell ? 1 : 0;
        jb.fast = bx.fast; jb.bx0 = bx.bx0; jb.by0 = bx.by0; jb.bnx = bx.bnx; jb.bny = bx.bny; jb.npack = bx.npack;
        jb.strip = bx.fast == 1 ? bx.strip : 0;
        if (jb.fast == 1 && jb.strip > 0) {
            // window records: allocated per geometry class after this loop (alloc_strip_records)
        } else if (jb.fast == 1) {
            size_t cells = (size_t)jb.npack * jb.bnx * jb.bny;
            size_t cells8 = (size_t)jb.npack * ((jb.bnx + 7) / 8 * 8) * jb.bny;
            jb.rec = (const float4*)h->arena.take(cells8 * 6 * sizeof(float4));
            jb.cf = (const unsigned char*)h->arena.take(cells);
            WPS_CHECK(jb.rec && jb.cf, -3, "device arena exhausted (packed records)");
        }
        bool has_search = false;
        for (int k = 0; k < WPSGPU_MAX_OPS && jb.ops[k]; ++k) if (jb.ops[k] == WPSGPU_SEARCH) has_search = true;
        jb.defer_base = defer_cursor;
        if (has_search) { defer_cursor += (unsigned long long)dg.npts() * groups[g].size(); any_search = true; }
        size_t rbytes = dg.npts() * sizeof(float), wbytes = (size_t)jb.ny * jb.nxw * sizeof(int);
        const std::vector<int>& grp = groups[g];
        const int ns = (int)grp.size(), s0 = slab_cursor;
        slab_cursor += ns;
        // one pass per kind of array, so that the arrays of neighbouring levels end up next to each other
        for (int k = 0; k < ns; ++k) {
            SlabPtrs& p = sp[s0 + k];
            p.decided = nullptr; p.valid = nullptr; p.job = j; p.pad_ = 0;
            const wpsgpu_slab& s = queue[grp[k]].s;
            if (host && fs.staged.count(s.slab) && !fs.reusable(s.slab, fs.sub[g])) {   // staged for another grid with a smaller rectangle
                fs.staged.erase(s.slab);
                fs.staged_sub.erase(s.slab);
            }
            if (s.raw_nx == 0 && host && fs.sub[g].on && fs.staged.find(s.slab) == fs.staged.end()) {
                // full-size device slab (indexing unchanged), but only the rectangle the chain can touch is transferred
                const FlushShared::Sub& sb = fs.sub[g];
                float* d = (float*)h->arena.take(sbytes);
                WPS_CHECK(d != nullptr, -3, "device arena exhausted (source slab)");
                if (poison) WPS_CUDA(cudaMemsetAsync(d, 0xff, sbytes, s_in));   // testing aid: stray reads see NaN
                const size_t pitch = (size_t)(jb.maxx - jb.minx + 1) * sizeof(float);
                const size_t off = (size_t)(sb.y0 - jb.miny) * (jb.maxx - jb.minx + 1) + (sb.x0 - jb.minx);
                const size_t wbytes_row = (size_t)(sb.x1 - sb.x0 + 1) * sizeof(float), rows = (size_t)(sb.y1 - sb.y0 + 1);
                WPS_CUDA(cudaMemcpy2DAsync(d + off, pitch, s.slab + off, pitch, wbytes_row, rows, cudaMemcpyHostToDevice, s_in));
                h->h2d_bytes += (int64_t)(wbytes_row * rows);
                fs.staged[s.slab] = d;
                fs.staged_sub[s.slab] = sb;
                p.src = d;
                continue;
            }
            if (s.raw_nx == 0) { WPS_TRY(stage_shared(s.slab, sbytes, &p.src)); continue; }
            // raw record: bring the bytes over (host mode) and let k_decode_records build the slab
            auto it = fs.staged.find(s.slab);
            if (it != fs.staged.end()) {   // the same record queued again: share the decoded slab
                if (it->second) p.src = (const float*)it->second; else decode_pending.push_back(std::make_pair(s0 + k, (const void*)s.slab));
                continue;
            }
            const size_t rawbytes = (size_t)s.raw_nx * (jb.maxy - jb.miny + 1) * sizeof(float);
            const void* d_raw = s.slab;
            DecodeTask t;
            t.raw_nx = s.raw_nx; t.ny = jb.maxy - jb.miny + 1;
            t.bdr = (jb.maxx - jb.minx + 1 - s.raw_nx) / 2; t.big_endian = s.raw_big_endian;
            t.row0 = 0; t.nrows = t.ny;
            if (host && fs.sub[g].on) {
                // only the rows the chain can touch (whole rows: the halo columns wrap around to the other end of the record)
                t.row0 = fs.sub[g].y0 - jb.miny; t.nrows = fs.sub[g].y1 - fs.sub[g].y0 + 1;
                FlushShared::Sub rows = fs.sub[g];
                rows.x0 = jb.minx; rows.x1 = jb.maxx;          // whole rows are decoded
                fs.staged_sub[s.slab] = rows;
                char* d = (char*)h->arena.take(rawbytes);
                WPS_CHECK(d != nullptr, -3, "device arena exhausted (raw records)");
                const size_t off = (size_t)t.row0 * s.raw_nx * sizeof(float), len = (size_t)t.nrows * s.raw_nx * sizeof(float);
                WPS_CUDA(cudaMemcpyAsync(d + off, reinterpret_cast<const char*>(s.slab) + off, len, cudaMemcpyHostToDevice, s_in));
                h->h2d_bytes += (int64_t)len;
                d_raw = d;
            } else if (host) { d_raw = take_run(in_runs, s.slab, rawbytes); WPS_CHECK(d_raw != nullptr, -3, "device arena exhausted (raw records)"); }
            t.raw = (const unsigned*)d_raw;
            decode_tasks.push_back(t);
            decode_slots.push_back(s0 + k);
            fs.staged[s.slab] = nullptr;   // filled in once the decoded slabs are placed (below)
            decode_hosts.push_back(s.slab);
        }
        if (jb.gcell) {
            gc_decided[j] = (int*)h->arena.take(wbytes);
            WPS_CHECK(gc_decided[j] != nullptr, -3, "device arena exhausted (average_gcell bits)");
            sp[s0].decided = gc_decided[j];
        }
        if (host) {
            for (int k = 0; k < ns; ++k) {
                const wpsgpu_slab& s = queue[grp[k]].s;
                sp[s0 + k].r_arr = (float*)take_run(out_r_runs, s.r_arr, rbytes);
                WPS_CHECK(sp[s0 + k].r_arr != nullptr, -3, "device arena exhausted (outputs)");
            }
            for (int k = 0; k < ns; ++k) {
                const wpsgpu_slab& s = queue[grp[k]].s;
                sp[s0 + k].new_pts = (int*)take_run(out_np_runs, s.new_pts, wbytes);
                WPS_CHECK(sp[s0 + k].new_pts != nullptr, -3, "device arena exhausted (outputs)");
            }
            for (int k = 0; k < ns; ++k) {
                const wpsgpu_slab& s = queue[grp[k]].s;
                if (!s.valid_mask) continue;   // r_arr carries lower-priority data only where valid bits are set
                void* d_v = h->arena.take(wbytes);
                WPS_CHECK(d_v != nullptr, -3, "device arena exhausted (valid mask)");
                WPS_CUDA(cudaMemcpyAsync(sp[s0 + k].r_arr, s.r_arr, rbytes, cudaMemcpyHostToDevice, s_in));
                WPS_CUDA(cudaMemcpyAsync(d_v, s.valid_mask, wbytes, cudaMemcpyHostToDevice, s_in));
                h->h2d_bytes += (int64_t)(rbytes + wbytes);
                sp[s0 + k].valid = (const int*)d_v;
            }
        } else {
            for (int k = 0; k < ns; ++k) {
                const wpsgpu_slab& s = queue[grp[k]].s;
                sp[s0 + k].r_arr = s.r_arr; sp[s0 + k].valid = s.valid_mask; sp[s0 + k].new_pts = s.new_pts;
            }
        }
    }
    // decoded slabs are placed after the raw bytes so that the raw records of neighbouring levels stay adjacent
    size_t decode_max = 0;
    for (size_t k = 0; k < decode_tasks.size(); ++k) {
        DecodeTask& t = decode_tasks[k];
        const size_t n = (size_t)(t.raw_nx + 2 * t.bdr) * t.ny;
        t.out = (float*)h->arena.take(n * sizeof(float));
        WPS_CHECK(t.out != nullptr, -3, "device arena exhausted (decoded slabs)");
        if (poison && t.nrows < t.ny) WPS_CUDA(cudaMemsetAsync(t.out, 0xff, n * sizeof(float), h->stream));
        sp[decode_slots[k]].src = t.out;
        fs.staged[decode_hosts[k]] = t.out;
        decode_max = std::max(decode_max, n);
    }
    WPS_TRY(alloc_strip_records(h, jobs));
    for (const auto& pr : decode_pending) sp[pr.first].src = (const float*)fs.staged[pr.second];
    for (const CopyRun& r : in_runs) { WPS_CUDA(cudaMemcpyAsync(r.d, r.h, r.bytes, cudaMemcpyHostToDevice, s_in)); h->h2d_bytes += (int64_t)r.bytes; }
    // All tables of this pass travel in ONE copy: [jobs | slab pointers | six int lists of njobs+1 entries]
    std::vector<int> pack_jobs;
    std::vector<PackWinTask> win_tasks;     // (job, pack) of the tolerance-mode pre-pass
    int pack_cells = 0, pack_max = 0, win_tiles = 0;
    for (int j : cls_jobs[1])
        if (jobs[j].fast == 1 && jobs[j].strip > 0) {
            for (int pk = 0; pk < jobs[j].npack; ++pk) { PackWinTask t; t.job = j; t.pack = pk; win_tasks.push_back(t); }
            win_tiles = std::max(win_tiles, ((jobs[j].bnx + 15) / 16) * ((jobs[j].bny + 15) / 16));
        } else if (jobs[j].fast == 1) {
            pack_jobs.push_back(j);
            pack_cells = std::max(pack_cells, jobs[j].bnx * jobs[j].bny);
            pack_max = std::max(pack_max, jobs[j].npack);
        }
    WPS_CHECK(pack_jobs.size() < 65536, -3, "too many packed jobs in one flush");
    const size_t off_sp = (sizeof(Job) * njobs + 255) & ~size_t(255);
    const size_t off_cls = (off_sp + sizeof(SlabPtrs) * nq + 255) & ~size_t(255);
    const size_t off_dec = (off_cls + sizeof(int) * 6 * (njobs + 1) + 255) & ~size_t(255);
    const size_t off_win = (off_dec + sizeof(DecodeTask) * decode_tasks.size() + 255) & ~size_t(255);
    const size_t tab_bytes = off_win + sizeof(PackWinTask) * win_tasks.size();
    WPS_CHECK(win_tasks.size() < 65536, -3, "too many window-record packs in one flush");
    std::vector<char>& tab = fs.tab;
    tab.assign(tab_bytes, 0);
    memcpy(tab.data(), jobs.data(), sizeof(Job) * njobs);
    memcpy(tab.data() + off_sp, sp.data(), sizeof(SlabPtrs) * nq);
    {
        int* cls = reinterpret_cast<int*>(tab.data() + off_cls);
        const std::vector<int> unused;
        const std::vector<int>* lists[6] = { &cls_jobs[0], &cls_tile0[0], &cls_jobs[1], &cls_tile0[1], &unused, &pack_jobs };
        for (int c = 0; c < 6; ++c)
            if (!lists[c]->empty()) memcpy(cls + (size_t)c * (njobs + 1), lists[c]->data(), sizeof(int) * lists[c]->size());
    }
    if (!decode_tasks.empty()) memcpy(tab.data() + off_dec, decode_tasks.data(), sizeof(DecodeTask) * decode_tasks.size());
    if (!win_tasks.empty()) memcpy(tab.data() + off_win, win_tasks.data(), sizeof(PackWinTask) * win_tasks.size());
    char* d_tab = (char*)h->arena.take(tab_bytes);
    WPS_CHECK(d_tab != nullptr, -3, "device arena exhausted (tables)");
    WPS_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab_bytes, cudaMemcpyHostToDevice, s_in));   // pageable source: staged before the call returns
    Job* d_jobs = (Job*)d_tab;
    SlabPtrs* d_sp = (SlabPtrs*)(d_tab + off_sp);
    int* d_cls = (int*)(d_tab + off_cls);
    int* d_cls_jobs[2] = { d_cls, d_cls + 2 * (njobs + 1) };
    int* d_cls_tile0[2] = { d_cls + (njobs + 1), d_cls + 3 * (njobs + 1) };
    int* d_pack_jobs = d_cls + 5 * (njobs + 1);
    // slow queue: room for 1/32 of the first-method outputs + 1 M entries (a whole all-missing slab fits); beyond that
    // the pass falls back to the generic kernel
    SlowQ sq;
    sq.q = nullptr; sq.cnt = reinterpret_cast<unsigned*>(h->d_counters + 16); sq.cap = 0;
    if (tile_cursor[1] > 0) {
        sq.cap = (unsigned)std::min(fast_pairs, fast_pairs / 32.0 + 1048576.0);
        sq.q = (int2*)h->arena.take(sizeof(int2) * (size_t)sq.cap);
        WPS_CHECK(sq.q != nullptr, -3, "device arena exhausted (slow queue)");
    }
    Deferred *d_def = nullptr, *d_def2 = nullptr;
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(h->d_counters);
    if (any_search) {
        d_def = (Deferred*)h->arena.take(sizeof(Deferred) * defer_cursor);
        d_def2 = (Deferred*)h->arena.take(sizeof(Deferred) * defer_cursor);
        WPS_CHECK(d_def && d_def2, -3, "device arena exhausted (deferred lists)");
    }
    // inputs staged on s_in must be complete before the compute stream touches them
    if (s_in != h->stream) {
        WPS_CUDA(cudaEventRecord(h->ev_copy, s_in));
        WPS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
    }
    if (any_search) WPS_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), h->stream));
    if (tile_cursor[1] > 0) WPS_CUDA(cudaMemsetAsync(sq.cnt, 0, 2 * sizeof(unsigned), h->stream));

    // ---- raw records -> halo'd slabs (one launch for all records of this pass) ----
    if (!decode_tasks.empty()) {
        unsigned gx = (unsigned)std::min<size_t>((decode_max + 255) / 256, 4096);
        k_decode_records<<<dim3(gx, (unsigned)decode_tasks.size()), 256, 0, h->stream>>>((const DecodeTask*)(d_tab + off_dec));
        h->launches++;
        WPS_CUDA(cudaGetLastError());
    }

    // ---- average_gcell pre-pass: cell averages into r_arr + "decided" bits, one slab per job ----
    for (int j = 0; j < njobs; ++j) {
        if (!jobs[j].gcell) continue;
        const Job& jb = jobs[j];
        const QueuedSlab& q0 = queue[groups[gsel[j]][0]];
        const int snx = jb.maxx - jb.minx + 1, sny = jb.maxy - jb.miny + 1;
        const size_t npts = (size_t)jb.nx * jb.ny;
        int2* d_wm = (int2*)h->arena.take(sizeof(int2) * (size_t)snx * sny);
        double* d_sum = (double*)h->arena.take(sizeof(double) * npts);
        int* d_gcnt = (int*)h->arena.take(sizeof(int) * npts);
        WPS_CHECK(d_wm && d_sum && d_gcnt, -3, "device arena exhausted (average_gcell buffers)");
        WPS_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double) * npts, h->stream));
        WPS_CUDA(cudaMemsetAsync(d_gcnt, 0, sizeof(int) * npts, h->stream));
        const float* d_gauss_dom = nullptr;
        if (q0.dom_proj->code == WPSGPU_PROJ_GAUSS) WPS_TRY(upload_gauss(h, *q0.dom_proj, &d_gauss_dom));
        DevProj dom = to_dev(*q0.dom_proj, d_gauss_dom);
        dim3 b2(32, 8);
        k_gc_wm<<<dim3((snx + 31) / 32, (sny + 7) / 8), b2, 0, h->stream>>>(jb.proj, dom, jb.minx, jb.miny, snx, sny, jb.sm1, jb.em1, jb.sm2, jb.em2, d_wm);
        const int bw = jb.maxx - jb.minx + 1 - 2 * jb.bdr;
        if (bw > 0)
            k_gc_accum<<<dim3((bw + 31) / 32, (sny + 7) / 8), b2, 0, h->stream>>>(d_wm, sp[jb.slab0].src, jb.mask[0], jb.mask_rel[0], jb.mask_val[0], jb.msg,
                                                                                 jb.minx, jb.maxx, jb.miny, jb.maxy, jb.bdr, jb.sm1, jb.sm2, jb.nx, d_sum, d_gcnt);
        k_gc_apply<<<dim3(jb.nxw, jb.ny), 32, 0, h->stream>>>(d_jobs, j, d_sp, d_sum, d_gcnt, gc_decided[j]);
        h->launches += 3;
        WPS_CUDA(cudaGetLastError());
    }

    // ---- launches ----
    dim3 block(TILE_X, TILE_Y);
    // The four_pt / nearest_neighbor first-method jobs (a handful of 2-d fields, 20-40 us per launch, latency-bound) run on
    // a side stream beside the sixteen_pt kernels instead of after them: they only share the slow queue (atomics).
    const bool have16 = !pack_jobs.empty() || !win_tasks.empty();
    bool have4 = false;
    for (int j : cls_jobs[1]) have4 = have4 || jobs[j].fast >= 2;
    const bool side = have16 && have4 && env_int("WPSGPU_SIDE_STREAM", 1, 0, 1) != 0;
    cudaStream_t s4 = h->stream;
    if (side) {
        if (!h->s_side) {
            WPS_CUDA(cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking));
            WPS_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
            WPS_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        }
        WPS_CUDA(cudaEventRecord(h->ev_fork, h->stream));      // inputs staged, tables uploaded, counters cleared
        WPS_CUDA(cudaStreamWaitEvent(h->s_side, h->ev_fork, 0));
        s4 = h->s_side;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[6], s4));
    for (int j : cls_jobs[1]) {
        if (jobs[j].fast < 2) continue;
        Fast4Params fp;
        fp.jb = jobs[j];
        for (int k = 0; k < fp.jb.nslab; ++k) {
            const SlabPtrs& p = sp[fp.jb.slab0 + k];
            fp.src[k] = p.src; fp.r_arr[k] = p.r_arr; fp.new_pts[k] = p.new_pts;
        }
        fp.sq = sq;
        if (jobs[j].fast == 2) k_metgrid_fast4<WPSGPU_FOUR_POINT><<<fp.jb.tiles_x * fp.jb.tiles_y, block, 0, s4>>>(fp);
        else k_metgrid_fast4<WPSGPU_N_NEIGHBOR><<<fp.jb.tiles_x * fp.jb.tiles_y, block, 0, s4>>>(fp);
        h->launches++;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[7], s4));
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[0], h->stream));
    if (!pack_jobs.empty()) {
        k_pack16<<<dim3((pack_cells + 255) / 256, pack_max, (unsigned)pack_jobs.size()), 256, 0, h->stream>>>(d_jobs, d_pack_jobs, d_sp);
        h->launches++;
    }
    if (!win_tasks.empty()) {
        k_pack_win<<<dim3((unsigned)win_tiles, (unsigned)win_tasks.size()), 256, 0, h->stream>>>(d_jobs, (const PackWinTask*)(d_tab + off_win), d_sp);
        h->launches++;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[1], h->stream));
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[2], h->stream));
    for (int j : cls_jobs[1]) {
        if (jobs[j].fast != 1 || jobs[j].strip > 0) continue;
        FastParams fp;
        fp.jb = jobs[j];
        for (int k = 0; k < fp.jb.nslab; ++k) {
            const SlabPtrs& p = sp[fp.jb.slab0 + k];
            fp.r_arr[k] = p.r_arr; fp.new_pts[k] = p.new_pts;
        }
        fp.sq = sq;
        k_metgrid_fast16<<<fp.jb.tiles_x * ((fp.jb.ny + F16_Y - 1) / F16_Y), dim3(TILE_X, F16_Y), 0, h->stream>>>(fp);
        h->launches++;
    }
    {
        // tolerance mode: ONE strip launch per (points per thread, masked=both or not) takes every sixteen_pt job of the
        // pass, cut into chunks of up to 8 packs (32 levels) -- 1-level surface fields no longer cost a launch each
        std::vector<int> sj;
        for (int j : cls_jobs[1]) if (jobs[j].fast == 1 && jobs[j].strip > 0) sj.push_back(j);
        float* dump_r = nullptr;
        int* dump_w = nullptr;
        if (!sj.empty()) {
            size_t mp = 0, mw = 0;
            for (int j : sj) { mp = std::max(mp, (size_t)jobs[j].nx * jobs[j].ny); mw = std::max(mw, (size_t)jobs[j].nxw * jobs[j].ny); }
            dump_r = (float*)h->arena.take(mp * sizeof(float));
            dump_w = (int*)h->arena.take(mw * sizeof(int));
            WPS_CHECK(dump_r && dump_w, -3, "device arena exhausted (strip dump area)");
        }
        WPS_TRY(launch_strips(h, jobs, sj, sp, dump_r, dump_w, sq));
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[3], h->stream));
    if (side) {   // join: the slow pass and everything after it see the side stream's results
        WPS_CUDA(cudaEventRecord(h->ev_join, h->s_side));
        WPS_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[8], h->stream));
    if (tile_cursor[1] > 0) {
        k_metgrid_slowlist<<<h->sm_count * SLOW_MINBLOCKS, 128, 0, h->stream>>>(d_jobs, njobs, d_sp, sq, d_def, d_cnt);
        // queue overflow (cnt[1] set by a producer): the generic kernel redoes every first-method job of the pass
        k_metgrid_interp<<<std::min(tile_cursor[1], h->sm_count * 8), block, 0, h->stream>>>(d_jobs, d_cls_jobs[1], d_cls_tile0[1], (int)cls_jobs[1].size(),
                                                                                          tile_cursor[1], d_sp, d_def, d_cnt, sq.cnt + 1);
        h->launches += 2;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[9], h->stream));
    if (env_int("WPSGPU_DEBUG_SLOW", 0, 0, 1) && tile_cursor[1] > 0) {   // diagnostic: size of the slow queue
        unsigned c[2] = { 0, 0 };
        WPS_CUDA(cudaStreamSynchronize(h->stream));
        WPS_CUDA(cudaMemcpy(c, sq.cnt, sizeof(c), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[wpsgpu] slow queue: %u entries of %.0f first-method outputs (capacity %u, overflow %u)\n", c[0], fast_pairs, sq.cap, c[1]);
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[4], h->stream));
    if (tile_cursor[0] > 0) {
        k_metgrid_interp<<<tile_cursor[0], block, 0, h->stream>>>(d_jobs, d_cls_jobs[0], d_cls_tile0[0], (int)cls_jobs[0].size(), tile_cursor[0], d_sp, d_def, d_cnt, nullptr);
        h->launches++;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[5], h->stream));
    WPS_CUDA(cudaGetLastError());
    if (last) {
        h->last_pts3[0] = fast16_points; h->last_pts3[1] = fast16_points; h->last_pts3[2] = cls_points[0]; h->last_pts3[3] = fast4_points;
        h->last_pts3[4] = cls_points[1]; h->last_pts3[5] = (double)defer_cursor;
        // the dominant kernel for the roofline line: whichever main kernel produced more outputs
        int dom = fast16_points >= cls_points[0] ? 1 : 2;
        h->ev_k0 = h->ev_t[2 * dom]; h->ev_k1 = h->ev_t[2 * dom + 1];
        h->last_points = dom == 1 ? fast16_points : cls_points[0];
        h->timed_flush = h->timing != 0;
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[10], h->stream));
    if (any_search) {
        int sblocks = h->sm_count * 8;
        k_metgrid_search<<<sblocks, 256, 0, h->stream>>>(d_jobs, njobs, d_sp, d_def, d_cnt, d_def2, d_cnt + 1);
        h->launches++;
        WPS_CUDA(cudaGetLastError());
        // literal fallback: few threads, big per-thread scratch
        int R = std::min(std::max(fs.max_depth, 1), std::max(fs.max_span, 1));
        size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
        int qcap = 8 * R + 64;
        size_t words_per_thread = ((vis_words + 1) & ~size_t(1)) + (size_t)qcap * sizeof(QNode) / sizeof(unsigned);
        int lit_threads = 256;
        size_t bytes = words_per_thread * sizeof(unsigned) * lit_threads;
        if (bytes > h->search_scratch_bytes) {
            WPS_CUDA(cudaStreamSynchronize(h->stream));
            if (h->d_search_scratch) cudaFree(h->d_search_scratch);
            WPS_CUDA(cudaMalloc(&h->d_search_scratch, bytes));
            h->search_scratch_bytes = bytes;
        }
        k_metgrid_search_literal<<<lit_threads / 64, 64, 0, h->stream>>>(d_jobs, njobs, d_sp, d_def2, d_cnt + 1,
                                                                      (unsigned*)h->d_search_scratch, words_per_thread, R, qcap);
        h->launches++;
        WPS_CUDA(cudaGetLastError());
    }
    if (h->timing) WPS_CUDA(cudaEventRecord(h->ev_t[11], h->stream));
    if (host) {
        if (s_out != h->stream) {
            WPS_CUDA(cudaEventRecord(h->ev_done, h->stream));
            WPS_CUDA(cudaStreamWaitEvent(s_out, h->ev_done, 0));
        }
        for (const CopyRun& r : out_r_runs) { WPS_CUDA(cudaMemcpyAsync((void*)r.h, r.d, r.bytes, cudaMemcpyDeviceToHost, s_out)); h->d2h_bytes += (int64_t)r.bytes; }
        for (const CopyRun& r : out_np_runs) { WPS_CUDA(cudaMemcpyAsync((void*)r.h, r.d, r.bytes, cudaMemcpyDeviceToHost, s_out)); h->d2h_bytes += (int64_t)r.bytes; }
    }
    return 0;
}

static int flush_impl(wpsgpu_ctx* h, bool wait);
int wpsgpu_metgrid_flush(wpsgpu_handle_t h) { return flush_impl(h, true); }
int wpsgpu_metgrid_flush_async(wpsgpu_handle_t h) { return flush_impl(h, false); }

static int flush_impl(wpsgpu_ctx* h, bool wait)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    if (h->queue.empty()) return 0;
    WPS_CUDA(cudaSetDevice(h->device));
    WPS_TRY(arena_begin(h));   // an earlier asynchronous flush may still be copying results out of the arena
    std::vector<QueuedSlab> queue;
    queue.swap(h->queue);
    h->queue.reserve(queue.size());
    struct PoolClear { wpsgpu_ctx* h; ~PoolClear() { h->proj_pool.clear(); } } pool_clear{h};   // `queue` points into the pool
    const bool host = h->ptr_mode == WPSGPU_PTR_HOST;
    const int nq = (int)queue.size();

    // ---- group slabs into jobs (stable: first-appearance order) ----
    // host-pointer mode keeps the groups short so that copies and kernels of neighbouring groups overlap
    const int group_cap = host ? env_int("WPSGPU_PIPE_GROUP", 16, 1, FAST_MAX_SLABS) : FAST_MAX_SLABS;
    std::vector<std::vector<int>> groups;
    for (int a = 0; a < nq; ++a) {
        bool placed = false;
        for (size_t g = 0; g < groups.size() && !placed; ++g)
            if ((int)groups[g].size() < group_cap && same_group(queue[groups[g][0]], queue[a])) { groups[g].push_back(a); placed = true; }
        if (!placed) groups.push_back(std::vector<int>(1, a));
    }
    const int njobs = (int)groups.size();
    FlushShared fs;
    fs.queue = &queue; fs.groups = &groups;
    fs.job_xy.resize(njobs); fs.job_gauss.assign(njobs, nullptr); fs.box.resize(njobs);

    // ---- coordinate cache: (rx, ry) of every destination point per (grid, source projection), computed once ----
    std::vector<int> job_bbox(4 * njobs, 0), job_rbbox(4 * njobs, 0);
    std::vector<CoordCacheEntry*> job_entry(njobs, nullptr);
    for (int g = 0; g < njobs; ++g) {
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        std::string key = cache_key(q0.s.grid_id, *q0.proj);
        auto it = h->coord_cache.find(key);
        if (it == h->coord_cache.end() || it->second.grid_version != dg.version) {
            CoordCacheEntry e;
            e.grid_version = dg.version;
            e.d_gauss = nullptr;
            WPS_CUDA(cudaMalloc((void**)&e.d_xy, dg.npts() * sizeof(float2)));
            if (q0.proj->code == WPSGPU_PROJ_GAUSS) {
                WPS_CUDA(cudaMalloc((void**)&e.d_gauss, sizeof(float) * WPSGPU_MAX_GAUSS));
                WPS_CUDA(cudaMemcpyAsync(e.d_gauss, q0.proj->gauss_lat, sizeof(float) * WPSGPU_MAX_GAUSS, cudaMemcpyHostToDevice, h->stream));
            }
            DevProj dp = to_dev(*q0.proj, e.d_gauss);
            int64_t n = (int64_t)dg.npts();
            k_coord_cache<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dp, dg.d_xlat, dg.d_xlon, n, e.d_xy);
            // source-cell bounding box of this destination grid (read back once per cached coordinate set)
            int init_bbox[4] = { INT_MAX, INT_MIN, INT_MAX, INT_MIN };
            int* d_bbox = h->d_counters + 8;
            WPS_CUDA(cudaMemcpyAsync(d_bbox, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice, h->stream));
            k_xy_bbox<<<h->sm_count * 4, 256, 0, h->stream>>>(e.d_xy, n, d_bbox);
            WPS_CUDA(cudaMemcpyAsync(e.bbox, d_bbox, sizeof(init_bbox), cudaMemcpyDeviceToHost, h->stream));
            k_xy_bbox_retry<<<h->sm_count * 4, 256, 0, h->stream>>>(dp, dg.d_xlat, dg.d_xlon, n, d_bbox);   // grows the same box
            h->launches += 3;
            WPS_CUDA(cudaGetLastError());
            WPS_CUDA(cudaMemcpyAsync(e.rbbox, d_bbox, sizeof(init_bbox), cudaMemcpyDeviceToHost, h->stream));
            WPS_CUDA(cudaStreamSynchronize(h->stream));
            if (it != h->coord_cache.end()) { free_coord_entry(it->second); h->coord_cache.erase(it); }
            it = h->coord_cache.insert(std::make_pair(key, e)).first;
        }
        fs.job_xy[g] = it->second.d_xy;
        fs.job_gauss[g] = it->second.d_gauss;
        job_entry[g] = &it->second;
        memcpy(&job_bbox[4 * g], it->second.bbox, 4 * sizeof(int));
        memcpy(&job_rbbox[4 * g], it->second.rbbox, 4 * sizeof(int));
    }

    // ---- classify jobs and size the arena ----
    // first-method kernels: chain starts with sixteen_pt / four_pt / nearest_neighbor (the latter two only when the
    // landmask logic selects a single interp mask, i.e. not masked=both); sixteen_pt additionally needs the
    // source-cell bounding box to be small against the destination grid (source coarser than the domain).
    const bool subrect_enabled = env_int("WPSGPU_SUBRECT", 1, 0, 1) != 0;
    fs.sub.resize(njobs);
    size_t need = (size_t)njobs * (sizeof(Job) + 256) + (size_t)nq * sizeof(SlabPtrs) + (1 << 20);
    size_t defer_entries = 0;
    std::map<const void*, size_t> uniq_in;  // host source/mask pointers staged once
    for (int g = 0; g < njobs; ++g) {
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        size_t sbytes = (size_t)(q0.s.maxx - q0.s.minx + 1) * (q0.s.maxy - q0.s.miny + 1) * sizeof(float);
        bool has_search = false;
        for (int k = 0; k < WPSGPU_MAX_OPS && q0.fo.ops[k]; ++k)
            if (q0.fo.ops[k] == WPSGPU_SEARCH) { has_search = true; fs.max_depth = std::max(fs.max_depth, q0.fo.opts[k]); }
        if (has_search) {
            defer_entries += dg.npts() * groups[g].size();
            fs.max_span = std::max(fs.max_span, std::max(q0.s.maxx - q0.s.minx, q0.s.maxy - q0.s.miny));
        }
        // Host-pointer mode: a chain without `search` reads the source within 2 cells of the cell a destination point
        // (or its lon+360 retry position) falls into, and the packed pre-pass within 6 cells of the bounding box; only
        // that rectangle (+8 cells) of the slab has to cross PCIe.  `search` can walk anywhere, average_gcell visits
        // every source cell: those slabs are copied whole.
        FlushShared::Sub& sb = fs.sub[g];
        sb.on = false;
        if (host && !has_search && !q0.gcell && subrect_enabled) {
            const int* rb = &job_rbbox[4 * g];
            if (rb[0] <= rb[1] && rb[2] <= rb[3]) {
                const int margin = 8;
                long x0 = std::max<long>(q0.s.minx, (long)rb[0] - margin), x1 = std::min<long>(q0.s.maxx, (long)rb[1] + margin);
                long y0 = std::max<long>(q0.s.miny, (long)rb[2] - margin), y1 = std::min<long>(q0.s.maxy, (long)rb[3] + margin);
                const double full = (double)(q0.s.maxx - q0.s.minx + 1) * (q0.s.maxy - q0.s.miny + 1);
                if (x1 >= x0 && y1 >= y0 && (double)(x1 - x0 + 1) * (double)(y1 - y0 + 1) < 0.7 * full) {
                    sb.on = true; sb.x0 = (int)x0; sb.x1 = (int)x1; sb.y0 = (int)y0; sb.y1 = (int)y1;
                }
            }
        }
        Box& b = fs.box[g];
        b.fast = 0; b.bx0 = b.by0 = b.bnx = b.bny = b.npack = 0;
        b.strip = 0;
        const int* bb = &job_bbox[4 * g];
        bool simple_mask = !(q0.s.use_landmask && q0.fo.masked == WPSGPU_MASKED_BOTH);
        if (q0.gcell) {   // where_maps_to, sums, counts, decided bits
            need += (size_t)(q0.s.maxx - q0.s.minx + 1) * (q0.s.maxy - q0.s.miny + 1) * sizeof(int2) + dg.npts() * (sizeof(double) + sizeof(int)) +
                    (size_t)dg.ny() * ((dg.nx() + 31) / 32) * sizeof(int) + 2048;
        } else if (!h->disable_fast && (simple_mask || q0.fo.ops[0] == WPSGPU_SIXTEEN_POINT)) {
            // (missing_value == 0 stays with the exact kernel: a weighted sum can be 0, and the strip kernel carries no such test)
            if (q0.fo.ops[0] == WPSGPU_SIXTEEN_POINT && bb[0] <= bb[1] && bb[2] <= bb[3] && h->precision == WPSGPU_PRECISION_TOLERANCE &&
                q0.fo.missing_value != 0.0f && (size_t)(bb[1] - bb[0] + 7) * (size_t)(bb[3] - bb[2] + 7) * 2 <= dg.npts()) {
                // window records over the source cells the grid's points fall into, +-2 / +4 for the 5 x 5 windows; not
                // clipped to the source array (the pre-pass clamps columns and rows as interp_module.F:1227-1241 does)
                b.fast = 1; b.bx0 = bb[0] - 2; b.by0 = bb[2] - 2; b.bnx = bb[1] - bb[0] + 7; b.bny = bb[3] - bb[2] + 7;
                b.npack = ((int)groups[g].size() + 3) / 4;
                // points per thread: a strip must stay within two source cells in each direction
                const double ppc = sqrt((double)dg.npts() / ((double)(bb[1] - bb[0] + 1) * (bb[3] - bb[2] + 1)));
                b.strip = ppc >= 5.0 ? 4 : (ppc >= 2.5 ? 2 : 1);
                b.strip = env_int("WPSGPU_STRIP", b.strip, 1, 4) == 3 ? 2 : env_int("WPSGPU_STRIP", b.strip, 1, 4);
                need += (size_t)b.npack * b.bnx * b.bny * (sizeof(float4) + sizeof(uint2)) + 1024;
                need += dg.npts() * sizeof(float) + (size_t)dg.ny() * ((dg.nx() + 31) / 32) * sizeof(int) + 512;   // dump area (per pass at most once)
            } else if (q0.fo.ops[0] == WPSGPU_SIXTEEN_POINT && bb[0] <= bb[1] && bb[2] <= bb[3]) {
                long x0 = std::max<long>(q0.s.minx, (long)bb[0] - 1), x1 = std::min<long>(q0.s.maxx, (long)bb[1] + 2);
                long y0 = std::max<long>(q0.s.miny, (long)bb[2] - 2), y1 = std::min<long>(q0.s.maxy, (long)bb[3] + 4);
                if (x1 >= x0 && y1 >= y0 && (size_t)(x1 - x0 + 1) * (size_t)(y1 - y0 + 1) * 2 <= dg.npts()) {
                    b.fast = 1; b.bx0 = (int)x0; b.by0 = (int)y0; b.bnx = (int)(x1 - x0 + 1); b.bny = (int)(y1 - y0 + 1);
                    b.npack = ((int)groups[g].size() + 3) / 4;
                    need += (size_t)b.npack * (b.bnx + 8) * b.bny * (6 * sizeof(float4) + 1) + 1024;
                }
            } else if (q0.fo.ops[0] == WPSGPU_FOUR_POINT) b.fast = 2;
            else if (q0.fo.ops[0] == WPSGPU_N_NEIGHBOR) b.fast = 3;
            if (b.fast) need += (size_t)((double)groups[g].size() * dg.npts() / 32.0) * sizeof(int2) + 256;  // share of the slow queue
        }
        if (host) {
            for (int m = 0; m < 3; ++m) if (q0.s.mask[m]) uniq_in[q0.s.mask[m]] = sbytes;
            for (int a : groups[g]) {
                uniq_in[queue[a].s.slab] = queue[a].s.raw_nx > 0 ? (size_t)queue[a].s.raw_nx * (q0.s.maxy - q0.s.miny + 1) * sizeof(float) : sbytes;
                need += dg.npts() * sizeof(float) + 256;                                     // r_arr
                need += 2 * ((size_t)dg.ny() * ((dg.nx() + 31) / 32) * sizeof(int) + 256);  // valid + new_pts
            }
        }
    }
    for (auto& kv : uniq_in) need += kv.second + 256;
    if (host) {
        // a host array that several groups use (a slab shared by two grids, a slab that is also a mask) can be staged again
        // in full when the first copy only held a rectangle: budget one more copy per additional user
        std::map<const void*, int> users;
        for (int g = 0; g < njobs; ++g) {
            const QueuedSlab& q0 = queue[groups[g][0]];
            std::vector<const void*> seen;
            for (int m = 0; m < 3; ++m) if (q0.s.mask[m]) seen.push_back(q0.s.mask[m]);
            for (int a : groups[g]) seen.push_back(queue[a].s.slab);
            std::sort(seen.begin(), seen.end());
            seen.erase(std::unique(seen.begin(), seen.end()), seen.end());
            for (const void* p : seen) users[p]++;
        }
        for (auto& kv : users)
            if (kv.second > 1) {
                auto it = uniq_in.find(kv.first);
                if (it != uniq_in.end()) need += (size_t)(kv.second - 1) * (it->second + 256);
            }
    }
    for (int a = 0; a < nq; ++a)   // decoded copies of raw records (both pointer modes) + their task entries
        if (queue[a].s.raw_nx > 0)
            need += (size_t)(queue[a].s.maxx - queue[a].s.minx + 1) * (queue[a].s.maxy - queue[a].s.miny + 1) * sizeof(float) + 256 + sizeof(DecodeTask);
    need += 2 * (defer_entries * sizeof(Deferred) + 256) + (size_t)njobs * 4096;
    const int chunk_slabs = env_int("WPSGPU_PIPE_CHUNK", 12, 1, 1 << 20);
    {
        // the slow queue's fixed 1 M entries, once per pass (host-pointer mode runs the groups in several passes)
        int npass = 1;
        if (host) {
            npass = 0;
            for (int g = 0, count = 0; g < njobs; ++g) {
                count += (int)groups[g].size();
                if (count >= chunk_slabs || g == njobs - 1) { ++npass; count = 0; }
            }
        }
        need += (size_t)npass * ((size_t(1) << 23) + 4096);
    }
    h->arena.reset();
    WPS_TRY(h->arena.reserve(need, h->stream));

    if (!host) {
        std::vector<int> all(njobs);
        for (int g = 0; g < njobs; ++g) all[g] = g;
        return run_groups(h, fs, all, h->stream, h->stream, true);
    }
    // Host-pointer mode: pipeline the job groups through three streams so that the H2D copies of group k+1, the kernels of
    // group k and the D2H copies of group k-1 overlap (PCIe is full duplex and is the end-to-end limiter).
    if (!h->s_in) {
        WPS_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        WPS_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        WPS_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
        WPS_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
        WPS_CUDA(cudaEventCreateWithFlags(&h->ev_out_done, cudaEventDisableTiming));
    }
    // everything previously enqueued on the handle's stream (set_grid copies, earlier calls) precedes the new copies
    WPS_CUDA(cudaEventRecord(h->ev_done, h->stream));
    WPS_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_done, 0));
    const bool trace = env_int("WPSGPU_PIPE_TRACE", 0, 0, 1) != 0;   // print where the pipeline's time goes
    cudaEvent_t tr[4] = { nullptr, nullptr, nullptr, nullptr };
    const auto host_t0 = std::chrono::steady_clock::now();
    if (trace) {
        for (int k = 0; k < 4; ++k) WPS_CUDA(cudaEventCreate(&tr[k]));
        WPS_CUDA(cudaEventRecord(tr[0], h->s_in));
    }
    std::vector<int> sel;
    int count = 0;
    for (int g = 0; g < njobs; ++g) {
        sel.push_back(g);
        count += (int)groups[g].size();
        if (count >= chunk_slabs || g == njobs - 1) {
            WPS_TRY(run_groups(h, fs, sel, h->s_in, h->s_out, g == njobs - 1));
            sel.clear();
            count = 0;
        }
    }
    const double issue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    if (trace) {
        WPS_CUDA(cudaEventRecord(tr[1], h->s_in));
        WPS_CUDA(cudaEventRecord(tr[2], h->stream));
        WPS_CUDA(cudaEventRecord(tr[3], h->s_out));
    }
    if (wait || trace) {
        WPS_CUDA(cudaStreamSynchronize(h->s_out));
        WPS_CUDA(cudaStreamSynchronize(h->stream));
    } else {   // results are complete after wpsgpu_synchronize; later work on the handle is ordered behind the copies
        WPS_CUDA(cudaEventRecord(h->ev_out_done, h->s_out));
        h->async_pending = true;
    }
    if (trace) {
        float a = 0.0f, b = 0.0f, c = 0.0f;
        cudaEventElapsedTime(&a, tr[0], tr[1]); cudaEventElapsedTime(&b, tr[0], tr[2]); cudaEventElapsedTime(&c, tr[0], tr[3]);
        fprintf(stderr, "[wpsgpu] pipeline: %d groups, host issue %.2f ms, H2D done %.2f ms, kernels done %.2f ms, D2H done %.2f ms\n", njobs, issue_ms, a, b, c);
        for (int k = 0; k < 4; ++k) cudaEventDestroy(tr[k]);
    }
    return 0;
}

int wpsgpu_metgrid_set_precision(wpsgpu_handle_t h, int mode)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    WPS_CHECK(mode == WPSGPU_PRECISION_EXACT || mode == WPSGPU_PRECISION_TOLERANCE, -1, "unknown precision mode %d", mode);
    h->precision = mode;
    return 0;
}

int wpsgpu_metgrid_set_timing(wpsgpu_handle_t h, int on)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    h->timing = on ? 1 : 0;
    return 0;
}

int wpsgpu_debug_disable_fast(wpsgpu_handle_t h, int disable)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    h->disable_fast = disable != 0;
    return 0;
}

}  // extern "C"
