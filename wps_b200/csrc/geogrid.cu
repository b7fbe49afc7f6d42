// geogrid.cu -- calc_field (process_tile_module.F:1208-1681) on device: per-tile accumulation with the
// reference's quadtree "four corners agree" rule (proc_point_module.F:218-458,553-779), per-point fallback
// (get_point, :788-855), post-processing, and the raw tile decode of read_geogrid.c.
#include <algorithm>
#include <climits>
#include "ctx.h"
#include "interp.cuh"

namespace wps {

#define GEO_OUTSIDE 100000000  // OUTSIDE_DOMAIN, misc_definitions_module.F:18
// where_maps_to of a source pixel: destination (x, y) as two 16-bit halves of one word (a field dimension beyond +-32767
// is refused at field_begin), so that "four corners agree" is one comparison per pair and the array is half the size
typedef int WmCell;
constexpr WmCell WM_OUTSIDE = 0x00008000;          // x = -32768, y = 0
__host__ __device__ __forceinline__ WmCell wm_pack(int x, int y) { return (WmCell)(((unsigned)y << 16) | ((unsigned)x & 0xffffu)); }
__host__ __device__ __forceinline__ int wm_x(WmCell w) { return (int)(short)(w & 0xffff); }
__host__ __device__ __forceinline__ int wm_y(WmCell w) { return w >> 16; }
constexpr int GEO_SETS = 10;   // staging buffer sets: one batch (GEO_NB = 8 tiles) + two tiles being staged for the next

struct GeoState {
    bool field_open = false, level_open = false;
    wpsgpu_geo_field f;
    wpsgpu_geo_level l;
    int nx = 0, ny = 0, nk = 0, nxw = 0;
    size_t npts = 0;
    float *d_field = nullptr, *d_xlat = nullptr, *d_xlon = nullptr, *d_landmask = nullptr, *d_count = nullptr;
    int *d_processed = nullptr, *d_level = nullptr;
    float2* d_src_xy = nullptr;       // lltoxy(SOURCE_PROJ) of every destination point (get_point / is_point_in_tile)
    int* d_acc_cnt = nullptr;         // per-tile accumulation: counts [nk][ny][nx]
    double* d_acc_sum = nullptr;      // per-tile accumulation: sums (continuous)
    int* d_touched = nullptr;         // per-tile: cell was reached by a credit attempt
    unsigned long long* d_invalid = nullptr;
    // Two sets of tile / where_maps_to buffers: in device-pointer mode the decode and where_maps_to kernels of tile k+1
    // run on s_pre while tile k is accumulated, merged and point-filled on the handle's stream.
    // GEO_SETS sets of tile / where_maps_to buffers: the decode and where_maps_to kernels of the coming tiles run on s_pre
    // while earlier tiles are accumulated, merged and point-filled on the handle's stream; a batch holds up to GEO_NB tiles.
    float* d_tile[GEO_SETS] = {}; size_t tile_cap[GEO_SETS] = {};
    WmCell* d_wm[GEO_SETS] = {}; size_t wm_cap[GEO_SETS] = {};
    cudaStream_t s_pre = nullptr;
    cudaEvent_t ev_pre[GEO_SETS] = {}, ev_used[GEO_SETS] = {};
    bool used_rec[GEO_SETS] = {};
    int parity = 0;
    void* d_raw[GEO_SETS] = {}; size_t raw_cap[GEO_SETS] = {};   // host-pointer mode: the file bytes of a tile
    struct Pending { int set, smx, sMx, smy, sMy, smz, sMz, bdr; const unsigned char* raw; int wordsize, is_signed, little, flip; };
    std::vector<Pending> pending;     // tiles staged (decoded, mapped) but not yet accumulated: the current batch
    int* d_first_tile = nullptr; size_t cap_first = 0;
    int* d_dstart = nullptr;          // per tile buffer set: first level of the block hierarchy with a uniform block (k_geo_dstart)
    float* d_terms[GEO_SETS] = {}; size_t terms_cap[GEO_SETS] = {};     // per set: separable where_maps_to terms, [tny] + 2 [tnx]
    int2* d_bounds[GEO_SETS] = {}; size_t bounds_cap[GEO_SETS] = {};   // per set: level-dstart block ranges of the tile's columns, then rows
    cudaEvent_t ev_h2d = nullptr;    // host-pointer mode: the caller's tile buffer has been read
    Deferred* d_def = nullptr; Deferred* d_def2 = nullptr;
    void* d_lit = nullptr; size_t lit_bytes = 0;
    // buffers kept across fields (grow-only): a geogrid run processes dozens of fields on the same patch
    size_t cap_count = 0, cap_acc_cnt = 0, cap_acc_sum = 0, cap_touched = 0, cap_bits = 0, cap_xy = 0, cap_def = 0, cap_blk = 0;
    float4* d_blk_xy = nullptr;       // per 256-point block: min/max of src_xy (lets the per-tile point pass skip far blocks)
    int* d_cand = nullptr; size_t cap_cand = 0;   // points of the current tile that need the interpolation chain
    int* d_tbbox = nullptr;           // destination rows/columns credited by the current tile: {jmin, imin, jmax, imax}
    const float* d_gauss_src = nullptr; const float* d_gauss_dom = nullptr;
    bool own_field = false, own_ll = false;
    bool has_search = false; int max_depth = 0;
};

struct GeoK {  // kernel-side view
    DevProj src, dom;
    int istagger, itype;
    int start_i, end_i, start_j, end_j, start_k, end_k, nx, ny, nk, nxw;
    size_t npts;
    float sr_x, sr_y, msg_val, msg_fill_val, mask_val;
    float* field; float* count; const float* landmask;
    int* processed; int* level;
    int* acc_cnt; double* acc_sum; int* touched; unsigned long long* invalid;
    int* tbbox; const float4* blk_xy;
    unsigned long long* counters;   // [0] deferred to the warp pass, [1] to the literal pass, [2] candidates of the point pass
    // current tile
    const float* tile; int smx, sMx, smy, sMy, smz, sMz, bdr, tnx, tny;
    WmCell* wm;
    int ops[WPSGPU_MAX_OPS], opts[WPSGPU_MAX_OPS];
};

__device__ __forceinline__ bool bit_test(const int* w, int nxw, int ii, int jj) { return (w[(size_t)jj * nxw + (ii >> 5)] >> (ii & 31)) & 1; }
__device__ __forceinline__ void bit_set_atomic(int* w, int nxw, int ii, int jj) { atomicOr(w + (size_t)jj * nxw + (ii >> 5), (int)(1u << (ii & 31))); }

// per-tile state consumed by the kernels that follow on the handle's stream (k_geo_wm of the NEXT tile may already be
// running on the other stream, so it cannot do this)
__global__ void k_geo_tile_reset(int* tbbox, unsigned long long* counters)
{
    tbbox[0] = INT_MAX; tbbox[1] = INT_MAX; tbbox[2] = INT_MIN; tbbox[3] = INT_MIN;
    counters[0] = 0ull; counters[1] = 0ull; counters[2] = 0ull;
}

// where_maps_to for every pixel of the tile (proc_point_module.F:245-310).  One block covers WM_T x WM_T pixels.
// For a regular lat-lon source read onto a Lambert or polar-stereographic domain the projection separates into a
// per-row term (latitude) and a per-column term (longitude): those 2*WM_T evaluations are shared by the WM_T^2
// pixels of the block through shared memory -- same expressions as llij_lc / llij_ps, so the same bits.
constexpr int WM_T = 64;
struct GeoWmTile { int smx, smy, tnx, tny; WmCell* wm; const float* terms; };
// (the tile comes as a small separate struct: assigning tile fields into a copy of the kernel's GeoK parameter makes every
// thread copy the whole parameter block to local memory)
__device__ __forceinline__ void geo_wm_body(const GeoK& g, const GeoWmTile& w)
{
    __shared__ float s_rm[WM_T], s_sin[WM_T], s_cos[WM_T];
    const int ti0 = blockIdx.x * WM_T, tj0 = blockIdx.y * WM_T, t = threadIdx.x;
    const bool separable = g.src.code == WPSGPU_PROJ_LATLON && (g.dom.code == WPSGPU_PROJ_LC || g.dom.code == WPSGPU_PROJ_PS);
    if (separable && w.terms) {   // computed once per tile row / column
        if (t < WM_T) { if (tj0 + t < w.tny) s_rm[t] = w.terms[tj0 + t]; }
        else if (t < 2 * WM_T) {
            const int c = t - WM_T;
            if (ti0 + c < w.tnx) { s_sin[c] = w.terms[w.tny + ti0 + c]; s_cos[c] = w.terms[w.tny + w.tnx + ti0 + c]; }
        }
        __syncthreads();
    } else if (separable) {
        if (t < WM_T) {   // row terms: ijll_latlon gives lat from j alone
            float lat, lon;
            xytoll(g.src, (float)(w.smx + ti0), (float)(w.smy + tj0 + t), lat, lon, WPSGPU_M);
            s_rm[t] = g.dom.code == WPSGPU_PROJ_LC ? llij_lc_rm(lat, g.dom) : llij_ps_rm(lat, g.dom);
        } else if (t < 2 * WM_T) {   // column terms: lon from i alone
            int c = t - WM_T;
            float lat, lon, sn, cs;
            xytoll(g.src, (float)(w.smx + ti0 + c), (float)(w.smy + tj0), lat, lon, WPSGPU_M);
            if (g.dom.code == WPSGPU_PROJ_LC) llij_lc_sc(lon, g.dom, sn, cs); else llij_ps_sc(lon, g.dom, sn, cs);
            s_sin[c] = sn; s_cos[c] = cs;
        }
        __syncthreads();
    }
    for (int e = t; e < WM_T * WM_T; e += 256) {
        int li = e % WM_T, lj = e / WM_T, ti = ti0 + li, tj = tj0 + lj;
        if (ti >= w.tnx || tj >= w.tny) continue;
        float rx, ry;
        if (separable) {
            if (g.dom.code == WPSGPU_PROJ_LC) llij_lc_combine(s_rm[lj], s_sin[li], s_cos[li], g.dom, rx, ry);
            else llij_ps_combine(s_rm[lj], s_sin[li], s_cos[li], g.dom, rx, ry);
            if (g.istagger == WPSGPU_U) rx = rx + 0.5f;
            else if (g.istagger == WPSGPU_V) ry = ry + 0.5f;
            else if (g.istagger == WPSGPU_CORNER) { rx = rx + 0.5f; ry = ry + 0.5f; }
        } else {
            float lat, lon;
            xytoll(g.src, (float)(w.smx + ti), (float)(w.smy + tj), lat, lon, WPSGPU_M);
            lltoxy(g.dom, lat, lon, rx, ry, g.istagger);
        }
        rx = (rx - 0.5f) * g.sr_x + 0.5f;
        ry = (ry - 0.5f) * g.sr_y + 0.5f;
        WmCell wv = WM_OUTSIDE;
        if ((float)g.start_i <= rx && rx <= (float)g.end_i && (float)g.start_j <= ry && ry <= (float)g.end_j) wv = wm_pack(f_nint(rx), f_nint(ry));
        w.wm[(size_t)tj * w.tnx + ti] = wv;
    }
}
__global__ void __launch_bounds__(256) k_geo_wm(GeoK g)
{
    GeoWmTile w;
    w.smx = g.smx; w.smy = g.smy; w.tnx = g.tnx; w.tny = g.tny; w.wm = g.wm; w.terms = nullptr;
    geo_wm_body(g, w);
}

// Quadtree emulation: walk the block hierarchy of process_*_block top-down for this pixel and credit the
// first block whose four corners agree, else the pixel's own cell at the <=2x2 leaf.
// Tile-specific part of the kernel arguments: a batch launch takes several of these (blockIdx.z picks one).
struct GeoTile {
    const float* tile; const WmCell* wm; int smx, sMx, smy, sMy, smz, sMz, bdr, tnx, tny; const int* dstart; const int2* xb; const int2* yb; float* terms;
    // categorical batches read the file's bytes directly (read_geogrid.c:91-128 + the row flip); `tile` is then decoded only if
    // the batch's point pass has candidates (k_geo_decode_b with a gate)
    const unsigned char* raw; int rws, rsigned, rlittle, rflip;
};
__device__ __forceinline__ float geo_raw_value(const GeoTile& t, int ix, int jy)
{
    const unsigned e = (unsigned)(t.rflip ? t.tny - 1 - jy : jy) * (unsigned)t.tnx + (unsigned)ix;     // one level of a tile: far below 2^31 bytes
    const unsigned char* b = t.raw + e * (unsigned)t.rws;
    int ival;
    if (t.rws == 1) { ival = b[0]; if (t.rsigned && ival > (1 << 7)) ival -= (1 << 8); }
    else if (t.rws == 2) { ival = t.rlittle ? ((b[1] << 8) | b[0]) : ((b[0] << 8) | b[1]); if (t.rsigned && ival > (1 << 15)) ival -= (1 << 16); }
    else if (t.rws == 3) { ival = t.rlittle ? ((b[2] << 16) | (b[1] << 8) | b[0]) : ((b[0] << 16) | (b[1] << 8) | b[2]); if (t.rsigned && ival > (1 << 23)) ival -= (1 << 24); }
    else { const unsigned u = t.rlittle ? (((unsigned)b[3] << 24) | (b[2] << 16) | (b[1] << 8) | b[0]) : (((unsigned)b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]); ival = (int)u; }
    return (float)ival;
}

// Shallowest level of the tile's block hierarchy (process_*_block, proc_point_module.F:218-458) that holds a block whose four
// corners map to one destination cell, looked for among the levels 0 .. GEO_DCAP (21845 blocks at most).  With a source much
// finer than the domain there is none up there: every pixel's walk then skips the corner tests -- and their loads from the
// far ends of the where_maps_to array -- of those levels.  One thread per (level, block); a block that the recursion cannot
// reach (its parent is a <= 2 x 2 leaf, or a right / upper half that does not exist) is skipped.
constexpr int GEO_DCAP = 7;
constexpr int GEO_DNODES = 21845;      // (4^(GEO_DCAP+1) - 1) / 3
__device__ __forceinline__ void geo_dstart_body(const WmCell* __restrict__ wm, int smx, int smy, int tnx, int I0, int I1, int J0, int J1, int* __restrict__ dstart)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= GEO_DNODES || I0 > I1 || J0 > J1) return;
    int d = 0, off = 0;
    while (off + (1 << (2 * d)) <= n) { off += 1 << (2 * d); ++d; }
    const int local = n - off, a = local & ((1 << d) - 1), b = local >> d;
    int min_i = I0, max_i = I1, min_j = J0, max_j = J1;
    for (int st = d - 1; st >= 0; --st) {
        if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) return;      // the parent is a leaf
        const int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
        if ((a >> st) & 1) { if (ci >= max_i) return; min_i = ci + 1; } else max_i = ci;
        if ((b >> st) & 1) { if (cj >= max_j) return; min_j = cj + 1; } else max_j = cj;
    }
    // a <= 2 x 2 leaf ends the recursion like a uniform block does: above the level found here every block splits in both
    // directions, so the x and y ranges of a pixel's block can be followed independently (k_geo_bounds)
    if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) { atomicMin(dstart, d); return; }
    auto W = [&](int i, int j) { return wm[(size_t)(j - smy) * tnx + (i - smx)]; };
    const WmCell p = W(min_i, min_j), q = W(max_i, max_j);
    if (p == q && p != WM_OUTSIDE) {
        const WmCell r = W(min_i, max_j), t = W(max_i, min_j);
        if (p == r && p == t) atomicMin(dstart, d);
    }
}
constexpr int GEO_NB = 8;     // tiles per batch
struct GeoBatch { GeoTile t[GEO_NB]; int n; };

__global__ void __launch_bounds__(256) k_geo_dstart(const WmCell* __restrict__ wm, int smx, int smy, int tnx, int I0, int I1, int J0, int J1, int* __restrict__ dstart)
{
    geo_dstart_body(wm, smx, smy, tnx, I0, I1, J0, J1, dstart);
}

// Range [lo, hi] of the level-dstart block that holds each pixel column / row of the tile's interior: the walk of
// process_*_block down to that level without looking at where_maps_to (no block above it is uniform or a leaf).
__device__ __forceinline__ void geo_bounds_body(const int* __restrict__ dstart, int I0, int I1, int J0, int J1, int2* __restrict__ xb, int2* __restrict__ yb)
{
    const int d0 = min(*dstart, GEO_DCAP + 1);
    const int nxi = I1 - I0 + 1, nyi = J1 - J0 + 1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nxi + nyi) return;
    const bool isx = e < nxi;
    const int c = isx ? I0 + e : J0 + (e - nxi);
    int lo = isx ? I0 : J0, hi = isx ? I1 : J1;
    for (int d = 0; d < d0; ++d) {
        const int m = (hi + lo) / 2;
        if (c <= m) hi = m; else lo = m + 1;
    }
    (isx ? xb[e] : yb[e - nxi]) = make_int2(lo, hi);
}
__global__ void __launch_bounds__(256) k_geo_bounds(const int* __restrict__ dstart, int I0, int I1, int J0, int J1, int2* __restrict__ xb, int2* __restrict__ yb)
{
    geo_bounds_body(dstart, I0, I1, J0, J1, xb, yb);
}

// Categorical accumulation of a batch of tiles (accum_categorical, proc_point_module.F:100-216, + process_categorical_block
// :218-458): a block takes 32 x 32 pixels of one tile (4 rows per thread).  Every pixel walks the block hierarchy from level
// dstart (k_geo_dstart / k_geo_bounds) to the first block whose corners agree or to its <= 2 x 2 leaf; the credits are
// counted in a SHARED-MEMORY histogram over the destination cells the block's pixels map to ([category + 1 plane of
// credit attempts][cell of the block's bounding box], integer atomics) and flushed once: one global atomic per (cell,
// category) the block met instead of one per pixel group.  Blocks whose bounding box does not fit (a domain much finer than
// the source) fall back to global atomics per pixel.
constexpr int GEO_HCAP = 4096;      // shared-memory counters per block
__global__ void __launch_bounds__(256) k_geo_accum_cat_b(GeoK g, GeoBatch bt, int* __restrict__ first_tile)
{
    const GeoTile& t = bt.t[blockIdx.z];
    const int I0 = t.smx + t.bdr, I1 = t.sMx - t.bdr, J0 = t.smy + t.bdr, J1 = t.sMy - t.bdr;
    if (I0 + (int)blockIdx.x * 32 > I1 || J0 + (int)blockIdx.y * 32 > J1) return;
    __shared__ int s_bb[4];          // {jmin, imin, jmax, imax} of the cells this block credits
    __shared__ int s_hist[GEO_HCAP];
    const int tid = threadIdx.y * 32 + threadIdx.x, order = (int)blockIdx.z;
    if (tid == 0) { s_bb[0] = INT_MAX; s_bb[1] = INT_MAX; s_bb[2] = INT_MIN; s_bb[3] = INT_MIN; }
    __syncthreads();
    const int d0 = min(__ldg(t.dstart), GEO_DCAP + 1);
    const int tnx = t.tnx;
    const WmCell* __restrict__ wm = t.wm - ((size_t)t.smy * tnx + t.smx);     // wm[j * tnx + i] in tile coordinates
    const int pi = I0 + blockIdx.x * 32 + threadIdx.x;
    const bool colok = pi <= I1;
    const int2 xr = colok ? __ldg(t.xb + (pi - I0)) : make_int2(0, 0);
    int cell[4];                     // destination cell of the thread's four pixels, -1 = no credit
    int cat[4];                      // category index (>= 0), -1 = missing value, -2 = invalid category
    int jmn = INT_MAX, imn = INT_MAX, jmx = INT_MIN, imx = INT_MIN;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int pj = J0 + blockIdx.y * 32 + threadIdx.y + 8 * r;
        cell[r] = -1; cat[r] = -1;
        if (!colok || pj > J1) continue;
        const int2 yr = __ldg(t.yb + (pj - J0));
        int min_i = xr.x, max_i = xr.y, min_j = yr.x, max_j = yr.y;
        WmCell dest;
        for (;;) {
            // four corners agree?  The diagonal pair is compared first: where it differs the other two need not be fetched.
            const WmCell a = wm[min_j * tnx + min_i], c = wm[max_j * tnx + max_i];
            if (a == c && a != WM_OUTSIDE) {
                const WmCell b = wm[max_j * tnx + min_i], d = wm[min_j * tnx + max_i];
                if (a == b && a == d) { dest = a; break; }
            }
            if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) { dest = wm[pj * tnx + pi]; break; }
            const int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
            if (pi <= ci) max_i = ci; else min_i = ci + 1;
            if (pj <= cj) max_j = cj; else min_j = cj + 1;
        }
        if (dest == WM_OUTSIDE) continue;
        const int ii = wm_x(dest) - g.start_i, jj = wm_y(dest) - g.start_j;
        if (bit_test(g.processed, g.nxw, ii, jj)) continue;
        cell[r] = (jj << 16) | ii;      // row and column of the destination cell (both < 65536)
        jmn = min(jmn, jj); jmx = max(jmx, jj); imn = min(imn, ii); imx = max(imx, ii);
        const float v = t.raw ? geo_raw_value(t, pi - t.smx, pj - t.smy) : t.tile[(size_t)(pj - t.smy) * tnx + (pi - t.smx)];
        const int cv = (int)v;
        if (v != g.msg_val) cat[r] = (cv >= g.start_k && cv <= g.end_k) ? cv - g.start_k : -2;
    }
    (void)d0;
    {   // bounding box of the credited cells: one shared-memory update per warp
        jmn = __reduce_min_sync(0xffffffffu, jmn); imn = __reduce_min_sync(0xffffffffu, imn);
        jmx = __reduce_max_sync(0xffffffffu, jmx); imx = __reduce_max_sync(0xffffffffu, imx);
        if (threadIdx.x == 0 && jmn != INT_MAX) { atomicMin(&s_bb[0], jmn); atomicMin(&s_bb[1], imn); atomicMax(&s_bb[2], jmx); atomicMax(&s_bb[3], imx); }
    }
    __syncthreads();
    if (s_bb[0] == INT_MAX) return;                 // the block credits nothing (warp-uniform: shared value)
    const int bj0 = s_bb[0], bi0 = s_bb[1], bw = s_bb[3] - bi0 + 1, bh = s_bb[2] - bj0 + 1;
    const int ncell = bw * bh, nplane = g.nk + 1;   // plane nk counts the credit attempts ("touched")
    const bool use_hist = (long long)ncell * nplane <= GEO_HCAP;
    if (use_hist) {
        for (int e = tid; e < ncell * nplane; e += 256) s_hist[e] = 0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (cell[r] < 0) continue;
            const int jj = cell[r] >> 16, ii = cell[r] & 0xffff;
            const int lc = (jj - bj0) * bw + (ii - bi0);
            atomicAdd(&s_hist[g.nk * ncell + lc], 1);
            if (cat[r] >= 0) atomicAdd(&s_hist[cat[r] * ncell + lc], 1);
            else if (cat[r] == -2) atomicAdd(g.invalid, 1ull);
        }
        __syncthreads();
        // counter index -> (plane, row, column) by multiplication with rounded-up reciprocals: exact for e < 4096 (GEO_HCAP)
        const unsigned rc_n = 0xffffffffu / (unsigned)ncell + 1u, rc_w = 0xffffffffu / (unsigned)bw + 1u;
        for (int e = tid; e < ncell * nplane; e += 256) {
            const int v = s_hist[e];
            if (v == 0) continue;
            const int k = ncell == 1 ? e : (int)__umulhi((unsigned)e, rc_n), lc = e - k * ncell;       // (the reciprocal of 1 does not fit)
            const int lr = bw == 1 ? lc : (int)__umulhi((unsigned)lc, rc_w);
            const int jj = bj0 + lr, ii = bi0 + (lc - lr * bw);
            const size_t gc = (size_t)jj * g.nx + ii;
            if (k == g.nk) g.touched[gc] = 1;
            else {
                atomicAdd(g.acc_cnt + (size_t)k * g.npts + gc, v);
                if (__ldcg(first_tile + gc) > order) atomicMin(first_tile + gc, order);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (cell[r] < 0) continue;
            const int gcell = (cell[r] >> 16) * g.nx + (cell[r] & 0xffff);
            g.touched[gcell] = 1;
            if (cat[r] >= 0) {
                atomicAdd(g.acc_cnt + (size_t)cat[r] * g.npts + gcell, 1);
                if (__ldcg(first_tile + gcell) > order) atomicMin(first_tile + gcell, order);
            } else if (cat[r] == -2) atomicAdd(g.invalid, 1ull);
        }
    }
    // rows / columns this batch reaches (k_geo_merge_b skips the rest of the domain)
    if (tid < 4) {
        const int v = s_bb[tid];
        const int cur = __ldcg(g.tbbox + tid);
        if (tid < 2) { if (v < cur) atomicMin(g.tbbox + tid, v); }
        else if (v > cur) atomicMax(g.tbbox + tid, v);
    }
}

// order / first_tile: batch launches record, per destination cell, the first tile of the batch (in the caller's order) that
// actually credited it -- the per-point pass needs to know whether a cell already held data when "its" tile was processed
__device__ __forceinline__ void geo_accum_body(GeoK g, const GeoTile& t, int order, int* __restrict__ first_tile)
{
    g.tile = t.tile; g.wm = const_cast<WmCell*>(t.wm);
    g.smx = t.smx; g.sMx = t.sMx; g.smy = t.smy; g.sMy = t.sMy; g.smz = t.smz; g.sMz = t.sMz; g.bdr = t.bdr; g.tnx = t.tnx; g.tny = t.tny;
    __shared__ int s_bb[4];   // {jmin, imin, jmax, imax} of the cells this block credits
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) { s_bb[0] = INT_MAX; s_bb[1] = INT_MAX; s_bb[2] = INT_MIN; s_bb[3] = INT_MIN; }
    __syncthreads();
    int I0 = g.smx + g.bdr, I1 = g.sMx - g.bdr, J0 = g.smy + g.bdr, J1 = g.sMy - g.bdr;
    int pi = I0 + blockIdx.x * blockDim.x + threadIdx.x, pj = J0 + blockIdx.y * blockDim.y + threadIdx.y;
    bool credit = pi <= I1 && pj <= J1;
    int ii = 0, jj = 0;
    if (credit) {
        auto W = [&](int i, int j) { return g.wm[(size_t)(j - g.smy) * g.tnx + (i - g.smx)]; };
        int min_i = I0, max_i = I1, min_j = J0, max_j = J1;
        WmCell dest;
        // levels above d0 hold no block whose corners agree (k_geo_dstart): the walk only narrows the rectangle there
        const int d0 = t.dstart ? min(__ldg(t.dstart), GEO_DCAP + 1) : 0;
        for (int lev = 0;; ++lev) {
            // four corners agree?  The diagonal pair is compared first: where it differs the other two corners need
            // not be fetched.
            if (lev >= d0) {
                const WmCell a = W(min_i, min_j), c = W(max_i, max_j);
                if (a == c && a != WM_OUTSIDE) {
                    const WmCell b = W(min_i, max_j), d = W(max_i, min_j);
                    if (a == b && a == d) { dest = a; break; }
                }
            }
            if ((max_i - min_i + 1) <= 2 && (max_j - min_j + 1) <= 2) { dest = W(pi, pj); break; }
            int ci = (max_i + min_i) / 2, cj = (max_j + min_j) / 2;
            if (pi <= ci) max_i = ci; else min_i = ci + 1;
            if (pj <= cj) max_j = cj; else min_j = cj + 1;
        }
        credit = dest != WM_OUTSIDE;
        if (credit) {
            ii = wm_x(dest) - g.start_i; jj = wm_y(dest) - g.start_j;
            credit = !bit_test(g.processed, g.nxw, ii, jj);
        }
    }
    if (credit) {
        size_t cell = (size_t)jj * g.nx + ii;
        g.touched[cell] = 1;
        {   // one shared-memory update per group of lanes that arrive here together
            const unsigned act = __activemask();
            const int jmn = __reduce_min_sync(act, jj), imn = __reduce_min_sync(act, ii);
            const int jmx = __reduce_max_sync(act, jj), imx = __reduce_max_sync(act, ii);
            if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1)) {
                atomicMin(&s_bb[0], jmn); atomicMin(&s_bb[1], imn); atomicMax(&s_bb[2], jmx); atomicMax(&s_bb[3], imx);
            }
        }
        size_t tn2 = (size_t)g.tnx * g.tny, toff = (size_t)(pj - g.smy) * g.tnx + (pi - g.smx);
        if (g.itype == WPSGPU_CATEGORICAL) {
            float v = g.tile[toff];
            int cat = (int)v;
            const bool has = v != g.msg_val, ok = has && cat >= g.start_k && cat <= g.end_k;
            if (has && !ok) atomicAdd(g.invalid, 1ull);
            // neighbouring pixels mostly fall into the same cell with the same category: one atomic per distinct
            // (category, cell) among the lanes that arrive here together, carrying their count
            const unsigned lane = threadIdx.x & 31;
            unsigned long long key = ok ? ((unsigned long long)(cat - g.start_k) * g.npts + cell) : (~0ull - lane);
            unsigned peers = __match_any_sync(__activemask(), key);
            if (ok && lane == (unsigned)(__ffs(peers) - 1)) {
                atomicAdd(g.acc_cnt + key, __popc(peers));
                if (first_tile && __ldcg(first_tile + cell) > order) atomicMin(first_tile + cell, order);
            }
        } else {
            bool any = false;
            for (int k = g.smz; k <= g.sMz; ++k) {
                float v = g.tile[(size_t)(k - g.smz) * tn2 + toff];
                if (v != g.msg_val) {
                    atomicAdd(g.acc_sum + (size_t)(k - g.start_k) * g.npts + cell, (double)v);
                    atomicAdd(g.acc_cnt + (size_t)(k - g.start_k) * g.npts + cell, 1);
                    any = true;
                }
            }
            if (any && first_tile && __ldcg(first_tile + cell) > order) atomicMin(first_tile + cell, order);
        }
    }
    // rows / columns this tile reaches (k_geo_merge skips the rest of the domain): one update per block, and only
    // when it would change the running range
    __syncthreads();
    if (tid < 4 && s_bb[0] != INT_MAX) {
        int v = s_bb[tid];
        int cur = __ldcg(g.tbbox + tid);
        if (tid < 2) { if (v < cur) atomicMin(g.tbbox + tid, v); }
        else if (v > cur) atomicMax(g.tbbox + tid, v);
    }
}

__device__ __forceinline__ GeoTile tile_of(const GeoK& g)
{
    GeoTile t;
    t.tile = g.tile; t.wm = g.wm; t.smx = g.smx; t.sMx = g.sMx; t.smy = g.smy; t.sMy = g.sMy; t.smz = g.smz; t.sMz = g.sMz;
    t.bdr = g.bdr; t.tnx = g.tnx; t.tny = g.tny; t.dstart = nullptr; t.xb = nullptr; t.yb = nullptr; t.terms = nullptr;
    t.raw = nullptr; t.rws = 0; t.rsigned = 0; t.rlittle = 0; t.rflip = 0;
    return t;
}
__global__ void __launch_bounds__(256) k_geo_accum(GeoK g) { geo_accum_body(g, tile_of(g), 0, nullptr); }
// batch: blockIdx.z = tile of the batch; the grid covers the largest tile
__global__ void __launch_bounds__(256) k_geo_accum_b(GeoK g, GeoBatch bt, int* __restrict__ first_tile)
{
    const GeoTile& t = bt.t[blockIdx.z];
    const int inx = t.tnx - 2 * t.bdr, iny = t.tny - 2 * t.bdr;
    if ((int)(blockIdx.x * blockDim.x) >= inx || (int)(blockIdx.y * blockDim.y) >= iny) return;
    geo_accum_body(g, t, (int)blockIdx.z, first_tile);
}

// Batch selection of the per-point pass (before the batch's accumulations are merged): the point belongs to the tile that
// contains it (is_point_in_tile, proc_point_module.F:911-929); it takes the interpolation chain if it holds no value of
// this level yet and no tile that comes BEFORE its own in the batch credited its cell -- exactly the points the
// sequential order (accumulate tile, merge, fill points, next tile) would hand to get_point.  A cell that a LATER tile
// credits keeps its chain value and the merge adds to it, as in the reference.
__global__ void __launch_bounds__(256)
k_geo_select_b(GeoK g, GeoBatch bt, const float2* __restrict__ src_xy, const int* __restrict__ first_tile, int* __restrict__ cand,
               unsigned long long* __restrict__ counters)
{
    {   // block-level rejection as in k_geo_merge_select: no tile of the batch can contain a point of this block
        const float4 b = g.blk_xy[blockIdx.x];
        bool any = false;
        for (int z = 0; z < bt.n; ++z) {
            const GeoTile& t = bt.t[z];
            any = any || (t.smx + t.bdr <= f_floor(b.y + 0.5f) && f_ceiling(b.x - 0.5f) <= t.sMx - t.bdr &&
                          t.smy + t.bdr <= f_floor(b.w + 0.5f) && f_ceiling(b.z - 0.5f) <= t.sMy - t.bdr);
        }
        if (!any) return;
    }
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= g.npts) return;
    const float2 xy = src_xy[cell];
    int owner = -1;
    for (int z = 0; z < bt.n && owner < 0; ++z) {
        const GeoTile& t = bt.t[z];
        if (t.smx + t.bdr <= f_floor(xy.x + 0.5f) && f_ceiling(xy.x - 0.5f) <= t.sMx - t.bdr &&
            t.smy + t.bdr <= f_floor(xy.y + 0.5f) && f_ceiling(xy.y - 0.5f) <= t.sMy - t.bdr) owner = z;
    }
    if (owner < 0) return;
    const int ii = (int)(cell % g.nx), jj = (int)(cell / g.nx);
    if (bit_test(g.level, g.nxw, ii, jj)) return;
    if (first_tile[cell] <= owner) return;          // its own tile or an earlier one gives the cell data: level bit set by then
    if (bit_test(g.processed, g.nxw, ii, jj)) { bit_set_atomic(g.level, g.nxw, ii, jj); return; }
    cand[(size_t)owner * g.npts + atomicAdd(counters + 3 * owner + 2, 1ull)] = (int)cell;
}

// Batch merge: k_geo_merge_select's first half for all tiles of the batch at once (runs after the batch's point pass, so a
// cell the chain has just filled is "not new" and the accumulation is added to it).  ilevel == 1 or continuous data only:
// the half-area rule of overlay levels needs the tile-by-tile order.
__global__ void __launch_bounds__(256) k_geo_merge_b(GeoK g, int* __restrict__ first_tile, int k0, int k1)
{
    {
        size_t c0 = (size_t)blockIdx.x * blockDim.x, c1 = min(c0 + blockDim.x - 1, g.npts - 1);
        int r0 = (int)(c0 / g.nx), r1 = (int)(c1 / g.nx);
        if (r1 < g.tbbox[0] || r0 > g.tbbox[2]) return;   // no cell of this block was reached by the batch
    }
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= g.npts || !g.touched[cell]) return;
    const int ii = (int)(cell % g.nx), jj = (int)(cell / g.nx);
    g.touched[cell] = 0;
    first_tile[cell] = INT_MAX;
    const bool isnew = bit_test(g.level, g.nxw, ii, jj);
    bool any = false;
    constexpr int CH = 8;
    if (g.itype == WPSGPU_CATEGORICAL) {
        for (int k0 = 0; k0 < g.nk; k0 += CH) {
            int c[CH]; float f[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                size_t o = (size_t)(k0 + u) * g.npts + cell;
                bool in = k0 + u < g.nk;
                c[u] = in ? g.acc_cnt[o] : 0;
                f[u] = (in && isnew) ? g.field[o] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                if (k0 + u >= g.nk) break;
                size_t o = (size_t)(k0 + u) * g.npts + cell;
                if (c[u] != 0) g.acc_cnt[o] = 0;
                g.field[o] = f[u] + (float)c[u];
                any |= c[u] > 0;
            }
        }
    } else {
        for (int kk = k0; kk <= k1; ++kk) {      // the level range every tile of the batch carries (src_min_z .. src_max_z)
            size_t o = (size_t)kk * g.npts + cell;
            const int c = g.acc_cnt[o];
            const double sm = g.acc_sum[o];
            g.acc_cnt[o] = 0; g.acc_sum[o] = 0.0;
            const float f = isnew ? g.field[o] : 0.0f, n = g.count[o];
            g.field[o] = f + (float)sm;
            g.count[o] = n + (float)c;
            any |= c > 0;
        }
    }
    if (any) bit_set_atomic(g.level, g.nxw, ii, jj);
}

// Fold the tile's accumulation into field / data_count with the reference's first-touch semantics
// (proc_point_module.F:326-346) and, for ilevel>1 categorical data, the half-area rule (:165-202).
// The same launch then lists the points of this tile that still have no value (is_point_in_tile, :911-929, and the
// level / processed tests of calc_field, :1410-1420) for the interpolation-chain kernel: a cell's level bit is only
// changed by its own thread here, so the decision uses the thread's own view of it.
__global__ void __launch_bounds__(256)
k_geo_merge_select(GeoK g, int ilevel, float src_area, float dom_area, int do_merge, const float2* __restrict__ src_xy,
                   int* __restrict__ cand, unsigned long long* __restrict__ cand_cnt)
{
    const bool all_cells = ilevel > 1 && g.itype == WPSGPU_CATEGORICAL;   // the half-area rule looks at every cell
    bool blk_merge = do_merge != 0;
    if (blk_merge && !all_cells) {
        size_t c0 = (size_t)blockIdx.x * blockDim.x, c1 = min(c0 + blockDim.x - 1, g.npts - 1);
        int r0 = (int)(c0 / g.nx), r1 = (int)(c1 / g.nx);
        if (r1 < g.tbbox[0] || r0 > g.tbbox[2]) blk_merge = false;   // no cell of this block was reached by the tile
    }
    bool blk_select;
    {   // is_point_in_tile is monotone in each coordinate: if it fails for the block's extreme coordinates it fails for all
        float4 b = g.blk_xy[blockIdx.x];   // {xmin, xmax, ymin, ymax}
        blk_select = g.smx + g.bdr <= f_floor(b.y + 0.5f) && f_ceiling(b.x - 0.5f) <= g.sMx - g.bdr &&
                     g.smy + g.bdr <= f_floor(b.w + 0.5f) && f_ceiling(b.z - 0.5f) <= g.sMy - g.bdr;
    }
    if (!blk_merge && !blk_select) return;
    size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= g.npts) return;
    int ii = (int)(cell % g.nx), jj = (int)(cell / g.nx);
    bool levbit = bit_test(g.level, g.nxw, ii, jj);
    if (blk_merge && (all_cells || (ii >= g.tbbox[1] && ii <= g.tbbox[3]))) {
        if (g.touched[cell]) {
            g.touched[cell] = 0;
            bool isnew = levbit;
            bool any = false;
            // loads of a chunk of levels are issued together before the stores (the arrays may alias as far as the
            // compiler knows, so a load-store-load sequence would pay one memory latency per level)
            constexpr int CH = 8;
            if (g.itype == WPSGPU_CATEGORICAL) {
                for (int k0 = 0; k0 < g.nk; k0 += CH) {
                    int c[CH]; float f[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        size_t o = (size_t)(k0 + u) * g.npts + cell;
                        bool in = k0 + u < g.nk;
                        c[u] = in ? g.acc_cnt[o] : 0;
                        f[u] = (in && isnew) ? g.field[o] : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        if (k0 + u >= g.nk) break;
                        size_t o = (size_t)(k0 + u) * g.npts + cell;
                        if (c[u] != 0) g.acc_cnt[o] = 0;
                        g.field[o] = f[u] + (float)c[u];
                        any |= c[u] > 0;
                    }
                }
            } else {
                int k0 = g.smz - g.start_k, k1 = g.sMz - g.start_k;
                for (int kk = k0; kk <= k1; kk += CH) {
                    int c[CH]; double sm[CH]; float f[CH], n[CH];
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        size_t o = (size_t)(kk + u) * g.npts + cell;
                        bool in = kk + u <= k1;
                        c[u] = in ? g.acc_cnt[o] : 0;
                        sm[u] = in ? g.acc_sum[o] : 0.0;
                        f[u] = (in && isnew) ? g.field[o] : 0.0f;
                        n[u] = in ? g.count[o] : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < CH; ++u) {
                        if (kk + u > k1) break;
                        size_t o = (size_t)(kk + u) * g.npts + cell;
                        g.acc_cnt[o] = 0; g.acc_sum[o] = 0.0;
                        g.field[o] = f[u] + (float)sm[u];
                        g.count[o] = n[u] + (float)c[u];
                        any |= c[u] > 0;
                    }
                }
            }
            if (any) { bit_set_atomic(g.level, g.nxw, ii, jj); levbit = true; }
        }
        if (all_cells) {
            if (levbit && !bit_test(g.processed, g.nxw, ii, jj)) {
                float rarea = 0.0f;
                for (int k = 0; k < g.nk; ++k) rarea = rarea + g.field[(size_t)k * g.npts + cell];
                rarea = rarea * src_area;
                if (dom_area > 2.0f * rarea) {
                    for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + cell] = 0.0f;
                    atomicAnd(g.level + (size_t)jj * g.nxw + (ii >> 5), (int)~(1u << (ii & 31)));
                    levbit = false;
                }
            }
        }
    }
    if (!blk_select) return;
    float2 xy = src_xy[cell];
    if (!(g.smx + g.bdr <= f_floor(xy.x + 0.5f) && f_ceiling(xy.x - 0.5f) <= g.sMx - g.bdr &&
          g.smy + g.bdr <= f_floor(xy.y + 0.5f) && f_ceiling(xy.y - 0.5f) <= g.sMy - g.bdr)) return;
    if (levbit) return;
    if (bit_test(g.processed, g.nxw, ii, jj)) { bit_set_atomic(g.level, g.nxw, ii, jj); return; }
    cand[atomicAdd(cand_cnt, 1ull)] = (int)cell;
}

// source-grid coordinates of every destination point (get_point / is_point_in_tile, :811-823,911-921)
// Also records, per block of 256 points, the range of those coordinates: a tile's point pass can then drop whole
// blocks whose points all lie outside the tile with one 16-byte load.
__global__ void __launch_bounds__(256)
k_geo_src_xy(DevProj src, const float* __restrict__ xlat, const float* __restrict__ xlon, size_t n, float2* __restrict__ xy,
             float4* __restrict__ blk_xy)
{
    __shared__ float s_red[4][8];
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    if (p < n) {
        float rlat = xlat[p], rlon = xlon[p], rx, ry;
        if (rlon >= 180.0f) rlon = rlon - 360.0f;
        lltoxy(src, rlat, rlon, rx, ry, WPSGPU_M);
        xy[p] = make_float2(rx, ry);
        // NaN coordinates fail every in-tile test, so they may be left out of the range (fminf/fmaxf drop them)
        xmin = xmax = rx; ymin = ymax = ry;
    }
    for (int off = 16; off > 0; off >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, off)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, off)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    }
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_red[0][w] = xmin; s_red[1][w] = xmax; s_red[2][w] = ymin; s_red[3][w] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            xmin = fminf(xmin, s_red[0][k]); xmax = fmaxf(xmax, s_red[1][k]); ymin = fminf(ymin, s_red[2][k]); ymax = fmaxf(ymax, s_red[3][k]);
        }
        // an all-NaN / empty block keeps (+inf, -inf) and is skipped by every tile; a NaN seen first is replaced by fminf/fmaxf
        blk_xy[blockIdx.x] = make_float4(xmin, xmax, ymin, ymax);
    }
}

// inner loop of calc_field for one destination point (process_tile_module.F:1408-1519)
template <int MODE>
__device__ __forceinline__ bool geo_point(const GeoK& g, const float2* __restrict__ src_xy, size_t p, const LitScratch* ls)
{
    int ii = (int)(p % g.nx), jj = (int)(p / g.nx);
    float2 xy = src_xy[p];
    // is_point_in_tile (proc_point_module.F:928-929)
    if (!(g.smx + g.bdr <= f_floor(xy.x + 0.5f) && f_ceiling(xy.x - 0.5f) <= g.sMx - g.bdr &&
          g.smy + g.bdr <= f_floor(xy.y + 0.5f) && f_ceiling(xy.y - 0.5f) <= g.sMy - g.bdr)) return false;
    if (bit_test(g.level, g.nxw, ii, jj)) return false;
    bool lead = (MODE != 1) || ((threadIdx.x & 31) == 0);
    if (bit_test(g.processed, g.nxw, ii, jj)) { if (lead) bit_set_atomic(g.level, g.nxw, ii, jj); return false; }
    bool lm_ok = true;
    if (g.landmask && (g.istagger == WPSGPU_M || g.istagger == WPSGPU_HH)) lm_ok = (g.landmask[p] != g.mask_val);
    Src s;
    s.m = nullptr; s.sx = g.smx; s.ex = g.sMx; s.sy = g.smy; s.ey = g.sMy; s.nx = g.tnx;
    s.msg = g.msg_val; s.maskval = 0.0f; s.rel = REL_EQ;
    size_t tn2 = (size_t)g.tnx * g.tny;
    if (g.itype != WPSGPU_CATEGORICAL) {
        if (lm_ok) {
            float vals[1];
            // first pass: evaluate (may defer before anything is written)
            bool setbit = false;
            for (int k = g.smz; k <= g.sMz; ++k) {
                s.a = g.tile + (size_t)(k - g.smz) * tn2;
                float temp;
                int st = interp_sequence<MODE>(s, g.ops, g.opts, xy.x, xy.y, temp, ls);
                if (st == ST_DEFER) return true;
                vals[0] = temp;
                if (lead) {
                    size_t o = (size_t)(k - g.start_k) * g.npts + p;
                    if (vals[0] != g.msg_val) {
                        g.field[o] = vals[0];
                        setbit = true;
                        if (g.itype == WPSGPU_SP_CONTINUOUS) g.count[o] = 1.0f;
                    } else g.field[o] = g.msg_fill_val;
                }
            }
            if (lead && setbit) bit_set_atomic(g.level, g.nxw, ii, jj);
        } else if (lead) {
            for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + p] = g.msg_fill_val;
        }
    } else {
        if (lm_ok) {
            s.a = g.tile;
            float temp;
            int st = interp_sequence<MODE>(s, g.ops, g.opts, xy.x, xy.y, temp, ls);
            if (st == ST_DEFER) return true;
            if (lead) {
                for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + p] = 0.0f;
                if (temp != g.msg_val) {
                    int c = (int)temp;
                    if (c >= g.start_k && c <= g.end_k) g.field[(size_t)(c - g.start_k) * g.npts + p] += 1.0f;
                    else atomicAdd(g.invalid, 1ull);
                    bit_set_atomic(g.level, g.nxw, ii, jj);
                }
            }
        } else if (lead) {
            for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + p] = 0.0f;
        }
    }
    return false;
}

// The per-point pass of a tile: k_geo_merge_select has listed the points that lie in the tile and have no value yet;
// k_geo_points runs the interpolation chain (a 220-register kernel) over that usually short list.
__global__ void __launch_bounds__(256)
k_geo_points(GeoK g, const float2* __restrict__ src_xy, const int* __restrict__ cand, const unsigned long long* __restrict__ cand_cnt,
             Deferred* def, unsigned long long* def_cnt)
{
    const unsigned long long n = *cand_cnt;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
        size_t p = (size_t)cand[e];
        if (geo_point<0>(g, src_xy, p, nullptr)) {
            unsigned long long slot = atomicAdd(def_cnt, 1ull);
            Deferred d; d.slab = 0; d.idx = (int)p;
            def[slot] = d;
        }
    }
}
__global__ void k_geo_points_warp(GeoK g, const float2* __restrict__ src_xy, const Deferred* def, const unsigned long long* def_cnt,
                                  Deferred* def2, unsigned long long* def2_cnt)
{
    unsigned long long n = *def_cnt;
    unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long e = warp; e < n; e += nw) {
        Deferred d = def[e];
        bool again = geo_point<1>(g, src_xy, (size_t)d.idx, nullptr);
        if (again && (threadIdx.x & 31) == 0) { unsigned long long slot = atomicAdd(def2_cnt, 1ull); def2[slot] = d; }
    }
}
__global__ void k_geo_points_literal(GeoK g, const float2* __restrict__ src_xy, const Deferred* def2, const unsigned long long* def2_cnt,
                                     unsigned* scratch, size_t words_per_thread, int R, int qcap)
{
    unsigned long long n = *def2_cnt, tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
    unsigned* mine = scratch + tid * words_per_thread;
    size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
    LitScratch ls;
    ls.visited = mine; ls.R = R; ls.q = reinterpret_cast<QNode*>(mine + ((vis_words + 1) & ~size_t(1))); ls.qcap = qcap;
    for (unsigned long long e = tid; e < n; e += nth) geo_point<2>(g, src_xy, (size_t)def2[e].idx, &ls);
}

// Batch versions: one launch walks the candidate lists of every tile of the batch (tile z's list sits at cand + z * npts,
// its length at counters[3 z + 2]); deferred searches carry the tile in Deferred::slab.
__device__ __forceinline__ void apply_tile(GeoK& g, const GeoTile& t)
{
    g.tile = t.tile; g.wm = const_cast<WmCell*>(t.wm);
    g.smx = t.smx; g.sMx = t.sMx; g.smy = t.smy; g.sMy = t.sMy; g.smz = t.smz; g.sMz = t.sMz; g.bdr = t.bdr; g.tnx = t.tnx; g.tny = t.tny;
}
// Pre-work of a whole batch, one launch per step (blockIdx.z = tile): decode of the raw file bytes, where_maps_to, the first
// level with a uniform block or a leaf, the level's block ranges.  A single tile of a geogrid data set (1200 x 1200 pixels)
// does not fill the machine; eight of them do.
struct GeoRaw { const unsigned char* bytes; float* out; int wordsize, is_signed, little, flip, tnx, tny, tnz; const unsigned long long* gate; };
struct GeoRawBatch { GeoRaw r[GEO_NB]; };
__device__ __forceinline__ void geo_decode_one(const unsigned char* __restrict__ c, int wordsize, int is_signed, int little, int flip,
                                               int tnx, int tny, size_t e, float* __restrict__ out);
__global__ void __launch_bounds__(256) k_geo_decode_b(GeoRawBatch rb)
{
    const GeoRaw& r = rb.r[blockIdx.z];
    if (!r.bytes) return;
    if (r.gate && *r.gate == 0ull) return;          // nobody is going to read the decoded tile
    const size_t n = (size_t)r.tnx * r.tny * r.tnz;
    if (r.wordsize == 1 && (r.tnx & 3) == 0 && (((size_t)r.bytes | (size_t)r.out) & 15) == 0) {
        // one-byte words (land use, soil categories): four pixels of a row per thread, uchar4 in, float4 out
        const size_t n4 = n >> 2;
        const int tnx4 = r.tnx >> 2;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
            const uchar4 b = reinterpret_cast<const uchar4*>(r.bytes)[q];
            int v[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
            for (int u = 0; u < 4; ++u) if (r.is_signed && v[u] > (1 << 7)) v[u] -= (1 << 8);      // read_geogrid.c:96-100
            const size_t i4 = q % tnx4, j = (q / tnx4) % r.tny, k = q / ((size_t)tnx4 * r.tny);
            const size_t jo = r.flip ? (size_t)(r.tny - 1) - j : j;
            reinterpret_cast<float4*>(r.out)[(k * r.tny + jo) * tnx4 + i4] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        }
        return;
    }
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        geo_decode_one(r.bytes, r.wordsize, r.is_signed, r.little, r.flip, r.tnx, r.tny, e, r.out);
}
// row terms (latitude) and column terms (longitude) of where_maps_to for a regular lat-lon source on a Lambert / polar
// stereographic domain: tny + tnx evaluations per tile instead of 128 per 64 x 64 block -- the same expressions, the same bits
__global__ void __launch_bounds__(256) k_geo_wm_terms_b(GeoK g, GeoBatch bt)
{
    const GeoTile& t = bt.t[blockIdx.z];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= t.tnx + t.tny) return;
    float lat, lon;
    if (e < t.tny) {
        xytoll(g.src, (float)t.smx, (float)(t.smy + e), lat, lon, WPSGPU_M);
        t.terms[e] = g.dom.code == WPSGPU_PROJ_LC ? llij_lc_rm(lat, g.dom) : llij_ps_rm(lat, g.dom);
    } else {
        const int c = e - t.tny;
        float sn, cs;
        xytoll(g.src, (float)(t.smx + c), (float)t.smy, lat, lon, WPSGPU_M);
        if (g.dom.code == WPSGPU_PROJ_LC) llij_lc_sc(lon, g.dom, sn, cs); else llij_ps_sc(lon, g.dom, sn, cs);
        t.terms[t.tny + c] = sn; t.terms[t.tny + t.tnx + c] = cs;
    }
}
// where_maps_to of a batch from the row / column terms alone: a kernel of its own, because in the general kernel the
// compiler hoists the loop-invariant double-precision set-up of every projection branch in front of the pixel loop -- a
// thousand instructions per thread that this path never needs.  A block takes WMS_ROWS rows of the tile: a thread keeps the
// column terms of its pixels and walks down the rows (a block per row lived too short to be worth scheduling).
constexpr int WMS_ROWS = 8;
__global__ void __launch_bounds__(256) k_geo_wm_sep_b(GeoK g, GeoBatch bt)
{
    const GeoTile& t = bt.t[blockIdx.z];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *const_cast<int*>(t.dstart) = INT_MAX;    // k_geo_dstart_b follows in stream order
    const int tj0 = blockIdx.y * WMS_ROWS;
    if (tj0 >= t.tny) return;
    const int tj1 = min(tj0 + WMS_ROWS, t.tny);
    const float* __restrict__ sn = t.terms + t.tny;
    const float* __restrict__ cs = t.terms + t.tny + t.tnx;
    const bool lc = g.dom.code == WPSGPU_PROJ_LC;
    const float si = (float)g.start_i, ei = (float)g.end_i, sj = (float)g.start_j, ej = (float)g.end_j;
    for (int ti = blockIdx.x * blockDim.x + threadIdx.x; ti < t.tnx; ti += gridDim.x * blockDim.x) {
        const float s = sn[ti], c = cs[ti];
        for (int tj = tj0; tj < tj1; ++tj) {
            const float rm = t.terms[tj];
            float rx, ry;
            if (lc) llij_lc_combine(rm, s, c, g.dom, rx, ry);
            else llij_ps_combine(rm, s, c, g.dom, rx, ry);
            if (g.istagger == WPSGPU_U) rx = rx + 0.5f;
            else if (g.istagger == WPSGPU_V) ry = ry + 0.5f;
            else if (g.istagger == WPSGPU_CORNER) { rx = rx + 0.5f; ry = ry + 0.5f; }
            rx = (rx - 0.5f) * g.sr_x + 0.5f;
            ry = (ry - 0.5f) * g.sr_y + 0.5f;
            WmCell wv = WM_OUTSIDE;
            if (si <= rx && rx <= ei && sj <= ry && ry <= ej) wv = wm_pack(f_nint(rx), f_nint(ry));
            const_cast<WmCell*>(t.wm)[(size_t)tj * t.tnx + ti] = wv;
        }
    }
}
__global__ void __launch_bounds__(256) k_geo_wm_b(GeoK g, GeoBatch bt, int have_terms)
{
    const GeoTile& t = bt.t[blockIdx.z];
    if ((int)blockIdx.x * WM_T >= t.tnx || (int)blockIdx.y * WM_T >= t.tny) return;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *const_cast<int*>(t.dstart) = INT_MAX;    // k_geo_dstart_b follows in stream order
    GeoWmTile w;
    w.smx = t.smx; w.smy = t.smy; w.tnx = t.tnx; w.tny = t.tny; w.wm = const_cast<WmCell*>(t.wm); w.terms = have_terms ? t.terms : nullptr;
    geo_wm_body(g, w);
}
__global__ void __launch_bounds__(256) k_geo_dstart_b(GeoBatch bt)
{
    const GeoTile& t = bt.t[blockIdx.z];
    geo_dstart_body(t.wm, t.smx, t.smy, t.tnx, t.smx + t.bdr, t.sMx - t.bdr, t.smy + t.bdr, t.sMy - t.bdr, const_cast<int*>(t.dstart));
}
__global__ void __launch_bounds__(256) k_geo_bounds_b(GeoBatch bt)
{
    const GeoTile& t = bt.t[blockIdx.z];
    geo_bounds_body(t.dstart, t.smx + t.bdr, t.sMx - t.bdr, t.smy + t.bdr, t.sMy - t.bdr, const_cast<int2*>(t.xb), const_cast<int2*>(t.yb));
}

__global__ void __launch_bounds__(256)
k_geo_points_b(GeoK g, GeoBatch bt, const float2* __restrict__ src_xy, const int* __restrict__ cand, unsigned long long* counters,
               Deferred* def)
{
    for (int z = 0; z < bt.n; ++z) {
        const unsigned long long n = counters[3 * z + 2];
        if (n == 0) continue;
        apply_tile(g, bt.t[z]);
        for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
            size_t p = (size_t)cand[(size_t)z * g.npts + e];
            if (geo_point<0>(g, src_xy, p, nullptr)) {
                unsigned long long slot = atomicAdd(counters, 1ull);
                Deferred d; d.slab = z; d.idx = (int)p;
                def[slot] = d;
            }
        }
    }
}
__global__ void k_geo_points_warp_b(GeoK g, GeoBatch bt, const float2* __restrict__ src_xy, const Deferred* def, unsigned long long* counters,
                                    Deferred* def2)
{
    unsigned long long n = counters[0];
    unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long e = warp; e < n; e += nw) {
        Deferred d = def[e];
        apply_tile(g, bt.t[d.slab]);
        bool again = geo_point<1>(g, src_xy, (size_t)d.idx, nullptr);
        if (again && (threadIdx.x & 31) == 0) { unsigned long long slot = atomicAdd(counters + 1, 1ull); def2[slot] = d; }
    }
}
__global__ void k_geo_points_literal_b(GeoK g, GeoBatch bt, const float2* __restrict__ src_xy, const Deferred* def2, const unsigned long long* counters,
                                       unsigned* scratch, size_t words_per_thread, int R, int qcap)
{
    unsigned long long n = counters[1], tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
    unsigned* mine = scratch + tid * words_per_thread;
    size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
    LitScratch ls;
    ls.visited = mine; ls.R = R; ls.q = reinterpret_cast<QNode*>(mine + ((vis_words + 1) & ~size_t(1))); ls.qcap = qcap;
    for (unsigned long long e = tid; e < n; e += nth) {
        apply_tile(g, bt.t[def2[e].slab]);
        geo_point<2>(g, src_xy, (size_t)def2[e].idx, &ls);
    }
}
__global__ void k_geo_batch_reset(int* tbbox, unsigned long long* counters)
{
    if (threadIdx.x == 0) { tbbox[0] = INT_MAX; tbbox[1] = INT_MAX; tbbox[2] = INT_MIN; tbbox[3] = INT_MIN; }
    if (threadIdx.x < 3 * GEO_NB) counters[threadIdx.x] = 0ull;
}

// calc_field post-processing (process_tile_module.F:1580-1659)
__global__ void k_geo_finish(GeoK g, int has_scale, float scale_factor)
{
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.npts) return;
    int ii = (int)(p % g.nx), jj = (int)(p / g.nx);
    bool proc = bit_test(g.processed, g.nxw, ii, jj), lev = bit_test(g.level, g.nxw, ii, jj);
    bool use_lm = g.landmask && (g.istagger == WPSGPU_M || g.istagger == WPSGPU_HH);
    if (g.itype == WPSGPU_SP_CONTINUOUS) {
        if (!use_lm || g.landmask[p] != g.mask_val) {
            for (int k = 0; k < g.nk; ++k) {
                size_t o = (size_t)k * g.npts + p;
                if (g.count[o] > 0.0f) g.field[o] = g.field[o] / g.count[o];
                else if (!proc) g.field[o] = g.msg_fill_val;
            }
        } else if (!proc) {
            for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + p] = g.msg_fill_val;
        }
    } else if (g.itype == WPSGPU_CATEGORICAL) {
        if (use_lm && g.landmask[p] == g.mask_val)
            for (int k = 0; k < g.nk; ++k) g.field[(size_t)k * g.npts + p] = 0.0f;
    }
    if (has_scale && lev && !proc)
        for (int k = 0; k < g.nk; ++k) {
            size_t o = (size_t)k * g.npts + p;
            if (g.field[o] != g.msg_fill_val) g.field[o] = g.field[o] * scale_factor;
        }
}
__global__ void k_geo_merge_bits(int* processed, const int* level, size_t nwords)
{
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nwords) processed[w] |= level[w];
}

// read_geogrid.c:91-128 + the TOP_BOTTOM row flip of source_data_module.F:2887-2898
__device__ __forceinline__ void geo_decode_one(const unsigned char* __restrict__ c, int wordsize, int is_signed, int little, int flip,
                                               int tnx, int tny, size_t e, float* __restrict__ out)
{
    int ival;
    const unsigned char* b = c + e * wordsize;
    if (wordsize == 1) { ival = b[0]; if (is_signed && ival > (1 << 7)) ival -= (1 << 8); }
    else if (wordsize == 2) { ival = little ? ((b[1] << 8) | b[0]) : ((b[0] << 8) | b[1]); if (is_signed && ival > (1 << 15)) ival -= (1 << 16); }
    else if (wordsize == 3) { ival = little ? ((b[2] << 16) | (b[1] << 8) | b[0]) : ((b[0] << 16) | (b[1] << 8) | b[2]); if (is_signed && ival > (1 << 23)) ival -= (1 << 24); }
    else { unsigned u = little ? (((unsigned)b[3] << 24) | (b[2] << 16) | (b[1] << 8) | b[0]) : (((unsigned)b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]); ival = (int)u; }
    size_t i = e % tnx, j = (e / tnx) % tny, k = e / ((size_t)tnx * tny);
    size_t jo = flip ? (size_t)(tny - 1) - j : j;
    out[(k * tny + jo) * tnx + i] = (float)ival;
}
__global__ void k_geo_decode(const unsigned char* __restrict__ c, int wordsize, int is_signed, int little, int flip,
                             int tnx, int tny, int tnz, float* __restrict__ out)
{
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)tnx * tny * tnz;
    if (e >= n) return;
    geo_decode_one(c, wordsize, is_signed, little, flip, tnx, tny, e, out);
}

__global__ void k_geo_scale(float* __restrict__ a, size_t n, float f)
{
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) a[e] = a[e] * f;
}

static int geo_grow(void** p, size_t* cap, size_t bytes, cudaStream_t st)
{
    if (bytes <= *cap) return 0;
    WPS_CUDA(cudaStreamSynchronize(st));
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    WPS_CUDA(cudaMalloc(p, bytes + bytes / 8));
    *cap = bytes + bytes / 8;
    return 0;
}

static GeoK make_k(wpsgpu_ctx* h)
{
    GeoState* s = h->geo;
    GeoK g;
    memset(&g, 0, sizeof(g));
    g.src = to_dev(s->l.src_proj, s->d_gauss_src); g.dom = to_dev(s->f.dom_proj, s->d_gauss_dom);
    g.istagger = s->f.istagger; g.itype = s->l.itype;
    g.start_i = s->f.start_i; g.end_i = s->f.end_i; g.start_j = s->f.start_j; g.end_j = s->f.end_j; g.start_k = s->f.start_k; g.end_k = s->f.end_k;
    g.nx = s->nx; g.ny = s->ny; g.nk = s->nk; g.nxw = s->nxw; g.npts = s->npts;
    g.sr_x = (float)s->l.sr_x; g.sr_y = (float)s->l.sr_y; g.msg_val = s->l.msg_val; g.msg_fill_val = s->l.msg_fill_val; g.mask_val = s->l.mask_val;
    g.field = s->d_field; g.count = s->d_count; g.landmask = s->d_landmask; g.processed = s->d_processed; g.level = s->d_level;
    g.acc_cnt = s->d_acc_cnt; g.acc_sum = s->d_acc_sum; g.touched = s->d_touched; g.invalid = s->d_invalid;
    g.tbbox = s->d_tbbox; g.blk_xy = s->d_blk_xy; g.counters = s->d_invalid + 1;
    memcpy(g.ops, s->l.ops, sizeof(g.ops)); memcpy(g.opts, s->l.opts, sizeof(g.opts));
    g.ops[WPSGPU_MAX_OPS - 1] = 0;
    return g;
}

static void geo_free_field(GeoState* s)
{
    if (s->own_field && s->d_field) cudaFree(s->d_field);
    if (s->own_ll) { if (s->d_xlat) cudaFree(s->d_xlat); if (s->d_xlon) cudaFree(s->d_xlon); if (s->d_landmask) cudaFree(s->d_landmask); }
    s->d_field = s->d_xlat = s->d_xlon = s->d_landmask = nullptr;
    s->own_field = s->own_ll = false;
    s->field_open = s->level_open = false;
}

// grow-only scratch; `zero` buffers are cleared when (re)allocated and are returned to all-zero by the kernels that
// consume them (k_geo_merge clears what k_geo_accum wrote), so they are not cleared again per field
static int geo_buf(void** p, size_t* cap, size_t bytes, bool zero, cudaStream_t st)
{
    if (bytes <= *cap) return 0;
    WPS_CUDA(cudaStreamSynchronize(st));
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    WPS_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    if (zero) WPS_CUDA(cudaMemsetAsync(*p, 0, bytes, st));
    return 0;
}

}  // namespace wps

using namespace wps;

extern "C" {

void wpsgpu_geo_free(wpsgpu_ctx* h)
{
    if (!h->geo) return;
    GeoState* s = h->geo;
    geo_free_field(s);
    for (int b = 0; b < GEO_SETS; ++b) {
        if (s->d_tile[b]) cudaFree(s->d_tile[b]);
        if (s->d_wm[b]) cudaFree(s->d_wm[b]);
        if (b == 0 && s->d_dstart) { cudaFree(s->d_dstart); s->d_dstart = nullptr; }
        if (s->d_bounds[b]) { cudaFree(s->d_bounds[b]); s->d_bounds[b] = nullptr; s->bounds_cap[b] = 0; }
        if (s->d_terms[b]) { cudaFree(s->d_terms[b]); s->d_terms[b] = nullptr; s->terms_cap[b] = 0; }
        if (s->ev_pre[b]) cudaEventDestroy(s->ev_pre[b]);
        if (s->ev_used[b]) cudaEventDestroy(s->ev_used[b]);
    }
    if (s->s_pre) cudaStreamDestroy(s->s_pre);
    for (int b = 0; b < GEO_SETS; ++b) if (s->d_raw[b]) cudaFree(s->d_raw[b]);
    if (s->d_first_tile) cudaFree(s->d_first_tile);
    if (s->ev_h2d) cudaEventDestroy(s->ev_h2d);
    if (s->d_lit) cudaFree(s->d_lit);
    if (s->d_invalid) cudaFree(s->d_invalid);
    void* ptrs[] = { s->d_count, s->d_processed, s->d_level, s->d_src_xy, s->d_acc_cnt, s->d_acc_sum, s->d_touched, s->d_def, s->d_def2,
                     s->d_blk_xy, s->d_tbbox, s->d_cand };
    for (void* p : ptrs) if (p) cudaFree(p);
    delete s;
    h->geo = nullptr;
}

int wpsgpu_geogrid_field_begin(wpsgpu_handle_t h, const wpsgpu_geo_field* f)
{
    WPS_CHECK(h && f && f->xlat && f->xlon && f->field, -1, "wpsgpu_geogrid_field_begin: NULL argument");
    WPS_CHECK(f->dom_proj.init, -1, "domain projection not initialised");
    WPS_CUDA(cudaSetDevice(h->device));
    if (!h->geo) { h->geo = new GeoState(); WPS_CUDA(cudaMalloc((void**)&h->geo->d_invalid, (1 + 3 * GEO_NB) * sizeof(unsigned long long))); }
    GeoState* s = h->geo;
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    geo_free_field(s);
    s->f = *f;
    s->nx = f->end_i - f->start_i + 1; s->ny = f->end_j - f->start_j + 1; s->nk = f->end_k - f->start_k + 1;
    WPS_CHECK(s->nx > 0 && s->ny > 0 && s->nk > 0, -1, "empty field");
    WPS_CHECK(f->start_i > -32768 && f->end_i <= 32767 && f->start_j > -32768 && f->end_j <= 32767, -1,
              "wpsgpu_geogrid_field_begin: field indices beyond +-32767 are not supported (where_maps_to keeps 16-bit cell indices)");
    s->nxw = (s->nx + 31) / 32; s->npts = (size_t)s->nx * s->ny;
    size_t b2 = s->npts * sizeof(float), b3 = b2 * s->nk, wb = (size_t)s->ny * s->nxw * sizeof(int);
    const bool host = h->ptr_mode == WPSGPU_PTR_HOST;
    if (host) {
        WPS_CUDA(cudaMalloc((void**)&s->d_field, b3)); s->own_field = true;
        WPS_CUDA(cudaMalloc((void**)&s->d_xlat, b2)); WPS_CUDA(cudaMalloc((void**)&s->d_xlon, b2)); s->own_ll = true;
        WPS_CUDA(cudaMemcpyAsync(s->d_field, f->field, b3, cudaMemcpyHostToDevice, h->stream));
        WPS_CUDA(cudaMemcpyAsync(s->d_xlat, f->xlat, b2, cudaMemcpyHostToDevice, h->stream));
        WPS_CUDA(cudaMemcpyAsync(s->d_xlon, f->xlon, b2, cudaMemcpyHostToDevice, h->stream));
        if (f->landmask) { WPS_CUDA(cudaMalloc((void**)&s->d_landmask, b2)); WPS_CUDA(cudaMemcpyAsync(s->d_landmask, f->landmask, b2, cudaMemcpyHostToDevice, h->stream)); }
    } else {
        s->d_field = f->field; s->own_field = false;
        s->d_xlat = const_cast<float*>(f->xlat); s->d_xlon = const_cast<float*>(f->xlon); s->d_landmask = const_cast<float*>(f->landmask); s->own_ll = false;
    }
    const size_t nblk = (s->npts + 255) / 256;
    {   // d_processed and d_level share one capacity counter: allocate both whenever it grows
        size_t cap = s->cap_bits;
        WPS_TRY(geo_buf((void**)&s->d_processed, &cap, wb, false, h->stream));
        WPS_TRY(geo_buf((void**)&s->d_level, &s->cap_bits, wb, false, h->stream));
    }
    WPS_CUDA(cudaMemsetAsync(s->d_processed, 0, wb, h->stream));  // bitarray_create at ilevel==1 (:1257)
    WPS_TRY(geo_buf((void**)&s->d_src_xy, &s->cap_xy, s->npts * sizeof(float2), false, h->stream));
    WPS_TRY(geo_buf((void**)&s->d_blk_xy, &s->cap_blk, nblk * sizeof(float4), false, h->stream));
    WPS_TRY(geo_buf((void**)&s->d_acc_cnt, &s->cap_acc_cnt, (size_t)s->nk * s->npts * sizeof(int), true, h->stream));
    WPS_TRY(geo_buf((void**)&s->d_touched, &s->cap_touched, s->npts * sizeof(int), true, h->stream));
    {
        size_t cap = s->cap_def;
        WPS_TRY(geo_buf((void**)&s->d_def, &cap, s->npts * sizeof(Deferred), false, h->stream));
        WPS_TRY(geo_buf((void**)&s->d_def2, &s->cap_def, s->npts * sizeof(Deferred), false, h->stream));
    }
    WPS_TRY(geo_buf((void**)&s->d_cand, &s->cap_cand, (size_t)GEO_NB * s->npts * sizeof(int), false, h->stream));
    if (s->npts * sizeof(int) > s->cap_first) {
        WPS_TRY(geo_buf((void**)&s->d_first_tile, &s->cap_first, s->npts * sizeof(int), false, h->stream));
        WPS_CUDA(cudaMemsetAsync(s->d_first_tile, 0x7f, s->cap_first, h->stream));   // "no tile yet"; k_geo_merge_b restores it
    }
    s->pending.clear();
    if (!s->d_tbbox) WPS_CUDA(cudaMalloc((void**)&s->d_tbbox, 4 * sizeof(int)));
    WPS_CUDA(cudaMemsetAsync(s->d_invalid, 0, (1 + 3 * GEO_NB) * sizeof(unsigned long long), h->stream));
    if (host) WPS_CUDA(cudaStreamSynchronize(h->stream));
    s->field_open = true;
    return 0;
}

int wpsgpu_geogrid_level_begin(wpsgpu_handle_t h, const wpsgpu_geo_level* l)
{
    WPS_CHECK(h && l && h->geo && h->geo->field_open, -1, "wpsgpu_geogrid_level_begin: no field open");
    WPS_CHECK(l->src_proj.init, -1, "source projection not initialised");
    WPS_CHECK(l->src_proj.code != WPSGPU_PROJ_GAUSS && h->geo->f.dom_proj.code != WPSGPU_PROJ_GAUSS, -1, "Gaussian projections are not supported for geogrid data");
    WPS_CUDA(cudaSetDevice(h->device));
    GeoState* s = h->geo;
    s->l = *l;
    size_t wb = (size_t)s->ny * s->nxw * sizeof(int);
    WPS_CUDA(cudaMemsetAsync(s->d_level, 0, wb, h->stream));     // bitarray_create(level_domain) (:1345)
    if (l->itype == WPSGPU_SP_CONTINUOUS) {
        WPS_TRY(geo_buf((void**)&s->d_count, &s->cap_count, s->npts * s->nk * sizeof(float), false, h->stream));
        WPS_TRY(geo_buf((void**)&s->d_acc_sum, &s->cap_acc_sum, s->npts * s->nk * sizeof(double), true, h->stream));
        WPS_CUDA(cudaMemsetAsync(s->d_count, 0, s->npts * s->nk * sizeof(float), h->stream));  // data_count = 0. (:1316)
    }
    s->has_search = false; s->max_depth = 0;
    for (int k = 0; k < WPSGPU_MAX_OPS && l->ops[k]; ++k) if (l->ops[k] == WPSGPU_SEARCH) { s->has_search = true; s->max_depth = std::max(s->max_depth, l->opts[k]); }
    GeoK g = make_k(h);
    k_geo_src_xy<<<(unsigned)((s->npts + 255) / 256), 256, 0, h->stream>>>(g.src, s->d_xlat, s->d_xlon, s->npts, s->d_src_xy, s->d_blk_xy);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    s->level_open = true;
    return 0;
}

// Buffers and stream of the coming tile: alternate between the two buffer sets and stage on s_pre (ordered behind the
// previous user of the set).  In host-pointer mode the call returns as soon as the caller's buffer has been read
// (geo_tile_staged), so the host reads the next tile from disk while this one is processed; wpsgpu_geogrid_level_end /
// field_finish run on the handle's stream behind every tile.
static int geo_tile_begin(wpsgpu_ctx* h, size_t tile_floats, int* set, cudaStream_t* pre)
{
    GeoState* s = h->geo;
    if (!s->s_pre) {
        WPS_CUDA(cudaStreamCreateWithFlags(&s->s_pre, cudaStreamNonBlocking));
        for (int b = 0; b < GEO_SETS; ++b) {
            WPS_CUDA(cudaEventCreateWithFlags(&s->ev_pre[b], cudaEventDisableTiming));
            WPS_CUDA(cudaEventCreateWithFlags(&s->ev_used[b], cudaEventDisableTiming));
        }
        WPS_CUDA(cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming | cudaEventBlockingSync));
    }
    *set = s->parity; s->parity = (s->parity + 1) % GEO_SETS; *pre = s->s_pre;
    if (s->used_rec[*set]) WPS_CUDA(cudaStreamWaitEvent(s->s_pre, s->ev_used[*set], 0));
    if (tile_floats * sizeof(float) > s->tile_cap[*set]) {
        WPS_CUDA(cudaStreamSynchronize(h->stream));
        WPS_TRY(geo_grow((void**)&s->d_tile[*set], &s->tile_cap[*set], tile_floats * sizeof(float), *pre));
    }
    return 0;
}

// host-pointer mode: wait until the copy out of the caller's buffer is done (not for the kernels)
static int geo_tile_staged(wpsgpu_ctx* h, cudaStream_t pre)
{
    if (h->ptr_mode != WPSGPU_PTR_HOST) return 0;
    WPS_CUDA(cudaEventRecord(h->geo->ev_h2d, pre));
    WPS_CUDA(cudaEventSynchronize(h->geo->ev_h2d));
    return 0;
}

// Tiles of a level whose results do not depend on the tile-by-tile order beyond what k_geo_select_b reproduces are
// accumulated GEO_NB at a time: one accumulate launch, one selection, one point pass and one merge per batch instead of per
// tile.  Overlay levels of categorical data (half-area rule after every tile, proc_point_module.F:165-202) keep the
// sequential path; WPSGPU_GEO_BATCH=0 forces it for everything (A/B measurements and the parity tests).
static bool geo_batchable(const GeoState* s)
{
    static const bool enabled = !(getenv("WPSGPU_GEO_BATCH") && atoi(getenv("WPSGPU_GEO_BATCH")) == 0);
    if (!enabled) return false;
    if (s->nx > 65535 || s->ny > 32767) return false;      // k_geo_accum_cat_b packs a cell's (row, column) into one int
    if (s->l.itype == WPSGPU_SP_CONTINUOUS) return true;
    return s->l.itype == WPSGPU_CATEGORICAL && s->l.ilevel == 1;
}

static int geo_flush_batch(wpsgpu_ctx* h)
{
    GeoState* s = h->geo;
    if (s->pending.empty()) return 0;
    GeoK g = make_k(h);
    GeoBatch bt;
    memset(&bt, 0, sizeof(bt));
    bt.n = (int)s->pending.size();
    int gx = 1, gy = 1, tmax = 1;
    for (int z = 0; z < bt.n; ++z) {
        const GeoState::Pending& p = s->pending[z];
        GeoTile& t = bt.t[z];
        t.tile = s->d_tile[p.set]; t.wm = s->d_wm[p.set]; t.dstart = s->d_dstart + p.set;
        t.xb = s->d_bounds[p.set]; t.yb = s->d_bounds[p.set] + std::max(0, (p.sMx - p.smx + 1) - 2 * p.bdr);
        t.terms = s->d_terms[p.set];
        t.smx = p.smx; t.sMx = p.sMx; t.smy = p.smy; t.sMy = p.sMy; t.smz = p.smz; t.sMz = p.sMz; t.bdr = p.bdr;
        t.tnx = p.sMx - p.smx + 1; t.tny = p.sMy - p.smy + 1;
        gx = std::max(gx, (t.tnx - 2 * p.bdr + 31) / 32); gy = std::max(gy, (t.tny - 2 * p.bdr + 7) / 8);
        tmax = std::max(tmax, std::max(t.tnx, t.tny));
    }
    // Categorical tiles that arrive as file bytes are not decoded up front: the accumulate kernel reads the bytes, and the decoded
    // copy is only made (after the selection) for tiles whose point pass has candidates -- none where the source is finer than
    // the domain.  One-level tiles only (a categorical tile has one level).
    static const bool lazy_enabled = !(getenv("WPSGPU_GEO_LAZY_DECODE") && atoi(getenv("WPSGPU_GEO_LAZY_DECODE")) == 0);
    const bool lazy = lazy_enabled && s->l.itype == WPSGPU_CATEGORICAL;
    GeoRawBatch rb, rb_lazy;
    bool any_lazy = false;
    size_t lazy_maxpix = 1;
    {   // pre-work of the whole batch on the staging stream (behind the tiles' copies): decode, where_maps_to, level tables
        memset(&rb, 0, sizeof(rb));
        memset(&rb_lazy, 0, sizeof(rb_lazy));
        bool any_raw = false;
        int wx = 1, wy = 1, bnd = 1;
        size_t maxpix = 1;
        for (int z = 0; z < bt.n; ++z) {
            const GeoState::Pending& p = s->pending[z];
            const GeoTile& t = bt.t[z];
            if (p.raw) {
                GeoRaw& r = lazy ? rb_lazy.r[z] : rb.r[z];
                r.bytes = p.raw; r.out = s->d_tile[p.set]; r.wordsize = p.wordsize; r.is_signed = p.is_signed; r.little = p.little; r.flip = p.flip;
                r.tnx = t.tnx; r.tny = t.tny; r.tnz = p.sMz - p.smz + 1;
                r.gate = lazy ? g.counters + 3 * z + 2 : nullptr;       // length of the tile's candidate list (k_geo_select_b)
                maxpix = std::max(maxpix, (size_t)t.tnx * t.tny * r.tnz);
                if (lazy) {
                    any_lazy = true;
                    bt.t[z].raw = p.raw; bt.t[z].rws = p.wordsize; bt.t[z].rsigned = p.is_signed; bt.t[z].rlittle = p.little; bt.t[z].rflip = p.flip;
                } else any_raw = true;
            }
            if (lazy && p.raw) lazy_maxpix = std::max(lazy_maxpix, (size_t)t.tnx * t.tny * (p.sMz - p.smz + 1));
            wx = std::max(wx, (t.tnx + WM_T - 1) / WM_T); wy = std::max(wy, (t.tny + WM_T - 1) / WM_T);
            bnd = std::max(bnd, std::max(0, t.tnx - 2 * t.bdr) + std::max(0, t.tny - 2 * t.bdr));
        }
        if (any_raw) { k_geo_decode_b<<<dim3((unsigned)std::min<size_t>((maxpix + 255) / 256, 4096), 1, bt.n), 256, 0, s->s_pre>>>(rb); h->launches++; }
        const bool separable = g.src.code == WPSGPU_PROJ_LATLON && (g.dom.code == WPSGPU_PROJ_LC || g.dom.code == WPSGPU_PROJ_PS);
        if (separable) {
            int tmx = 1;
            for (int z = 0; z < bt.n; ++z) tmx = std::max(tmx, bt.t[z].tnx + bt.t[z].tny);
            k_geo_wm_terms_b<<<dim3((tmx + 255) / 256, 1, bt.n), 256, 0, s->s_pre>>>(g, bt);
            h->launches++;
        }
        if (separable) {
            int mx = 1, my = 1;
            for (int z = 0; z < bt.n; ++z) { mx = std::max(mx, bt.t[z].tnx); my = std::max(my, bt.t[z].tny); }
            k_geo_wm_sep_b<<<dim3((mx + 255) / 256, (my + WMS_ROWS - 1) / WMS_ROWS, bt.n), 256, 0, s->s_pre>>>(g, bt);
        } else k_geo_wm_b<<<dim3(wx, wy, bt.n), 256, 0, s->s_pre>>>(g, bt, 0);
        k_geo_dstart_b<<<dim3((GEO_DNODES + 255) / 256, 1, bt.n), 256, 0, s->s_pre>>>(bt);
        k_geo_bounds_b<<<dim3((bnd + 255) / 256, 1, bt.n), 256, 0, s->s_pre>>>(bt);
        h->launches += 3;
        WPS_CUDA(cudaEventRecord(s->ev_pre[s->pending[0].set], s->s_pre));
        WPS_CUDA(cudaStreamWaitEvent(h->stream, s->ev_pre[s->pending[0].set], 0));   // the tiles and their mappings are in place
    }
    unsigned long long* cnt = g.counters;
    k_geo_batch_reset<<<1, 32, 0, h->stream>>>(g.tbbox, cnt);
    if (s->l.itype == WPSGPU_CATEGORICAL) k_geo_accum_cat_b<<<dim3(gx, (gy + 3) / 4, bt.n), dim3(32, 8), 0, h->stream>>>(g, bt, s->d_first_tile);
    else k_geo_accum_b<<<dim3(gx, gy, bt.n), dim3(32, 8), 0, h->stream>>>(g, bt, s->d_first_tile);
    const unsigned nblk = (unsigned)((s->npts + 255) / 256);
    k_geo_select_b<<<nblk, 256, 0, h->stream>>>(g, bt, s->d_src_xy, s->d_first_tile, s->d_cand, cnt);
    // (a small grid: the gate is shut for almost every tile, and 32 k blocks that only look at it cost 20 us)
    if (any_lazy) { k_geo_decode_b<<<dim3((unsigned)std::min<size_t>((lazy_maxpix + 255) / 256, 148), 1, bt.n), 256, 0, h->stream>>>(rb_lazy); h->launches++; }
    k_geo_points_b<<<h->sm_count * 2, 256, 0, h->stream>>>(g, bt, s->d_src_xy, s->d_cand, cnt, s->d_def);
    h->launches += 4;
    if (s->has_search) {
        k_geo_points_warp_b<<<h->sm_count * 8, 256, 0, h->stream>>>(g, bt, s->d_src_xy, s->d_def, cnt, s->d_def2);
        int R = std::min(std::max(s->max_depth, 1), tmax);
        size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
        int qcap = 8 * R + 64;
        size_t wpt = ((vis_words + 1) & ~size_t(1)) + (size_t)qcap * sizeof(QNode) / sizeof(unsigned);
        int lit_threads = 256;
        WPS_TRY(geo_grow(&s->d_lit, &s->lit_bytes, wpt * sizeof(unsigned) * lit_threads, h->stream));
        k_geo_points_literal_b<<<lit_threads / 64, 64, 0, h->stream>>>(g, bt, s->d_src_xy, s->d_def2, cnt, (unsigned*)s->d_lit, wpt, R, qcap);
        h->launches += 2;
    }
    k_geo_merge_b<<<nblk, 256, 0, h->stream>>>(g, s->d_first_tile, s->pending[0].smz - s->f.start_k, s->pending[0].sMz - s->f.start_k);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    for (const GeoState::Pending& p : s->pending) { WPS_CUDA(cudaEventRecord(s->ev_used[p.set], h->stream)); s->used_rec[p.set] = true; }
    s->pending.clear();
    return 0;
}

static int geo_process_tile(wpsgpu_ctx* h, int set, cudaStream_t pre, int smx, int sMx, int smy, int sMy, int smz, int sMz, int bdr,
                            const GeoRaw* raw = nullptr)
{
    GeoState* s = h->geo;
    int tnx = sMx - smx + 1, tny = sMy - smy + 1;
    GeoK g = make_k(h);
    g.tile = s->d_tile[set]; g.smx = smx; g.sMx = sMx; g.smy = smy; g.sMy = sMy; g.smz = smz; g.sMz = sMz; g.bdr = bdr; g.tnx = tnx; g.tny = tny;
    if (s->l.itype != WPSGPU_CATEGORICAL)
        WPS_CHECK(smz >= s->f.start_k && sMz <= s->f.end_k, -1, "source levels %d..%d outside field levels %d..%d", smz, sMz, s->f.start_k, s->f.end_k);
    dim3 b(32, 8);
    const bool accum = s->l.itype == WPSGPU_CATEGORICAL || s->l.itype == WPSGPU_SP_CONTINUOUS;
    if (accum) {
        if ((size_t)tnx * tny * sizeof(WmCell) > s->wm_cap[set]) {
            WPS_CUDA(cudaStreamSynchronize(h->stream));
            WPS_TRY(geo_grow((void**)&s->d_wm[set], &s->wm_cap[set], (size_t)tnx * tny * sizeof(WmCell), pre));
        }
        g.wm = s->d_wm[set];
        if (geo_batchable(s) && pre != h->stream) {
            // batch path: where_maps_to and the level tables are computed for the whole batch at once (geo_flush_batch)
            if (!s->d_dstart) WPS_CUDA(cudaMalloc((void**)&s->d_dstart, GEO_SETS * sizeof(int)));
            const int inx = std::max(0, tnx - 2 * bdr), iny = std::max(0, tny - 2 * bdr);
            if ((size_t)(inx + iny + 1) * sizeof(int2) > s->bounds_cap[set]) {
                WPS_CUDA(cudaStreamSynchronize(h->stream));
                WPS_TRY(geo_grow((void**)&s->d_bounds[set], &s->bounds_cap[set], (size_t)(inx + iny + 1) * sizeof(int2), pre));
            }
            if ((size_t)(2 * tnx + tny) * sizeof(float) > s->terms_cap[set]) {
                WPS_CUDA(cudaStreamSynchronize(h->stream));
                WPS_TRY(geo_grow((void**)&s->d_terms[set], &s->terms_cap[set], (size_t)(2 * tnx + tny) * sizeof(float), pre));
            }
        } else {
            k_geo_wm<<<dim3((tnx + WM_T - 1) / WM_T, (tny + WM_T - 1) / WM_T), 256, 0, pre>>>(g);
            h->launches++;
        }
    }
    if (accum && geo_batchable(s) && pre != h->stream) {
        // batch path: the tile waits, staged, until GEO_NB tiles have come together (or the level ends)
        if (!s->pending.empty() && (s->pending[0].smz != smz || s->pending[0].sMz != sMz)) WPS_TRY(geo_flush_batch(h));
        GeoState::Pending p;
        p.set = set; p.smx = smx; p.sMx = sMx; p.smy = smy; p.sMy = sMy; p.smz = smz; p.sMz = sMz; p.bdr = bdr;
        p.raw = raw ? raw->bytes : nullptr;
        p.wordsize = raw ? raw->wordsize : 0; p.is_signed = raw ? raw->is_signed : 0; p.little = raw ? raw->little : 0; p.flip = raw ? raw->flip : 0;
        s->pending.push_back(p);
        if ((int)s->pending.size() == GEO_NB) WPS_TRY(geo_flush_batch(h));
        return 0;
    }
    WPS_TRY(geo_flush_batch(h));   // anything pending precedes a tile of the sequential path
    if (pre != h->stream) {   // the handle's stream takes over once the tile and its mapping are in place
        WPS_CUDA(cudaEventRecord(s->ev_pre[set], pre));
        WPS_CUDA(cudaStreamWaitEvent(h->stream, s->ev_pre[set], 0));
    }
    if (accum) {
        k_geo_tile_reset<<<1, 1, 0, h->stream>>>(g.tbbox, g.counters);
        int inx = tnx - 2 * bdr, iny = tny - 2 * bdr;
        if (inx > 0 && iny > 0) k_geo_accum<<<dim3((inx + 31) / 32, (iny + 7) / 8), b, 0, h->stream>>>(g);
        // areas of the half-area rule (proc_point_module.F:178-197)
        const wpsgpu_proj &sp = s->l.src_proj, &dp = s->f.dom_proj;
        float sa = (sp.dx < 0.0f) ? (sp.latinc * 111000.0f) * (sp.latinc * 111000.0f) : sp.dx * sp.dx;
        float da = (dp.dx < 0.0f) ? (dp.latinc * 111000.0f) * (dp.latinc * 111000.0f) : dp.dx * dp.dx;
        k_geo_merge_select<<<(unsigned)((s->npts + 255) / 256), 256, 0, h->stream>>>(g, s->l.ilevel, sa, da, 1, s->d_src_xy, s->d_cand, g.counters + 2);
        h->launches += 3;
        WPS_CUDA(cudaGetLastError());
    } else {   // point-wise types: no accumulation, only the point pass
        WPS_CUDA(cudaMemsetAsync(g.counters, 0, 3 * sizeof(unsigned long long), h->stream));
        k_geo_merge_select<<<(unsigned)((s->npts + 255) / 256), 256, 0, h->stream>>>(g, s->l.ilevel, 0.0f, 0.0f, 0, s->d_src_xy, s->d_cand, g.counters + 2);
        h->launches++;
    }
    unsigned long long* cnt = g.counters;
    k_geo_points<<<h->sm_count * 2, 256, 0, h->stream>>>(g, s->d_src_xy, s->d_cand, cnt + 2, s->d_def, cnt);
    h->launches++;
    if (s->has_search) {
        k_geo_points_warp<<<h->sm_count * 8, 256, 0, h->stream>>>(g, s->d_src_xy, s->d_def, cnt, s->d_def2, cnt + 1);
        int R = std::min(std::max(s->max_depth, 1), std::max(std::max(tnx, tny), 1));
        size_t vis_words = ((size_t)(2 * R + 1) * (2 * R + 1) + 31) / 32;
        int qcap = 8 * R + 64;
        size_t wpt = ((vis_words + 1) & ~size_t(1)) + (size_t)qcap * sizeof(QNode) / sizeof(unsigned);
        int lit_threads = 256;
        WPS_TRY(geo_grow(&s->d_lit, &s->lit_bytes, wpt * sizeof(unsigned) * lit_threads, h->stream));
        k_geo_points_literal<<<lit_threads / 64, 64, 0, h->stream>>>(g, s->d_src_xy, s->d_def2, cnt + 1, (unsigned*)s->d_lit, wpt, R, qcap);
        h->launches += 2;
    }
    WPS_CUDA(cudaGetLastError());
    if (pre != h->stream) { WPS_CUDA(cudaEventRecord(s->ev_used[set], h->stream)); s->used_rec[set] = true; }
    return 0;
}

int wpsgpu_geogrid_add_tile(wpsgpu_handle_t h, const float* tile, int src_min_x, int src_max_x, int src_min_y,
                            int src_max_y, int src_min_z, int src_max_z, int bdr)
{
    WPS_CHECK(h && tile && h->geo && h->geo->level_open, -1, "wpsgpu_geogrid_add_tile: no level open");
    WPS_CUDA(cudaSetDevice(h->device));
    GeoState* s = h->geo;
    size_t n = (size_t)(src_max_x - src_min_x + 1) * (src_max_y - src_min_y + 1) * (src_max_z - src_min_z + 1);
    WPS_CHECK(n > 0, -1, "empty tile");
    int set = 0;
    cudaStream_t pre = h->stream;
    WPS_TRY(geo_tile_begin(h, n, &set, &pre));
    WPS_CUDA(cudaMemcpyAsync(s->d_tile[set], tile, n * sizeof(float),
                             h->ptr_mode == WPSGPU_PTR_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, pre));
    WPS_TRY(geo_tile_staged(h, pre));
    WPS_TRY(geo_process_tile(h, set, pre, src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr));
    return 0;
}

int wpsgpu_geogrid_add_tile_raw(wpsgpu_handle_t h, const void* bytes, int wordsize, int is_signed, int endian,
                                int row_order, int src_min_x, int src_max_x, int src_min_y, int src_max_y,
                                int src_min_z, int src_max_z, int bdr)
{
    WPS_CHECK(h && bytes && h->geo && h->geo->level_open, -1, "wpsgpu_geogrid_add_tile_raw: no level open");
    WPS_CHECK(wordsize >= 1 && wordsize <= 4, -1, "wordsize must be 1..4");
    WPS_CUDA(cudaSetDevice(h->device));
    GeoState* s = h->geo;
    int tnx = src_max_x - src_min_x + 1, tny = src_max_y - src_min_y + 1, tnz = src_max_z - src_min_z + 1;
    size_t n = (size_t)tnx * tny * tnz;
    WPS_CHECK(n > 0, -1, "empty tile");
    int set = 0;
    cudaStream_t pre = h->stream;
    WPS_TRY(geo_tile_begin(h, n, &set, &pre));
    const unsigned char* d_bytes = (const unsigned char*)bytes;
    if (h->ptr_mode == WPSGPU_PTR_HOST) {
        if (n * wordsize > s->raw_cap[set]) {
            WPS_CUDA(cudaStreamSynchronize(h->stream));
            WPS_TRY(geo_grow(&s->d_raw[set], &s->raw_cap[set], n * wordsize, pre));
        }
        WPS_CUDA(cudaMemcpyAsync(s->d_raw[set], bytes, n * wordsize, cudaMemcpyHostToDevice, pre));
        WPS_TRY(geo_tile_staged(h, pre));
        d_bytes = (const unsigned char*)s->d_raw[set];
    }
    const bool accum = s->l.itype == WPSGPU_CATEGORICAL || s->l.itype == WPSGPU_SP_CONTINUOUS;
    if (accum && geo_batchable(s) && pre != h->stream) {
        // batch path: the bytes are decoded together with the other tiles of the batch
        GeoRaw raw;
        memset(&raw, 0, sizeof(raw));
        raw.bytes = d_bytes; raw.wordsize = wordsize; raw.is_signed = is_signed; raw.little = endian == 1; raw.flip = row_order == 2;
        WPS_TRY(geo_process_tile(h, set, pre, src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr, &raw));
        return 0;
    }
    k_geo_decode<<<(unsigned)((n + 255) / 256), 256, 0, pre>>>(d_bytes, wordsize, is_signed, endian == 1, row_order == 2, tnx, tny, tnz, s->d_tile[set]);
    h->launches++;
    WPS_CUDA(cudaGetLastError());
    WPS_TRY(geo_process_tile(h, set, pre, src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr));
    return 0;
}

int wpsgpu_geogrid_decode_tile(wpsgpu_handle_t h, const void* bytes, int wordsize, int is_signed, int endian, int row_order,
                               int nx, int ny, int nz, float scalefactor, float* rarray)
{
    WPS_CHECK(h && bytes && rarray, -1, "wpsgpu_geogrid_decode_tile: NULL argument");
    WPS_CHECK(wordsize >= 1 && wordsize <= 4, -1, "wordsize must be 1..4");
    WPS_CHECK(nx > 0 && ny > 0 && nz > 0, -1, "empty tile");
    WPS_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)nx * ny * nz;
    WPS_TRY(arena_begin(h));
    h->arena.reset();
    WPS_TRY(h->arena.reserve(n * (wordsize + sizeof(float)) + 1024, h->stream));
    void* d_in = nullptr;
    void* d_out = nullptr;
    WPS_TRY(stage_in(h, bytes, n * wordsize, &d_in));
    WPS_TRY(alloc_out(h, rarray, n * sizeof(float), &d_out));
    k_geo_decode<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((const unsigned char*)d_in, wordsize, is_signed, endian == 1, row_order == 2,
                                                                    nx, ny, nz, (float*)d_out);
    h->launches++;
    if (scalefactor != 1.0f) {      // read_geogrid.c:133-137
        k_geo_scale<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((float*)d_out, n, scalefactor);
        h->launches++;
    }
    WPS_CUDA(cudaGetLastError());
    WPS_TRY(copy_out(h, rarray, d_out, n * sizeof(float)));
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wpsgpu_geogrid_level_end(wpsgpu_handle_t h)
{
    WPS_CHECK(h && h->geo && h->geo->level_open, -1, "wpsgpu_geogrid_level_end: no level open");
    WPS_CUDA(cudaSetDevice(h->device));
    GeoState* s = h->geo;
    WPS_TRY(geo_flush_batch(h));
    GeoK g = make_k(h);
    k_geo_finish<<<(unsigned)((s->npts + 255) / 256), 256, 0, h->stream>>>(g, s->l.has_scale, s->l.scale_factor);
    size_t nwords = (size_t)s->ny * s->nxw;
    k_geo_merge_bits<<<(unsigned)((nwords + 255) / 256), 256, 0, h->stream>>>(s->d_processed, s->d_level, nwords);
    h->launches += 2;
    WPS_CUDA(cudaGetLastError());
    s->level_open = false;
    return 0;
}

int wpsgpu_geogrid_field_finish(wpsgpu_handle_t h, int32_t* processed_domain_out, int64_t* n_invalid_category)
{
    WPS_CHECK(h && h->geo && h->geo->field_open, -1, "wpsgpu_geogrid_field_finish: no field open");
    WPS_CHECK(!h->geo->level_open, -1, "wpsgpu_geogrid_field_finish: a level is still open");
    WPS_CUDA(cudaSetDevice(h->device));
    GeoState* s = h->geo;
    size_t b3 = s->npts * s->nk * sizeof(float), wb = (size_t)s->ny * s->nxw * sizeof(int);
    if (h->ptr_mode == WPSGPU_PTR_HOST) WPS_CUDA(cudaMemcpyAsync(s->f.field, s->d_field, b3, cudaMemcpyDeviceToHost, h->stream));
    if (processed_domain_out)
        WPS_CUDA(cudaMemcpyAsync(processed_domain_out, s->d_processed, wb,
                                 h->ptr_mode == WPSGPU_PTR_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    unsigned long long inv = 0;
    WPS_CUDA(cudaMemcpyAsync(&inv, s->d_invalid, sizeof(inv), cudaMemcpyDeviceToHost, h->stream));
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    if (n_invalid_category) *n_invalid_category = (int64_t)inv;
    geo_free_field(s);
    return 0;
}

}  // extern "C"
