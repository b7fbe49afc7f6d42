// interp.cuh -- device restatement of WPS interp_module.F operators (interp_sequence and its eight
// methods).  Same operation order as the Fortran, binary32, no FMA contraction (-fmad=false), no FTZ.
#pragma once
#include "proj.cuh"

namespace wps {

enum { ST_VAL = 0, ST_NEXT = 1, ST_FINAL_MSG = 2, ST_DEFER = 3 };
enum { REL_EQ = 0, REL_LT = 1, REL_GT = 2 };
struct Deferred { int slab; int idx; };  // deferred (search) work item: global slab index, point index

// One 2-D source array with Fortran bounds array(sx:ex, sy:ey) (+ optional mask of the same shape).
struct Src {
    const float* __restrict__ a;
    const float* __restrict__ m;  // nullptr = no mask arguments
    int sx, ex, sy, ey, nx;
    float msg, maskval;
    int rel;
};

__device__ __forceinline__ float s_at(const Src& s, int i, int j) { return __ldg(s.a + (size_t)(j - s.sy) * s.nx + (i - s.sx)); }
__device__ __forceinline__ float m_at(const Src& s, int i, int j) { return __ldg(s.m + (size_t)(j - s.sy) * s.nx + (i - s.sx)); }

// source point unusable: '<': mask<val, '>': mask>val, ' ': mask==val (interp_module.F:417-421,695-709)
__device__ __forceinline__ bool masked_out(const Src& s, int i, int j)
{
    if (!s.m) return false;
    float m = m_at(s, i, j);
    return s.rel == REL_GT ? (m > s.maskval) : (s.rel == REL_LT ? (m < s.maskval) : (m == s.maskval));
}
// the same test on a mask value that has been loaded already (several loads in flight instead of one after the other)
__device__ __forceinline__ bool masked_val(const Src& s, float m)
{
    return s.rel == REL_GT ? (m > s.maskval) : (s.rel == REL_LT ? (m < s.maskval) : (m == s.maskval));
}
// source point valid in search_extrap / sixteen_pt's node branch: '>': mask<=val, '<': mask>=val,
// ' ': mask/=val (interp_module.F:505-517,1302-1311)
__device__ __forceinline__ bool mask_valid(const Src& s, int i, int j)
{
    if (!s.m) return true;
    float m = m_at(s, i, j);
    return s.rel == REL_GT ? (m <= s.maskval) : (s.rel == REL_LT ? (m >= s.maskval) : (m != s.maskval));
}

// nearest_neighbor, interp_module.F:376-442
__device__ __forceinline__ int op_nearest(const Src& s, float xx, float yy, float& r)
{
    int ix = f_nint(xx), iy = f_nint(yy);
    if (ix < s.sx || ix > s.ex || iy < s.sy || iy > s.ey) return ST_FINAL_MSG;  // returns msgval WITHOUT trying the next method
    if (masked_out(s, ix, iy)) return ST_NEXT;
    r = s_at(s, ix, iy);
    return (r == s.msg) ? ST_NEXT : ST_VAL;
}

// four_pt, interp_module.F:1055-1171
__device__ __forceinline__ int op_four_pt(const Src& s, float xx, float yy, float& r)
{
    int min_x = f_floor(xx), min_y = f_floor(yy), max_x = f_ceiling(xx), max_y = f_ceiling(yy);
    if (min_x < s.sx || max_x > s.ex) return ST_NEXT;
    if (min_y < s.sy || max_y > s.ey) return ST_NEXT;
    float a00 = s_at(s, min_x, min_y), a10 = s_at(s, max_x, min_y), a01 = s_at(s, min_x, max_y), a11 = s_at(s, max_x, max_y);
    if (a00 == s.msg || a10 == s.msg || a01 == s.msg || a11 == s.msg) return ST_NEXT;
    if (s.m) {
        // the four mask values together: the short-circuit chain of the source would wait for one load after the other
        const float m00 = m_at(s, min_x, min_y), m10 = m_at(s, max_x, min_y), m01 = m_at(s, min_x, max_y), m11 = m_at(s, max_x, max_y);
        if (masked_val(s, m00) || masked_val(s, m10) || masked_val(s, m01) || masked_val(s, m11)) return ST_NEXT;
    }
    if (min_x == max_x) {
        if (min_y == max_y) r = a00;
        else r = a00 * ((float)max_y - yy) + a01 * (yy - (float)min_y);
    } else if (min_y == max_y) {
        r = a00 * ((float)max_x - xx) + a10 * (xx - (float)min_x);
    } else {
        r = (yy - (float)min_y) * (a01 * ((float)max_x - xx) + a11 * (xx - (float)min_x)) +
            ((float)max_y - yy) * (a00 * ((float)max_x - xx) + a10 * (xx - (float)min_x));
    }
    return ST_VAL;
}

// four_pt_average (:623-734) and wt_four_pt_average (:742-853, with the `ifx < start_y` slip at :795)
template <bool WT>
__device__ __forceinline__ int op_avg4(const Src& s, float xx, float yy, float& r)
{
    float fxfy = 1.0f, fxcy = 1.0f, cxfy = 1.0f, cxcy = 1.0f;
    int ifx = f_floor(xx), icx = f_ceiling(xx), ify = f_floor(yy), icy = f_ceiling(yy);
    if (WT) {
        float dxf = xx - (float)ifx, dxc = xx - (float)icx, dyf = yy - (float)ify, dyc = yy - (float)icy;
        fxfy = fmaxf(0.0f, 1.0f - t_sqrt(dxf * dxf + dyf * dyf));
        fxcy = fmaxf(0.0f, 1.0f - t_sqrt(dxf * dxf + dyc * dyc));
        cxfy = fmaxf(0.0f, 1.0f - t_sqrt(dxc * dxc + dyf * dyf));
        cxcy = fmaxf(0.0f, 1.0f - t_sqrt(dxc * dxc + dyc * dyc));
    }
    if (ifx < s.sx || icx > s.ex || ify < s.sy || icy > s.ey) {
        if (xx > (float)s.sx - 0.5f && ifx < s.sx) { ifx = s.sx; icx = s.sx; }
        else if (xx < (float)s.ex + 0.5f && icx > s.ex) { ifx = s.ex; icx = s.ex; }
        if (yy > (float)s.sy - 0.5f && (WT ? (ifx < s.sy) : (ify < s.sy))) { ify = s.sy; icy = s.sy; }
        else if (yy < (float)s.ey + 0.5f && icy > s.ey) { ify = s.ey; icy = s.ey; }
        if (ifx < s.sx || icx > s.ex || ify < s.sy || icy > s.ey) return ST_FINAL_MSG;
    }
    float a_ff = s_at(s, ifx, ify), a_fc = s_at(s, ifx, icy), a_cf = s_at(s, icx, ify), a_cc = s_at(s, icx, icy);
    bool m_ff = false, m_fc = false, m_cf = false, m_cc = false;
    if (s.m) {
        const float v_ff = m_at(s, ifx, ify), v_fc = m_at(s, ifx, icy), v_cf = m_at(s, icx, ify), v_cc = m_at(s, icx, icy);
        m_ff = masked_val(s, v_ff); m_fc = masked_val(s, v_fc); m_cf = masked_val(s, v_cf); m_cc = masked_val(s, v_cc);
    }
    if (a_ff == s.msg || m_ff) fxfy = 0.0f;
    if (a_fc == s.msg || m_fc) fxcy = 0.0f;
    if (a_cf == s.msg || m_cf) cxfy = 0.0f;
    if (a_cc == s.msg || m_cc) cxcy = 0.0f;
    if (fxfy == 0.0f && fxcy == 0.0f && cxfy == 0.0f && cxcy == 0.0f) return ST_NEXT;
    r = (fxfy * a_ff + fxcy * a_fc + cxfy * a_cf + cxcy * a_cc) / (fxfy + fxcy + cxfy + cxcy);
    return ST_VAL;
}

// sixteen_pt_average (:861-950) and wt_sixteen_pt_average (:958-1047); summation order i=1..4, j=1..4 over
// the points (ifx+3-i, ify+3-j)
template <bool WT>
__device__ __forceinline__ int op_avg16(const Src& s, float xx, float yy, float& r)
{
    int ifx = f_floor(xx), ify = f_floor(yy);
    if (ifx < s.sx + 1 || ifx > s.ex - 2 || ify < s.sy + 1 || ify > s.ey - 2) return ST_NEXT;
    float sum_weight = 0.0f, sum = 0.0f;
    // the reference computes all weights first, then the weighted sum; both accumulate in the same order,
    // so two independent running sums reproduce it exactly
#pragma unroll
    for (int i = 1; i <= 4; ++i) {
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            int px = ifx + 3 - i, py = ify + 3 - j;
            float v = s_at(s, px, py), w;
            if (masked_out(s, px, py)) w = 0.0f;
            else if (WT) {
                float ddx = xx - (float)px, ddy = yy - (float)py;
                w = fmaxf(0.0f, 2.0f - t_sqrt(ddx * ddx + ddy * ddy));
            } else w = 1.0f;
            if (v == s.msg) w = 0.0f;
            sum_weight = sum_weight + w;
            sum = sum + w * v;
        }
    }
    if (sum_weight == 0.0f) return ST_NEXT;
    r = sum / sum_weight;
    return ST_VAL;
}

// oned, interp_module.F:1342-1374
__device__ __forceinline__ float oned(float x, float a, float b, float c, float d)
{
    float r = 0.0f;
    if (x == 0.0f) r = b; else if (x == 1.0f) r = c;
    if (b * c != 0.0f) {
        if (a * d == 0.0f) {
            if (a == 0.0f && d == 0.0f) r = b * (1.0f - x) + c * x;
            else if (a != 0.0f) r = b + x * (0.5f * (c - a) + x * (0.5f * (c + a) - b));
            else if (d != 0.0f) r = c + (1.0f - x) * (0.5f * (b - d) + (1.0f - x) * (0.5f * (b + d) - c));
        } else {
            r = (1.0f - x) * (b + x * (0.5f * (c - a) + x * (0.5f * (c + a) - b))) +
                x * (c + (1.0f - x) * (0.5f * (b - d) + (1.0f - x) * (0.5f * (b + d) - c)));
        }
    }
    return r;
}

// sixteen_pt, interp_module.F:1179-1334
__device__ __forceinline__ int op_sixteen_pt(const Src& s, float xx, float yy, float& r)
{
    if ((int)xx < s.sx || (int)xx > s.ex || (int)yy < s.sy || (int)yy > s.ey) return ST_NEXT;
    int i = (int)(xx + 0.00001f), j = (int)(yy + 0.00001f);
    float x = xx - (float)i, y = yy - (float)j;
    if (fabsf(x) > 0.0001f || fabsf(y) > 0.0001f) {
        float stl[4][4];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int kk = min(max(i + k - 1, s.sx), s.ex);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                int ll = min(max(j + l - 1, s.sy), s.ey);
                float v = s_at(s, kk, ll);
                if (s.m && masked_out(s, kk, ll)) bad = true;
                if (v == 0.0f && s.msg != 0.0f) v = 1.0e-20f;
                if (v == s.msg) bad = true;
                stl[k][l] = v;
            }
        }
        if (bad) return ST_NEXT;
        float a = oned(x, stl[0][0], stl[1][0], stl[2][0], stl[3][0]);
        float b = oned(x, stl[0][1], stl[1][1], stl[2][1], stl[3][1]);
        float c = oned(x, stl[0][2], stl[1][2], stl[2][2], stl[3][2]);
        float d = oned(x, stl[0][3], stl[1][3], stl[2][3], stl[3][3]);
        r = oned(y, a, b, c, d);
        if (r == 1.0e-20f) r = 0.0f;
        return ST_VAL;
    }
    if (i >= s.sx && i <= s.ex && j >= s.sy && j <= s.ey && mask_valid(s, i, j)) {
        r = s_at(s, i, j);
        if (r != s.msg) return ST_VAL;
    }
    return ST_NEXT;
}

// ---- search_extrap (interp_module.F:451-615), warp-cooperative closed form ----
// The BFS of the reference explores cells irrespective of their values, so its pop order is geometric:
// Manhattan ring by ring; within ring k: dx=-k..-1 ascending (dy<0 before dy>0), dx=+k..+1 descending
// (dy<0 first), then (0,-k), (0,+k); cells outside the array are simply absent.  A ring-(k+1) cell is
// pushed by (dx, dy-sgn dy) when dy/=0, else by (dx-sgn dx, 0).  When the first valid cell F (ring k,
// position t_F) is popped, the queue holds the rest of ring k and the ring-(k+1) cells pushed by ring-k
// cells up to and including F; the result is the closest valid one among F and those (strict <, FIFO
// order).  This is exact as long as the cumulative depth counter (<= ring+3) never reaches maxdepth,
// i.e. for k_F + 4 <= maxdepth; otherwise the caller falls back to the literal queue simulation.
__device__ __forceinline__ void ring_node(int k, int t, int& dx, int& dy)
{
    if (k == 0) { dx = 0; dy = 0; return; }
    int na = 2 * k - 1;
    if (t < na) {
        if (t == 0) { dx = -k; dy = 0; }
        else { int u = t - 1; dx = -k + 1 + (u >> 1); int m = k + dx; dy = (u & 1) ? m : -m; }
    } else if (t < 2 * na) {
        int tp = t - na;
        if (tp == 0) { dx = k; dy = 0; }
        else { int u = tp - 1; dx = k - 1 - (u >> 1); int m = k - dx; dy = (u & 1) ? m : -m; }
    } else {
        dx = 0; dy = (t == 2 * na) ? -k : k;
    }
}
__device__ __forceinline__ int ring_index(int k, int dx, int dy)
{
    if (k == 0) return 0;
    int na = 2 * k - 1;
    if (dx < 0) { if (dy == 0) return 0; return 1 + 2 * (dx + k - 1) + (dy > 0 ? 1 : 0); }
    if (dx > 0) { if (dy == 0) return na; return na + 1 + 2 * (k - 1 - dx) + (dy > 0 ? 1 : 0); }
    return 2 * na + (dy > 0 ? 1 : 0);
}
__device__ __forceinline__ bool cell_valid(const Src& s, int i, int j)
{
    if (i < s.sx || i > s.ex || j < s.sy || j > s.ey) return false;
    return s_at(s, i, j) != s.msg && mask_valid(s, i, j);
}

// returns: 0 found (r set), 1 not found and exact (r = msg), 2 closed form not applicable -> literal fallback
static __device__ int warp_search(const Src& s, float xx, float yy, int maxdepth, float& r)
{
    const unsigned FULL = 0xffffffffu;
    int lane = threadIdx.x & 31;
    int x0 = f_nint(xx), y0 = f_nint(yy);
    if (x0 < s.sx || x0 > s.ex || y0 < s.sy || y0 > s.ey) { r = s.msg; return 1; }
    int kmax_geo = max(x0 - s.sx, s.ex - x0) + max(y0 - s.sy, s.ey - y0);  // farthest cell of the array
    for (int k = 0; k <= kmax_geo; ++k) {
        if (k + 4 > maxdepth) return 2;
        int n = (k == 0) ? 1 : 4 * k;
        int tF = -1;
        for (int base = 0; base < n; base += 32) {
            int t = base + lane;
            bool v = false;
            if (t < n) { int dx, dy; ring_node(k, t, dx, dy); v = cell_valid(s, x0 + dx, y0 + dy); }
            unsigned b = __ballot_sync(FULL, v);
            if (b) { tF = base + __ffs(b) - 1; break; }
        }
        if (tF < 0) continue;
        // candidates: F, ring-k cells after F, ring-(k+1) cells whose discoverer index <= tF
        int fdx, fdy;
        ring_node(k, tF, fdx, fdy);
        float best_d = ((float)(x0 + fdx) - xx) * ((float)(x0 + fdx) - xx) + ((float)(y0 + fdy) - yy) * ((float)(y0 + fdy) - yy);
        float my_d = best_d; int my_seq = 0x7fffffff; int my_x = x0 + fdx, my_y = y0 + fdy;
        if (lane == 0) my_seq = 0;
        for (int t = tF + 1 + lane; t < n; t += 32) {
            int dx, dy; ring_node(k, t, dx, dy);
            int ci = x0 + dx, cj = y0 + dy;
            if (cell_valid(s, ci, cj)) {
                float d = ((float)ci - xx) * ((float)ci - xx) + ((float)cj - yy) * ((float)cj - yy);
                int seq = t - tF;
                if (d < my_d || (d == my_d && seq < my_seq)) { my_d = d; my_seq = seq; my_x = ci; my_y = cj; }
            }
        }
        int n1 = 4 * (k + 1);
        for (int t = lane; t < n1; t += 32) {
            int dx, dy; ring_node(k + 1, t, dx, dy);
            int pdx = dx, pdy = dy;
            if (dy != 0) pdy = dy - (dy > 0 ? 1 : -1); else pdx = dx - (dx > 0 ? 1 : -1);
            if (ring_index(k, pdx, pdy) > tF) continue;
            int ci = x0 + dx, cj = y0 + dy;
            if (cell_valid(s, ci, cj)) {
                float d = ((float)ci - xx) * ((float)ci - xx) + ((float)cj - yy) * ((float)cj - yy);
                int seq = n + t;
                if (d < my_d || (d == my_d && seq < my_seq)) { my_d = d; my_seq = seq; my_x = ci; my_y = cj; }
            }
        }
        // warp arg-min on (distance, sequence)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float od = __shfl_xor_sync(FULL, my_d, off);
            int os = __shfl_xor_sync(FULL, my_seq, off);
            int ox = __shfl_xor_sync(FULL, my_x, off), oy = __shfl_xor_sync(FULL, my_y, off);
            if (od < my_d || (od == my_d && os < my_seq)) { my_d = od; my_seq = os; my_x = ox; my_y = oy; }
        }
        r = s_at(s, my_x, my_y);
        return 0;
    }
    r = s.msg;
    return 1;
}

// Literal single-thread simulation of the reference's queue (exact for any maxdepth).  Scratch: a visited
// bitmap over the (2R+1)^2 window around the origin and a ring-buffer queue of capacity qcap.
struct QNode { short dx, dy; int depth; };
static __device__ float literal_search(const Src& s, float xx, float yy, int maxdepth, unsigned* visited, int R, QNode* q, int qcap)
{
    int x0 = f_nint(xx), y0 = f_nint(yy);
    if (x0 < s.sx || x0 > s.ex || y0 < s.sy || y0 > s.ey) return s.msg;
    int W = 2 * R + 1;
    size_t nwords = ((size_t)W * W + 31) / 32;
    for (size_t w = 0; w < nwords; ++w) visited[w] = 0u;
    auto vis_test = [&](int dx, int dy) { size_t b = (size_t)(dy + R) * W + (dx + R); return (visited[b >> 5] >> (b & 31)) & 1u; };
    auto vis_set = [&](int dx, int dy) { size_t b = (size_t)(dy + R) * W + (dx + R); visited[b >> 5] |= 1u << (b & 31); };
    long head = 0, tail = 0;
    auto push = [&](int dx, int dy, int depth) { QNode nd; nd.dx = (short)dx; nd.dy = (short)dy; nd.depth = depth; q[tail % qcap] = nd; ++tail; };
    push(0, 0, 0); vis_set(0, 0);
    bool found = false;
    int fi = 0, fj = 0;
    while (head < tail && !found) {
        QNode nd = q[head % qcap]; ++head;
        int dx = nd.dx, dy = nd.dy, depth = nd.depth, i = x0 + dx, j = y0 + dy;
        if (s_at(s, i, j) != s.msg && mask_valid(s, i, j)) { found = true; fi = i; fj = j; }
        if (i - 1 >= s.sx && depth < maxdepth && !vis_test(dx - 1, dy)) { depth++; push(dx - 1, dy, depth); vis_set(dx - 1, dy); }
        if (i + 1 <= s.ex && depth < maxdepth && !vis_test(dx + 1, dy)) { depth++; push(dx + 1, dy, depth); vis_set(dx + 1, dy); }
        if (j - 1 >= s.sy && depth < maxdepth && !vis_test(dx, dy - 1)) { depth++; push(dx, dy - 1, depth); vis_set(dx, dy - 1); }
        if (j + 1 <= s.ey && depth < maxdepth && !vis_test(dx, dy + 1)) { depth++; push(dx, dy + 1, depth); vis_set(dx, dy + 1); }
    }
    if (!found) return s.msg;
    float distance = ((float)fi - xx) * ((float)fi - xx) + ((float)fj - yy) * ((float)fj - yy);
    float r = s_at(s, fi, fj);
    while (head < tail) {
        QNode nd = q[head % qcap]; ++head;
        int i = x0 + nd.dx, j = y0 + nd.dy;
        float v = s_at(s, i, j);
        if (v != s.msg && mask_valid(s, i, j)) {
            float d = ((float)i - xx) * ((float)i - xx) + ((float)j - yy) * ((float)j - yy);
            if (d < distance) { distance = d; r = v; }
        }
    }
    return r;
}

// interp_sequence, interp_module.F:304-367, as a loop over the 0-terminated method list.
// MODE 0: one thread per point; a SEARCH method defers the point (ST_DEFER).
// MODE 1: one warp per point (all lanes hold identical arguments); SEARCH runs warp_search and may
//         request the literal fallback (ST_DEFER again).
// MODE 2: one thread per point with the literal queue simulation (scratch in `ls`).
struct LitScratch { unsigned* visited; int R; QNode* q; int qcap; };
template <int MODE>
__device__ __forceinline__ int interp_sequence(const Src& s, const int* __restrict__ ops, const int* __restrict__ opts,
                                               float xx, float yy, float& r, const LitScratch* ls = nullptr, int first = 0)
{
    // `first` > 0: the caller already knows that methods 0..first-1 hand over to their successor at this point
    for (int idx = first; idx < WPSGPU_MAX_OPS; ++idx) {
        int op = ops[idx];
        if (op == 0) break;
        int st;
        switch (op) {
        case WPSGPU_FOUR_POINT: st = op_four_pt(s, xx, yy, r); break;
        case WPSGPU_AVERAGE4: st = op_avg4<false>(s, xx, yy, r); break;
        case WPSGPU_W_AVERAGE4: st = op_avg4<true>(s, xx, yy, r); break;
        case WPSGPU_N_NEIGHBOR: st = op_nearest(s, xx, yy, r); break;
        case WPSGPU_SIXTEEN_POINT: st = op_sixteen_pt(s, xx, yy, r); break;
        case WPSGPU_AVERAGE16: st = op_avg16<false>(s, xx, yy, r); break;
        case WPSGPU_W_AVERAGE16: st = op_avg16<true>(s, xx, yy, r); break;
        case WPSGPU_SEARCH:
            if (MODE == 0) return ST_DEFER;
            else if (MODE == 2) {
                r = literal_search(s, xx, yy, opts[idx], ls->visited, ls->R, ls->q, ls->qcap);
                return (r == s.msg) ? ST_FINAL_MSG : ST_VAL;
            } else {
                int w = warp_search(s, xx, yy, opts[idx], r);
                if (w == 2) return ST_DEFER;
                // search_extrap never tries the next method: value found or msgval
                if (w == 1) { r = s.msg; return ST_FINAL_MSG; }
                return ST_VAL;
            }
        default: st = ST_NEXT; break;
        }
        if (st == ST_VAL) return ST_VAL;
        if (st == ST_FINAL_MSG) { r = s.msg; return ST_FINAL_MSG; }
    }
    r = s.msg;
    return ST_FINAL_MSG;
}

}  // namespace wps
