// metgrid.cu -- batched replacement of metgrid's interp_met_field (process_domain_module.F:2205-2669):
// coordinate cache, one launch over (destination tile x slab group), deferred search passes.
#include <algorithm>
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
};

struct SlabPtrs {
    const float* src;
    float* r_arr;
    const int* valid;
    int* new_pts;
};


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

// ---- per-point logic of interp_met_field (non-gcell branch, :2487-2579) ------------------------
// out flags
enum { PF_WRITE = 1, PF_BIT = 2, PF_DEFER = 4 };

template <int MODE>
__device__ __forceinline__ int interp_to_latlon(const Job& jb, const float* __restrict__ src, int sel, bool use_mask,
                                                float2 xy, float rlat, float rlon, float& r, const LitScratch* ls)
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
        st = interp_sequence<MODE>(s, jb.ops, jb.opts, xy.x, xy.y, r, ls);
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
                                             float2 xy, bool valid_bit, float& out, const LitScratch* ls = nullptr)
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
        if (jb.masked == WPSGPU_MASKED_BOTH) {
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
        int st = interp_to_latlon<MODE>(jb, src, sel, jb.mask[sel] != nullptr, xy, rlat, rlon, temp, ls);
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
__global__ void __launch_bounds__(TILE_X * TILE_Y)
k_metgrid_interp(const Job* __restrict__ jobs, int njobs, const SlabPtrs* __restrict__ slabs,
                 Deferred* __restrict__ deferred, unsigned long long* __restrict__ defer_count)
{
    __shared__ Job jb;
    __shared__ int s_job;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int lo = 0, hi = njobs - 1, b = blockIdx.x;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].tile0 <= b) lo = mid; else hi = mid - 1;
        }
        s_job = lo;
    }
    __syncthreads();
    {
        const int* gsrc = reinterpret_cast<const int*>(jobs + s_job);
        int* sdst = reinterpret_cast<int*>(&jb);
        int tid = threadIdx.y * TILE_X + threadIdx.x;
        for (int w = tid; w < (int)(sizeof(Job) / sizeof(int)); w += TILE_X * TILE_Y) sdst[w] = gsrc[w];
    }
    __syncthreads();
    int t = blockIdx.x - jb.tile0;
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
            flags = process_point<0>(jb, sp.src, i, j, idx, xy, vbit, out);
            if (flags & PF_DEFER) {
                unsigned long long slot = atomicAdd(defer_count, 1ull);
                Deferred d; d.slab = jb.slab0 + sl; d.idx = (int)idx;
                deferred[slot] = d;
            } else if (flags & PF_WRITE) sp.r_arr[idx] = out;
        }
        unsigned word = __ballot_sync(0xffffffffu, (flags & PF_BIT) != 0);
        if (threadIdx.x == 0 && jj < jb.ny) sp.new_pts[widx] = (int)word;
    }
}

// ---- deferred passes: warp per point (closed-form search), then thread per point (literal) ---------
__device__ __forceinline__ int job_of_slab(const Job* jobs, int njobs, int slab)
{
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].slab0 <= slab) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void finish_point(const Job& jb, const SlabPtrs& sp, int idx, int flags, float out)
{
    if (flags & PF_WRITE) sp.r_arr[idx] = out;
    if (flags & PF_BIT) {
        int ii = idx % jb.nx, jj = idx / jb.nx;
        atomicOr(sp.new_pts + (size_t)jj * jb.nxw + (ii >> 5), (int)(1u << (ii & 31)));
    }
}

__global__ void __launch_bounds__(256)
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
        const Job& jb = jobs[job_of_slab(jobs, njobs, d.slab)];
        const SlabPtrs sp = slabs[d.slab];
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
        const Job& jb = jobs[job_of_slab(jobs, njobs, d.slab)];
        const SlabPtrs sp = slabs[d.slab];
        int ii = d.idx % jb.nx, jj = d.idx / jb.nx;
        bool vbit = false;
        if (sp.valid) vbit = (__ldg(sp.valid + (size_t)jj * jb.nxw + (ii >> 5)) >> (ii & 31)) & 1;
        float out = 0.0f;
        int flags = process_point<2>(jb, sp.src, jb.sm1 + ii, jb.sm2 + jj, (size_t)d.idx, __ldg(jb.xy + d.idx), vbit, out, &ls);
        finish_point(jb, sp, d.idx, flags, out);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
static std::string cache_key(int grid_id, const wpsgpu_proj& p)
{
    std::string k((const char*)&grid_id, sizeof(int));
    size_t nbytes = offsetof(wpsgpu_proj, gauss_lat);
    k.append((const char*)&p, nbytes);
    return k;
}

static int rel_code(char c) { return c == '<' ? REL_LT : (c == '>' ? REL_GT : REL_EQ); }

static bool same_group(const QueuedSlab& a, const QueuedSlab& b)
{
    if (memcmp(&a.proj, &b.proj, offsetof(wpsgpu_proj, gauss_lat)) != 0) return false;
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
        if (gid == grid_id) { cudaFree(it->second.d_xy); it = h->coord_cache.erase(it); } else ++it;
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
    WPS_CHECK(!slab->use_landmask || h->grids[slab->grid_id].d_landmask, -1, "use_landmask set but no landmask given for grid %d", slab->grid_id);
    QueuedSlab q;
    q.proj = *src_proj; q.fo = *fo; q.s = *slab;
    h->queue.push_back(q);
    return 0;
}

int wpsgpu_metgrid_last_kernel(wpsgpu_handle_t h, float* ms, double* points)
{
    WPS_CHECK(h && ms && points, -1, "NULL argument");
    WPS_CUDA(cudaEventSynchronize(h->ev_k1));
    WPS_CUDA(cudaEventElapsedTime(ms, h->ev_k0, h->ev_k1));
    *points = h->last_points;
    return 0;
}

int wpsgpu_metgrid_flush(wpsgpu_handle_t h)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    if (h->queue.empty()) return 0;
    WPS_CUDA(cudaSetDevice(h->device));
    std::vector<QueuedSlab> queue;
    queue.swap(h->queue);
    const bool host = h->ptr_mode == WPSGPU_PTR_HOST;
    const int nq = (int)queue.size();

    // ---- group slabs into jobs (stable: first-appearance order) ----
    std::vector<int> job_of(nq, -1);
    std::vector<std::vector<int>> groups;
    for (int a = 0; a < nq; ++a) {
        bool placed = false;
        for (size_t g = 0; g < groups.size() && !placed; ++g)
            if (same_group(queue[groups[g][0]], queue[a])) { groups[g].push_back(a); placed = true; }
        if (!placed) groups.push_back(std::vector<int>(1, a));
    }
    const int njobs = (int)groups.size();

    // ---- arena sizing ----
    size_t need = (size_t)njobs * sizeof(Job) + (size_t)nq * sizeof(SlabPtrs) + (1 << 20);
    size_t defer_entries = 0;
    int max_depth = 0, max_span = 0;
    std::map<const void*, size_t> uniq_in;  // host source/mask pointers staged once
    for (int g = 0; g < njobs; ++g) {
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        size_t sbytes = (size_t)(q0.s.maxx - q0.s.minx + 1) * (q0.s.maxy - q0.s.miny + 1) * sizeof(float);
        bool has_search = false;
        for (int k = 0; k < WPSGPU_MAX_OPS && q0.fo.ops[k]; ++k)
            if (q0.fo.ops[k] == WPSGPU_SEARCH) { has_search = true; max_depth = std::max(max_depth, q0.fo.opts[k]); }
        if (has_search) {
            defer_entries += dg.npts() * groups[g].size();
            max_span = std::max(max_span, std::max(q0.s.maxx - q0.s.minx, q0.s.maxy - q0.s.miny));
        }
        if (q0.proj.code == WPSGPU_PROJ_GAUSS) need += sizeof(float) * WPSGPU_MAX_GAUSS + 256;
        if (host) {
            for (int m = 0; m < 3; ++m) if (q0.s.mask[m]) uniq_in[q0.s.mask[m]] = sbytes;
            for (int a : groups[g]) {
                uniq_in[queue[a].s.slab] = sbytes;
                need += dg.npts() * sizeof(float) + 256;                       // r_arr
                need += 2 * ((size_t)dg.ny() * ((dg.nx() + 31) / 32) * sizeof(int) + 256);  // valid + new_pts
            }
        }
    }
    for (auto& kv : uniq_in) need += kv.second + 256;
    need += 2 * (defer_entries * sizeof(Deferred) + 256);
    h->arena.reset();
    WPS_TRY(h->arena.reserve(need, h->stream));

    // ---- coordinate cache ----
    std::vector<const float2*> job_xy(njobs);
    std::vector<const float*> job_gauss(njobs, nullptr);
    for (int g = 0; g < njobs; ++g) {
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        WPS_TRY(upload_gauss(h, q0.proj, &job_gauss[g]));
        std::string key = cache_key(q0.s.grid_id, q0.proj);
        auto it = h->coord_cache.find(key);
        if (it == h->coord_cache.end() || it->second.grid_version != dg.version) {
            CoordCacheEntry e;
            e.grid_version = dg.version;
            WPS_CUDA(cudaMalloc((void**)&e.d_xy, dg.npts() * sizeof(float2)));
            DevProj dp = to_dev(q0.proj, job_gauss[g]);
            int64_t n = (int64_t)dg.npts();
            k_coord_cache<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dp, dg.d_xlat, dg.d_xlon, n, e.d_xy);
            h->launches++;
            WPS_CUDA(cudaGetLastError());
            if (it != h->coord_cache.end()) { cudaFree(it->second.d_xy); h->coord_cache.erase(it); }
            it = h->coord_cache.insert(std::make_pair(key, e)).first;
        }
        job_xy[g] = it->second.d_xy;
    }

    // ---- stage inputs, build tables ----
    std::map<const void*, void*> staged;
    auto stage_shared = [&](const float* p, size_t bytes, const float** out) -> int {
        if (!p) { *out = nullptr; return 0; }
        if (!host) { *out = p; return 0; }
        auto it = staged.find(p);
        if (it != staged.end()) { *out = (const float*)it->second; return 0; }
        void* d;
        WPS_TRY(stage_in(h, p, bytes, &d));
        staged[p] = d;
        *out = (const float*)d;
        return 0;
    };
    std::vector<Job> jobs(njobs);
    std::vector<SlabPtrs> sp(nq);
    struct OutRec { float* h_r; void* d_r; int* h_np; void* d_np; size_t rbytes, wbytes; };
    std::vector<OutRec> outs;
    int slab_cursor = 0, tile_cursor = 0;
    unsigned long long defer_cursor = 0;
    bool any_search = false;
    for (int g = 0; g < njobs; ++g) {
        const QueuedSlab& q0 = queue[groups[g][0]];
        const DstGrid& dg = h->grids[q0.s.grid_id];
        Job& jb = jobs[g];
        memset(&jb, 0, sizeof(Job));
        jb.xy = job_xy[g]; jb.xlat = dg.d_xlat; jb.xlon = dg.d_xlon;
        jb.landmask = q0.s.use_landmask ? dg.d_landmask : nullptr;
        jb.sm1 = dg.sm1; jb.em1 = dg.em1; jb.sm2 = dg.sm2; jb.em2 = dg.em2;
        jb.nx = dg.nx(); jb.ny = dg.ny(); jb.nxw = (jb.nx + 31) / 32;
        jb.sd1 = h->sd1; jb.ed1 = h->ed1; jb.sd2 = h->sd2; jb.ed2 = h->ed2;
        // process_width, process_domain_module.F:2270-2278
        if (q0.s.process_bdy_width == 0) jb.pw = std::max(h->ed1 - h->sd1 + 1, h->ed2 - h->sd2 + 1);
        else { jb.pw = q0.s.process_bdy_width; if (q0.s.field_stagger != WPSGPU_M) jb.pw += 2; }
        jb.proj = to_dev(q0.proj, job_gauss[g]);
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
        jb.tile0 = tile_cursor;
        tile_cursor += jb.tiles_x * jb.tiles_y;
        bool has_search = false;
        for (int k = 0; k < WPSGPU_MAX_OPS && jb.ops[k]; ++k) if (jb.ops[k] == WPSGPU_SEARCH) has_search = true;
        jb.defer_base = defer_cursor;
        if (has_search) { defer_cursor += (unsigned long long)dg.npts() * groups[g].size(); any_search = true; }
        size_t rbytes = dg.npts() * sizeof(float), wbytes = (size_t)jb.ny * jb.nxw * sizeof(int);
        for (int a : groups[g]) {
            const wpsgpu_slab& s = queue[a].s;
            SlabPtrs& p = sp[slab_cursor++];
            WPS_TRY(stage_shared(s.slab, sbytes, &p.src));
            if (host) {
                void* d_r; void* d_np; void* d_v = nullptr;
                if (s.valid_mask) {  // r_arr carries lower-priority data only where valid bits are set
                    WPS_TRY(stage_in(h, s.r_arr, rbytes, &d_r));
                    WPS_TRY(stage_in(h, s.valid_mask, wbytes, &d_v));
                } else WPS_TRY(alloc_out(h, s.r_arr, rbytes, &d_r));
                WPS_TRY(alloc_out(h, s.new_pts, wbytes, &d_np));
                p.r_arr = (float*)d_r; p.valid = (const int*)d_v; p.new_pts = (int*)d_np;
                OutRec o; o.h_r = s.r_arr; o.d_r = d_r; o.h_np = s.new_pts; o.d_np = d_np; o.rbytes = rbytes; o.wbytes = wbytes;
                outs.push_back(o);
            } else {
                p.r_arr = s.r_arr; p.valid = s.valid_mask; p.new_pts = s.new_pts;
            }
        }
    }
    Job* d_jobs = (Job*)h->arena.take(sizeof(Job) * njobs);
    SlabPtrs* d_sp = (SlabPtrs*)h->arena.take(sizeof(SlabPtrs) * nq);
    WPS_CHECK(d_jobs && d_sp, -3, "device arena exhausted (tables)");
    WPS_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(Job) * njobs, cudaMemcpyHostToDevice, h->stream));
    WPS_CUDA(cudaMemcpyAsync(d_sp, sp.data(), sizeof(SlabPtrs) * nq, cudaMemcpyHostToDevice, h->stream));
    Deferred *d_def = nullptr, *d_def2 = nullptr;
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(h->d_counters);
    if (any_search) {
        d_def = (Deferred*)h->arena.take(sizeof(Deferred) * defer_cursor);
        d_def2 = (Deferred*)h->arena.take(sizeof(Deferred) * defer_cursor);
        WPS_CHECK(d_def && d_def2, -3, "device arena exhausted (deferred lists)");
        WPS_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), h->stream));
    }

    // ---- launches ----
    dim3 block(TILE_X, TILE_Y);
    WPS_CUDA(cudaEventRecord(h->ev_k0, h->stream));
    k_metgrid_interp<<<tile_cursor, block, 0, h->stream>>>(d_jobs, njobs, d_sp, d_def, d_cnt);
    WPS_CUDA(cudaEventRecord(h->ev_k1, h->stream));
    h->launches++;
    h->last_points = 0.0;
    for (int g = 0; g < njobs; ++g) h->last_points += (double)h->grids[queue[groups[g][0]].s.grid_id].npts() * groups[g].size();
    WPS_CUDA(cudaGetLastError());
    if (any_search) {
        int sblocks = h->sm_count * 8;
        k_metgrid_search<<<sblocks, 256, 0, h->stream>>>(d_jobs, njobs, d_sp, d_def, d_cnt, d_def2, d_cnt + 1);
        h->launches++;
        WPS_CUDA(cudaGetLastError());
        // literal fallback: few threads, big per-thread scratch
        int R = std::min(std::max(max_depth, 1), std::max(max_span, 1));
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
    if (host) {
        for (const OutRec& o : outs) {
            WPS_TRY(copy_out(h, o.h_r, o.d_r, o.rbytes));
            WPS_TRY(copy_out(h, o.h_np, o.d_np, o.wbytes));
        }
        WPS_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

}  // extern "C"
