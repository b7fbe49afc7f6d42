// ctx.h -- handle, error plumbing and the device arena of libwpsgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <deque>
#include <map>
#include <string>
#include <vector>
#include "../../include/wpsgpu.h"
#include "proj.cuh"

namespace wps {

void set_error(const char* fmt, ...);

#define WPS_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            wps::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return -100;                                                                      \
        }                                                                                     \
    } while (0)
#define WPS_CHECK(cond, code, ...)                 \
    do {                                           \
        if (!(cond)) { wps::set_error(__VA_ARGS__); return (code); } \
    } while (0)
#define WPS_TRY(expr)                   \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != 0) return rc_;       \
    } while (0)

// Grow-only bump arena in device memory; reset at the start of every batch.  All per-batch staging
// (source slabs, masks, outputs, job tables) lives here so a flush does O(1) cudaMalloc calls.
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    std::vector<void*> retired;  // older, smaller blocks still referenced by in-flight work
    int reserve(size_t bytes, cudaStream_t st);
    void* take(size_t bytes)
    {
        size_t a = (off + 255) & ~size_t(255);
        if (a + bytes > cap) return nullptr;
        off = a + bytes;
        return base + a;
    }
    // continue the previous block without alignment padding (lets neighbouring host arrays be copied in one piece)
    void* take_tail(size_t bytes)
    {
        if (off + bytes > cap) return nullptr;
        off += bytes;
        return base + off - bytes;
    }
    char* tail() const { return base + off; }
    void reset() { off = 0; }
    void release();
};

struct DstGrid {
    bool set = false;
    int sm1 = 0, em1 = 0, sm2 = 0, em2 = 0;
    float *d_xlat = nullptr, *d_xlon = nullptr, *d_landmask = nullptr;
    bool own = false;            // device copies owned by the handle (host pointer mode)
    uint64_t version = 0;
    int nx() const { return em1 - sm1 + 1; }
    int ny() const { return em2 - sm2 + 1; }
    size_t npts() const { return (size_t)nx() * ny(); }
};

struct CoordCacheEntry {
    uint64_t grid_version;
    float2* d_xy;  // (rx, ry) of lltoxy(xlat, xlon) with the source projection, M stagger
    int bbox[4];   // min/max of floor(rx), floor(ry) over the destination grid (source-cell bounding box)
    int rbbox[4];  // the same including the lon+360 retry position of interp_to_latlon (:2640-2660) for points with lon < 0
    float* d_gauss; // device copy of the Gaussian latitude table (PROJ_GAUSS sources), else nullptr
};

inline void free_coord_entry(CoordCacheEntry& e)
{
    cudaFree(e.d_xy);
    if (e.d_gauss) cudaFree(e.d_gauss);
}

struct QueuedSlab {
    const wpsgpu_proj* proj;      // interned in wpsgpu_ctx::proj_pool (the descriptor carries an 8 KB Gaussian table)
    wpsgpu_field_opts fo;
    wpsgpu_slab s;
    bool gcell = false;           // average_gcell branch: cell averages first, the chain for cells no source pixel reached
    const wpsgpu_proj* dom_proj = nullptr;   // projection of the destination domain (gcell only), interned
};

struct GeoState;  // geogrid.cu
struct MpasState; // mpas.cu

}  // namespace wps

struct wpsgpu_ctx {
    int device = 0;
    int ptr_mode = WPSGPU_PTR_HOST;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
    wps::Arena arena;
    wps::DstGrid grids[8];
    int sd1 = 0, ed1 = 0, sd2 = 0, ed2 = 0;
    bool domain_set = false;
    uint64_t grid_version_counter = 0;
    std::map<std::string, wps::CoordCacheEntry> coord_cache;
    std::vector<wps::QueuedSlab> queue;
    std::deque<wpsgpu_proj> proj_pool;   // projections referenced by the queue; cleared when a flush has consumed it
    wps::GeoState* geo = nullptr;
    wps::MpasState* mpas = nullptr;
    // small persistent device scratch (counters etc.)
    int* d_counters = nullptr;
    // search fallback scratch
    void* d_search_scratch = nullptr;
    size_t search_scratch_bytes = 0;
    // CUDA events bracketing the dominant kernel of the last metgrid flush (roofline measurement)
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // aliases into ev_t[] of the dominant kernel
    cudaEvent_t ev_t[12] = {};  // pack / fast16 / generic / fast4 / slow-bitmap / search brackets
    double last_points = 0.0;
    double last_pts3[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the host-pointer pipeline
    cudaStream_t s_side = nullptr;                  // four_pt / nearest_neighbor first-method kernels beside the sixteen_pt kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_done = nullptr, ev_out_done = nullptr;
    int64_t h2d_bytes = 0, d2h_bytes = 0;           // bytes moved by the host-pointer paths of the metgrid flush
    bool async_pending = false;                     // wpsgpu_metgrid_flush_async: D2H copies may still read the arena
    int precision = WPSGPU_PRECISION_EXACT;   // wpsgpu_metgrid_set_precision
    int timing = 0;             // wpsgpu_metgrid_set_timing: record the CUDA events that bracket the kernels of a flush
    bool timed_flush = false;   // the last flush recorded them
    bool disable_fast = false;  // testing aid: route everything through the generic kernel
};

namespace wps {
// copy helpers honouring the pointer mode: returns a device pointer for `p` (count floats/ints);
// in host mode stages through the arena with an async H2D on the handle's stream.
int stage_in(wpsgpu_ctx* h, const void* p, size_t bytes, void** dptr);
int alloc_out(wpsgpu_ctx* h, void* user_ptr, size_t bytes, void** dptr);
int copy_out(wpsgpu_ctx* h, void* user_ptr, const void* dptr, size_t bytes);
int upload_gauss(wpsgpu_ctx* h, const wpsgpu_proj& p, const float** d_gauss);
int arena_begin(wpsgpu_ctx* h);   // call before re-using the arena: orders the handle's stream behind a pending async flush
}  // namespace wps
