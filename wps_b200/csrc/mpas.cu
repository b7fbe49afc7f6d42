// mpas.cu -- MPAS source remapping for metgrid: remapper.F's remap_info_setup / remap_field and the
// per-level extraction of process_mpas_fields, one thread per target point.
//
// Reference (metgrid/src): remapper.F:83-286 (tables), :417-663 (remap_field), :781-897 (greedy walks), :900-916
// (sphere_distance), :934-1121 (Wachspress coordinates on the sphere), :1124-1326 (search_for_cells with its bounded queue
// and bit dictionary); process_domain_module.F:1700-1956 (planes per level, interface averaging, landmask fill).
//
// Arithmetic: binary32 in the reference's operation order, no contraction (-fmad=false); transcendentals in binary64
// rounded once (proj.cuh t_*, the policy of the projections).  The set-up is a one-off per (mesh, target grid) and is
// cached by the handle; the per-field remap is the part that runs for every field, level and time and is a gather
// bounded by its 4-byte store per output value.
#include "ctx.h"

namespace wps {

struct MpasMeshDev {
    int nCells = 0, nVertices = 0, nEdges = 0, maxEdges = 0;
    int *landmask = nullptr, *nEdgesOnCell = nullptr, *cellsOnCell = nullptr, *verticesOnCell = nullptr, *edgesOnCell = nullptr;
    int *cellsOnVertex = nullptr, *cellsOnEdge = nullptr;
    float *latCell = nullptr, *lonCell = nullptr, *latVertex = nullptr, *lonVertex = nullptr, *latEdge = nullptr, *lonEdge = nullptr;
};

struct MpasTables {            // remap_info_type for one target grid
    bool set = false;
    int nlon = 0, nlat = 0;
    int* nearest[3] = { nullptr, nullptr, nullptr };            // cells, vertices, edges
    int* src[4] = { nullptr, nullptr, nullptr, nullptr };      // cells, vertices, edges, masked cells
    float* w[4] = { nullptr, nullptr, nullptr, nullptr };
    int width[4] = { 3, 0, 0, 3 };
    int64_t nsearch = 0;                                        // points that needed search_for_cells
    int passes = 0;                                             // passes of the sequential recurrence the set-up took
};

struct MpasState {
    bool mesh_set = false;
    MpasMeshDev m;
    MpasTables grid[3];
};

static void free_tables(MpasTables& t)
{
    for (int k = 0; k < 3; ++k) cudaFree(t.nearest[k]);
    for (int k = 0; k < 4; ++k) { cudaFree(t.src[k]); cudaFree(t.w[k]); }
    t = MpasTables();
}
static void free_mesh(MpasMeshDev& m)
{
    cudaFree(m.landmask); cudaFree(m.nEdgesOnCell); cudaFree(m.cellsOnCell); cudaFree(m.verticesOnCell); cudaFree(m.edgesOnCell);
    cudaFree(m.cellsOnVertex); cudaFree(m.cellsOnEdge);
    cudaFree(m.latCell); cudaFree(m.lonCell); cudaFree(m.latVertex); cudaFree(m.lonVertex); cudaFree(m.latEdge); cudaFree(m.lonEdge);
    m = MpasMeshDev();
}

// ---- geometry (remapper.F:900-1121) ------------------------------------------------------------------------------------
__device__ __forceinline__ float sphere_distance(float lat1, float lon1, float lat2, float lon2, float radius)
{
    const float s1 = t_sin(0.5f * (lat2 - lat1));
    const float s2 = t_sin(0.5f * (lon2 - lon1));
    const float arg1 = t_sqrt(s1 * s1 + t_cos(lat1) * t_cos(lat2) * (s2 * s2));
    return 2.0f * radius * t_asin(arg1);
}
__device__ __forceinline__ void convert_lx(float lat, float lon, float radius, float v[3])
{
    v[0] = radius * t_cos(lon) * t_cos(lat);
    v[1] = radius * t_sin(lon) * t_cos(lat);
    v[2] = radius * t_sin(lat);
}
__device__ __forceinline__ float arc_length(float ax, float ay, float az, float bx, float by, float bz)
{
    const float cx = bx - ax, cy = by - ay, cz = bz - az;
    const float r = t_sqrt(ax * ax + ay * ay + az * az);
    const float c = t_sqrt(cx * cx + cy * cy + cz * cz);
    return r * 2.0f * t_asin(c / (2.0f * r));
}
__device__ float triangle_signed_area_sphere(const float a[3], const float b[3], const float c[3], float radius)
{
    const float ab = arc_length(a[0], a[1], a[2], b[0], b[1], b[2]) / radius;
    const float bc = arc_length(b[0], b[1], b[2], c[0], c[1], c[2]) / radius;
    const float ca = arc_length(c[0], c[1], c[2], a[0], a[1], a[2]) / radius;
    const float semiperim = 0.5f * (ab + bc + ca);
    const float prod = t_tan(0.5f * semiperim) * t_tan(0.5f * (semiperim - ab)) * t_tan(0.5f * (semiperim - bc)) * t_tan(0.5f * (semiperim - ca));
    const float tanqe = t_sqrt(prod > 0.0f ? prod : 0.0f);
    float area = 4.0f * radius * radius * t_atan(tanqe);
    const float abl[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, acl[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
    const float d0 = (abl[1] * acl[2]) - (abl[2] * acl[1]);
    const float d1 = -((abl[0] * acl[2]) - (abl[2] * acl[0]));
    const float d2 = (abl[0] * acl[1]) - (abl[1] * acl[0]);
    if ((d0 * a[0] + d1 * a[1] + d2 * a[2]) < 0.0f) area = -area;
    return area;
}
// mpas_wachspress_coordinates for nVertices = 3, the only value a call site passes (:167, :213, :238, :264)
__device__ void wachspress3(const float vert[3][3], const float p[3], float w[3])
{
    const float radiusLocal = t_sqrt(vert[0][0] * vert[0][0] + vert[0][1] * vert[0][1] + vert[0][2] * vert[0][2]);
    float areaA[3], areaB[3], wach[3];
#pragma unroll
    for (int i = 1; i <= 3; ++i) {
        const int im1 = (3 + i - 2) % 3, i0 = (3 + i - 1) % 3, ip1 = (3 + i) % 3;
        areaB[i - 1] = triangle_signed_area_sphere(vert[im1], vert[i0], vert[ip1], radiusLocal);
        areaA[i0] = triangle_signed_area_sphere(p, vert[i0], vert[ip1], radiusLocal);
    }
#pragma unroll
    for (int i = 1; i <= 3; ++i) wach[i - 1] = areaB[i - 1] * areaA[(3 + i) % 3];       // j = i + 1 only: areaA(mod(nV + j - 1, nV) + 1)
    float total = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) total = total + wach[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = wach[i] / total;
}

// ---- greedy walks (remapper.F:781-897) -----------------------------------------------------------------------------------
__device__ int nearest_cell(const MpasMeshDev& m, float tlat, float tlon, int start)
{
    int nearest = start, current = -1;
    while (nearest != current) {
        current = nearest;
        float nd = sphere_distance(m.latCell[current - 1], m.lonCell[current - 1], tlat, tlon, 1.0f);
        const int ne = m.nEdgesOnCell[current - 1];
        for (int i = 0; i < ne; ++i) {
            const int ic = m.cellsOnCell[(size_t)(current - 1) * m.maxEdges + i];
            if (ic <= m.nCells) {
                const float d = sphere_distance(m.latCell[ic - 1], m.lonCell[ic - 1], tlat, tlon, 1.0f);
                if (d < nd) { nearest = ic; nd = d; }
            }
        }
    }
    return nearest;
}
template <int DEG>
__device__ int nearest_node(const MpasMeshDev& m, float tlat, float tlon, int start)
{
    const int* nodesOnCell = DEG == 3 ? m.verticesOnCell : m.edgesOnCell;
    const int* cellsOnNode = DEG == 3 ? m.cellsOnVertex : m.cellsOnEdge;
    const float *latN = DEG == 3 ? m.latVertex : m.latEdge, *lonN = DEG == 3 ? m.lonVertex : m.lonEdge;
    int nearest = start, current = -1;
    while (nearest != current) {
        current = nearest;
        float nd = sphere_distance(latN[current - 1], lonN[current - 1], tlat, tlon, 1.0f);
        const int c1 = cellsOnNode[(size_t)(current - 1) * DEG], c2 = cellsOnNode[(size_t)(current - 1) * DEG + 1];
        const float d1 = sphere_distance(m.latCell[c1 - 1], m.lonCell[c1 - 1], tlat, tlon, 1.0f);
        const float d2 = sphere_distance(m.latCell[c2 - 1], m.lonCell[c2 - 1], tlat, tlon, 1.0f);
        int ic;
        if (DEG == 3) {
            const int c3 = cellsOnNode[(size_t)(current - 1) * DEG + 2];
            const float d3 = sphere_distance(m.latCell[c3 - 1], m.lonCell[c3 - 1], tlat, tlon, 1.0f);
            if (d1 < d2) ic = d1 < d3 ? c1 : c3;
            else ic = d2 < d3 ? c2 : c3;
        } else ic = d1 < d2 ? c1 : c2;
        const int ne = m.nEdgesOnCell[ic - 1];
        for (int i = 0; i < ne; ++i) {
            const int iv = nodesOnCell[(size_t)(ic - 1) * m.maxEdges + i];
            const float d = sphere_distance(latN[iv - 1], lonN[iv - 1], tlat, tlon, 1.0f);
            if (d < nd) { nearest = iv; nd = d; }
        }
    }
    return nearest;
}

// The reference scans the target grid sequentially and starts each walk at the previous point's result (idx, :111-145), so
// R(p) = walk(start = R(p-1), target p), R(-1) = 1.  Where a walk has more than one fixed point (3 % of the points for
// edges, < 0.1 % for vertices, binary32 ties for cells) the result depends on that start, so the sequence is reproduced,
// not approximated: a first guess (a coarse lattice of points walks from cell 1, every point then starts from its lattice
// point's cell; vertices / edges start from one of that cell's) is followed by passes of r'(p) = walk(r(p-1), p) over all
// points in parallel until nothing changes.  By induction a fixed point of that pass IS the sequential result; it is
// reached after a handful of passes because a wrong start only matters at the few start-sensitive points.
constexpr int MPAS_SEED_STRIDE = 16;
__global__ void k_mpas_seed_cells(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, int nlon, int nlat,
                                  int ncx, int ncy, int* __restrict__ seeds)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncx * ncy) return;
    const int ix = min((t % ncx) * MPAS_SEED_STRIDE, nlon - 1), iy = min((t / ncx) * MPAS_SEED_STRIDE, nlat - 1);
    const size_t p = (size_t)iy * nlon + ix;
    seeds[t] = nearest_cell(m, lat[p], lon[p], 1);
}
__global__ void k_mpas_first_guess(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, int nlon, int nlat,
                                   int ncx, const int* __restrict__ seeds, int* __restrict__ ncell, int* __restrict__ nvert, int* __restrict__ nedge)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= nlon || iy >= nlat) return;
    const size_t p = (size_t)iy * nlon + ix;
    const float tlat = lat[p], tlon = lon[p];
    const int c = nearest_cell(m, tlat, tlon, seeds[(iy / MPAS_SEED_STRIDE) * ncx + ix / MPAS_SEED_STRIDE]);
    ncell[p] = c;
    nvert[p] = nearest_node<3>(m, tlat, tlon, m.verticesOnCell[(size_t)(c - 1) * m.maxEdges]);
    nedge[p] = nearest_node<2>(m, tlat, tlon, m.edgesOnCell[(size_t)(c - 1) * m.maxEdges]);
}
// one pass of the sequential recurrence; KIND 0 cells, 1 vertices, 2 edges
template <int KIND>
__global__ void k_mpas_chain(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, size_t npts,
                             const int* __restrict__ in, int* __restrict__ out, int* __restrict__ changed)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const int start = p ? in[p - 1] : 1;
    int r;
    if (KIND == 0) r = nearest_cell(m, lat[p], lon[p], start);
    else if (KIND == 1) r = nearest_node<3>(m, lat[p], lon[p], start);
    else r = nearest_node<2>(m, lat[p], lon[p], start);
    out[p] = r;
    if (r != in[p]) *changed = 1;
}
// The masked cell weights of a point (:264-270): Wachspress weights of the three cells of its vertex, water cells zeroed, the
// rest divided by their sum.  Returns false -- and leaves w3 unspecified -- when no positive land weight remains: the point then
// goes to search_for_cells (:271-281).
__device__ __forceinline__ bool masked_weights(const MpasMeshDev& m, const int sc[3], float tlat, float tlon, float w3[3])
{
    int nland = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) nland += m.landmask[sc[j] - 1] == 1;
    if (nland == 0) return false;                   // sumWeights = 0 whatever the weights are
    float pt[3], vc[3][3];
    convert_lx(tlat, tlon, 6371229.0f, pt);
#pragma unroll
    for (int j = 0; j < 3; ++j) convert_lx(m.latCell[sc[j] - 1], m.lonCell[sc[j] - 1], 6371229.0f, vc[j]);
    wachspress3(vc, pt, w3);
    float sumWeights = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (m.landmask[sc[j] - 1] == 1) sumWeights = sumWeights + w3[j];
        else w3[j] = 0.0f;
    }
    if (!(sumWeights > 0.0f)) return false;
#pragma unroll
    for (int j = 0; j < 3; ++j) w3[j] = w3[j] / sumWeights;
    return true;
}

// The masked table's scan (:251-283) is a recurrence of its own: after a point that goes to search_for_cells, idx holds the
// CELL index nearest_cell returned (:272) and is the next point's start VERTEX, as written.  next[p] = idx after point p.
__global__ void k_mpas_chain_masked(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, size_t npts,
                                    const int* __restrict__ next_in, int* __restrict__ next_out, int* __restrict__ vert_out, int* __restrict__ changed)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float tlat = lat[p], tlon = lon[p];
    const int v = nearest_node<3>(m, tlat, tlon, p ? next_in[p - 1] : 1);
    int sc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sc[j] = m.cellsOnVertex[(size_t)(v - 1) * 3 + j];
    float w3[3];
    const int next = masked_weights(m, sc, tlat, tlon, w3) ? v : nearest_cell(m, tlat, tlon, sc[0]);
    vert_out[p] = v;
    next_out[p] = next;
    if (next != next_in[p]) *changed = 1;
}

// ---- weight tables (remapper.F:158-283) ------------------------------------------------------------------------------------
__global__ void k_mpas_weights(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, size_t npts,
                               const int* __restrict__ ncell, const int* __restrict__ nvert, const int* __restrict__ nvertM,
                               int* __restrict__ srcC, float* __restrict__ wC, int* __restrict__ srcV, float* __restrict__ wV,
                               int* __restrict__ srcE, float* __restrict__ wE, int* __restrict__ srcM, float* __restrict__ wM,
                               int* __restrict__ search_list, unsigned long long* __restrict__ nsearch)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float tlat = lat[p], tlon = lon[p];
    float pt[3], vc[3][3], w3[3];
    convert_lx(tlat, tlon, 6371229.0f, pt);
    // cells of the nearest vertex (:158-175)
    {
        const int v = nvert[p];
        int sc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            sc[j] = m.cellsOnVertex[(size_t)(v - 1) * 3 + j];
            convert_lx(m.latCell[sc[j] - 1], m.lonCell[sc[j] - 1], 6371229.0f, vc[j]);
        }
        wachspress3(vc, pt, w3);
#pragma unroll
        for (int j = 0; j < 3; ++j) { srcC[p * 3 + j] = sc[j]; wC[p * 3 + j] = w3[j]; }
    }
    // the masked variant (:251-283), from the vertex of its own scan
    {
        const int v = nvertM[p];
        int sc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) sc[j] = m.cellsOnVertex[(size_t)(v - 1) * 3 + j];
        if (masked_weights(m, sc, tlat, tlon, w3)) {
#pragma unroll
            for (int j = 0; j < 3; ++j) { srcM[p * 3 + j] = sc[j]; wM[p * 3 + j] = w3[j]; }
        } else {
            // search_for_cells: deferred to k_mpas_search; its start is nearest_cell from sourceMaskedCells(1) (:272)
            srcM[p * 3] = nearest_cell(m, tlat, tlon, sc[0]);
            const unsigned long long k = atomicAdd(nsearch, 1ull);
            search_list[k] = (int)p;
        }
    }
    // vertices (:185-207) and edges (:219-241) of the nearest cell; the Wachspress call passes nVertices = 3 (:204, :238)
    const int c = ncell[p];
    const int nn = m.nEdgesOnCell[c - 1], me = m.maxEdges;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int* nodesOnCell = pass == 0 ? m.verticesOnCell : m.edgesOnCell;
        const float *latN = pass == 0 ? m.latVertex : m.latEdge, *lonN = pass == 0 ? m.lonVertex : m.lonEdge;
        int* srcN = pass == 0 ? srcV : srcE;
        float* wN = pass == 0 ? wV : wE;
        for (int j = 0; j < me; ++j) {
            int v = 1;
            if (j < nn) v = nodesOnCell[(size_t)(c - 1) * me + j];
            srcN[p * me + j] = v;
            if (j < 3 && j < nn) convert_lx(latN[v - 1], lonN[v - 1], 6371229.0f, vc[j]);
        }
        wachspress3(vc, pt, w3);
        for (int j = 0; j < me; ++j) wN[p * me + j] = (j < 3 && j < nn) ? w3[j] : 0.0f;
    }
}

// search_for_cells (:1124-1210) with the module's queue (:1216-1263, max_queue_length 2700: an insert into a full queue is
// dropped) and bit dictionary (:1269-1326, max_dictionary_size 82000 words: cells beyond it are never recorded), one
// thread per point, scratch per worker thread: [queue 2700 | touched-word log | dictionary words]
constexpr int MPAS_MAXQ = 2700, MPAS_MAXDICT = 82000, MPAS_LOG = 8192;
__global__ void k_mpas_search(MpasMeshDev m, const float* __restrict__ lat, const float* __restrict__ lon, const int* __restrict__ list,
                              unsigned long long nlist, int ndw, int* __restrict__ scratch, int* __restrict__ srcM, float* __restrict__ wM)
{
    const size_t worker = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nworkers = (size_t)gridDim.x * blockDim.x;
    int* q = scratch + worker * (size_t)(MPAS_MAXQ + MPAS_LOG + ndw);
    int* log = q + MPAS_MAXQ;
    unsigned* dict = (unsigned*)(log + MPAS_LOG);
    for (unsigned long long e = worker; e < nlist; e += nworkers) {
        const size_t p = (size_t)list[e];
        const float tlat = lat[p], tlon = lon[p];
        const int start = srcM[p * 3];
        int head = 0, tail = 0, size = 0, nlog = 0;
        auto q_insert = [&](int i) {
            if (size == MPAS_MAXQ) return;
            ++size; q[head] = i; head = (head + 1) % MPAS_MAXQ;
        };
        auto d_insert = [&](int i) {
            const int n_integer = ((i - 1) / 32) + 1, n_bit = (i - 1) % 32;
            if (n_integer > MPAS_MAXDICT) return;
            const unsigned old = dict[n_integer - 1];
            if (old == 0u) { if (nlog < MPAS_LOG) log[nlog] = n_integer - 1; ++nlog; }
            dict[n_integer - 1] = old | (1u << n_bit);
        };
        auto d_search = [&](int i) -> bool {
            const int n_integer = ((i - 1) / 32) + 1, n_bit = (i - 1) % 32;
            if (n_integer > MPAS_MAXDICT) return false;
            return (dict[n_integer - 1] >> n_bit) & 1u;
        };
        bool no_more_queueing = false;
        int unscanned = 0, best = 1;
        float d_min = 1000.0f, wbest = 0.0f;
        q_insert(start);
        d_insert(start);
        while (size > 0) {
            --size;
            const int scan = q[tail];
            tail = (tail + 1) % MPAS_MAXQ;
            if (m.landmask[scan - 1] == 1) {
                const float d = sphere_distance(tlat, tlon, m.latCell[scan - 1], m.lonCell[scan - 1], 1.0f);
                if (d < d_min) {
                    d_min = d; best = scan; wbest = 1.0f;
                    if (!no_more_queueing) unscanned = size;
                    no_more_queueing = true;
                }
            }
            if (!no_more_queueing || unscanned > 0) {
                const int ne = m.nEdgesOnCell[scan - 1];
                for (int i = 0; i < ne; ++i) {
                    const int nb = m.cellsOnCell[(size_t)(scan - 1) * m.maxEdges + i];
                    if (!d_search(nb)) { q_insert(nb); d_insert(nb); }
                }
            }
            if (no_more_queueing) unscanned = unscanned - 1;
        }
        srcM[p * 3] = best; srcM[p * 3 + 1] = 1; srcM[p * 3 + 2] = 1;
        wM[p * 3] = wbest; wM[p * 3 + 1] = 0.0f; wM[p * 3 + 2] = 0.0f;
        // dictionary_reset: only the words this search touched (all of them if the log overflowed)
        if (nlog <= MPAS_LOG) for (int k = 0; k < nlog; ++k) dict[log[k]] = 0u;
        else for (int k = 0; k < ndw; ++k) dict[k] = 0u;
    }
}

// ---- remap_field (:417-663) ------------------------------------------------------------------------------------------------
// reference layout, level index fastest: thread e = point * nlev + level, so a warp reads consecutive levels of the same
// few source columns and writes consecutive outputs
template <typename T>
__global__ void k_mpas_remap_levels(const T* __restrict__ src, int nlev, const int* __restrict__ nodes, const float* __restrict__ w,
                                    int tw, int nw, size_t ntot, T* __restrict__ dst)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntot) return;
    const size_t p = e / (size_t)nlev;
    const int k = (int)(e - p * (size_t)nlev);
    T acc = (T)0;
    for (int j = 0; j < nw; ++j) acc = acc + (T)w[p * tw + j] * src[(size_t)(nodes[p * tw + j] - 1) * nlev + k];
    dst[e] = acc;
}
__global__ void k_mpas_remap_nearest(const int* __restrict__ src, int nlev, const int* __restrict__ nearest, size_t ntot, int* __restrict__ dst)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntot) return;
    const size_t p = e / (size_t)nlev;
    dst[e] = src[(size_t)(nearest[p] - 1) * nlev + (int)(e - p * (size_t)nlev)];
}
// planes per level (met_em layout): a thread owns one target point and walks the levels of its three source columns
// (neighbouring points share them: the loads are L1 hits), every store of a warp is one 128-byte row segment of a plane
constexpr int MPAS_MAX_PLANES = 256;
struct MpasPlanes { float* p[MPAS_MAX_PLANES]; };
template <int MODE>
__global__ void __launch_bounds__(256) k_mpas_remap_planes(const float* __restrict__ src, int nlev, const int* __restrict__ nodes,
                                                           const float* __restrict__ w, int tw, size_t npts, const float* __restrict__ landmask,
                                                           float fill, const __grid_constant__ MpasPlanes out, int k0, int k1)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const int n0 = nodes[p * tw], n1 = nodes[p * tw + 1], n2 = nodes[p * tw + 2];
    const float w0 = w[p * tw], w1 = w[p * tw + 1], w2 = w[p * tw + 2];
    const float* s0 = src + (size_t)(n0 - 1) * nlev;
    const float* s1 = src + (size_t)(n1 - 1) * nlev;
    const float* s2 = src + (size_t)(n2 - 1) * nlev;
    const bool water = landmask != nullptr && landmask[p] == 0.0f;
    auto level = [&](int k) -> float {
        float acc = 0.0f;
        acc = acc + w0 * __ldg(s0 + k);
        acc = acc + w1 * __ldg(s1 + k);
        acc = acc + w2 * __ldg(s2 + k);
        return acc;
    };
    float prev = 0.0f;
    if (MODE == 1 && k0 > 0) prev = level(k0 - 1);
    for (int k = k0; k < k1; ++k) {
        const float cur = level(k);
        float v = cur;
        if (MODE == 1 && k > 0) v = 0.5f * (prev + cur);
        prev = cur;
        out.p[k - k0][p] = water ? fill : v;
    }
}

// derive_mpas_fields (process_domain_module.F:1957-2053): TT = theta * (pressure / P0) ** (RD / CP) on the stored planes
struct MpasDerive { float* theta[MPAS_MAX_PLANES]; const float* pressure[MPAS_MAX_PLANES]; };
__global__ void __launch_bounds__(256) k_mpas_derive_tt(const __grid_constant__ MpasDerive a, size_t npts)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float P0 = 1.0e5f, RD = 287.0f, CP = 1004.0f;       // constants_module.F:28-30
    const float exner = t_pow(a.pressure[blockIdx.y][p] / P0, RD / CP);
    a.theta[blockIdx.y][p] = a.theta[blockIdx.y][p] * exner;
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
template <typename T>
static int upload(T** d, const T* hsrc, size_t n, cudaStream_t st)
{
    WPS_CUDA(cudaMalloc((void**)d, sizeof(T) * (n ? n : 1)));
    WPS_CUDA(cudaMemcpyAsync(*d, hsrc, sizeof(T) * n, cudaMemcpyHostToDevice, st));
    return 0;
}
static int check_grid(wpsgpu_ctx* h, int grid_id, const char* who)
{
    WPS_CHECK(h != nullptr, -1, "NULL handle");
    WPS_CHECK(h->mpas && h->mpas->mesh_set, -2, "%s: no MPAS mesh (wpsgpu_mpas_set_mesh)", who);
    WPS_CHECK(grid_id >= 0 && grid_id < 3, -1, "%s: grid_id %d outside 0..2", who, grid_id);
    return 0;
}

}  // namespace wps

using namespace wps;

extern "C" void wpsgpu_mpas_free(wpsgpu_ctx* h)
{
    if (!h->mpas) return;
    for (auto& t : h->mpas->grid) free_tables(t);
    free_mesh(h->mpas->m);
    delete h->mpas;
    h->mpas = nullptr;
}

extern "C" {

int wpsgpu_mpas_set_mesh(wpsgpu_handle_t h, const wpsgpu_mpas_mesh* mesh)
{
    WPS_CHECK(h && mesh, -1, "wpsgpu_mpas_set_mesh: NULL argument");
    WPS_CHECK(mesh->nCells > 0 && mesh->nVertices > 0 && mesh->nEdges > 0 && mesh->maxEdges >= 3, -1, "wpsgpu_mpas_set_mesh: bad dimensions");
    WPS_CHECK(mesh->landmask && mesh->nEdgesOnCell && mesh->cellsOnCell && mesh->verticesOnCell && mesh->edgesOnCell && mesh->cellsOnVertex &&
              mesh->cellsOnEdge && mesh->latCell && mesh->lonCell && mesh->latVertex && mesh->lonVertex && mesh->latEdge && mesh->lonEdge, -1,
              "wpsgpu_mpas_set_mesh: NULL mesh array");
    WPS_CUDA(cudaSetDevice(h->device));
    // connectivity is followed blindly on the device: validate it here (the reference would read out of bounds)
    const size_t nc = mesh->nCells, nv = mesh->nVertices, ne = mesh->nEdges, me = mesh->maxEdges;
    for (size_t c = 0; c < nc; ++c) {
        const int n = mesh->nEdgesOnCell[c];
        WPS_CHECK(n >= 3 && n <= (int)me, -1, "wpsgpu_mpas_set_mesh: nEdgesOnCell(%zu) = %d outside 3..maxEdges", c + 1, n);
        for (int i = 0; i < n; ++i) {
            const int a = mesh->cellsOnCell[c * me + i], b = mesh->verticesOnCell[c * me + i], d = mesh->edgesOnCell[c * me + i];
            WPS_CHECK(a >= 1 && b >= 1 && b <= (int)nv && d >= 1 && d <= (int)ne, -1, "wpsgpu_mpas_set_mesh: connectivity of cell %zu out of range", c + 1);
            // cellsOnCell may exceed nCells on limited-area meshes (nearest_cell skips those, :808); search_for_cells would follow them
            WPS_CHECK(a <= (int)nc, -1, "wpsgpu_mpas_set_mesh: limited-area meshes (cellsOnCell > nCells) are not supported");
        }
    }
    for (size_t v = 0; v < nv * 3; ++v) WPS_CHECK(mesh->cellsOnVertex[v] >= 1 && mesh->cellsOnVertex[v] <= (int)nc, -1, "wpsgpu_mpas_set_mesh: cellsOnVertex out of range");
    for (size_t e = 0; e < ne * 2; ++e) WPS_CHECK(mesh->cellsOnEdge[e] >= 1 && mesh->cellsOnEdge[e] <= (int)nc, -1, "wpsgpu_mpas_set_mesh: cellsOnEdge out of range");
    if (!h->mpas) h->mpas = new MpasState();
    MpasState* s = h->mpas;
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    for (auto& t : s->grid) free_tables(t);
    free_mesh(s->m);
    s->mesh_set = false;
    MpasMeshDev& m = s->m;
    m.nCells = mesh->nCells; m.nVertices = mesh->nVertices; m.nEdges = mesh->nEdges; m.maxEdges = mesh->maxEdges;
    WPS_TRY(upload(&m.landmask, mesh->landmask, nc, h->stream));
    WPS_TRY(upload(&m.nEdgesOnCell, mesh->nEdgesOnCell, nc, h->stream));
    WPS_TRY(upload(&m.cellsOnCell, mesh->cellsOnCell, nc * me, h->stream));
    WPS_TRY(upload(&m.verticesOnCell, mesh->verticesOnCell, nc * me, h->stream));
    WPS_TRY(upload(&m.edgesOnCell, mesh->edgesOnCell, nc * me, h->stream));
    WPS_TRY(upload(&m.cellsOnVertex, mesh->cellsOnVertex, nv * 3, h->stream));
    WPS_TRY(upload(&m.cellsOnEdge, mesh->cellsOnEdge, ne * 2, h->stream));
    WPS_TRY(upload(&m.latCell, mesh->latCell, nc, h->stream));
    WPS_TRY(upload(&m.lonCell, mesh->lonCell, nc, h->stream));
    WPS_TRY(upload(&m.latVertex, mesh->latVertex, nv, h->stream));
    WPS_TRY(upload(&m.lonVertex, mesh->lonVertex, nv, h->stream));
    WPS_TRY(upload(&m.latEdge, mesh->latEdge, ne, h->stream));
    WPS_TRY(upload(&m.lonEdge, mesh->lonEdge, ne, h->stream));
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    s->mesh_set = true;
    return 0;
}

int wpsgpu_mpas_remap_setup(wpsgpu_handle_t h, int grid_id, const float* lat_rad, const float* lon_rad, int nlon, int nlat)
{
    WPS_TRY(check_grid(h, grid_id, "wpsgpu_mpas_remap_setup"));
    WPS_CHECK(lat_rad && lon_rad && nlon > 0 && nlat > 0, -1, "wpsgpu_mpas_remap_setup: bad arguments");
    MpasState* s = h->mpas;
    const MpasMeshDev& m = s->m;
    const size_t npts = (size_t)nlon * nlat, me = m.maxEdges;
    WPS_TRY(begin_call(h, 2 * sizeof(float) * npts + 8192));
    void *dlat, *dlon;
    WPS_TRY(stage_in(h, lat_rad, sizeof(float) * npts, &dlat));
    WPS_TRY(stage_in(h, lon_rad, sizeof(float) * npts, &dlon));
    MpasTables& t = s->grid[grid_id];
    free_tables(t);
    t.nlon = nlon; t.nlat = nlat;
    t.width[0] = 3; t.width[1] = (int)me; t.width[2] = (int)me; t.width[3] = 3;
    for (int k = 0; k < 3; ++k) WPS_CUDA(cudaMalloc((void**)&t.nearest[k], sizeof(int) * npts));
    for (int k = 0; k < 4; ++k) {
        WPS_CUDA(cudaMalloc((void**)&t.src[k], sizeof(int) * npts * t.width[k]));
        WPS_CUDA(cudaMalloc((void**)&t.w[k], sizeof(float) * npts * t.width[k]));
    }
    const int ncx = (nlon + MPAS_SEED_STRIDE - 1) / MPAS_SEED_STRIDE, ncy = (nlat + MPAS_SEED_STRIDE - 1) / MPAS_SEED_STRIDE;
    int *d_seeds = nullptr, *d_list = nullptr;
    unsigned long long* d_n = nullptr;
    WPS_CUDA(cudaMalloc((void**)&d_seeds, sizeof(int) * ncx * ncy));
    WPS_CUDA(cudaMalloc((void**)&d_list, sizeof(int) * npts));
    WPS_CUDA(cudaMalloc((void**)&d_n, sizeof(unsigned long long)));
    WPS_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), h->stream));
    k_mpas_seed_cells<<<(ncx * ncy + 63) / 64, 64, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, nlon, nlat, ncx, ncy, d_seeds);
    k_mpas_first_guess<<<dim3((nlon + 31) / 32, (nlat + 3) / 4), dim3(32, 4), 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, nlon, nlat, ncx, d_seeds,
                                                                                             t.nearest[0], t.nearest[1], t.nearest[2]);
    h->launches += 2;
    // passes of the sequential recurrence until nothing changes (d_list doubles as the second buffer, d_tmp2 / d_vm for the masked scan)
    int *d_tmp2 = nullptr, *d_vm = nullptr, *d_changed = nullptr;
    WPS_CUDA(cudaMalloc((void**)&d_tmp2, sizeof(int) * npts));
    WPS_CUDA(cudaMalloc((void**)&d_vm, sizeof(int) * npts));
    WPS_CUDA(cudaMalloc((void**)&d_changed, sizeof(int)));
    const unsigned nbp = (unsigned)((npts + 127) / 128);
    int rc = 0;
    t.passes = 0;
    auto converge = [&](int kind, int* cur, int* other) -> int {     // result ends in `cur`
        for (int it = 0; it < 4096; ++it) {
            WPS_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), h->stream));
            if (kind == 0) k_mpas_chain<0><<<nbp, 128, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, npts, cur, other, d_changed);
            else if (kind == 1) k_mpas_chain<1><<<nbp, 128, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, npts, cur, other, d_changed);
            else if (kind == 2) k_mpas_chain<2><<<nbp, 128, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, npts, cur, other, d_changed);
            else k_mpas_chain_masked<<<nbp, 128, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, npts, cur, other, d_vm, d_changed);
            h->launches++; t.passes++;
            int changed = 0;
            WPS_CUDA(cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            WPS_CUDA(cudaStreamSynchronize(h->stream));
            if (!changed) return 0;                                   // other == cur
            WPS_CUDA(cudaMemcpyAsync(cur, other, sizeof(int) * npts, cudaMemcpyDeviceToDevice, h->stream));
        }
        set_error("wpsgpu_mpas_remap_setup: the sequential scan did not settle in 4096 passes");
        return -4;
    };
    for (int kind = 0; kind < 3 && rc == 0; ++kind) rc = converge(kind, t.nearest[kind], d_list);
    if (rc == 0) {
        // masked scan: first guess = no point of the scan goes to search_for_cells (idx = the vertex scan)
        cudaMemcpyAsync(d_tmp2, t.nearest[1], sizeof(int) * npts, cudaMemcpyDeviceToDevice, h->stream);
        rc = converge(3, d_tmp2, d_list);
    }
    if (rc == 0) {
        k_mpas_weights<<<nbp, 128, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, npts, t.nearest[0], t.nearest[1], d_vm,
                                                   t.src[0], t.w[0], t.src[1], t.w[1], t.src[2], t.w[2], t.src[3], t.w[3], d_list, d_n);
        h->launches++;
    }
    cudaFree(d_tmp2); cudaFree(d_vm); cudaFree(d_changed);
    if (rc != 0) { cudaFree(d_seeds); cudaFree(d_list); cudaFree(d_n); return rc; }
    unsigned long long nsearch = 0;
    WPS_CUDA(cudaMemcpyAsync(&nsearch, d_n, sizeof(nsearch), cudaMemcpyDeviceToHost, h->stream));
    WPS_CUDA(cudaStreamSynchronize(h->stream));
    t.nsearch = (int64_t)nsearch;
    if (nsearch > 0) {
        const int ndw = std::min((m.nCells + 31) / 32, MPAS_MAXDICT);
        const size_t per = (size_t)(MPAS_MAXQ + MPAS_LOG + ndw) * sizeof(int);
        size_t workers = std::min<size_t>(nsearch, 16384);
        while (workers > 64 && workers * per > ((size_t)1 << 30)) workers /= 2;     // scratch of at most 1 GiB
        workers = (workers + 63) / 64 * 64;
        int* d_scratch = nullptr;
        if (cudaMalloc((void**)&d_scratch, workers * per) != cudaSuccess) { set_error("wpsgpu_mpas_remap_setup: cannot allocate the search scratch"); rc = -3; }
        else {
            cudaMemsetAsync(d_scratch, 0, workers * per, h->stream);
            k_mpas_search<<<(unsigned)(workers / 64), 64, 0, h->stream>>>(m, (const float*)dlat, (const float*)dlon, d_list, nsearch, ndw, d_scratch, t.src[3], t.w[3]);
            h->launches++;
            if (cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("wpsgpu_mpas_remap_setup: search kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -100; }
            cudaFree(d_scratch);
        }
    }
    cudaFree(d_seeds); cudaFree(d_list); cudaFree(d_n);
    if (rc != 0) return rc;
    t.set = true;
    return end_call(h);
}

int wpsgpu_mpas_get_tables(wpsgpu_handle_t h, int grid_id, int kind, int32_t* nearest, int32_t* sources, float* weights)
{
    WPS_TRY(check_grid(h, grid_id, "wpsgpu_mpas_get_tables"));
    WPS_CHECK(kind >= 0 && kind <= 3, -1, "wpsgpu_mpas_get_tables: bad kind %d", kind);
    const MpasTables& t = h->mpas->grid[grid_id];
    WPS_CHECK(t.set, -2, "wpsgpu_mpas_get_tables: wpsgpu_mpas_remap_setup has not run for grid %d", grid_id);
    WPS_CHECK(!(nearest && kind == WPSGPU_MPAS_MASKED_CELLS), -1, "wpsgpu_mpas_get_tables: the masked cell table has no nearest array");
    WPS_CUDA(cudaSetDevice(h->device));
    const size_t npts = (size_t)t.nlon * t.nlat;
    const cudaMemcpyKind dir = h->ptr_mode == WPSGPU_PTR_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (nearest) WPS_CUDA(cudaMemcpyAsync(nearest, t.nearest[kind], sizeof(int) * npts, dir, h->stream));
    if (sources) WPS_CUDA(cudaMemcpyAsync(sources, t.src[kind], sizeof(int) * npts * t.width[kind], dir, h->stream));
    if (weights) WPS_CUDA(cudaMemcpyAsync(weights, t.w[kind], sizeof(float) * npts * t.width[kind], dir, h->stream));
    return end_call(h);
}

int wpsgpu_mpas_remap_field(wpsgpu_handle_t h, int grid_id, int decomp, int xtype, int masked, const void* src, int nlev, int nw, void* dst)
{
    WPS_TRY(check_grid(h, grid_id, "wpsgpu_mpas_remap_field"));
    WPS_CHECK(src && dst && nlev > 0, -1, "wpsgpu_mpas_remap_field: bad arguments");
    WPS_CHECK(decomp >= WPSGPU_MPAS_CELLS && decomp <= WPSGPU_MPAS_EDGES, -1, "Remap exception: unhandled decomposed dim");
    WPS_CHECK(xtype >= WPSGPU_MPAS_REAL && xtype <= WPSGPU_MPAS_INTEGER, -1, "Remap exception: unhandled type");
    const MpasTables& t = h->mpas->grid[grid_id];
    WPS_CHECK(t.set, -2, "wpsgpu_mpas_remap_field: wpsgpu_mpas_remap_setup has not run for grid %d", grid_id);
    const MpasMeshDev& m = h->mpas->m;
    const int kind = (decomp == WPSGPU_MPAS_CELLS && masked) ? WPSGPU_MPAS_MASKED_CELLS : decomp;     // :445-461
    const int tw = t.width[kind];
    if (nw == 0) nw = nlev == 1 ? tw : 3;
    WPS_CHECK(nw >= 1 && nw <= tw, -1, "wpsgpu_mpas_remap_field: nw %d outside 1..%d", nw, tw);
    const size_t nnodes = decomp == WPSGPU_MPAS_CELLS ? m.nCells : (decomp == WPSGPU_MPAS_VERTICES ? m.nVertices : m.nEdges);
    const size_t npts = (size_t)t.nlon * t.nlat, ntot = npts * nlev;
    const size_t esz = xtype == WPSGPU_MPAS_DOUBLE ? 8 : 4;
    WPS_TRY(begin_call(h, esz * (nnodes * nlev + ntot) + 8192));
    void *dsrc, *ddst;
    WPS_TRY(stage_in(h, src, esz * nnodes * nlev, &dsrc));
    WPS_TRY(alloc_out(h, dst, esz * ntot, &ddst));
    const unsigned nb = (unsigned)((ntot + 255) / 256);
    if (xtype == WPSGPU_MPAS_REAL) k_mpas_remap_levels<float><<<nb, 256, 0, h->stream>>>((const float*)dsrc, nlev, t.src[kind], t.w[kind], tw, nw, ntot, (float*)ddst);
    else if (xtype == WPSGPU_MPAS_DOUBLE) k_mpas_remap_levels<double><<<nb, 256, 0, h->stream>>>((const double*)dsrc, nlev, t.src[kind], t.w[kind], tw, nw, ntot, (double*)ddst);
    else k_mpas_remap_nearest<<<nb, 256, 0, h->stream>>>((const int*)dsrc, nlev, t.nearest[decomp], ntot, (int*)ddst);
    h->launches++;
    WPS_TRY(copy_out(h, dst, ddst, esz * ntot));
    return end_call(h);
}

int wpsgpu_mpas_remap_planes(wpsgpu_handle_t h, int grid_id, int decomp, int masked, const float* src, int nlev, int mode,
                             const float* landmask, float fill_missing, float* const* planes)
{
    WPS_TRY(check_grid(h, grid_id, "wpsgpu_mpas_remap_planes"));
    WPS_CHECK(src && planes && nlev > 0, -1, "wpsgpu_mpas_remap_planes: bad arguments");
    WPS_CHECK(decomp >= WPSGPU_MPAS_CELLS && decomp <= WPSGPU_MPAS_EDGES, -1, "Remap exception: unhandled decomposed dim");
    WPS_CHECK(mode == 0 || mode == 1, -1, "wpsgpu_mpas_remap_planes: mode %d", mode);
    const MpasTables& t = h->mpas->grid[grid_id];
    WPS_CHECK(t.set, -2, "wpsgpu_mpas_remap_planes: wpsgpu_mpas_remap_setup has not run for grid %d", grid_id);
    const MpasMeshDev& m = h->mpas->m;
    const int kind = (decomp == WPSGPU_MPAS_CELLS && masked) ? WPSGPU_MPAS_MASKED_CELLS : decomp;
    const size_t nnodes = decomp == WPSGPU_MPAS_CELLS ? m.nCells : (decomp == WPSGPU_MPAS_VERTICES ? m.nVertices : m.nEdges);
    const size_t npts = (size_t)t.nlon * t.nlat;
    for (int k = 0; k < nlev; ++k) WPS_CHECK(planes[k] != nullptr, -1, "wpsgpu_mpas_remap_planes: NULL plane %d", k);
    WPS_TRY(begin_call(h, sizeof(float) * (nnodes * nlev + npts * (nlev + 1)) + 512 * (size_t)nlev + 8192));
    void *dsrc, *dlm;
    WPS_TRY(stage_in(h, src, sizeof(float) * nnodes * nlev, &dsrc));
    WPS_TRY(stage_in(h, landmask, sizeof(float) * npts, &dlm));
    std::vector<float*> dpl(nlev);
    for (int k = 0; k < nlev; ++k) { void* d; WPS_TRY(alloc_out(h, planes[k], sizeof(float) * npts, &d)); dpl[k] = (float*)d; }
    const unsigned nb = (unsigned)((npts + 255) / 256);
    for (int k0 = 0; k0 < nlev; k0 += MPAS_MAX_PLANES) {
        const int k1 = std::min(nlev, k0 + MPAS_MAX_PLANES);
        MpasPlanes out;
        memset(&out, 0, sizeof(out));
        for (int k = k0; k < k1; ++k) out.p[k - k0] = dpl[k];
        if (mode == 0) k_mpas_remap_planes<0><<<nb, 256, 0, h->stream>>>((const float*)dsrc, nlev, t.src[kind], t.w[kind], t.width[kind], npts, (const float*)dlm, fill_missing, out, k0, k1);
        else k_mpas_remap_planes<1><<<nb, 256, 0, h->stream>>>((const float*)dsrc, nlev, t.src[kind], t.w[kind], t.width[kind], npts, (const float*)dlm, fill_missing, out, k0, k1);
        h->launches++;
    }
    for (int k = 0; k < nlev; ++k) WPS_TRY(copy_out(h, planes[k], dpl[k], sizeof(float) * npts));
    return end_call(h);
}

int wpsgpu_mpas_derive_tt(wpsgpu_handle_t h, float* const* theta, const float* const* pressure, int nlev, int64_t npts)
{
    WPS_CHECK(h && theta && pressure && nlev >= 0 && npts > 0, -1, "wpsgpu_mpas_derive_tt: bad arguments");
    if (nlev == 0) return 0;
    const size_t bytes = sizeof(float) * (size_t)npts;
    WPS_TRY(begin_call(h, 2 * (bytes + 256) * (size_t)nlev + 8192));
    std::vector<float*> dt(nlev);
    std::vector<const float*> dp(nlev);
    for (int k = 0; k < nlev; ++k) {
        WPS_CHECK(theta[k] && pressure[k], -1, "wpsgpu_mpas_derive_tt: NULL level array");
        void *a, *b;
        WPS_TRY(stage_in(h, theta[k], bytes, &a));
        WPS_TRY(stage_in(h, pressure[k], bytes, &b));
        dt[k] = (float*)a; dp[k] = (const float*)b;
    }
    for (int k0 = 0; k0 < nlev; k0 += MPAS_MAX_PLANES) {
        const int n = std::min(nlev - k0, MPAS_MAX_PLANES);
        MpasDerive a;
        memset(&a, 0, sizeof(a));
        for (int k = 0; k < n; ++k) { a.theta[k] = dt[k0 + k]; a.pressure[k] = dp[k0 + k]; }
        k_mpas_derive_tt<<<dim3((unsigned)((npts + 255) / 256), n), 256, 0, h->stream>>>(a, (size_t)npts);
        h->launches++;
    }
    for (int k = 0; k < nlev; ++k) WPS_TRY(copy_out(h, theta[k], dt[k], bytes));
    return end_call(h);
}

}  // extern "C"
