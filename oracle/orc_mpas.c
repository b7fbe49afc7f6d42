/* orc_mpas.c -- CPU ORACLE (test infrastructure; never linked into or called by the product path).
 * Restates the MPAS remapping path of metgrid: metgrid/src/remapper.F (remap_info_setup :83-286, remap_field :417-663,
 * nearest_cell :781-820, nearest_vertex :823-897, sphere_distance :900-916, mpas_wachspress_coordinates :934-1020,
 * mpas_triangle_signed_area_sphere :1035-1078, mpas_arc_length :1088-1106, convert_lx :1109-1121, search_for_cells
 * :1124-1210, queue / dictionary :1216-1326) and the per-level extraction of process_mpas_fields
 * (process_domain_module.F:1700-1962).  Default REAL arithmetic, operation order as written; connectivity is 1-based as
 * in the MPAS file.  PARITY UNPINNED: the reference holds no vectors for this path and cannot be compiled here.
 *
 * One thing the reference leaves undefined is stated in orc_mpas_remap_setup: vertexWeights / edgeWeights(1:nn) are
 * assigned from a THREE-element function result (:213, :238 call mpas_wachspress_coordinates with nVertices = 3), a
 * shape mismatch for nn /= 3.  Elements 1..3 are the Wachspress coordinates in the triangle of the cell's first three
 * vertices / edges; elements 4..nn are whatever the compiler reads past the temporary.  Here they are 0. */
#include "orc_internal.h"
#include <stdlib.h>
#include <string.h>

#define MAX_QUEUE_LENGTH 2700      /* remapper.F:7 */
#define MAX_DICTIONARY_SIZE 82000  /* remapper.F:8 */

/* remapper.F:900-916 */
static float sphere_distance(float lat1, float lon1, float lat2, float lon2, float radius)
{
    float s1 = o_sin(0.5f * (lat2 - lat1));
    float s2 = o_sin(0.5f * (lon2 - lon1));
    float arg1 = o_sqrt(s1 * s1 + o_cos(lat1) * o_cos(lat2) * (s2 * s2));
    return 2.0f * radius * o_asin(arg1);
}

/* remapper.F:781-820 */
int orc_mpas_nearest_cell(const orc_mpas_mesh *m, float tlat, float tlon, int start_cell)
{
    int nearest = start_cell, current = -1;
    while (nearest != current) {
        current = nearest;
        float nd = sphere_distance(m->latCell[current - 1], m->lonCell[current - 1], tlat, tlon, 1.0f);
        for (int i = 1; i <= m->nEdgesOnCell[current - 1]; ++i) {
            int ic = m->cellsOnCell[(size_t)(current - 1) * m->maxEdges + (i - 1)];
            if (ic <= m->nCells) {
                float d = sphere_distance(m->latCell[ic - 1], m->lonCell[ic - 1], tlat, tlon, 1.0f);
                if (d < nd) { nearest = ic; nd = d; }
            }
        }
    }
    return nearest;
}

/* remapper.F:823-897; degree 3: vertices (cellsOnVertex, verticesOnCell), degree 2: edges (cellsOnEdge, edgesOnCell) */
int orc_mpas_nearest_vertex(const orc_mpas_mesh *m, float tlat, float tlon, int start, int degree)
{
    const int *nodesOnCell = degree == 3 ? m->verticesOnCell : m->edgesOnCell;
    const int *cellsOnNode = degree == 3 ? m->cellsOnVertex : m->cellsOnEdge;
    const float *latN = degree == 3 ? m->latVertex : m->latEdge, *lonN = degree == 3 ? m->lonVertex : m->lonEdge;
    int nearest = start, current = -1;
    while (nearest != current) {
        current = nearest;
        float nd = sphere_distance(latN[current - 1], lonN[current - 1], tlat, tlon, 1.0f);
        int c1 = cellsOnNode[(size_t)(current - 1) * degree + 0], c2 = cellsOnNode[(size_t)(current - 1) * degree + 1], c3 = 0;
        float d1 = sphere_distance(m->latCell[c1 - 1], m->lonCell[c1 - 1], tlat, tlon, 1.0f);
        float d2 = sphere_distance(m->latCell[c2 - 1], m->lonCell[c2 - 1], tlat, tlon, 1.0f);
        float d3 = 0.0f;
        int ic;
        if (degree == 3) {
            c3 = cellsOnNode[(size_t)(current - 1) * degree + 2];
            d3 = sphere_distance(m->latCell[c3 - 1], m->lonCell[c3 - 1], tlat, tlon, 1.0f);
            if (d1 < d2) ic = d1 < d3 ? c1 : c3;
            else ic = d2 < d3 ? c2 : c3;
        } else ic = d1 < d2 ? c1 : c2;
        for (int i = 1; i <= m->nEdgesOnCell[ic - 1]; ++i) {
            int iv = nodesOnCell[(size_t)(ic - 1) * m->maxEdges + (i - 1)];
            float d = sphere_distance(latN[iv - 1], lonN[iv - 1], tlat, tlon, 1.0f);
            if (d < nd) { nearest = iv; nd = d; }
        }
    }
    return nearest;
}

/* remapper.F:1109-1121 */
static void convert_lx(float lat, float lon, float radius, float v[3])
{
    v[0] = radius * o_cos(lon) * o_cos(lat);
    v[1] = radius * o_sin(lon) * o_cos(lat);
    v[2] = radius * o_sin(lat);
}

/* remapper.F:1088-1106 */
static float arc_length(float ax, float ay, float az, float bx, float by, float bz)
{
    float cx = bx - ax, cy = by - ay, cz = bz - az;
    float r = o_sqrt(ax * ax + ay * ay + az * az);
    float c = o_sqrt(cx * cx + cy * cy + cz * cz);
    return r * 2.0f * o_asin(c / (2.0f * r));
}

/* remapper.F:1035-1078 */
static float triangle_signed_area_sphere(const float a[3], const float b[3], const float c[3], float radius)
{
    float ab = arc_length(a[0], a[1], a[2], b[0], b[1], b[2]) / radius;
    float bc = arc_length(b[0], b[1], b[2], c[0], c[1], c[2]) / radius;
    float ca = arc_length(c[0], c[1], c[2], a[0], a[1], a[2]) / radius;
    float semiperim = 0.5f * (ab + bc + ca);
    float prod = o_tan(0.5f * semiperim) * o_tan(0.5f * (semiperim - ab)) * o_tan(0.5f * (semiperim - bc)) * o_tan(0.5f * (semiperim - ca));
    float tanqe = o_sqrt(prod > 0.0f ? prod : 0.0f);     /* max(0.0, prod): a NaN product gives 0 with gfortran's MAX as well */
    float area = 4.0f * radius * radius * o_atan(tanqe);
    float abl[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] }, acl[3] = { c[0] - a[0], c[1] - a[1], c[2] - a[2] };
    float d0 = (abl[1] * acl[2]) - (abl[2] * acl[1]);
    float d1 = -((abl[0] * acl[2]) - (abl[2] * acl[0]));
    float d2 = (abl[0] * acl[1]) - (abl[1] * acl[0]);
    if ((d0 * a[0] + d1 * a[1] + d2 * a[2]) < 0.0f) area = -area;
    return area;
}

/* remapper.F:934-1020 for nVertices = 3 -- the only value any call site passes (:167, :213, :238, :264) */
void orc_mpas_wachspress3(const float vert[3][3], const float p[3], float w[3])
{
    float radiusLocal = o_sqrt(vert[0][0] * vert[0][0] + vert[0][1] * vert[0][1] + vert[0][2] * vert[0][2]);
    float areaA[3], areaB[3], wach[3];
    for (int i = 1; i <= 3; ++i) {
        int im1 = (3 + i - 2) % 3, i0 = (3 + i - 1) % 3, ip1 = (3 + i) % 3;     /* 0-based */
        areaB[i - 1] = triangle_signed_area_sphere(vert[im1], vert[i0], vert[ip1], radiusLocal);
    }
    for (int i = 1; i <= 3; ++i) {
        int i0 = (3 + i - 1) % 3, ip1 = (3 + i) % 3;
        areaA[i0] = triangle_signed_area_sphere(p, vert[i0], vert[ip1], radiusLocal);
    }
    for (int i = 1; i <= 3; ++i) {
        wach[i - 1] = areaB[i - 1];
        for (int j = i + 1; j <= i + 3 - 2; ++j) wach[i - 1] = wach[i - 1] * areaA[(3 + j - 1) % 3];
    }
    float total = 0.0f;
    for (int i = 0; i < 3; ++i) total = total + wach[i];
    for (int i = 0; i < 3; ++i) w[i] = wach[i] / total;
}

/* remapper.F:1124-1210 with the module's queue (:1216-1263) and dictionary (:1269-1326) */
typedef struct { int head, tail, size; int q[MAX_QUEUE_LENGTH]; unsigned *dict; } search_state;
static void q_insert(search_state *s, int i)
{
    if (s->size == MAX_QUEUE_LENGTH) return;              /* 'Error: queue overrun': the cell is dropped */
    s->size++; s->q[s->head] = i; s->head = (s->head + 1) % MAX_QUEUE_LENGTH;
}
static int q_remove(search_state *s)
{
    s->size--;
    int v = s->q[s->tail];
    s->tail = (s->tail + 1) % MAX_QUEUE_LENGTH;
    return v;
}
static void d_insert(search_state *s, int i)
{
    int n_integer = ((i - 1) / 32) + 1, n_bit = (i - 1) % 32;
    if (n_integer > MAX_DICTIONARY_SIZE) return;
    s->dict[n_integer - 1] |= 1u << n_bit;
}
static int d_search(const search_state *s, int i)
{
    int n_integer = ((i - 1) / 32) + 1, n_bit = (i - 1) % 32;
    if (n_integer > MAX_DICTIONARY_SIZE) return 0;
    return (s->dict[n_integer - 1] >> n_bit) & 1u;
}
void orc_mpas_search_for_cells(const orc_mpas_mesh *m, float tlat, float tlon, int start, int cells[3], float w[3])
{
    search_state *s = (search_state *)malloc(sizeof(search_state));
    s->head = s->tail = s->size = 0;
    s->dict = (unsigned *)calloc(MAX_DICTIONARY_SIZE, sizeof(unsigned));
    int no_more_queueing = 0, unscanned = 0;
    float d_min = 1000.0f;
    cells[0] = cells[1] = cells[2] = 1;
    w[0] = w[1] = w[2] = 0.0f;
    q_insert(s, start);
    d_insert(s, start);
    while (s->size > 0) {
        int scan = q_remove(s);
        if (m->landmask[scan - 1] == 1) {
            float d = sphere_distance(tlat, tlon, m->latCell[scan - 1], m->lonCell[scan - 1], 1.0f);
            if (d < d_min) {
                d_min = d;
                cells[0] = scan;
                w[0] = 1.0f;
                if (!no_more_queueing) unscanned = s->size;
                no_more_queueing = 1;
            }
        }
        if (!no_more_queueing || unscanned > 0) {
            for (int i = 1; i <= m->nEdgesOnCell[scan - 1]; ++i) {
                int nb = m->cellsOnCell[(size_t)(scan - 1) * m->maxEdges + (i - 1)];
                if (!d_search(s, nb)) { q_insert(s, nb); d_insert(s, nb); }
            }
        }
        if (no_more_queueing) unscanned = unscanned - 1;
    }
    free(s->dict);
    free(s);
}

/* remapper.F:83-286.  lat / lon: the target grid in radians, (ix, iy) with ix fastest (irank = 1, lat2d / lon2d of
 * target_mesh.F).  The sequential scan hands each point the previous point's result as the walk's start (idx). */
void orc_mpas_remap_setup(const orc_mpas_mesh *m, const float *lat, const float *lon, int nlon, int nlat, orc_mpas_remap *r)
{
    const long n = (long)nlon * nlat;
    const int me = m->maxEdges;
    int idx;
    idx = 1;
    for (long p = 0; p < n; ++p) r->nearestCell[p] = idx = orc_mpas_nearest_cell(m, lat[p], lon[p], idx);
    idx = 1;
    for (long p = 0; p < n; ++p) r->nearestVertex[p] = idx = orc_mpas_nearest_vertex(m, lat[p], lon[p], idx, 3);
    idx = 1;
    for (long p = 0; p < n; ++p) r->nearestEdge[p] = idx = orc_mpas_nearest_vertex(m, lat[p], lon[p], idx, 2);
    /* :158-175 cell weights */
    idx = 1;
    for (long p = 0; p < n; ++p) {
        idx = orc_mpas_nearest_vertex(m, lat[p], lon[p], idx, 3);
        float vc[3][3], pt[3];
        convert_lx(lat[p], lon[p], 6371229.0f, pt);
        for (int j = 0; j < 3; ++j) {
            int c = m->cellsOnVertex[(size_t)(idx - 1) * 3 + j];
            r->sourceCells[p * 3 + j] = c;
            convert_lx(m->latCell[c - 1], m->lonCell[c - 1], 6371229.0f, vc[j]);
        }
        orc_mpas_wachspress3((const float(*)[3])vc, pt, &r->cellWeights[p * 3]);
    }
    /* :185-207 vertex weights, :219-241 edge weights */
    for (int pass = 0; pass < 2; ++pass) {
        const int *nodesOnCell = pass == 0 ? m->verticesOnCell : m->edgesOnCell;
        const float *latN = pass == 0 ? m->latVertex : m->latEdge, *lonN = pass == 0 ? m->lonVertex : m->lonEdge;
        int *srcN = pass == 0 ? r->sourceVertices : r->sourceEdges;
        float *wN = pass == 0 ? r->vertexWeights : r->edgeWeights;
        idx = 1;
        for (long p = 0; p < n; ++p) {
            idx = orc_mpas_nearest_cell(m, lat[p], lon[p], idx);
            int nn = m->nEdgesOnCell[idx - 1];
            float vc[3][3], pt[3];
            convert_lx(lat[p], lon[p], 6371229.0f, pt);
            for (int j = 0; j < me; ++j) { srcN[p * me + j] = 1; wN[p * me + j] = 0.0f; }
            for (int j = 0; j < nn; ++j) {
                int v = nodesOnCell[(size_t)(idx - 1) * me + j];
                srcN[p * me + j] = v;
                if (j < 3) convert_lx(latN[v - 1], lonN[v - 1], 6371229.0f, vc[j]);
            }
            float w3[3];
            orc_mpas_wachspress3((const float(*)[3])vc, pt, w3);
            for (int j = 0; j < 3 && j < nn; ++j) wN[p * me + j] = w3[j];      /* elements 4..nn: see the header */
        }
    }
    /* :251-283 masked cell weights */
    idx = 1;
    for (long p = 0; p < n; ++p) {
        idx = orc_mpas_nearest_vertex(m, lat[p], lon[p], idx, 3);
        float vc[3][3], pt[3];
        convert_lx(lat[p], lon[p], 6371229.0f, pt);
        int *sc = &r->sourceMaskedCells[p * 3];
        float *w = &r->cellMaskedWeights[p * 3];
        for (int j = 0; j < 3; ++j) {
            sc[j] = m->cellsOnVertex[(size_t)(idx - 1) * 3 + j];
            convert_lx(m->latCell[sc[j] - 1], m->lonCell[sc[j] - 1], 6371229.0f, vc[j]);
        }
        orc_mpas_wachspress3((const float(*)[3])vc, pt, w);
        float sumWeights = 0.0f;
        for (int j = 0; j < 3; ++j) {
            if (m->landmask[sc[j] - 1] == 1) sumWeights = sumWeights + w[j];
            else w[j] = 0.0f;
        }
        if (sumWeights > 0.0f) {
            for (int j = 0; j < 3; ++j) w[j] = w[j] / sumWeights;
        } else {
            /* :272: idx now holds a CELL index and is the next point's start VERTEX (as written) */
            idx = orc_mpas_nearest_cell(m, lat[p], lon[p], sc[0]);
            orc_mpas_search_for_cells(m, lat[p], lon[p], idx, sc, w);
        }
    }
}

/* remap_field, REAL fields (:497-545): dst(:, ix, iy) = sum_j nodeWeights(j, ix, iy) * src(:, sourceNodes(j, ix, iy)), level
 * index fastest in both arrays; nw = 3 for 3-d / 4-d fields, size(sourceNodes, 1) for 2-d ones. */
void orc_mpas_remap_field_real(const int *sourceNodes, const float *nodeWeights, int tw, int nw, long npts, const float *src,
                               int nlev, float *dst)
{
    for (long p = 0; p < npts; ++p)
        for (int k = 0; k < nlev; ++k) {
            float acc = 0.0f;
            for (int j = 0; j < nw; ++j) acc = acc + nodeWeights[p * tw + j] * src[(size_t)(sourceNodes[p * tw + j] - 1) * nlev + k];
            dst[p * nlev + k] = acc;
        }
}
/* DOUBLE fields (:563-620): the REAL weight is promoted */
void orc_mpas_remap_field_double(const int *sourceNodes, const float *nodeWeights, int tw, int nw, long npts, const double *src,
                                 int nlev, double *dst)
{
    for (long p = 0; p < npts; ++p)
        for (int k = 0; k < nlev; ++k) {
            double acc = 0.0;
            for (int j = 0; j < nw; ++j) acc = acc + (double)nodeWeights[p * tw + j] * src[(size_t)(sourceNodes[p * tw + j] - 1) * nlev + k];
            dst[p * nlev + k] = acc;
        }
}
/* INTEGER fields (:621-655): nearest node */
void orc_mpas_remap_field_int(const int *nearest, long npts, const int *src, int nlev, int *dst)
{
    for (long p = 0; p < npts; ++p)
        for (int k = 0; k < nlev; ++k) dst[p * nlev + k] = src[(size_t)(nearest[p] - 1) * nlev + k];
}

/* process_mpas_fields' per-level extraction of a remapped REAL field wrf(k, ix, iy) (process_domain_module.F):
 *   mode 0 (nVertLevels :1700-1731, nSoilLevels :1811-1890, 2-d fields :1927-1956): plane k = wrf(k, :, :)
 *   mode 1 (nVertLevelsP1 :1733-1776): plane 0 = wrf(1, :, :) [level 200100], plane k = 0.5 * (wrf(k) + wrf(k+1)), k = 1..nlev-1
 * then, for masked=water fields, `where (landmask == 0) r_arr = fill_missing` (:1854-1858, :1948-1952). */
void orc_mpas_extract_planes(const float *wrf, int nlev, long npts, int mode, const float *landmask, float fill, float *planes)
{
    for (int k = 0; k < nlev; ++k)
        for (long p = 0; p < npts; ++p) {
            float v;
            if (mode == 1 && k > 0) v = 0.5f * (wrf[p * nlev + (k - 1)] + wrf[p * nlev + k]);
            else v = wrf[p * nlev + k];
            if (landmask && landmask[p] == 0.0f) v = fill;
            planes[(size_t)k * npts + p] = v;
        }
}

/* derive_mpas_fields (process_domain_module.F:1957-2053), per level: exner = (pressure / P0) ** (RD / CP); theta = theta * exner
 * (the stored field TT holds potential temperature until this step); constants_module.F:28-30 */
void orc_mpas_derive_tt(float *theta, const float *pressure, long n)
{
    const float P0 = 1.0e5f, RD = 287.0f, CP = 1004.0f;
    for (long k = 0; k < n; ++k) {
        float exner = o_pow(pressure[k] / P0, RD / CP);
        theta[k] = theta[k] * exner;
    }
}
