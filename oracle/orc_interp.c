/*
 * orc_interp.c -- CPU ORACLE (test infrastructure): interpolation operators.
 * Restates geogrid/src/interp_module.F (+ queue_module.F, bitarray_module.F as used by
 * search_extrap) of WPS v4.6.0.  PARITY UNPINNED (see wps_oracle.h).  Not product code.
 */
#include "orc_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* interp_array_from_string / interp_options_from_string / get_search_depth
 * interp_module.F:20-295 */

static int tok_is(const char *s, int p1, int p2m1, const char *name)
{
    /* index(tok,name) /= 0 .and. len_trim(tok) == len_trim(name); tok = s(p1:p2m1), 1-based inclusive */
    int len = p2m1 - p1 + 1;
    if (len < 0) len = 0;
    while (len > 0 && s[p1 - 1 + len - 1] == ' ') --len; /* len_trim */
    int nl = (int)strlen(name);
    if (len != nl) return 0;
    /* with equal trimmed lengths, index /= 0 means equality unless leading blanks shift it */
    return strncmp(s + p1 - 1, name, (size_t)nl) == 0;
}
static int tok_has(const char *s, int p1, int p2m1, const char *name)
{
    int len = p2m1 - p1 + 1, nl = (int)strlen(name);
    for (int k = 0; k + nl <= len; ++k)
        if (strncmp(s + p1 - 1 + k, name, (size_t)nl) == 0) return 1;
    return 0;
}

/* get_search_depth, interp_module.F:254-295 */
static int get_search_depth(const char *tok, int len, int *depth)
{
    *depth = 1200;
    int i = -1, nl = 6;
    for (int k = 0; k + nl <= len; ++k) if (strncmp(tok + k, "search", 6) == 0) { i = k + 1; break; }
    if (i > 0) {
        int p = 0;
        for (int k = i; k <= len; ++k) if (tok[k - 1] == '+') { p = k - i + 1; break; }
        if (p == 0) { p = len; while (p > 0 && tok[p - 1] == ' ') --p; }
        int p1 = 0, p2 = 0;
        for (int k = i; k <= p; ++k) if (tok[k - 1] == '(') { p1 = k - i + 1; break; }
        for (int k = i; k <= p; ++k) if (tok[k - 1] == ')') { p2 = k - i + 1; break; }
        if (p1 != 0 && p2 != 0) {
            /* read(interp_opt(p1+1:p2-1),*) depth -- p1,p2 used as absolute positions */
            char buf[64]; int n = p2 - 1 - (p1 + 1) + 1;
            if (n <= 0 || n >= 63) return -1;
            memcpy(buf, tok + p1, (size_t)n); buf[n] = 0;
            char *end; long v = strtol(buf, &end, 10);
            while (*end == ' ') ++end;
            if (*end != 0) return -1; /* ERROR: must be an integer */
            *depth = (int)v;
        } else if (p1 == 0 && p2 == 0) {
        } else {
            *depth = 1200; /* WARN */
        }
    }
    return 0;
}

static int parse_chain(const char *str, int *codes, int *opts, int maxn)
{
    int iend = (int)strlen(str);
    while (iend > 0 && str[iend - 1] == ' ') --iend; /* len_trim */
    int num_methods = 1;
    for (int j = 0; j < iend; ++j) if (str[j] == '+') ++num_methods;
    if (num_methods + 1 > maxn) return -1;
    for (int j = 0; j < num_methods + 1; ++j) { if (codes) codes[j] = 0; if (opts) opts[j] = 0; }
    int p1 = 1, p2 = 0, j = 1;
    for (int k = 1; k <= iend; ++k) if (str[k - 1] == '+') { p2 = k; break; }
    int trailing = 0;
    for (;;) {
        int e; /* token end (inclusive, 1-based) */
        if (!trailing) {
            if (!(p2 >= p1)) { trailing = 1; continue; }
            e = p2 - 1;
        } else {
            p2 = iend + 1;
            if (!(p1 < iend)) break;
            e = p2 - 1;
        }
        int code = 0, depth = 0, matched = 1;
        if (tok_is(str, p1, e, "nearest_neighbor")) code = ORC_N_NEIGHBOR;
        else if (tok_is(str, p1, e, "average_4pt")) code = ORC_AVERAGE4;
        else if (tok_is(str, p1, e, "average_16pt")) code = ORC_AVERAGE16;
        else if (tok_is(str, p1, e, "wt_average_4pt")) code = ORC_W_AVERAGE4;
        else if (tok_is(str, p1, e, "wt_average_16pt")) code = ORC_W_AVERAGE16;
        else if (tok_is(str, p1, e, "four_pt")) code = ORC_FOUR_POINT;
        else if (tok_is(str, p1, e, "sixteen_pt")) code = ORC_SIXTEEN_POINT;
        else if (tok_has(str, p1, e, "search")) {
            code = ORC_SEARCH;
            if (get_search_depth(str + p1 - 1, e - p1 + 1, &depth) != 0) return -2;
        } else matched = 0; /* average_gcell: silently skipped; anything else: WARN */
        if (matched) { if (codes) codes[j - 1] = code; if (opts) opts[j - 1] = depth; ++j; }
        if (trailing) break;
        p1 = p2 + 1;
        p2 = 0;
        for (int k = p1; k <= iend; ++k) if (str[k - 1] == '+') { p2 = k; break; }
        if (p2 == 0) p2 = p1 - 1; /* index()=0 => p2 = p1-1 */
    }
    return num_methods + 1;
}

int orc_interp_array_from_string(const char *s, int *codes, int maxn) { return parse_chain(s, codes, NULL, maxn); }
int orc_interp_options_from_string(const char *s, int *opts, int maxn) { return parse_chain(s, NULL, opts, maxn); }

/* get_gcell_threshold, metgrid/src/interp_option_module.F:692-727 (same in source_data_module.F:2662) */
int orc_get_gcell_threshold(const char *s, float *threshold)
{
    *threshold = 1.0f;
    const char *g = strstr(s, "average_gcell");
    if (g) {
        const char *q1 = strchr(g, '('), *q2 = strchr(g, ')');
        if (q1 && q2) {
            /* p1,p2 relative to i are used as absolute positions: read(interp_opt(p1+1:p2-1)) */
            long p1 = q1 - g + 1, p2 = q2 - g + 1;
            long n = p2 - 1 - (p1 + 1) + 1;
            char buf[64];
            if (n <= 0 || n > 63 || p1 + n > (long)strlen(s)) return -1;
            memcpy(buf, s + p1, (size_t)n); buf[n] = 0;
            char *end; float v = strtof(buf, &end);
            if (end == buf) return -1;
            *threshold = v;
        } else *threshold = 1.0f;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* operators */

typedef struct {
    const float *array; int sx, ex, sy, ey, sz, ez;
    float msgval; const int *list; const int *opts;
    int mask_mode; char rel; float maskval; const float *mask;
} ictx;

#define A3(c, i, j, k) ((c)->array[((long)((k) - (c)->sz) * ((c)->ey - (c)->sy + 1) + ((j) - (c)->sy)) * ((c)->ex - (c)->sx + 1) + ((i) - (c)->sx)])
#define MK(c, i, j) ((c)->mask[(long)((j) - (c)->sy) * ((c)->ex - (c)->sx + 1) + ((i) - (c)->sx)])

static float seq(const ictx *c, float xx, float yy, int izz, int idx);

/* "unusable" test shared by nearest_neighbor / four_pt / averages / sixteen_pt:
 * '<': mask<val; '>': mask>val; ' ': mask==val  (interp_module.F:417-421,695-709) */
static inline int masked_out(const ictx *c, int i, int j)
{
    if (!c->mask_mode) return 0;
    float m = MK(c, i, j);
    if (c->rel == '>') return m > c->maskval;
    if (c->rel == '<') return m < c->maskval;
    return m == c->maskval;
}
/* "valid" test of search_extrap / sixteen_pt near-node branch:
 * '>': mask<=val; '<': mask>=val; ' ': mask/=val (interp_module.F:505-517,1302-1311) */
static inline int mask_valid(const ictx *c, int i, int j)
{
    if (!c->mask_mode) return 1;
    float m = MK(c, i, j);
    if (c->rel == '>') return m <= c->maskval;
    if (c->rel == '<') return m >= c->maskval;
    if (c->rel == ' ') return m != c->maskval;
    return 0;
}

/* interp_module.F:376-442 */
static float nearest_neighbor(const ictx *c, float xx, float yy, int izz, int idx)
{
    int ix = o_nint(xx), iy = o_nint(yy);
    float r;
    if (ix < c->sx || ix > c->ex) return c->msgval;
    if (iy < c->sy || iy > c->ey) return c->msgval;
    if (c->mask_mode) {
        float m = MK(c, ix, iy);
        if (c->rel == '<' && m < c->maskval) r = c->msgval;
        else if (c->rel == '>' && m > c->maskval) r = c->msgval;
        else if (c->rel == ' ' && m == c->maskval) r = c->msgval;
        else r = A3(c, ix, iy, izz);
    } else r = A3(c, ix, iy, izz);
    if (r == c->msgval) r = seq(c, xx, yy, izz, idx);
    return r;
}

/* interp_module.F:451-615 with queue_module.F FIFO semantics and bitarray visited set.
 * idx here is the index of the NEXT method (as passed by interp_sequence), so the depth option
 * is opts[idx-1] in Fortran 1-based terms == c->opts[idx-2] 0-based. */
typedef struct { int x, y, depth; } qd;
static float search_extrap(const ictx *c, float xx, float yy, int izz, int idx)
{
    int nx = c->ex - c->sx + 1, ny = c->ey - c->sy + 1;
    if (o_nint(xx) < c->sx || o_nint(xx) > c->ex || o_nint(yy) < c->sy || o_nint(yy) > c->ey) return c->msgval;
    int maxdepth = c->opts[idx - 2];
    /* bitarray_create (interp_module.F:490, bitarray_module.F:25-50): one BIT per source cell */
    const size_t nwords = ((size_t)nx * (size_t)ny + 31) / 32;
    unsigned *b = (unsigned *)calloc(nwords, sizeof(unsigned));
    long cap = 1024, head = 0, tail = 0;
    qd *q = (qd *)malloc(sizeof(qd) * (size_t)cap);
#define QPUSH(v) do { if (tail == cap) { cap *= 2; q = (qd *)realloc(q, sizeof(qd) * (size_t)cap); } q[tail++] = (v); } while (0)
#define BPOS(i, j) ((size_t)((j) - c->sy) * (size_t)nx + (size_t)((i) - c->sx))
#define BT(i, j) ((b[BPOS(i, j) >> 5] >> (BPOS(i, j) & 31)) & 1u)
#define BSET(i, j) (b[BPOS(i, j) >> 5] |= 1u << (BPOS(i, j) & 31))
    int found_valid = 0, i = 0, j = 0;
    qd d; d.x = o_nint(xx); d.y = o_nint(yy); d.depth = 0;
    QPUSH(d); BSET(d.x, d.y);
    while (head < tail && !found_valid) {
        d = q[head++];
        i = d.x; j = d.y;
        if (A3(c, i, j, izz) != c->msgval && mask_valid(c, i, j)) found_valid = 1;
        /* neighbours are pushed even when this node is valid; d.depth accumulates across siblings */
        if (i - 1 >= c->sx && !BT(i - 1, j) && d.depth < maxdepth) { d.x = i - 1; d.y = j; d.depth = d.depth + 1; QPUSH(d); BSET(i - 1, j); }
        if (i + 1 <= c->ex && !BT(i + 1, j) && d.depth < maxdepth) { d.x = i + 1; d.y = j; d.depth = d.depth + 1; QPUSH(d); BSET(i + 1, j); }
        if (j - 1 >= c->sy && !BT(i, j - 1) && d.depth < maxdepth) { d.x = i; d.y = j - 1; d.depth = d.depth + 1; QPUSH(d); BSET(i, j - 1); }
        if (j + 1 <= c->ey && !BT(i, j + 1) && d.depth < maxdepth) { d.x = i; d.y = j + 1; d.depth = d.depth + 1; QPUSH(d); BSET(i, j + 1); }
    }
    float r;
    if (found_valid) {
        float distance = ((float)i - xx) * ((float)i - xx) + ((float)j - yy) * ((float)j - yy);
        r = A3(c, i, j, izz);
        while (head < tail) {
            d = q[head++];
            if (A3(c, d.x, d.y, izz) != c->msgval && mask_valid(c, d.x, d.y)) {
                float dd = ((float)d.x - xx) * ((float)d.x - xx) + ((float)d.y - yy) * ((float)d.y - yy);
                if (dd < distance) { distance = dd; r = A3(c, d.x, d.y, izz); }
            }
        }
    } else r = c->msgval;
    free(q); free(b);
    return r;
#undef QPUSH
#undef BT
#undef BSET
#undef BPOS
}

/* pop order of the BFS (ignores data); used to test the closed-form ring order of SURVEY 8(a8) */
int orc_search_order(int x0, int y0, int sx, int ex, int sy, int ey, int maxdepth, int *xs, int *ys, int maxn)
{
    int nx = ex - sx + 1, ny = ey - sy + 1;
    unsigned char *b = (unsigned char *)calloc((size_t)nx * (size_t)ny, 1);
    long cap = 1024, head = 0, tail = 0; int n = 0;
    qd *q = (qd *)malloc(sizeof(qd) * (size_t)cap);
#define QPUSH(v) do { if (tail == cap) { cap *= 2; q = (qd *)realloc(q, sizeof(qd) * (size_t)cap); } q[tail++] = (v); } while (0)
#define BT(i, j) b[(long)((j) - sy) * nx + ((i) - sx)]
    qd d; d.x = x0; d.y = y0; d.depth = 0; QPUSH(d); BT(x0, y0) = 1;
    while (head < tail) {
        d = q[head++];
        int i = d.x, j = d.y;
        if (n < maxn) { xs[n] = i; ys[n] = j; }
        ++n;
        if (i - 1 >= sx && !BT(i - 1, j) && d.depth < maxdepth) { d.x = i - 1; d.y = j; d.depth++; QPUSH(d); BT(i - 1, j) = 1; }
        if (i + 1 <= ex && !BT(i + 1, j) && d.depth < maxdepth) { d.x = i + 1; d.y = j; d.depth++; QPUSH(d); BT(i + 1, j) = 1; }
        if (j - 1 >= sy && !BT(i, j - 1) && d.depth < maxdepth) { d.x = i; d.y = j - 1; d.depth++; QPUSH(d); BT(i, j - 1) = 1; }
        if (j + 1 <= ey && !BT(i, j + 1) && d.depth < maxdepth) { d.x = i; d.y = j + 1; d.depth++; QPUSH(d); BT(i, j + 1) = 1; }
    }
    free(q); free(b);
    return n;
#undef QPUSH
#undef BT
}

/* interp_module.F:623-734 (wt=0) and :742-853 (wt=1, including the `ifx < start_y` bug at :795) */
static float four_pt_avg(const ictx *c, float xx, float yy, int izz, int idx, int wt)
{
    float fxfy = 1.0f, fxcy = 1.0f, cxfy = 1.0f, cxcy = 1.0f;
    int ifx = o_floor(xx), icx = o_ceiling(xx), ify = o_floor(yy), icy = o_ceiling(yy);
    if (wt) {
        float dxf = xx - (float)ifx, dxc = xx - (float)icx, dyf = yy - (float)ify, dyc = yy - (float)icy;
        fxfy = fmaxf(0.0f, 1.0f - o_sqrt(dxf * dxf + dyf * dyf));
        fxcy = fmaxf(0.0f, 1.0f - o_sqrt(dxf * dxf + dyc * dyc));
        cxfy = fmaxf(0.0f, 1.0f - o_sqrt(dxc * dxc + dyf * dyf));
        cxcy = fmaxf(0.0f, 1.0f - o_sqrt(dxc * dxc + dyc * dyc));
    }
    if (ifx < c->sx || icx > c->ex || ify < c->sy || icy > c->ey) {
        if (xx > (float)c->sx - 0.5f && ifx < c->sx) { ifx = c->sx; icx = c->sx; }
        else if (xx < (float)c->ex + 0.5f && icx > c->ex) { ifx = c->ex; icx = c->ex; }
        if (yy > (float)c->sy - 0.5f && (wt ? (ifx < c->sy) : (ify < c->sy))) { ify = c->sy; icy = c->sy; }
        else if (yy < (float)c->ey + 0.5f && icy > c->ey) { ify = c->ey; icy = c->ey; }
        if (ifx < c->sx || icx > c->ex || ify < c->sy || icy > c->ey) return c->msgval;
    }
    if (A3(c, ifx, ify, izz) == c->msgval || masked_out(c, ifx, ify)) fxfy = 0.0f;
    if (A3(c, ifx, icy, izz) == c->msgval || masked_out(c, ifx, icy)) fxcy = 0.0f;
    if (A3(c, icx, ify, izz) == c->msgval || masked_out(c, icx, ify)) cxfy = 0.0f;
    if (A3(c, icx, icy, izz) == c->msgval || masked_out(c, icx, icy)) cxcy = 0.0f;
    if (fxfy == 0.0f && fxcy == 0.0f && cxfy == 0.0f && cxcy == 0.0f) return seq(c, xx, yy, izz, idx);
    return (fxfy * A3(c, ifx, ify, izz) + fxcy * A3(c, ifx, icy, izz) + cxfy * A3(c, icx, ify, izz) +
            cxcy * A3(c, icx, icy, izz)) / (fxfy + fxcy + cxfy + cxcy);
}

/* interp_module.F:861-950 (wt=0) and :958-1047 (wt=1) */
static float sixteen_pt_avg(const ictx *c, float xx, float yy, int izz, int idx, int wt)
{
    int ifx = o_floor(xx), ify = o_floor(yy);
    float weights[5][5];
    if (ifx < c->sx + 1 || ifx > c->ex - 2 || ify < c->sy + 1 || ify > c->ey - 2) return seq(c, xx, yy, izz, idx);
    float sum_weight = 0.0f;
    for (int i = 1; i <= 4; ++i)
        for (int j = 1; j <= 4; ++j) {
            int px = ifx + 3 - i, py = ify + 3 - j;
            float w;
            if (masked_out(c, px, py)) w = 0.0f;
            else if (wt) {
                float ddx = xx - (float)px, ddy = yy - (float)py;
                w = fmaxf(0.0f, 2.0f - o_sqrt(ddx * ddx + ddy * ddy));
            } else w = 1.0f;
            if (A3(c, px, py, izz) == c->msgval) w = 0.0f;
            weights[i][j] = w;
            sum_weight = sum_weight + w;
        }
    if (sum_weight == 0.0f) return seq(c, xx, yy, izz, idx);
    float sum = 0.0f;
    for (int i = 1; i <= 4; ++i)
        for (int j = 1; j <= 4; ++j) sum = sum + weights[i][j] * A3(c, ifx + 3 - i, ify + 3 - j, izz);
    return sum / sum_weight;
}

/* interp_module.F:1055-1171 */
static float four_pt(const ictx *c, float xx, float yy, int izz, int idx)
{
    int min_x = o_floor(xx), min_y = o_floor(yy), max_x = o_ceiling(xx), max_y = o_ceiling(yy);
    if (min_x < c->sx || max_x > c->ex) return seq(c, xx, yy, izz, idx);
    if (min_y < c->sy || max_y > c->ey) return seq(c, xx, yy, izz, idx);
    if (A3(c, min_x, min_y, izz) == c->msgval || masked_out(c, min_x, min_y) ||
        A3(c, max_x, min_y, izz) == c->msgval || masked_out(c, max_x, min_y) ||
        A3(c, min_x, max_y, izz) == c->msgval || masked_out(c, min_x, max_y) ||
        A3(c, max_x, max_y, izz) == c->msgval || masked_out(c, max_x, max_y))
        return seq(c, xx, yy, izz, idx);
    if (min_x == max_x) {
        if (min_y == max_y) return A3(c, min_x, min_y, izz);
        return A3(c, min_x, min_y, izz) * ((float)max_y - yy) + A3(c, min_x, max_y, izz) * (yy - (float)min_y);
    } else if (min_y == max_y) {
        return A3(c, min_x, min_y, izz) * ((float)max_x - xx) + A3(c, max_x, min_y, izz) * (xx - (float)min_x);
    }
    return (yy - (float)min_y) * (A3(c, min_x, max_y, izz) * ((float)max_x - xx) + A3(c, max_x, max_y, izz) * (xx - (float)min_x)) +
           ((float)max_y - yy) * (A3(c, min_x, min_y, izz) * ((float)max_x - xx) + A3(c, max_x, min_y, izz) * (xx - (float)min_x));
}

/* interp_module.F:1342-1374 */
float orc_oned(float x, float a, float b, float c, float d)
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

/* interp_module.F:1179-1334 */
static float sixteen_pt(const ictx *c, float xx, float yy, int izz, int idx)
{
    float stl[5][5];
    int is_masked = 0;
    if ((int)xx < c->sx || (int)xx > c->ex || (int)yy < c->sy || (int)yy > c->ey) return seq(c, xx, yy, izz, idx);
    int i = (int)(xx + 0.00001f), j = (int)(yy + 0.00001f);
    float x = xx - (float)i, y = yy - (float)j;
    if (fabsf(x) > 0.0001f || fabsf(y) > 0.0001f) {
        for (int k = 1; k <= 4; ++k) {
            int kk = i + k - 2;
            if (kk < c->sx) kk = c->sx; else if (kk > c->ex) kk = c->ex;
            for (int l = 1; l <= 4; ++l) {
                int ll = j + l - 2;
                if (ll < c->sy) ll = c->sy; else if (ll > c->ey) ll = c->ey;
                stl[k][l] = A3(c, kk, ll, izz);
                if (masked_out(c, kk, ll)) is_masked = 1;
                if (stl[k][l] == 0.0f && c->msgval != 0.0f) stl[k][l] = 1.0e-20f;
            }
        }
        for (int k = 1; k <= 4; ++k)
            for (int l = 1; l <= 4; ++l)
                if (stl[k][l] == c->msgval || (c->mask_mode && is_masked)) return seq(c, xx, yy, izz, idx);
        float a = orc_oned(x, stl[1][1], stl[2][1], stl[3][1], stl[4][1]);
        float b = orc_oned(x, stl[1][2], stl[2][2], stl[3][2], stl[4][2]);
        float cc = orc_oned(x, stl[1][3], stl[2][3], stl[3][3], stl[4][3]);
        float d = orc_oned(x, stl[1][4], stl[2][4], stl[3][4], stl[4][4]);
        float r = orc_oned(y, a, b, cc, d);
        /* n is always 16: the (n /= 16) averaging branch at :1291-1297 is dead */
        if (r == 1.0e-20f) r = 0.0f;
        return r;
    }
    if (i >= c->sx && i <= c->ex && j >= c->sy && j <= c->ey && mask_valid(c, i, j) && A3(c, i, j, izz) != c->msgval)
        return A3(c, i, j, izz);
    return seq(c, xx, yy, izz, idx);
}

/* interp_sequence, interp_module.F:304-367.  idx is 1-based as in the Fortran. */
static float seq(const ictx *c, float xx, float yy, int izz, int idx)
{
    int m = c->list[idx - 1];
    if (m == 0) return c->msgval;
    switch (m) {
    case ORC_FOUR_POINT: return four_pt(c, xx, yy, izz, idx + 1);
    case ORC_AVERAGE4: return four_pt_avg(c, xx, yy, izz, idx + 1, 0);
    case ORC_W_AVERAGE4: return four_pt_avg(c, xx, yy, izz, idx + 1, 1);
    case ORC_N_NEIGHBOR: return nearest_neighbor(c, xx, yy, izz, idx + 1);
    case ORC_SIXTEEN_POINT: return sixteen_pt(c, xx, yy, izz, idx + 1);
    case ORC_SEARCH: return search_extrap(c, xx, yy, izz, idx + 1);
    case ORC_AVERAGE16: return sixteen_pt_avg(c, xx, yy, izz, idx + 1, 0);
    case ORC_W_AVERAGE16: return sixteen_pt_avg(c, xx, yy, izz, idx + 1, 1);
    default: return 0.0f; /* function result left undefined in the Fortran */
    }
}

float orc_interp_sequence(float xx, float yy, int izz, const float *array, int sx, int ex, int sy, int ey,
                          int sz, int ez, float msgval, const int *list, const int *opts, int idx,
                          int mask_mode, char rel, float maskval, const float *mask)
{
    ictx c = { array, sx, ex, sy, ey, sz, ez, msgval, list, opts, mask_mode, rel, maskval, mask };
    return seq(&c, xx, yy, izz, idx);
}

void orc_interp_sequence_array(long n, const float *xx, const float *yy, int izz, const float *array,
                               int sx, int ex, int sy, int ey, int sz, int ez, float msgval,
                               const int *list, const int *opts, int mask_mode, char rel, float maskval,
                               const float *mask, float *out)
{
    ictx c = { array, sx, ex, sy, ey, sz, ez, msgval, list, opts, mask_mode, rel, maskval, mask };
    for (long k = 0; k < n; ++k) out[k] = seq(&c, xx[k], yy[k], izz, 1);
}
