/*
 * Oracle (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * Plain-C restatement of the integer/selection stages of the HydroDEM
 * conditioning hot path, for parity checks at sizes where the NumPy oracle
 * (oracle/stencils.py) would need too much memory.  Each function names the
 * reference lines it follows under /root/reference/cguerrero/hydrodem/.
 * The NumPy and C oracles are checked against each other and against the
 * reference goldens in tests/test_oracle_golden.py.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC) -> oracle/_build/libhydro_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_float_nan_last(const void *pa, const void *pb)
{
    float a = *(const float *)pa, b = *(const float *)pb;
    int na = isnan(a), nb = isnan(b);
    if (na || nb) return na - nb;
    return (a > b) - (a < b);
}

/* gather the window around (j,i); corners dropped when circular.  returns count */
static int gather(const float *g, int64_t nx, int64_t j, int64_t i, int ws, int circular, float *buf)
{
    int h = ws / 2, n = 0;
    for (int dy = -h; dy <= h; ++dy)
        for (int dx = -h; dx <= h; ++dx) {
            if (circular && abs(dy) == h && abs(dx) == h) continue; /* sliding_window.py:499 */
            buf[n++] = g[(j + dy) * nx + (i + dx)];
        }
    return n;
}

/* MajorityFilter.apply, filters/custom_filters.py:48-73.  out must be zeroed float64. */
void ho_majority(const float *g, double *out, int64_t ny, int64_t nx, int ws, int min_count)
{
    int h = ws / 2;
#pragma omp parallel
    {
        float *buf = (float *)malloc(sizeof(float) * ws * ws);
#pragma omp for schedule(static)
        for (int64_t j = h; j < ny - h; ++j)
            for (int64_t i = h; i < nx - h; ++i) {
                int n = gather(g, nx, j, i, ws, 1, buf);
                qsort(buf, n, sizeof(float), cmp_float_nan_last);
                /* a value has count >= m iff sorted[k+m-1] == sorted[k]; NaN never equal (:69) */
                for (int k = 0; k + min_count - 1 < n; ++k)
                    if (buf[k] == buf[k + min_count - 1]) { out[j * nx + i] = buf[k]; break; }
            }
        free(buf);
    }
}

/* ExpandFilter.apply, filters/custom_filters.py:102-125.  out must be zeroed float64. */
void ho_expand(const float *g, double *out, int64_t ny, int64_t nx, int ws)
{
    int h = ws / 2;
#pragma omp parallel for schedule(static)
    for (int64_t j = h; j < ny - h; ++j)
        for (int64_t i = h; i < nx - h; ++i) {
            int hit = 0;
            for (int dy = -h; dy <= h && !hit; ++dy)
                for (int dx = -h; dx <= h; ++dx) {
                    if (abs(dy) == h && abs(dx) == h) continue;
                    if (g[(j + dy) * nx + (i + dx)] > 0.0f) { hit = 1; break; } /* NaN > 0 false (:122-123) */
                }
            if (hit) out[j * nx + i] = 1.0;
        }
}

/* NEW stage N1 (parity unpinned): nanmedian over the window, border copied from the input. */
void ho_median(const float *g, float *out, int64_t ny, int64_t nx, int ws, int circular)
{
    int h = ws / 2;
    memcpy(out, g, sizeof(float) * ny * nx);
#pragma omp parallel
    {
        float *buf = (float *)malloc(sizeof(float) * ws * ws);
#pragma omp for schedule(static)
        for (int64_t j = h; j < ny - h; ++j)
            for (int64_t i = h; i < nx - h; ++i) {
                int n = gather(g, nx, j, i, ws, circular, buf);
                qsort(buf, n, sizeof(float), cmp_float_nan_last);
                int k = 0;
                while (k < n && !isnan(buf[k])) ++k;
                float r;
                if (k == 0) r = NAN;
                else if (k & 1) r = buf[k / 2];
                else r = (buf[k / 2 - 1] + buf[k / 2]) / 2.0f;
                out[j * nx + i] = r;
            }
        free(buf);
    }
}

/* ---- NEW stage N2 (parity unpinned): depression filling by priority-flood ---------------
 * Fixed point of Planchon-Darboux (2001) with eps = 0 and 8-connectivity:
 *   W = z on the frame, NaN cells are outlets at -inf (output NaN),
 *   W(c) = max(z(c), min over N8 of W(n)) elsewhere.
 * Priority-flood (Barnes et al. 2014) reaches the same unique fixed point
 * (minimax path elevation to an outlet). */
typedef struct { float lev; int64_t idx; } hnode;
typedef struct { hnode *a; int64_t n, cap; } heap;
static void hpush(heap *h, float lev, int64_t idx)
{
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->a = (hnode *)realloc(h->a, sizeof(hnode) * h->cap); }
    int64_t k = h->n++;
    while (k > 0) {
        int64_t p = (k - 1) / 2;
        if (h->a[p].lev <= lev) break;
        h->a[k] = h->a[p]; k = p;
    }
    h->a[k].lev = lev; h->a[k].idx = idx;
}
static hnode hpop(heap *h)
{
    hnode top = h->a[0], last = h->a[--h->n];
    int64_t k = 0;
    for (;;) {
        int64_t c = 2 * k + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->a[c + 1].lev < h->a[c].lev) ++c;
        if (h->a[c].lev >= last.lev) break;
        h->a[k] = h->a[c]; k = c;
    }
    if (h->n) h->a[k] = last;
    return top;
}

void ho_priority_flood(const float *z, float *w, int64_t ny, int64_t nx)
{
    uint8_t *seen = (uint8_t *)calloc((size_t)(ny * nx), 1);
    heap h = {0, 0, 0};
    for (int64_t j = 0; j < ny; ++j)
        for (int64_t i = 0; i < nx; ++i) {
            int64_t c = j * nx + i;
            if (isnan(z[c])) { w[c] = NAN; seen[c] = 1; hpush(&h, -INFINITY, c); }
            else if (j == 0 || i == 0 || j == ny - 1 || i == nx - 1) { w[c] = z[c]; seen[c] = 1; hpush(&h, z[c], c); }
        }
    while (h.n) {
        hnode t = hpop(&h);
        int64_t j = t.idx / nx, i = t.idx % nx;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int64_t y = j + dy, x = i + dx;
                if ((!dy && !dx) || y < 0 || x < 0 || y >= ny || x >= nx) continue;
                int64_t n = y * nx + x;
                if (seen[n]) continue;
                seen[n] = 1;
                w[n] = z[n] > t.lev ? z[n] : t.lev;
                hpush(&h, w[n], n);
            }
    }
    free(h.a); free(seen);
}

/* ---- NEW stage N3 (parity unpinned): D8 flow direction, ESRI codes -----------------------
 * E=1 SE=2 S=4 SW=8 W=16 NW=32 N=64 NE=128; steepest positive drop, diagonal drops scaled by
 * 0.70710678f in float32, ties -> first in that order, frame cells / NaN centre / no drop -> 0. */
void ho_d8(const float *w, uint8_t *out, int64_t ny, int64_t nx)
{
    static const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    memset(out, 0, (size_t)(ny * nx));
#pragma omp parallel for schedule(static)
    for (int64_t j = 1; j < ny - 1; ++j)
        for (int64_t i = 1; i < nx - 1; ++i) {
            float c = w[j * nx + i], best = 0.0f;
            uint8_t code = 0;
            for (int k = 0; k < 8; ++k) {
                volatile float drop = c - w[(j + dy[k]) * nx + (i + dx[k])];
                if (k & 1) drop = drop * 0.70710678f;
                if (drop > best) { best = drop; code = (uint8_t)(1u << k); }
            }
            out[j * nx + i] = code;
        }
}


/* RouteRivers.apply, filters/custom_filters.py:165-199, window_size = 3: raster scan over interior cells of the mask with
 * int(v) == 1 (float32 snapshot); minimum of the 3x3 window of the float32 working DEM g (np.amin: NaN poisons it), every
 * cell equal to it becomes river and 10000 in g.  g is modified in place; out (uint8) must be zeroed by the caller. */
void ho_route_rivers(const float* mask, float* g, unsigned char* out, long ny, long nx)
{
    for (long j = 1; j < ny - 1; ++j)
        for (long i = 1; i < nx - 1; ++i) {
            const float mv = mask[j * nx + i];
            if (!(mv == mv) || (int)mv != 1) continue;
            float m = g[(j - 1) * nx + (i - 1)];
            int nan = 0;
            for (int a = -1; a <= 1; ++a)
                for (int b = -1; b <= 1; ++b) {
                    const float v = g[(j + a) * nx + (i + b)];
                    if (v != v) nan = 1;
                    if (v < m) m = v;
                }
            if (nan) continue;
            for (int a = -1; a <= 1; ++a)
                for (int b = -1; b <= 1; ++b)
                    if (g[(j + a) * nx + (i + b)] == m) { g[(j + a) * nx + (i + b)] = 10000.0f; out[(j + a) * nx + (i + b)] = 1; }
        }
}
