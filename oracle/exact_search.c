/* ORACLE - TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * Exact float64 haversine neighbour search on the CPU, for checking WHOLE graphs at sizes where sklearn's ball
 * tree needs minutes (O1280 -> TriNodes 7: 6.6 M queries).  What is restated:
 *
 *   - the distance arithmetic of scikit-learn 1.9.0 `HaversineDistance64.rdist`
 *     (sklearn/metrics/_dist_metrics.pyx.tp:2639-2648: `sin_0 = sin(0.5*(x1[0]-x2[0]))`,
 *     `sin_1 = sin(0.5*(x1[1]-x2[1]))`, `sin_0*sin_0 + cos(x1[0])*cos(x2[0])*sin_1*sin_1`, x1 = query,
 *     x2 = tree point) with the C library's sin / cos - the calls sklearn's compiled code makes; its binary
 *     contains no fused multiply-add, hence -ffp-contract=off in the build recipe;
 *   - `NearestNeighbors.kneighbors` (call site /root/reference/src/anemoi/graphs/edges/builder.py:259-265): the
 *     k smallest rdist per query.  Returned are the kk >= k smallest, sorted by (rdist, source index), so the
 *     caller can apply either tie rule (lower source index, or compare with sklearn's own choice);
 *   - `NearestNeighbors.radius_neighbors_graph` (edges/builder.py:364-366): every source with
 *     `rdist <= sin(0.5*r)^2` (sklearn/neighbors/_binary_tree.pxi.tp:1952-1957, `_dist_to_rdist` :2656-2658).
 *
 * The candidate enumeration is NOT sklearn's tree: sources are binned into latitude bands x longitude bins and a
 * query scans every bin that meets the bounding box of a spherical cap, widening the cap until it provably
 * contains the kk-th neighbour.  Exact nearest neighbours do not depend on the enumeration; ties do, and are
 * reported to the caller instead of being decided here.
 *
 * Pinned by tests/test_oracle_exact_search.py against sklearn itself (random clouds, O96 / N320 grids, poles,
 * the date line, duplicates) - "-m 'not gpu'".
 *
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/exact_search.c -o oracle/_build/libexact_search.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define TWO_PI 6.28318530717958647692

typedef struct {
    int64_t n;
    int n_bands;
    double band_h;   /* latitude height of a band */
    int* n_lon;      /* bins per band */
    int64_t* band_off; /* first cell of a band; [n_bands + 1] */
    int64_t* cell_start; /* [n_cells + 1] */
    int32_t* idx;    /* source index, grouped by cell, ascending inside a cell */
    double* lat;     /* float64 copies in cell order */
    double* lon;
    double* coslat;
} grid_t;

static double wrap_2pi(double lon) {
    double w = fmod(lon, TWO_PI);
    if (w < 0) w += TWO_PI;
    if (w >= TWO_PI) w = 0.0;
    return w;
}

static int band_of(const grid_t* g, double lat) {
    int b = (int)floor((lat + 0.5 * PI) / g->band_h);
    if (b < 0) b = 0;
    if (b >= g->n_bands) b = g->n_bands - 1;
    return b;
}

void* oracle_grid_build(const float* latlon, int64_t n, double cell_rad) {
    grid_t* g = (grid_t*)calloc(1, sizeof(grid_t));
    g->n = n;
    g->n_bands = (int)ceil(PI / cell_rad);
    if (g->n_bands < 1) g->n_bands = 1;
    g->band_h = PI / g->n_bands;
    g->n_lon = (int*)malloc(sizeof(int) * g->n_bands);
    g->band_off = (int64_t*)malloc(sizeof(int64_t) * (g->n_bands + 1));
    int64_t cells = 0;
    for (int b = 0; b < g->n_bands; ++b) {
        double lo = -0.5 * PI + b * g->band_h, hi = lo + g->band_h;
        double c = (lo <= 0 && hi >= 0) ? 1.0 : fmax(cos(lo), cos(hi));
        int nl = (int)ceil(TWO_PI * c / cell_rad);
        if (nl < 1) nl = 1;
        g->n_lon[b] = nl;
        g->band_off[b] = cells;
        cells += nl;
    }
    g->band_off[g->n_bands] = cells;
    g->cell_start = (int64_t*)calloc(cells + 1, sizeof(int64_t));
    int64_t* cell_of = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        double lat = (double)latlon[2 * i], lon = wrap_2pi((double)latlon[2 * i + 1]);
        int b = band_of(g, lat);
        int l = (int)(lon / TWO_PI * g->n_lon[b]);
        if (l >= g->n_lon[b]) l = g->n_lon[b] - 1;
        cell_of[i] = g->band_off[b] + l;
        g->cell_start[cell_of[i] + 1]++;
    }
    for (int64_t c = 0; c < cells; ++c) g->cell_start[c + 1] += g->cell_start[c];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (cells > 0 ? cells : 1));
    memcpy(fill, g->cell_start, sizeof(int64_t) * cells);
    g->idx = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    g->lat = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
    g->lon = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
    g->coslat = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { /* ascending i => ascending index inside a cell */
        int64_t p = fill[cell_of[i]]++;
        g->idx[p] = (int32_t)i;
        g->lat[p] = (double)latlon[2 * i];
        g->lon[p] = (double)latlon[2 * i + 1];
        g->coslat[p] = cos(g->lat[p]);
    }
    free(fill);
    free(cell_of);
    return g;
}

void oracle_grid_free(void* h) {
    grid_t* g = (grid_t*)h;
    if (!g) return;
    free(g->n_lon); free(g->band_off); free(g->cell_start); free(g->idx); free(g->lat); free(g->lon); free(g->coslat);
    free(g);
}

/* sklearn HaversineDistance64.rdist, x1 = query (lat1, lon1), x2 = tree point; cos(x1[0]), cos(x2[0]) passed in */
static inline double rdist64(double lat1, double lon1, double c1, double lat2, double lon2, double c2) {
    double sin_0 = sin(0.5 * (lat1 - lat2));
    double sin_1 = sin(0.5 * (lon1 - lon2));
    return (sin_0 * sin_0 + c1 * c2 * sin_1 * sin_1);
}

typedef void (*visit_fn)(void* ctx, int32_t idx, double rd);

/* visit every source in the bins meeting the lat/lon bounding box of the cap of angular radius r around the query */
static void scan_cap(const grid_t* g, double qlat, double qlon, double r, visit_fn visit, void* ctx) {
    double c1 = cos(qlat);
    double pad = 1e-9;
    double lat_lo = qlat - r - pad, lat_hi = qlat + r + pad;
    int b0 = band_of(g, lat_lo), b1 = band_of(g, lat_hi);
    int whole = (lat_lo <= -0.5 * PI) || (lat_hi >= 0.5 * PI) || r >= 0.5 * PI;
    double half = PI;
    if (!whole) {
        double s = sin(r) / c1; /* half-width in longitude of the cap's bounding box */
        if (s >= 1.0) whole = 1; else half = asin(s) + pad;
    }
    double qw = wrap_2pi(qlon);
    for (int b = b0; b <= b1; ++b) {
        int nl = g->n_lon[b];
        int64_t base = g->band_off[b];
        int l0 = 0, cnt = nl;
        if (!whole && 2.0 * half < TWO_PI) {
            double w = TWO_PI / nl;
            int lo = (int)floor((qw - half) / w), hi = (int)floor((qw + half) / w);
            cnt = hi - lo + 1;
            if (cnt >= nl) { cnt = nl; l0 = 0; } else { l0 = ((lo % nl) + nl) % nl; }
        }
        for (int t = 0; t < cnt; ++t) {
            int l = l0 + t; if (l >= nl) l -= nl;
            int64_t s = g->cell_start[base + l], e = g->cell_start[base + l + 1];
            for (int64_t p = s; p < e; ++p)
                visit(ctx, g->idx[p], rdist64(qlat, qlon, c1, g->lat[p], g->lon[p], g->coslat[p]));
        }
    }
}

/* ---- k nearest ------------------------------------------------------------------------------------------ */
typedef struct { int kk, n; double* rd; int32_t* id; } topk_t;

static void topk_visit(void* ctx, int32_t idx, double rd) {
    topk_t* t = (topk_t*)ctx;
    int n = t->n;
    if (n == t->kk) { /* full: compare with the current last (largest (rd, idx)) */
        if (rd > t->rd[n - 1] || (rd == t->rd[n - 1] && idx > t->id[n - 1])) return;
        n--;
    }
    int p = n;
    while (p > 0 && (t->rd[p - 1] > rd || (t->rd[p - 1] == rd && t->id[p - 1] > idx))) {
        t->rd[p] = t->rd[p - 1]; t->id[p] = t->id[p - 1]; --p;
    }
    t->rd[p] = rd; t->id[p] = idx;
    t->n = n + 1;
}

/* kk smallest (rdist, index) per query, ascending.  r0 = first cap radius (radians).  Returns 0. */
int oracle_knn(void* h, const float* q, int64_t nq, int kk, double r0, int32_t* out_idx, double* out_rd) {
    const grid_t* g = (const grid_t*)h;
    if (kk > g->n) return -1;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < nq; ++i) {
        topk_t t; t.kk = kk; t.rd = out_rd + i * kk; t.id = out_idx + i * kk;
        double qlat = (double)q[2 * i], qlon = (double)q[2 * i + 1];
        double r = r0;
        for (;;) {
            t.n = 0;
            scan_cap(g, qlat, qlon, r, topk_visit, &t);
            if (r >= PI) break;
            if (t.n == kk) { /* complete iff the kk-th distance lies strictly inside the cap (with slack) */
                double s = sin(0.5 * r * 0.999);
                if (t.rd[kk - 1] < s * s) break;
            }
            r *= 2.0; if (r > PI) r = PI;
        }
    }
    return 0;
}

/* ---- radius --------------------------------------------------------------------------------------------- */
typedef struct { double thr, tau; int64_t count, near; int32_t* out; } rad_t;

static void rad_visit(void* ctx, int32_t idx, double rd) {
    rad_t* t = (rad_t*)ctx;
    if (fabs(rd - t->thr) <= t->tau * t->thr) t->near++;
    if (rd <= t->thr) { if (t->out) t->out[t->count] = idx; t->count++; }
}

/* counts[i] = sources with rdist <= sin(0.5 r)^2; near[i] = sources within tau (relative) of that threshold */
int oracle_radius_count(void* h, const float* q, int64_t nq, double radius, double tau, int64_t* counts, int64_t* near) {
    const grid_t* g = (const grid_t*)h;
    double tmp = sin(0.5 * radius), thr = tmp * tmp;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
        rad_t t = {thr, tau, 0, 0, NULL};
        scan_cap(g, (double)q[2 * i], (double)q[2 * i + 1], radius * 1.001 + 1e-9, rad_visit, &t);
        counts[i] = t.count; if (near) near[i] = t.near;
    }
    return 0;
}

static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x > y) - (x < y); }

/* out_src[offsets[i] .. offsets[i+1]) = the sources of query i, ascending */
int oracle_radius_fill(void* h, const float* q, int64_t nq, double radius, const int64_t* offsets, int32_t* out_src) {
    const grid_t* g = (const grid_t*)h;
    double tmp = sin(0.5 * radius), thr = tmp * tmp;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
        rad_t t = {thr, 0.0, 0, 0, out_src + offsets[i]};
        scan_cap(g, (double)q[2 * i], (double)q[2 * i + 1], radius * 1.001 + 1e-9, rad_visit, &t);
        qsort(out_src + offsets[i], (size_t)t.count, sizeof(int32_t), cmp_i32);
    }
    return 0;
}

/* rdist of explicit (query, source) pairs - for bit-level comparisons in the boundary-case report */
int oracle_pair_rdist(const float* q, const float* s, int64_t n, double* out) {
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        double lat1 = q[2 * i], lon1 = q[2 * i + 1], lat2 = s[2 * i], lon2 = s[2 * i + 1];
        out[i] = rdist64(lat1, lon1, cos(lat1), lat2, lon2, cos(lat2));
    }
    return 0;
}
