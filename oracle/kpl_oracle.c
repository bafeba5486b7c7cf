/*
 * kpl_oracle.c -- CPU oracle for the Keypoint-Learning detection hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (keypoint_learning_b200/csrc) never includes, links or calls anything in this file.
 *
 * PARITY PIN: the reference (CVLAB-Unibo/Keypoint-Learning) ships no tests, no golden outputs and its
 * trained forests are absent from the checkout, and it cannot be built as a whole here (PCL 1.8.0 / FLANN /
 * Eigen / OpenCV 3.2 C++ are not installed).  Its OWN code on this path can be: oracle/_ref compiles
 * src/KeypointLearning.cpp (findAnnulusPair / findBinPair) and the detector templates of
 * include/KeypointLearning.h + include/impl/KeypointLearning.hpp from the mounted reference tree against a
 * stand-in environment (oracle/ref_stubs/kplref_env.h), and tests/test_oracle.py holds this file to it BIT FOR
 * BIT: the binning helpers over millions of inputs, computePointFeatures rows, runForest scores, the
 * threshold + local-maximum NMS and the draws-remove branch on crops of the bundled views.
 * What is NOT the reference's own code -- and therefore restated here from the upstream versions the
 * reference pins (README.md:66-67), their sources not being in the container -- is third-party:
 * PCL 1.8.0 NormalEstimation / computeMeanAndCovarianceMatrix / eigen33 / computeRoots, FLANN L2_Simple +
 * strict radius test, OpenCV DTreesImpl::predictTrees with PREDICT_SUM, Eigen's small reductions.  Those are
 * pinned independently in tests/: cv2.ml.RTrees (the real OpenCV) for the forest stage, scipy cKDTree for
 * neighbour sets, float64 PCA for normals.  This file follows
 *   include/impl/KeypointLearning.hpp:179-263  (detectKeypoints: threshold + local-max NMS)
 *   include/impl/KeypointLearning.hpp:267-296  (runForest: score = 1 - sum/ntrees)
 *   include/impl/KeypointLearning.hpp:321-376  (computePointFeatures: annuli x bins histogram)
 *   src/KeypointLearning.cpp:41-65, 68-92      (findAnnulusPair, findBinPair)
 *   src/main_test_detector.cpp:162-169         (k-NN(10) PCA normals, viewpoint 0,0,0)
 *
 * Arithmetic contract (shared with the CUDA kernels, which implement it independently):
 *   - all per-pair math in IEEE binary32, round-to-nearest, NO fused multiply-add
 *     (build with -ffp-contract=off; x86-64 SSE2 has no excess precision);
 *   - squared distance  d2 = ((dx*dx) + dy*dy) + dz*dz, neighbour iff d2 < (float)(r*r), r double;
 *   - Eigen fixed-size-3 reductions are a0 + (a1 + a2)   (Eigen redux_novec_unroller);
 *   - atan2/cos/sin inside computeRoots are evaluated by libm-independent double-precision
 *     series (below) and rounded to float once, so that CPU and GPU agree bit-for-bit; they are
 *     within 1 ulp of a correctly rounded libm (the reference's MSVC CRT is not reproducible);
 *   - histogram votes are accumulated in float in a DEFINED neighbour order:
 *       order 0: ascending point index (grid-free restatement);
 *       order 1: ascending (canonical cell key, point index) -- the order the GPU uses;
 *       order 2: oracle traversal order (unsorted; timing only).
 *     The reference's own order (FLANN unsorted traversal) cannot be reproduced by anyone.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KPLO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* libm-independent trig (double series, no FMA).  Domain: finite inputs; cos/sin for |x|<=1.6  */
/* ------------------------------------------------------------------------------------------ */
static double atan_series(double z) /* |z| <= tan(pi/8) */
{
    double z2 = z * z;
    double s = 1.0 / 39.0;
    for (int k = 18; k >= 0; --k) s = 1.0 / (double)(2 * k + 1) - z2 * s;
    return z * s;
}
static double atan_unit(double a) /* a in [0,1] */
{
    if (a > 0.41421356237309503) {
        double z = (a - 1.0) / (a + 1.0);
        return 0.78539816339744828 + atan_series(z);
    }
    return atan_series(a);
}
KPLO_API float kplo_atan2f(float y, float x)
{
    if (isnan(x) || isnan(y)) return NAN;
    double ax = fabs((double)x), ay = fabs((double)y), r;
    if (ax == 0.0 && ay == 0.0) r = 0.0;
    else if (ax >= ay) r = atan_unit(ay / ax);
    else r = 1.5707963267948966 - atan_unit(ax / ay);
    if (signbit(x)) r = 3.1415926535897931 - r;
    if (signbit(y)) r = -r;
    return (float)r;
}
KPLO_API float kplo_cosf(float xf)
{
    double x = (double)xf, x2 = x * x, s = 1.0;
    for (int k = 13; k >= 1; --k) s = 1.0 - (x2 / (double)((2 * k - 1) * (2 * k))) * s;
    return (float)s;
}
KPLO_API float kplo_sinf(float xf)
{
    double x = (double)xf, x2 = x * x, s = 1.0;
    for (int k = 13; k >= 1; --k) s = 1.0 - (x2 / (double)((2 * k) * (2 * k + 1))) * s;
    return (float)(x * s);
}

/* ------------------------------------------------------------------------------------------ */
/* src/KeypointLearning.cpp:41-65 and :68-92 (abs() there is float-abs on the authors' MSVC)    */
/* ------------------------------------------------------------------------------------------ */
KPLO_API void kplo_find_annulus_pair(int n_annulus, float distance, float support,
                                     int* annulus_index, int* annulus_index_pair, float* annulus_weight)
{
    float dim = support / (float)n_annulus;
    int a = (int)floorf(distance / dim);
    if (a == n_annulus) a--;
    float center = ((float)a * dim) + (dim / 2.0f);
    float w = distance - center;
    w = w / dim;
    int p = (w > 0) ? a + 1 : a - 1;
    if (p == -1) p = 0;
    if (p == n_annulus) p = a;
    *annulus_index = a; *annulus_index_pair = p; *annulus_weight = fabsf(w);
}
KPLO_API void kplo_find_bin_pair(int n_bins, float cosine, int* bin_index, int* bin_index_pair, float* bin_weight)
{
    if (cosine < 0) cosine = 0;
    if (cosine > 2) cosine = 2;
    float dim = 2.0f / (float)n_bins;
    int b = (int)floorf(cosine / dim);
    if (b == n_bins) b--;
    float center = ((float)b * dim) + (dim / 2.0f);
    float w = cosine - center;
    w = w / dim;
    int p = (w > 0) ? b + 1 : b - 1;
    if (p == -1) p = 0;
    if (p == n_bins) p = b;
    *bin_index = b; *bin_index_pair = p; *bin_weight = fabsf(w);
}

/* whole arrays at once (tests sweep millions of inputs against oracle/_ref) */
KPLO_API void kplo_annulus_sweep(int n_annulus, float support, const float* distance, int64_t n, int* index, int* pair, float* weight)
{
    for (int64_t i = 0; i < n; ++i) kplo_find_annulus_pair(n_annulus, distance[i], support, index + i, pair + i, weight + i);
}
KPLO_API void kplo_bin_sweep(int n_bins, const float* cosine, int64_t n, int* index, int* pair, float* weight)
{
    for (int64_t i = 0; i < n; ++i) kplo_find_bin_pair(n_bins, cosine[i], index + i, pair + i, weight + i);
}

/* FLANN L2_Simple<float>: result += diff*diff over x,y,z, FP32, no FMA */
static inline float dist2(const float* a, const float* b)
{
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return ((dx * dx) + dy * dy) + dz * dz;
}

/* ------------------------------------------------------------------------------------------ */
/* The oracle's own spatial index: dense uniform grid, counting sort (stable => ascending index  */
/* inside a cell).  Independent of the product's canonical grid.                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t n; const float* xyz;
    double org[3], h; int dim[3];
    int32_t* start;  /* ncell+1 */
    int32_t* item;   /* n, point indices grouped by cell */
    int32_t* cc;     /* n*3 cell coords per point */
} ogrid;

static void ogrid_free(ogrid* g) { free(g->start); free(g->item); free(g->cc); memset(g, 0, sizeof *g); }

static int ogrid_build(ogrid* g, const float* xyz, int64_t n, double h)
{
    memset(g, 0, sizeof *g);
    g->n = n; g->xyz = xyz;
    if (n <= 0 || !(h > 0)) return -1;
    double lo[3] = { xyz[0], xyz[1], xyz[2] }, hi[3] = { xyz[0], xyz[1], xyz[2] };
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            double v = xyz[3 * i + a];
            if (!isfinite(v)) return -2;
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    for (;;) {
        double nc = 1;
        for (int a = 0; a < 3; ++a) { g->dim[a] = (int)floor((hi[a] - lo[a]) / h) + 1; nc *= g->dim[a]; }
        if (nc <= 134217728.0) break;
        h *= 2;
    }
    g->h = h; for (int a = 0; a < 3; ++a) g->org[a] = lo[a];
    int64_t ncell = (int64_t)g->dim[0] * g->dim[1] * g->dim[2];
    g->start = (int32_t*)calloc((size_t)ncell + 1, sizeof(int32_t));
    g->item = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    g->cc = (int32_t*)malloc((size_t)n * 3 * sizeof(int32_t));
    if (!g->start || !g->item || !g->cc) { ogrid_free(g); return -3; }
    for (int64_t i = 0; i < n; ++i) {
        int c[3];
        for (int a = 0; a < 3; ++a) {
            c[a] = (int)floor(((double)xyz[3 * i + a] - g->org[a]) / h);
            if (c[a] < 0) c[a] = 0;
            if (c[a] >= g->dim[a]) c[a] = g->dim[a] - 1;
            g->cc[3 * i + a] = c[a];
        }
        g->start[((int64_t)c[2] * g->dim[1] + c[1]) * g->dim[0] + c[0] + 1]++;
    }
    for (int64_t k = 0; k < ncell; ++k) g->start[k + 1] += g->start[k];
    int32_t* fill = (int32_t*)malloc((size_t)ncell * sizeof(int32_t));
    if (!fill) { ogrid_free(g); return -3; }
    memcpy(fill, g->start, (size_t)ncell * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* c = g->cc + 3 * i;
        g->item[fill[((int64_t)c[2] * g->dim[1] + c[1]) * g->dim[0] + c[0]]++] = (int32_t)i;
    }
    free(fill);
    return 0;
}

typedef struct { int32_t idx; float d2; } nb_t;
typedef struct { nb_t* v; int64_t n, cap; } nbvec;
static inline int nbvec_push(nbvec* b, int32_t idx, float d2)
{
    if (b->n == b->cap) {
        int64_t nc = b->cap ? b->cap * 2 : 4096;
        nb_t* p = (nb_t*)realloc(b->v, (size_t)nc * sizeof(nb_t));
        if (!p) return -1;
        b->v = p; b->cap = nc;
    }
    b->v[b->n].idx = idx; b->v[b->n].d2 = d2; b->n++;
    return 0;
}

/* all j (self included) with d2 < r2, oracle traversal order */
static int ogrid_radius(const ogrid* g, int64_t q, double radius, float r2, nbvec* out)
{
    out->n = 0;
    const float* p = g->xyz + 3 * q;
    int reach = (int)floor(radius * (1.0 + 1e-6) / g->h) + 1;
    const int32_t* c = g->cc + 3 * q;
    double rr = radius * (1.0 + 1e-5); rr *= rr;
    for (int z = c[2] - reach; z <= c[2] + reach; ++z) {
        if (z < 0 || z >= g->dim[2]) continue;
        double z0 = g->org[2] + z * g->h, gz = p[2] < z0 ? z0 - p[2] : (p[2] > z0 + g->h ? p[2] - z0 - g->h : 0.0);
        for (int y = c[1] - reach; y <= c[1] + reach; ++y) {
            if (y < 0 || y >= g->dim[1]) continue;
            double y0 = g->org[1] + y * g->h, gy = p[1] < y0 ? y0 - p[1] : (p[1] > y0 + g->h ? p[1] - y0 - g->h : 0.0);
            if (gz * gz + gy * gy > rr) continue;
            for (int x = c[0] - reach; x <= c[0] + reach; ++x) {
                if (x < 0 || x >= g->dim[0]) continue;
                /* border cells also hold clamped points: never cull them by geometry */
                int border = (x == 0 || y == 0 || z == 0 || x == g->dim[0] - 1 || y == g->dim[1] - 1 || z == g->dim[2] - 1);
                double x0 = g->org[0] + x * g->h, gx = p[0] < x0 ? x0 - p[0] : (p[0] > x0 + g->h ? p[0] - x0 - g->h : 0.0);
                if (!border && gz * gz + gy * gy + gx * gx > rr) continue;
                int64_t cell = ((int64_t)z * g->dim[1] + y) * g->dim[0] + x;
                for (int32_t s = g->start[cell]; s < g->start[cell + 1]; ++s) {
                    int32_t j = g->item[s];
                    float d2 = dist2(p, g->xyz + 3 * (int64_t)j);
                    if (d2 < r2) if (nbvec_push(out, j, d2)) return -1;
                }
            }
        }
    }
    return 0;
}

static int cmp_idx(const void* a, const void* b)
{
    int32_t x = ((const nb_t*)a)->idx, y = ((const nb_t*)b)->idx;
    return (x > y) - (x < y);
}
static int cmp_d2_idx(const void* a, const void* b)
{
    const nb_t *x = (const nb_t*)a, *y = (const nb_t*)b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

static double auto_cell(const float* xyz, int64_t n, double scale)
{
    /* 2.5D heuristic: sqrt(area of the two largest bbox extents / n) * scale */
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int64_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {
        double v = xyz[3 * i + a]; if (v < lo[a]) lo[a] = v; if (v > hi[a]) hi[a] = v; }
    double e[3] = { hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] };
    double mn = e[0]; if (e[1] < mn) mn = e[1]; if (e[2] < mn) mn = e[2];
    double area = (mn > 0) ? e[0] * e[1] * e[2] / mn : 0.0;
    if (mn <= 0) { double mx = e[0] > e[1] ? e[0] : e[1]; if (e[2] > mx) mx = e[2]; area = mx * mx; }
    double h = sqrt(area / (double)(n > 0 ? n : 1)) * scale;
    if (!(h > 0)) h = 1.0;
    return h;
}

/* ------------------------------------------------------------------------------------------ */
/* F3: radius neighbour sets (FLANN radiusSearch semantics: strict <, self included)            */
/* ------------------------------------------------------------------------------------------ */
KPLO_API int kplo_radius_counts(const float* xyz, int64_t n, double radius, int32_t* counts)
{
    ogrid g; int rc = ogrid_build(&g, xyz, n, radius / 2.0);
    if (rc) return rc;
    float r2 = (float)(radius * radius);
    int err = 0;
#pragma omp parallel
    {
        nbvec nb = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            if (ogrid_radius(&g, i, radius, r2, &nb)) { err = 1; continue; }
            counts[i] = (int32_t)nb.n;
        }
        free(nb.v);
    }
    ogrid_free(&g);
    return err ? -3 : 0;
}

/* neighbour lists (ascending index, self included) for m query indices; two-call protocol:
 * out_idx == NULL fills offsets only. */
KPLO_API int kplo_radius_neighbors(const float* xyz, int64_t n, double radius, const int32_t* qidx, int64_t m,
                                   int64_t* offsets, int32_t* out_idx)
{
    ogrid g; int rc = ogrid_build(&g, xyz, n, radius / 2.0);
    if (rc) return rc;
    float r2 = (float)(radius * radius);
    nbvec nb = { 0, 0, 0 };
    int64_t off = 0;
    for (int64_t k = 0; k < m; ++k) {
        int64_t q = qidx ? qidx[k] : k;
        if (ogrid_radius(&g, q, radius, r2, &nb)) { free(nb.v); ogrid_free(&g); return -3; }
        qsort(nb.v, (size_t)nb.n, sizeof(nb_t), cmp_idx);
        if (out_idx) for (int64_t t = 0; t < nb.n; ++t) out_idx[off + t] = nb.v[t].idx;
        offsets[k] = off; off += nb.n;
    }
    offsets[m] = off;
    free(nb.v); ogrid_free(&g);
    return 0;
}

/* O(n*m) brute force counts: the independent check of the grid search on small crops */
KPLO_API void kplo_radius_counts_brute(const float* xyz, int64_t n, double radius, int32_t* counts)
{
    float r2 = (float)(radius * radius);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int32_t c = 0;
        for (int64_t j = 0; j < n; ++j) if (dist2(xyz + 3 * i, xyz + 3 * j) < r2) c++;
        counts[i] = c;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* F2: PCL 1.8.0 NormalEstimation: computeMeanAndCovarianceMatrix (single pass, un-centred,     */
/* FP32, neighbour order) -> eigen33 -> flipNormalTowardsViewpoint      [3P-recalled]           */
/* ------------------------------------------------------------------------------------------ */
static void compute_roots2(float b, float c, float roots[3])
{
    roots[0] = 0.0f;
    float d = (float)((double)(b * b) - 4.0 * (double)c);
    if (d < 0.0f) d = 0.0f;
    float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}
static void compute_roots(const float m[9], float roots[3])
{
    float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (fabsf(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    float theta = kplo_atan2f(sqrtf(-q), half_b) * s_inv3;
    float cos_theta = kplo_cosf(theta);
    float sin_theta = kplo_sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
        t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}
static inline void cross3(const float* a, const float* b, float* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float sqnorm3(const float* v) { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); }

static void eigen33_smallest(const float cov[9], float* eigenvalue, float ev[3])
{
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) { float a = fabsf(cov[i]); if (a > scale) scale = a; }
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[9];
    for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
    float roots[3];
    compute_roots(m, roots);
    *eigenvalue = roots[0] * scale;
    m[0] -= roots[0]; m[4] -= roots[0]; m[8] -= roots[0];
    float v1[3], v2[3], v3[3];
    cross3(m + 0, m + 3, v1);
    cross3(m + 0, m + 6, v2);
    cross3(m + 3, m + 6, v3);
    float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
    const float* v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float s = sqrtf(l);
    ev[0] = v[0] / s; ev[1] = v[1] / s; ev[2] = v[2] / s;
}

/* neighbours given in accumulation order */
static void point_normal(const float* xyz, const nb_t* nb, int cnt, const float* p, const float vp[3], float out[4])
{
    if (cnt < 3) { out[0] = out[1] = out[2] = out[3] = NAN; return; }
    float accu[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int t = 0; t < cnt; ++t) {
        const float* c = xyz + 3 * (int64_t)nb[t].idx;
        accu[0] += c[0] * c[0]; accu[1] += c[0] * c[1]; accu[2] += c[0] * c[2];
        accu[3] += c[1] * c[1]; accu[4] += c[1] * c[2]; accu[5] += c[2] * c[2];
        accu[6] += c[0]; accu[7] += c[1]; accu[8] += c[2];
    }
    float fn = (float)cnt;
    for (int i = 0; i < 9; ++i) accu[i] = accu[i] / fn;
    float cov[9];
    cov[0] = accu[0] - accu[6] * accu[6];
    cov[1] = accu[1] - accu[6] * accu[7];
    cov[2] = accu[2] - accu[6] * accu[8];
    cov[4] = accu[3] - accu[7] * accu[7];
    cov[5] = accu[4] - accu[7] * accu[8];
    cov[8] = accu[5] - accu[8] * accu[8];
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    float ev[3], lambda;
    eigen33_smallest(cov, &lambda, ev);
    float eig_sum = cov[0] + cov[4] + cov[8];
    float curv = (eig_sum != 0) ? fabsf(lambda / eig_sum) : 0.0f;
    float vx = vp[0] - p[0], vy = vp[1] - p[1], vz = vp[2] - p[2];
    float cos_theta = (vx * ev[0] + vy * ev[1] + vz * ev[2]);
    if (cos_theta < 0) { ev[0] *= -1; ev[1] *= -1; ev[2] *= -1; }
    out[0] = ev[0]; out[1] = ev[1]; out[2] = ev[2]; out[3] = curv;
}

/* exact k nearest (self included) by expanding rings; result sorted by (d2, index) */
static int ogrid_knn(const ogrid* g, int64_t q, int k, nb_t* best /* k */)
{
    const float* p = g->xyz + 3 * q;
    const int32_t* c = g->cc + 3 * q;
    int cnt = 0;
    int maxdim = g->dim[0]; if (g->dim[1] > maxdim) maxdim = g->dim[1]; if (g->dim[2] > maxdim) maxdim = g->dim[2];
    for (int R = 0; R <= maxdim; ++R) {
        for (int z = c[2] - R; z <= c[2] + R; ++z) {
            if (z < 0 || z >= g->dim[2]) continue;
            for (int y = c[1] - R; y <= c[1] + R; ++y) {
                if (y < 0 || y >= g->dim[1]) continue;
                int shell_zy = (z == c[2] - R || z == c[2] + R || y == c[1] - R || y == c[1] + R);
                for (int x = c[0] - R; x <= c[0] + R; ++x) {
                    if (x < 0 || x >= g->dim[0]) continue;
                    if (!shell_zy && x != c[0] - R && x != c[0] + R) continue;
                    int64_t cell = ((int64_t)z * g->dim[1] + y) * g->dim[0] + x;
                    for (int32_t s = g->start[cell]; s < g->start[cell + 1]; ++s) {
                        nb_t cand; cand.idx = g->item[s]; cand.d2 = dist2(p, g->xyz + 3 * (int64_t)cand.idx);
                        if (cnt == k && cmp_d2_idx(&cand, &best[k - 1]) >= 0) continue;
                        int pos = (cnt < k) ? cnt++ : k - 1;
                        while (pos > 0 && cmp_d2_idx(&cand, &best[pos - 1]) < 0) { best[pos] = best[pos - 1]; pos--; }
                        best[pos] = cand;
                    }
                }
            }
        }
        if (R >= 1 && cnt == k) {
            /* every unscanned point is farther than (R*h) minus clamping slack; border cells hold
             * clamped points only when the point lies outside the bbox, which cannot happen here */
            double guard = (double)R * g->h; guard = guard * guard * (1.0 - 1e-6);
            if ((double)best[k - 1].d2 < guard) break;
        }
        if (c[0] - R <= 0 && c[1] - R <= 0 && c[2] - R <= 0 &&
            c[0] + R >= g->dim[0] - 1 && c[1] + R >= g->dim[1] - 1 && c[2] + R >= g->dim[2] - 1) break;
    }
    return cnt;
}

/* normals4: n x (nx,ny,nz,curvature).  src/main_test_detector.cpp:162-169 with k = 10. */
KPLO_API int kplo_normals_knn(const float* xyz, int64_t n, int k, const float vp[3], double cell, float* normals4)
{
    if (k < 1 || k > 64) return -1;
    ogrid g; int rc = ogrid_build(&g, xyz, n, cell > 0 ? cell : auto_cell(xyz, n, 3.0));
    if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        nb_t best[64];
        int cnt = ogrid_knn(&g, i, k, best);
        point_normal(xyz, best, cnt, xyz + 3 * i, vp, normals4 + 4 * i);
    }
    ogrid_free(&g);
    return 0;
}

/* indices of the k nearest neighbours (for cross-checks against scipy) */
KPLO_API int kplo_knn_indices(const float* xyz, int64_t n, int k, double cell, int32_t* out_idx, float* out_d2)
{
    if (k < 1 || k > 64) return -1;
    ogrid g; int rc = ogrid_build(&g, xyz, n, cell > 0 ? cell : auto_cell(xyz, n, 3.0));
    if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        nb_t best[64];
        int cnt = ogrid_knn(&g, i, k, best);
        for (int t = 0; t < k; ++t) {
            out_idx[i * k + t] = t < cnt ? best[t].idx : -1;
            out_d2[i * k + t] = t < cnt ? best[t].d2 : INFINITY;
        }
    }
    ogrid_free(&g);
    return 0;
}

/* F2': include/impl/KeypointLearning.hpp:130-137, NormalEstimation.setRadiusSearch(search_radius_):
 * PCA over the whole ball (query included).  PCL accumulates the un-centred FP32 moments in the order its
 * sorted search returns, ascending (d2, index): order 0.  Sorting ~2500 neighbours per point is not
 * something a GPU should do, so the device path accumulates in the canonical (cell key, index) order it
 * already uses for the histograms: order 1 (cpr = cells per radius of the canonical grid).  The two differ
 * only by FP32 re-association of the moment sums (tests/test_oracle.py bounds the angle between them). */
KPLO_API int kplo_canon_grid(const float* xyz, int64_t n, double r_feat, int cpr, double org[3], double* cell, int32_t dims[3]);
KPLO_API void kplo_canon_keys(const float* xyz, int64_t n, const double org[3], double cell, const int32_t dims[3], int64_t* keys);
typedef struct { int64_t key; nb_t nb; } knb_t;
static int cmp_key_idx(const void* a, const void* b)
{
    const knb_t *x = (const knb_t*)a, *y = (const knb_t*)b;
    if (x->key != y->key) return (x->key > y->key) - (x->key < y->key);
    return (x->nb.idx > y->nb.idx) - (x->nb.idx < y->nb.idx);
}
KPLO_API int kplo_normals_radius_ordered(const float* xyz, int64_t n, double radius, const float vp[3], int order, int cpr, float* normals4)
{
    ogrid g; int rc = ogrid_build(&g, xyz, n, radius / 2.0);
    if (rc) return rc;
    int64_t* ckey = NULL;
    if (order == 1) {
        double org[3], cell; int32_t dims[3];
        kplo_canon_grid(xyz, n, radius, cpr, org, &cell, dims);
        ckey = (int64_t*)malloc((size_t)n * sizeof(int64_t));
        if (!ckey) { ogrid_free(&g); return -3; }
        kplo_canon_keys(xyz, n, org, cell, dims, ckey);
    }
    float r2 = (float)(radius * radius);
    int err = 0;
#pragma omp parallel
    {
        nbvec nb = { 0, 0, 0 };
        knb_t* kb = NULL; int64_t kcap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            if (ogrid_radius(&g, i, radius, r2, &nb)) { err = 1; continue; }
            if (order == 1) {
                if (nb.n > kcap) { free(kb); kcap = nb.n * 2; kb = (knb_t*)malloc((size_t)kcap * sizeof(knb_t)); }
                if (!kb) { err = 1; kcap = 0; continue; }
                for (int64_t t = 0; t < nb.n; ++t) { kb[t].key = ckey[nb.v[t].idx]; kb[t].nb = nb.v[t]; }
                qsort(kb, (size_t)nb.n, sizeof(knb_t), cmp_key_idx);
                for (int64_t t = 0; t < nb.n; ++t) nb.v[t] = kb[t].nb;
            } else qsort(nb.v, (size_t)nb.n, sizeof(nb_t), cmp_d2_idx);
            point_normal(xyz, nb.v, (int)nb.n, xyz + 3 * i, vp, normals4 + 4 * i);
        }
        free(nb.v); free(kb);
    }
    free(ckey); ogrid_free(&g);
    return err ? -3 : 0;
}
KPLO_API int kplo_normals_radius(const float* xyz, int64_t n, double radius, const float vp[3], float* normals4)
{
    return kplo_normals_radius_ordered(xyz, n, radius, vp, 0, 4, normals4);
}

/* ------------------------------------------------------------------------------------------ */
/* Canonical grid (the accumulation-order definition shared with the GPU path, restated here):  */
/*   cell = r_feat*(1+2^-20)/cells_per_radius (double); origin = per-axis float min as double;  */
/*   coord = clamp(floor(((double)v - origin)/cell), 0, dim-1); dim = floor((max-origin)/cell)+1*/
/*   key = (cz*dimy + cy)*dimx + cx                                                             */
/* ------------------------------------------------------------------------------------------ */
KPLO_API int kplo_canon_grid(const float* xyz, int64_t n, double r_feat, int cpr, double org[3], double* cell, int32_t dims[3])
{
    if (n <= 0 || cpr < 1) return -1;
    double lo[3] = { xyz[0], xyz[1], xyz[2] }, hi[3] = { xyz[0], xyz[1], xyz[2] };
    for (int64_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {
        double v = xyz[3 * i + a]; if (v < lo[a]) lo[a] = v; if (v > hi[a]) hi[a] = v; }
    double c = r_feat * (1.0 + 9.5367431640625e-07) / (double)cpr;
    *cell = c;
    for (int a = 0; a < 3; ++a) { org[a] = lo[a]; dims[a] = (int32_t)floor((hi[a] - lo[a]) / c) + 1; }
    return 0;
}
KPLO_API void kplo_canon_keys(const float* xyz, int64_t n, const double org[3], double cell, const int32_t dims[3], int64_t* keys)
{
    for (int64_t i = 0; i < n; ++i) {
        int64_t c[3];
        for (int a = 0; a < 3; ++a) {
            c[a] = (int64_t)floor(((double)xyz[3 * i + a] - org[a]) / cell);
            if (c[a] < 0) c[a] = 0;
            if (c[a] >= dims[a]) c[a] = dims[a] - 1;
        }
        keys[i] = (c[2] * dims[1] + c[1]) * dims[0] + c[0];
    }
}


/* ------------------------------------------------------------------------------------------ */
/* F4: include/impl/KeypointLearning.hpp:321-376.  normals4 = (nx,ny,nz,*).  The query's own     */
/* index is dropped ("skip slot 0" = author intent); neighbours whose normal is    */
/* non-finite are skipped (:338).  Output m x (A*B), index a*B + b (:366-369).                   */
/* canon_* may be NULL when order != 1 (then derived from the cloud when order == 1).            */
/* ------------------------------------------------------------------------------------------ */
/* Eigen's `row.normalize()` (hpp:360-365): Eigen >= 3.3 divides by the norm, Eigen 3.2.x multiplies by 1/norm. */
static int g_normalize_reciprocal = 0;
KPLO_API void kplo_set_normalize_mode(int reciprocal) { g_normalize_reciprocal = reciprocal; }

static void feature_row(const float* xyz, const float* normals4, int64_t q, nb_t* nb, int64_t cnt,
                        float support, int A, int B, float* hist /* A*B */, float* out)
{
    (void)xyz;
    for (int t = 0; t < A * B; ++t) hist[t] = 0.0f;
    const float* ni = normals4 + 4 * q;
    if (!(isfinite(ni[0]) && isfinite(ni[1]) && isfinite(ni[2]))) {
        /* runForest never calls computePointFeatures for such a point (hpp:277): zero row, no score */
        for (int t = 0; t < A * B; ++t) out[t] = 0.0f;
        return;
    }
    for (int64_t t = 0; t < cnt; ++t) {
        int64_t j = nb[t].idx;
        if (j == q) continue;
        const float* nj = normals4 + 4 * j;
        if (!(isfinite(nj[0]) && isfinite(nj[1]) && isfinite(nj[2]))) continue;
        float dot = ni[0] * nj[0] + (ni[1] * nj[1] + ni[2] * nj[2]);
        float cosine = 1.0f - dot;
        int a, ap, b, bp; float wa, wb;
        kplo_find_annulus_pair(A, sqrtf(nb[t].d2), support, &a, &ap, &wa);
        kplo_find_bin_pair(B, cosine, &b, &bp, &wb);
        hist[a * B + b] += ((1.0f - wb) * (1.0f - wa));
        hist[a * B + bp] += ((wb) * (1.0f - wa));
        hist[ap * B + b] += ((1.0f - wb) * (wa));
        hist[ap * B + bp] += ((wb) * (wa));
    }
    for (int a = 0; a < A; ++a) {
        float ss = 0.0f;
        for (int b = 0; b < B; ++b) ss += hist[a * B + b] * hist[a * B + b];
        float norm = sqrtf(ss);
        const float inv = 1.0f / norm;
        for (int b = 0; b < B; ++b)
            out[a * B + b] = (norm > 0) ? (g_normalize_reciprocal ? hist[a * B + b] * inv : hist[a * B + b] / norm) : hist[a * B + b];
    }
}

KPLO_API int kplo_features(const float* xyz, const float* normals4, int64_t n, double r_feat, int A, int B,
                           int order, int cpr, const double* canon_org, double canon_cell, const int32_t* canon_dims,
                           const int32_t* qidx, int64_t m, float* out)
{
    if (A < 1 || B < 1 || A * B > 4096) return -1;
    ogrid g; int rc = ogrid_build(&g, xyz, n, r_feat / 2.0);
    if (rc) return rc;
    int64_t* ckey = NULL;
    if (order == 1) {
        double org[3], cell; int32_t dims[3];
        if (canon_org && canon_dims && canon_cell > 0) { memcpy(org, canon_org, sizeof org); cell = canon_cell; memcpy(dims, canon_dims, sizeof dims); }
        else kplo_canon_grid(xyz, n, r_feat, cpr, org, &cell, dims);
        ckey = (int64_t*)malloc((size_t)n * sizeof(int64_t));
        if (!ckey) { ogrid_free(&g); return -3; }
        kplo_canon_keys(xyz, n, org, cell, dims, ckey);
    }
    float r2 = (float)(r_feat * r_feat);
    float support = (float)r_feat;
    int err = 0;
#pragma omp parallel
    {
        nbvec nb = { 0, 0, 0 };
        knb_t* kb = NULL; int64_t kcap = 0;
        float* hist = (float*)malloc((size_t)A * B * sizeof(float));
#pragma omp for schedule(dynamic, 32)
        for (int64_t k = 0; k < m; ++k) {
            int64_t q = qidx ? qidx[k] : k;
            if (!hist || ogrid_radius(&g, q, r_feat, r2, &nb)) { err = 1; continue; }
            if (order == 0) qsort(nb.v, (size_t)nb.n, sizeof(nb_t), cmp_idx);
            else if (order == 1) {
                if (nb.n > kcap) { free(kb); kcap = nb.n * 2; kb = (knb_t*)malloc((size_t)kcap * sizeof(knb_t)); }
                if (!kb) { err = 1; kcap = 0; continue; }
                for (int64_t t = 0; t < nb.n; ++t) { kb[t].key = ckey[nb.v[t].idx]; kb[t].nb = nb.v[t]; }
                qsort(kb, (size_t)nb.n, sizeof(knb_t), cmp_key_idx);
                for (int64_t t = 0; t < nb.n; ++t) nb.v[t] = kb[t].nb;
            }
            feature_row(xyz, normals4, q, nb.v, nb.n, support, A, B, hist, out + k * (int64_t)(A * B));
        }
        free(nb.v); free(kb); free(hist);
    }
    free(ckey); ogrid_free(&g);
    return err ? -3 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* F2' (organized clouds): pcl::IntegralImageNormalEstimation with SIMPLE_3D_GRADIENT and          */
/* setNormalSmoothingSize(5.0), the branch KeypointLearningDetector::initCompute takes for an        */
/* organized surface (impl/KeypointLearning.hpp:138-145).  [3P-recalled] PCL 1.8.0                   */
/* features/impl/integral_image_normal.hpp (computeFeature, computeFeatureFull with                  */
/* BORDER_POLICY_IGNORE and no depth-dependent smoothing, computePointNormal) and                    */
/* features/impl/integral_image2D.hpp (IntegralImage2D<float,3>: double sums, NaN elements skipped). */
/* xyz: height x width points (row-major, 3 floats each, NaN = no measurement).                      */
/* normals4: (nx, ny, nz, curvature = NaN) per pixel, NaN where PCL leaves the normal undefined.     */
/* The out-of-row reads of PCL's two distance-map passes (previous_row[ci + 1] at the last column,   */
/* next_row[ci - 1] at the first) are kept: they are reads of the neighbouring row inside the array. */
/* ------------------------------------------------------------------------------------------ */
KPLO_API int kplo_normals_integral_image(const float* xyz, int width, int height, float smoothing_size, const float vp[3], float* normals4)
{
    if (width < 1 || height < 1) return -1;
    const int64_t n = (int64_t)width * height;
    const int W = width, H = height;
    const float max_depth_change_factor = 20.0f * 0.001f;                 /* constructor default, never changed by the reference */
    unsigned char* change = (unsigned char*)malloc((size_t)n);
    float* dist = (float*)malloc((size_t)n * sizeof(float));
    double* I = (double*)calloc((size_t)(W + 1) * (H + 1) * 3, sizeof(double));
    if (!change || !dist || !I) { free(change); free(dist); free(I); return -3; }
    memset(change, 255, (size_t)n);
    for (int ri = 0; ri < H - 1; ++ri)
        for (int ci = 0; ci < W - 1; ++ci) {
            const int64_t index = (int64_t)ri * W + ci;
            const float depth = xyz[3 * index + 2], depthR = xyz[3 * (index + 1) + 2], depthD = xyz[3 * (index + W) + 2];
            const float lim = (max_depth_change_factor * (fabsf(depth) + 1.0f) * 2.0f);
            if (fabsf(depth - depthR) > lim || !isfinite(depth) || !isfinite(depthR)) { change[index] = 0; change[index + 1] = 0; }
            if (fabsf(depth - depthD) > lim || !isfinite(depth) || !isfinite(depthD)) { change[index] = 0; change[index + W] = 0; }
        }
    for (int64_t i = 0; i < n; ++i) dist[i] = change[i] == 0 ? 0.0f : (float)(W + H);
    for (int ri = 1; ri < H; ++ri) {                                       /* first pass */
        const float* prev = dist + (int64_t)(ri - 1) * W;
        float* cur = dist + (int64_t)ri * W;
        for (int ci = 1; ci < W; ++ci) {
            const float upLeft = prev[ci - 1] + 1.4f, up = prev[ci] + 1.0f, upRight = prev[ci + 1] + 1.4f, left = cur[ci - 1] + 1.0f;
            const float a = upLeft < up ? upLeft : up, b = left < upRight ? left : upRight;   /* std::min(std::min(upLeft, up), std::min(left, upRight)) */
            const float m = a < b ? a : b;
            if (m < cur[ci]) cur[ci] = m;
        }
    }
    for (int ri = H - 2; ri >= 0; --ri) {                                  /* second pass */
        const float* next = dist + (int64_t)(ri + 1) * W;
        float* cur = dist + (int64_t)ri * W;
        for (int ci = W - 2; ci >= 0; --ci) {
            const float lowerLeft = next[ci - 1] + 1.4f, lower = next[ci] + 1.0f, lowerRight = next[ci + 1] + 1.4f, right = cur[ci + 1] + 1.0f;
            const float a = lowerLeft < lower ? lowerLeft : lower, b = right < lowerRight ? right : lowerRight;
            const float m = a < b ? a : b;
            if (m < cur[ci]) cur[ci] = m;
        }
    }
    /* IntegralImage2D<float,3>::computeIntegralImages, first order only */
    const int S = W + 1;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const float* e = xyz + 3 * ((int64_t)r * W + c);
            const int fin = isfinite(e[0] + (e[1] + e[2]));                /* pcl_isfinite(element->sum()) */
            for (int a = 0; a < 3; ++a) {
                double v = I[((int64_t)r * S + (c + 1)) * 3 + a] + I[((int64_t)(r + 1) * S + c) * 3 + a] - I[((int64_t)r * S + c) * 3 + a];
                if (fin) v += (double)e[a];
                I[((int64_t)(r + 1) * S + (c + 1)) * 3 + a] = v;
            }
        }
    for (int64_t i = 0; i < 4 * n; ++i) normals4[i] = NAN;
    const int border = (int)smoothing_size;
    for (int ri = border; ri < H - border; ++ri)
        for (int ci = border; ci < W - border; ++ci) {
            const int64_t index = (int64_t)ri * W + ci;
            if (!isfinite(xyz[3 * index + 2])) continue;
            const float smoothing = dist[index] < smoothing_size ? dist[index] : smoothing_size;
            if (!(smoothing > 2.0f)) continue;
            const int rw = (int)smoothing, rh = (int)smoothing, rw2 = rw / 2, rh2 = rh / 2;
            double gx[3], gy[3];
#define KPLO_II_SUM(sx, sy, w, h, a) (I[((int64_t)((sy) + (h)) * S + (sx) + (w)) * 3 + (a)] + I[((int64_t)(sy) * S + (sx)) * 3 + (a)] - \
                                      I[((int64_t)(sy) * S + (sx) + (w)) * 3 + (a)] - I[((int64_t)((sy) + (h)) * S + (sx)) * 3 + (a)])
            for (int a = 0; a < 3; ++a) {
                gx[a] = KPLO_II_SUM(ci + rw2, ri - rh2, 1, rh, a) - KPLO_II_SUM(ci - rw2, ri - rh2, 1, rh, a);
                gy[a] = KPLO_II_SUM(ci - rw2, ri + rh2, rw, 1, a) - KPLO_II_SUM(ci - rw2, ri - rh2, rw, 1, a);
            }
#undef KPLO_II_SUM
            double nv[3] = { gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0] };   /* gradient_y.cross(gradient_x) */
            const double len = nv[0] * nv[0] + (nv[1] * nv[1] + nv[2] * nv[2]);
            if (len == 0.0) continue;
            const double s = sqrt(len);
            float nx = (float)(nv[0] / s), ny = (float)(nv[1] / s), nz = (float)(nv[2] / s);
            const float vx = vp[0] - xyz[3 * index], vy = vp[1] - xyz[3 * index + 1], vz = vp[2] - xyz[3 * index + 2];
            const float cos_theta = (vx * nx + vy * ny + vz * nz);
            if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
            normals4[4 * index] = nx; normals4[4 * index + 1] = ny; normals4[4 * index + 2] = nz;   /* curvature stays NaN (bad_point) */
        }
    free(change); free(dist); free(I);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* F7: OpenCV DTreesImpl::predictTrees(PREDICT_SUM) on flat arrays + KeypointLearning.hpp:287    */
/* var[i] < 0 => leaf.  go left iff x[var] <= thr.  sum in double, returned as float.            */
/* ------------------------------------------------------------------------------------------ */
KPLO_API void kplo_forest_sum(const int32_t* roots, int ntrees, const int32_t* var, const float* thr,
                              const int32_t* left, const int32_t* right, const float* value,
                              const float* feat, int64_t m, int F, float* sums)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; ++i) {
        const float* x = feat + i * (int64_t)F;
        double sum = 0;
        for (int t = 0; t < ntrees; ++t) {
            int32_t nidx = roots[t];
            while (var[nidx] >= 0) nidx = (x[var[nidx]] <= thr[nidx]) ? left[nidx] : right[nidx];
            sum += (double)value[nidx];
        }
        sums[i] = (float)sum;
    }
}
/* "Fragile split" report: per row, 1 if any split DECIDED while forest_->predict walks the trees
 * (hpp:281) had |x[var] - thr| <= eps in FP32 -- a feature difference of the tolerated size would send that tree
 * the other way.  Returns the number of flagged rows. */
KPLO_API int64_t kplo_forest_fragile(const int32_t* roots, int ntrees, const int32_t* var, const float* thr,
                                     const int32_t* left, const int32_t* right,
                                     const float* feat, int64_t m, int F, float eps, uint8_t* flags)
{
    int64_t cnt = 0;
#pragma omp parallel for schedule(static) reduction(+ : cnt)
    for (int64_t i = 0; i < m; ++i) {
        const float* x = feat + i * (int64_t)F;
        int frag = 0;
        for (int t = 0; t < ntrees; ++t) {
            int32_t nidx = roots[t];
            while (var[nidx] >= 0) {
                const float d = x[var[nidx]] - thr[nidx];
                if (fabsf(d) <= eps) frag = 1;
                nidx = (x[var[nidx]] <= thr[nidx]) ? left[nidx] : right[nidx];
            }
        }
        flags[i] = (uint8_t)frag;
        cnt += frag;
    }
    return cnt;
}
KPLO_API void kplo_scores(const float* sums, int64_t m, int ntrees, float* scores)
{
    for (int64_t i = 0; i < m; ++i) scores[i] = 1 - (sums[i] / ((float)ntrees * 1.0f));
}
/* runForest skips points whose normal is not finite (hpp:277): they get no score.  (The reference then
 * mis-aligns response indices; the restatement keeps the alignment and marks the score NaN.) */
KPLO_API int64_t kplo_mask_unscored(const float* normals4, int64_t n, float* scores)
{
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* v = normals4 + 4 * i;
        if (!(isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]))) { scores[i] = NAN; cnt++; }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* F8: include/impl/KeypointLearning.hpp:202-256 with non_maxima_draws_remove_ == false:        */
/* keypoint iff score >= th (compare in double, :207) and no point with d2 < r_nms^2 has a       */
/* strictly larger score (:219).  Output ascending index.  Returns the count.                   */
/* ------------------------------------------------------------------------------------------ */
KPLO_API int64_t kplo_nms(const float* xyz, const float* scores, int64_t n, double r_nms, double th, int32_t* kp_idx)
{
    ogrid g; int rc = ogrid_build(&g, xyz, n, r_nms > 0 ? r_nms : 1.0);
    if (rc) return rc;
    float r2 = (float)(r_nms * r_nms);
    uint8_t* flag = (uint8_t*)calloc((size_t)n, 1);
    if (!flag) { ogrid_free(&g); return -3; }
#pragma omp parallel
    {
        nbvec nb = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            if (!isfinite(scores[i]) || (double)scores[i] < th) continue;
            if (ogrid_radius(&g, i, r_nms, r2, &nb)) continue;
            int is_max = 1;
            for (int64_t t = 0; t < nb.n; ++t) if (scores[i] < scores[nb.v[t].idx]) { is_max = 0; break; }
            flag[i] = (uint8_t)is_max;
        }
        free(nb.v);
    }
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) if (flag[i]) kp_idx[cnt++] = (int32_t)i;
    free(flag); ogrid_free(&g);
    return cnt;
}

/* draws-remove variant (:233-250), sequential skipList semantics. [N4] */
KPLO_API int64_t kplo_nms_draws(const float* xyz, const float* scores, int64_t n, double r_nms, double th,
                                float draws_thr, int32_t* kp_idx)
{
    ogrid g; int rc = ogrid_build(&g, xyz, n, r_nms > 0 ? r_nms : 1.0);
    if (rc) return rc;
    float r2 = (float)(r_nms * r_nms);
    uint8_t* skip = (uint8_t*)calloc((size_t)n, 1);
    nbvec nb = { 0, 0, 0 };
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!isfinite(scores[i]) || (double)scores[i] < th) continue;
        if (ogrid_radius(&g, i, r_nms, r2, &nb)) break;
        int is_max = 1, has_draw = 0;
        for (int64_t t = 0; t < nb.n; ++t) if (scores[i] < scores[nb.v[t].idx]) { is_max = 0; break; }
        if (!is_max) continue;
        for (int64_t t = 0; t < nb.n; ++t) if (nb.v[t].idx != i && scores[i] == scores[nb.v[t].idx]) has_draw = 1;
        if (has_draw) {
            if (skip[i]) continue;
            int survive = 0;
            for (int64_t t = 0; t < nb.n; ++t) {
                int32_t j = nb.v[t].idx;
                if (j == i || scores[i] != scores[j]) continue;
                float dx = xyz[3 * i] - xyz[3 * (int64_t)j], dy = xyz[3 * i + 1] - xyz[3 * (int64_t)j + 1], dz = xyz[3 * i + 2] - xyz[3 * (int64_t)j + 2];
                float distance = sqrtf(dx * dx + (dy * dy + dz * dz));
                if (distance < draws_thr) { survive = 1; skip[j] = 1; }
            }
            if (survive) kp_idx[cnt++] = (int32_t)i;
        } else kp_idx[cnt++] = (int32_t)i;
    }
    free(nb.v); free(skip); ogrid_free(&g);
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* Whole TestDetector pipeline (main_test_detector.cpp:162-187) for the CPU baseline timing.    */
/* normals_mode: 0 = given in normals4, 1 = kNN(k), 2 = radius(r_feat).  order as above.        */
/* stage_ms[5] = normals, features(search+hist), forest, nms, total  (wall clock, ms)           */
/* ------------------------------------------------------------------------------------------ */
static double now_ms(void)
{
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}
KPLO_API int64_t kplo_detect(const float* xyz, float* normals4, int64_t n, int normals_mode, int k, const float vp[3], int flip,
                             double r_feat, double r_nms, double th, int A, int B, int order, int cpr,
                             const int32_t* roots, int ntrees, const int32_t* var, const float* thr,
                             const int32_t* left, const int32_t* right, const float* value,
                             float* features /* n*A*B scratch/out */, float* scores, int32_t* kp_idx, double* stage_ms)
{
    double t0 = now_ms(), t1;
    int rc = 0;
    if (normals_mode == 1) rc = kplo_normals_knn(xyz, n, k, vp, 0.0, normals4);
    else if (normals_mode == 2) rc = kplo_normals_radius_ordered(xyz, n, r_feat, vp, order == 1 ? 1 : 0, cpr, normals4);
    if (rc) return rc;
    if (flip && normals_mode != 0) for (int64_t i = 0; i < n; ++i) { normals4[4 * i] *= -1; normals4[4 * i + 1] *= -1; normals4[4 * i + 2] *= -1; }
    t1 = now_ms(); if (stage_ms) stage_ms[0] = t1 - t0;
    rc = kplo_features(xyz, normals4, n, r_feat, A, B, order, cpr, NULL, 0.0, NULL, NULL, n, features);
    if (rc) return rc;
    double t2 = now_ms(); if (stage_ms) stage_ms[1] = t2 - t1;
    kplo_forest_sum(roots, ntrees, var, thr, left, right, value, features, n, A * B, scores);
    kplo_scores(scores, n, ntrees, scores);
    kplo_mask_unscored(normals4, n, scores);
    double t3 = now_ms(); if (stage_ms) stage_ms[2] = t3 - t2;
    int64_t cnt = kplo_nms(xyz, scores, n, r_nms, th, kp_idx);
    double t4 = now_ms(); if (stage_ms) { stage_ms[3] = t4 - t3; stage_ms[4] = t4 - t0; }
    return cnt;
}

KPLO_API int kplo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
KPLO_API void kplo_set_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
