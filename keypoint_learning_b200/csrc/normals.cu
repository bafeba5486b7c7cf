// normals.cu -- PCA normals on the cell-sorted cloud.
//
// kNN mode replaces pcl::NormalEstimation with setKSearch(10) over a sorted search::KdTree
// (reference call site src/main_test_detector.cpp:162-169): the k nearest points (the query
// itself included) ordered by (d2, index), PCL 1.8.0's single-pass un-centred FP32 moment sums in
// that order, eigen33, viewpoint flip.  One thread per point keeps the whole sequential
// accumulation in registers, which is what makes the result bit-identical to the oracle.
#include <cstdlib>
#include "kpl_internal.h"
#include "kpl_math.cuh"

namespace kpl {

__device__ __forceinline__ bool less_d2_idx(float d2, uint32_t idx, float bd2, uint32_t bidx)
{
    return d2 < bd2 || (d2 == bd2 && idx < bidx);
}

__device__ __forceinline__ void key_to_cell(uint32_t key, const GridDesc& g, int& cx, int& cy, int& cz)
{
    uint32_t t = key / (uint32_t)g.dim[0];
    cx = (int)(key - t * (uint32_t)g.dim[0]);
    cz = (int)(t / (uint32_t)g.dim[1]);
    cy = (int)(t - (uint32_t)cz * (uint32_t)g.dim[1]);
}


// Per-lane view of the grid: the whole grid for a single cloud, the point's own view (origin, z range) inside a
// stacked batch grid (kpl_internal.h: ViewDesc).
struct LocalGrid {
    double org[3];
    int zlo, zhi;      // z layers the search may visit
};
__device__ __forceinline__ LocalGrid local_grid(const GridDesc& g, int cz)
{
    LocalGrid L;
    L.org[0] = g.org[0]; L.org[1] = g.org[1]; L.org[2] = g.org[2];
    L.zlo = 0; L.zhi = g.dim[2] - 1;
    if (g.layer_view) {
        const int v = __ldg(g.layer_view + cz);
        if (v >= 0) {
            const ViewDesc* V = g.views + v;
            L.org[0] = V->org[0]; L.org[1] = V->org[1]; L.org[2] = V->org[2];
            L.zlo = V->zoff; L.zhi = V->zoff + V->dimz - 1;
        }
    }
    return L;
}
// fractional position of coordinate v inside its cell cc along axis a, in [0, 1)
__device__ __forceinline__ double cell_fraction(const GridDesc& g, const LocalGrid& L, int a, float v, int cc)
{
    const int local = (a == 2) ? cc - L.zlo : cc;      // zlo = the view's first stacked layer (0 for a single cloud)
    return __ddiv_rn(__dsub_rn((double)v, L.org[a]), g.cell) - (double)g.off[a] - (double)local;
}
// Slab of a larger cloud: the k nearest points found inside the slab are the k nearest of the whole cloud only if
// the ball that holds them does not reach an x face behind which the cloud continues.  Points in the outermost
// guard_cells columns at such a face are expected to be clipped (nothing that is kept depends on their normals).
__device__ __forceinline__ bool knn_clipped(const GridDesc& g, int cx, float px, float kth_d2)
{
    bool clipped = false;
    if (g.interior_lo && cx >= g.guard_cells) {
        const double d = (double)px - (g.org[0] + (double)g.off[0] * g.cell);
        clipped |= !((double)kth_d2 < d * d * (1.0 - 1e-6));
    }
    if (g.interior_hi && cx < g.dim[0] - g.guard_cells) {
        const double d = (g.org[0] + (double)(g.off[0] + g.dim[0]) * g.cell) - (double)px;
        clipped |= !((double)kth_d2 < d * d * (1.0 - 1e-6));
    }
    return clipped;
}

// K > 0: compile-time k, candidates in registers.  K == 0: runtime k <= 64, candidates in local memory.
//
// Search order: the query's own cell, then those of the 26 surrounding cells whose box can still hold a
// point closer than the current k-th best (conservative lower bound on the distance to the cell), then --
// only if the k-th best is not provably inside the 3x3x3 block -- growing rings that rescan everything.
// The result is the exact k smallest candidates under the total order (d2, index), whatever the visiting
// order, so it equals the sorted-kd-tree result the reference consumes.
template <int K>
__global__ void __launch_bounds__(128) normals_knn_kernel(const float4* __restrict__ s_pos, const uint32_t* __restrict__ skey,
                                                          const int32_t* __restrict__ cell_start, const float4* __restrict__ xyz,
                                                          GridDesc g, int n, int k_rt, float vpx, float vpy, float vpz,
                                                          float4* __restrict__ s_nrm, unsigned long long* __restrict__ counters)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int KM = (K > 0) ? K : 64;
    const int k = (K > 0) ? K : k_rt;
    float bd2[KM];
    uint32_t bi[KM];
    const float4 p = s_pos[i];
    int cx, cy, cz;
    key_to_cell(skey[i], g, cx, cy, cz);
    const LocalGrid LG = local_grid(g, cz);
    int seen = 0;
#pragma unroll
    for (int t = 0; t < KM; ++t) { bd2[t] = CUDART_INF_F; bi[t] = 0xFFFFFFFFu; }

    auto scan = [&](int s, int e) {
        for (int j = s; j < e; ++j) {
            float4 c = __ldg(s_pos + j);
            float d2 = dist2(p.x, p.y, p.z, c.x, c.y, c.z);
            uint32_t oi = __float_as_uint(c.w);
            seen++;
            if (less_d2_idx(d2, oi, bd2[k - 1], bi[k - 1])) {
                if (K > 0) {
#pragma unroll
                    for (int t = KM - 1; t >= 0; --t) {
                        float pd = (t > 0) ? bd2[t > 0 ? t - 1 : 0] : -CUDART_INF_F;
                        uint32_t pi = (t > 0) ? bi[t > 0 ? t - 1 : 0] : 0u;
                        if (less_d2_idx(d2, oi, pd, pi)) { bd2[t] = pd; bi[t] = pi; }
                        else if (less_d2_idx(d2, oi, bd2[t], bi[t])) { bd2[t] = d2; bi[t] = oi; }
                    }
                } else {
                    int t = k - 1;
                    while (t > 0 && less_d2_idx(d2, oi, bd2[t - 1], bi[t - 1])) { bd2[t] = bd2[t - 1]; bi[t] = bi[t - 1]; --t; }
                    bd2[t] = d2; bi[t] = oi;
                }
            }
        }
    };

    // ---- ring 1 with per-cell culling ------------------------------------------------------------
    // lo[a] / hi[a]: squared lower bounds on the distance to the neighbouring cell on the low / high side
    // of axis a, shrunk by 1e-5 so that rounding in the cell assignment or in d2 can never hide a candidate.
    float lo2[3], hi2[3];
    {
        const float v[3] = {p.x, p.y, p.z};
        const int cc[3] = {cx, cy, cz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double q = cell_fraction(g, LG, a, v[a], cc[a]);   // in [0, 1)
            const float dl = fmaxf((float)(q * g.cell) * 0.99999f - 1e-30f, 0.0f);
            const float dh = fmaxf((float)((1.0 - q) * g.cell) * 0.99999f - 1e-30f, 0.0f);
            lo2[a] = dl * dl * 0.99999f; hi2[a] = dh * dh * 0.99999f;
        }
    }
    {
        const int64_t base0 = ((int64_t)cz * g.dim[1] + cy) * g.dim[0];
        scan(__ldg(cell_start + base0 + cx), __ldg(cell_start + base0 + cx + 1));
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = cz + dz;
            if (z < LG.zlo || z > LG.zhi) continue;
            const float gz2 = dz < 0 ? lo2[2] : (dz > 0 ? hi2[2] : 0.0f);
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= g.dim[1]) continue;
                const float gyz2 = gz2 + (dy < 0 ? lo2[1] : (dy > 0 ? hi2[1] : 0.0f));
                if (!(gyz2 <= bd2[k - 1])) continue;
                const int l_ok = (cx > 0 && gyz2 + lo2[0] <= bd2[k - 1]) ? 1 : 0;
                const int r_ok = (cx < g.dim[0] - 1 && gyz2 + hi2[0] <= bd2[k - 1]) ? 1 : 0;
                const int64_t base = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                if (dz == 0 && dy == 0) {
                    if (l_ok) scan(__ldg(cell_start + base + cx - 1), __ldg(cell_start + base + cx));
                    if (r_ok) scan(__ldg(cell_start + base + cx + 1), __ldg(cell_start + base + cx + 2));
                } else {
                    scan(__ldg(cell_start + base + cx - l_ok), __ldg(cell_start + base + cx + r_ok + 1));
                }
            }
        }
    }
    // ---- exactness guard; growing rings (rescan from scratch) when the block was not enough ---------
    const int maxdim = max(g.dim[0], max(g.dim[1], g.dim[2]));
    for (int R = 1; R <= maxdim; ++R) {
        const int z0 = max(cz - R, LG.zlo), z1 = min(cz + R, LG.zhi);
        const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dim[1] - 1);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.dim[0] - 1);
        if (R > 1) {
#pragma unroll
            for (int t = 0; t < KM; ++t) { bd2[t] = CUDART_INF_F; bi[t] = 0xFFFFFFFFu; }
            seen = 0;
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const int64_t base = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                    scan(__ldg(cell_start + base + x0), __ldg(cell_start + base + x1 + 1));
                }
        }
        const bool all = (z0 == LG.zlo && y0 == 0 && x0 == 0 && z1 == LG.zhi && y1 == g.dim[1] - 1 && x1 == g.dim[0] - 1);
        if (all) break;
        if (bd2[k - 1] < CUDART_INF_F) {       // k candidates found (culling never skips a cell while fewer than k are known)
            double guard = (double)R * g.cell;
            guard = guard * guard * (1.0 - 1e-6);
            if ((double)bd2[k - 1] < guard) break;
        }
    }
    // fewer than k points in the whole cloud: count the filled slots
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < KM; ++t) cnt += (t < k && bi[t] != 0xFFFFFFFFu) ? 1 : 0;
    (void)seen;
    if ((g.interior_lo | g.interior_hi) && knn_clipped(g, cx, p.x, bd2[k - 1])) atomicAdd(counters + 9, 1ull);
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < KM; ++t) {
        if (t < cnt) {
            float4 c = __ldg(xyz + bi[t]);
            accu[0] = __fadd_rn(accu[0], __fmul_rn(c.x, c.x));
            accu[1] = __fadd_rn(accu[1], __fmul_rn(c.x, c.y));
            accu[2] = __fadd_rn(accu[2], __fmul_rn(c.x, c.z));
            accu[3] = __fadd_rn(accu[3], __fmul_rn(c.y, c.y));
            accu[4] = __fadd_rn(accu[4], __fmul_rn(c.y, c.z));
            accu[5] = __fadd_rn(accu[5], __fmul_rn(c.z, c.z));
            accu[6] = __fadd_rn(accu[6], c.x);
            accu[7] = __fadd_rn(accu[7], c.y);
            accu[8] = __fadd_rn(accu[8], c.z);
        }
    }
    s_nrm[i] = normal_from_moments(accu, cnt, p.x, p.y, p.z, vpx, vpy, vpz);
}

// Branch-free insertion into a list kept sorted by the 64-bit key (d2 bits, index): for d2 >= 0 the IEEE bit
// pattern orders like the value, so the unsigned key order IS the (d2, index) order of the sorted search.
// Every slot is rewritten with selects; a candidate that is not smaller than the last slot changes nothing.
template <int K>
__device__ __forceinline__ void knn_insert(uint32_t (&hi)[K], uint32_t (&lo)[K], uint32_t chi, uint32_t clo)
{
    bool lt_prev = false;
    uint32_t phi = 0u, plo = 0u;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        const uint32_t ohi = hi[t], olo = lo[t];
        const bool lt = chi < ohi || (chi == ohi && clo < olo);
        hi[t] = lt_prev ? phi : (lt ? chi : ohi);
        lo[t] = lt_prev ? plo : (lt ? clo : olo);
        lt_prev = lt; phi = ohi; plo = olo;
    }
}

// Warp-cooperative form of the k = 10 search (the TestDetector setting).  The per-thread kernel above spends
// most of its issue slots on idle lanes (ncu: 7.75 of 32 threads active per instruction) because every
// lane walks its own cell ranges.  Here a warp owns 32 consecutive points of the Hilbert order of all points
// (grid.cu: build_lists -- a compact patch of the surface, a few millimetres across where the cells are 5 mm);
// the lanes within a cell of each other form a group whose candidate rows -- the cell rows of the group plus
// one on every side -- are staged 32 candidates at a time in shared memory.  A packed-FP32 pass marks the
// candidates that can still enter a lane's list (d2 <= its current k-th best); only those are inserted, in the
// exact (d2, index) order.  Rows, and the cells at the two ends of a row, that no member lane can still use are
// skipped warp-uniformly: the k-th best of a lane only shrinks, so what is skipped could never have entered its
// list.  Lanes whose k-th best is not provably inside their 3 x 3 x 3 block (sparse regions) fall back to the
// growing-ring scan of the per-thread kernel.
template <int K>
__global__ void __launch_bounds__(32) normals_knn_coop_kernel(const float4* __restrict__ s_pos, const uint32_t* __restrict__ skey,
                                                              const int32_t* __restrict__ cell_start, const float4* __restrict__ xyz,
                                                              GridDesc g, const int32_t* __restrict__ qorder, const int32_t* __restrict__ warp_starts,
                                                              uint64_t one2, float ball_factor, float vpx, float vpy, float vpz, float4* __restrict__ s_nrm,
                                                              unsigned long long* __restrict__ counters)
{
    __shared__ __align__(16) float tile[128];
    float* sx = tile; float* sy = tile + 32; float* sz = tile + 64;
    uint32_t* si = reinterpret_cast<uint32_t*>(tile + 96);
    const int lane = threadIdx.x;
    const int q0 = __ldg(warp_starts + blockIdx.x);           // entries [q0, q0 + count) of the order, count <= 32
    const bool active = lane < __ldg(warp_starts + blockIdx.x + 1) - q0;
    const int q = active ? __ldg(qorder + q0 + lane) : 0;
    float4 p = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    if (active) { p = __ldg(s_pos + q); key_to_cell(__ldg(skey + q), g, cx, cy, cz); }
    uint32_t bd2[K];          // bit patterns of the k smallest squared distances, ascending by (d2, index)
    uint32_t bi[K];
    const uint64_t QY = pack2(p.y, p.y), QZ = pack2(p.z, p.z);
    // First attempt with a BOUNDED list: most of the kernel's time went into list insertions (ncu: 65 % of the
    // instructions) because a list that starts at +inf takes the first 32 candidates whole and ~50 more while it
    // converges.  The lane's own cell tells the local density: with n_c points in a cell the surface crosses, a ball of
    // r0^2 = ball_factor cell^2 / n_c (ball_factor = 24 / pi) holds ~24 of them.  The list starts full of sentinels (r0^2, no index), so only
    // candidates inside that ball are ever inserted; if fewer than K turn up, the lane repeats the search unbounded.
    // Either way the list ends as the K smallest (d2, index) of everything scanned.
    float r0sq = CUDART_INF_F;
    if (active) {
        const int64_t cell = ((int64_t)cz * g.dim[1] + cy) * g.dim[0] + cx;
        const int nc = __ldg(cell_start + cell + 1) - __ldg(cell_start + cell);
        const float guess = ball_factor * (float)g.cell * (float)g.cell / (float)max(nc, 1);
        if (nc >= 8 && guess < 0.9f * (float)g.cell * (float)g.cell) r0sq = guess;     // else: unbounded from the start
    }
#pragma unroll
    for (int t = 0; t < K; ++t) { bd2[t] = __float_as_uint(r0sq); bi[t] = 0xFFFFFFFFu; }   // bound (or +inf), no index

    // lower bounds on the distance to the faces of the lane's own cell on the low / high side of every axis
    // (shrunk by 1e-5: rounding in the cell assignment or in d2 can never hide a candidate)
    float lo1[3], hi1[3];
    const float cellf = (float)g.cell * 0.99999f;
    {
        // every lane of a warp lies in ONE view of a batch (the list is cut where the curve jumps, and views are layers apart)
        const LocalGrid LGq = local_grid(g, cz);
        const float v[3] = {p.x, p.y, p.z};
        const int cc[3] = {cx, cy, cz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double f = cell_fraction(g, LGq, a, v[a], cc[a]);   // in [0, 1)
            lo1[a] = active ? fmaxf((float)(f * g.cell) * 0.99999f - 1e-30f, 0.0f) : CUDART_INF_F;
            hi1[a] = active ? fmaxf((float)((1.0 - f) * g.cell) * 0.99999f - 1e-30f, 0.0f) : CUDART_INF_F;
        }
    }
    // squared lower bound on the distance between this lane's point and the cells at offset d (in cells) along axis a
    auto gap_sq = [&](int a, int d) -> float {
        if (d == 0) return 0.0f;
        const float t = (d < 0 ? lo1[a] : hi1[a]) + (float)(abs(d) - 1) * cellf;
        return t * t * 0.99999f;
    };

    bool searching = active;          // lanes taking part in the current attempt
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned remaining = __ballot_sync(0xFFFFFFFFu, searching);
    while (remaining) {
        // a group: the lanes within one cell of the first remaining lane whose cells span at most two per axis
        const int leader = __ffs(remaining) - 1;
        const int lx = __shfl_sync(0xFFFFFFFFu, cx, leader), ly = __shfl_sync(0xFFFFFFFFu, cy, leader), lz = __shfl_sync(0xFFFFFFFFu, cz, leader);
        const bool near = searching && ((remaining >> lane) & 1u) && (unsigned)(cx - lx + 1) <= 2u && (unsigned)(cy - ly + 1) <= 2u &&
                          (unsigned)(cz - lz + 1) <= 2u;
        const int gx0 = __reduce_min_sync(0xFFFFFFFFu, near ? cx : lx), gy0 = __reduce_min_sync(0xFFFFFFFFu, near ? cy : ly);
        const int gz0 = __reduce_min_sync(0xFFFFFFFFu, near ? cz : lz);
        const bool member = near && cx - gx0 <= 1 && cy - gy0 <= 1 && cz - gz0 <= 1;
        remaining &= ~__ballot_sync(0xFFFFFFFFu, member);
        const int gy1 = __reduce_max_sync(0xFFFFFFFFu, member ? cy : ly), gz1 = __reduce_max_sync(0xFFFFFFFFu, member ? cz : lz);
        const LocalGrid LG = local_grid(g, lz);
        const float px = member ? p.x : CUDART_NAN_F;
        const uint64_t QX = pack2(px, px);
        const int ny = gy1 - gy0 + 3, nrows = ny * (gz1 - gz0 + 3);
        // the group's own rows first (they tighten every lane's k-th best), then the rows around them
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
            for (int r = 0; r < nrows; ++r) {
                const int y = gy0 - 1 + r % ny, z = gz0 - 1 + r / ny;
                const bool own = y >= gy0 && y <= gy1 && z >= gz0 && z <= gz1;
                if (own != (pass == 0)) continue;
                if (y < 0 || y >= g.dim[1] || z < LG.zlo || z > LG.zhi) continue;
                const float worst0 = __uint_as_float(bd2[K - 1]);
                const float gap2 = gap_sq(1, y - cy) + gap_sq(2, z - cz);
                const bool use_row = member && gap2 <= worst0;
                if (!__any_sync(0xFFFFFFFFu, use_row)) continue;                           // no member can use this row
                // cells of the row some member can still use: its own, and the one on either side if close enough
                const int xa = max(__reduce_min_sync(0xFFFFFFFFu, use_row ? cx - ((gap2 + gap_sq(0, -1) <= worst0) ? 1 : 0) : 0x7FFFFFFF), 0);
                const int xb = min(__reduce_max_sync(0xFFFFFFFFu, use_row ? cx + ((gap2 + gap_sq(0, 1) <= worst0) ? 1 : 0) : -1), g.dim[0] - 1);
                const int64_t base = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                const int rs = __ldg(cell_start + base + xa), re = __ldg(cell_start + base + xb + 1);
                for (int tb = rs; tb < re; tb += 32) {
                    float4 c = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
                    if (tb + lane < re) c = __ldg(s_pos + tb + lane);
                    __syncwarp();
                    sx[lane] = c.x; sy[lane] = c.y; sz[lane] = c.z; si[lane] = __float_as_uint(c.w);
                    __syncwarp();
                    const int cnt = min(32, re - tb);
                    const float worst = __uint_as_float(bd2[K - 1]);
                    uint32_t mask = 0;
#pragma unroll
                    for (int k0 = 0; k0 < 32; k0 += 8) {
                        if (k0 < cnt) {
#pragma unroll
                            for (int u = 0; u < 8; u += 4) {
                                const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(sx + k0 + u);
                                const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(sy + k0 + u);
                                const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(sz + k0 + u);
                                float d0, d1, d2, d3;
                                unpack2(dist2_x2(QX, QY, QZ, X.x, Y.x, Z.x, one2), d0, d1);
                                unpack2(dist2_x2(QX, QY, QZ, X.y, Y.y, Z.y, one2), d2, d3);
                                if (d0 <= worst) mask |= 0x80000000u >> (k0 + u);
                                if (d1 <= worst) mask |= 0x80000000u >> (k0 + u + 1);
                                if (d2 <= worst) mask |= 0x80000000u >> (k0 + u + 2);
                                if (d3 <= worst) mask |= 0x80000000u >> (k0 + u + 3);
                            }
                        }
                    }
                    while (mask) {
                        const int m = 31 - __clz(mask);
                        mask ^= 1u << m;
                        const int k = 31 - m;
                        const float d2 = dist2(p.x, p.y, p.z, sx[k], sy[k], sz[k]);
                        knn_insert<K>(bd2, bi, __float_as_uint(d2), si[k]);
                    }
                }
            }
        }
    }
    // second attempt, unbounded, for the lanes whose ball held fewer than K points
    searching = active && attempt == 0 && bi[K - 1] == 0xFFFFFFFFu && bd2[K - 1] != 0x7F800000u;
    if (!__any_sync(0xFFFFFFFFu, searching)) break;
    if (searching) {
#pragma unroll
        for (int t = 0; t < K; ++t) { bd2[t] = 0x7F800000u; bi[t] = 0xFFFFFFFFu; }
    }
    }
    if (!active) return;
    const LocalGrid LG = local_grid(g, cz);
    // exactness guard of the 3 x 3 x 3 block; growing rings (per lane, rescanning) where it does not hold
    {
        const bool all1 = (cz - 1 <= LG.zlo && cy - 1 <= 0 && cx - 1 <= 0 && cz + 1 >= LG.zhi && cy + 1 >= g.dim[1] - 1 && cx + 1 >= g.dim[0] - 1);
        double guard = g.cell;
        guard = guard * guard * (1.0 - 1e-6);
        const bool exact = all1 || (bd2[K - 1] < 0x7F800000u && (double)__uint_as_float(bd2[K - 1]) < guard);
        if (!exact) {
            const int maxdim = max(g.dim[0], max(g.dim[1], g.dim[2]));
            for (int R = 2; R <= maxdim; ++R) {
                const int z0 = max(cz - R, LG.zlo), z1 = min(cz + R, LG.zhi);
                const int y0 = max(cy - R, 0), y1 = min(cy + R, g.dim[1] - 1);
                const int x0 = max(cx - R, 0), x1 = min(cx + R, g.dim[0] - 1);
#pragma unroll
                for (int t = 0; t < K; ++t) { bd2[t] = 0x7F800000u; bi[t] = 0xFFFFFFFFu; }
                for (int z = z0; z <= z1; ++z)
                    for (int y = y0; y <= y1; ++y) {
                        const int64_t base = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                        const int s = __ldg(cell_start + base + x0), e = __ldg(cell_start + base + x1 + 1);
                        for (int j = s; j < e; ++j) {
                            const float4 c = __ldg(s_pos + j);
                            const float d2 = dist2(p.x, p.y, p.z, c.x, c.y, c.z);
                            if (d2 <= __uint_as_float(bd2[K - 1])) knn_insert<K>(bd2, bi, __float_as_uint(d2), __float_as_uint(c.w));
                        }
                    }
                if (z0 == LG.zlo && y0 == 0 && x0 == 0 && z1 == LG.zhi && y1 == g.dim[1] - 1 && x1 == g.dim[0] - 1) break;
                if (bd2[K - 1] < 0x7F800000u) {
                    double gr = (double)R * g.cell;
                    gr = gr * gr * (1.0 - 1e-6);
                    if ((double)__uint_as_float(bd2[K - 1]) < gr) break;
                }
            }
        }
    }
    if ((g.interior_lo | g.interior_hi) && knn_clipped(g, cx, p.x, __uint_as_float(bd2[K - 1]))) atomicAdd(counters + 9, 1ull);
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < K; ++t) cnt += (bi[t] != 0xFFFFFFFFu) ? 1 : 0;
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < K; ++t) {
        if (t < cnt) {
            const float4 c = __ldg(xyz + bi[t]);
            accu[0] = __fadd_rn(accu[0], __fmul_rn(c.x, c.x));
            accu[1] = __fadd_rn(accu[1], __fmul_rn(c.x, c.y));
            accu[2] = __fadd_rn(accu[2], __fmul_rn(c.x, c.z));
            accu[3] = __fadd_rn(accu[3], __fmul_rn(c.y, c.y));
            accu[4] = __fadd_rn(accu[4], __fmul_rn(c.y, c.z));
            accu[5] = __fadd_rn(accu[5], __fmul_rn(c.z, c.z));
            accu[6] = __fadd_rn(accu[6], c.x);
            accu[7] = __fadd_rn(accu[7], c.y);
            accu[8] = __fadd_rn(accu[8], c.z);
        }
    }
    s_nrm[q] = normal_from_moments(accu, cnt, p.x, p.y, p.z, vpx, vpy, vpz);
}

bool normals_knn_uses_work_list(const kpl_params& P)
{
    bool coop = P.k_normals == 10;
#ifdef KPL_EXPERIMENTS
    if (getenv("KPL_NORMALS_PER_THREAD")) coop = false;
#endif
    return coop;
}

// The query order and warp list of the warp-cooperative kernel are built by the caller together with the feature
// kernel's (grid.cu: build_lists, one host synchronisation for both).
cudaError_t launch_normals_knn(kpl_ctx* c, int64_t n)
{
    const kpl_params& P = c->params;
    int blocks = (int)((n + 127) / 128);
    const float4* xyz = c->cur_xyz;
    if (normals_knn_uses_work_list(P)) {
        float ball_factor = 7.6f;           // 24 / pi: ~24 expected points in the first attempt's ball (normals_knn_coop_kernel;
                                            // 5.1 / 6.4 / 7.6 / 9.55 measured 8.23 / 6.91 / 6.72 / 6.93 ms on the 10 M scene)
#ifdef KPL_EXPERIMENTS
        if (const char* ev = getenv("KPL_KNN_BALL")) ball_factor = (float)atof(ev);
#endif
        if (c->nwarps_norm > 0)
            normals_knn_coop_kernel<10><<<(unsigned)c->nwarps_norm, 32, 0, c->stream>>>(c->s_pos.p, c->key_b.p, c->cell_start.p, xyz, c->grid,
                                                                                        c->qorder_all.p, c->warp_starts_n.p,
                                                                                        0x3F8000003F800000ull, ball_factor, P.viewpoint[0], P.viewpoint[1],
                                                                                        P.viewpoint[2], c->s_nrm.p, c->counters.p);
    }
    else if (P.k_normals == 10)
        normals_knn_kernel<10><<<blocks, 128, 0, c->stream>>>(c->s_pos.p, c->key_b.p, c->cell_start.p, xyz, c->grid, (int)n, 10,
                                                               P.viewpoint[0], P.viewpoint[1], P.viewpoint[2], c->s_nrm.p, c->counters.p);
    else
        normals_knn_kernel<0><<<blocks, 128, 0, c->stream>>>(c->s_pos.p, c->key_b.p, c->cell_start.p, xyz, c->grid, (int)n, P.k_normals,
                                                              P.viewpoint[0], P.viewpoint[1], P.viewpoint[2], c->s_nrm.p, c->counters.p);
    c->launches++;
    return cudaGetLastError();
}

// --flipNormals (src/main_test_detector.cpp:173-179): negate the three components.
__global__ void __launch_bounds__(256) flip_normals_kernel(float4* __restrict__ s_nrm, int64_t n)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = s_nrm[i];
    v.x = __fmul_rn(v.x, -1.0f); v.y = __fmul_rn(v.y, -1.0f); v.z = __fmul_rn(v.z, -1.0f);
    s_nrm[i] = v;
}
cudaError_t launch_flip_normals(kpl_ctx* c, int64_t n)
{
    flip_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->s_nrm.p, n);
    c->launches++;
    return cudaGetLastError();
}

// The reference skips points whose normal is not finite in runForest (impl/KeypointLearning.hpp:277)
// and thereby mis-aligns its response cloud against the indices NMS uses (:203-253).  We keep the
// alignment instead: such a point gets no score (NaN), is never a keypoint and never suppresses a
// neighbour; counters[4] counts them (kpl_stats.n_unscored).
__global__ void __launch_bounds__(256) check_normals_kernel(const float4* __restrict__ s_nrm, const uint8_t* __restrict__ s_role,
                                                            int64_t n, unsigned long long* __restrict__ counters)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n && (!s_role || (s_role[i] & 1))) {
        float4 v = s_nrm[i];
        bad = !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z));
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, bad);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(counters + 4, (unsigned long long)__popc(m));
}
cudaError_t launch_check_normals(kpl_ctx* c, int64_t n, bool use_role)
{
    check_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->s_nrm.p, use_role ? c->s_role.p : nullptr, n, c->counters.p);
    c->launches++;
    return cudaGetLastError();
}

// Radius mode replaces the fallback of KeypointLearningDetector::initCompute (impl/KeypointLearning.hpp:130-137):
// pcl::NormalEstimation with setRadiusSearch(search_radius_), i.e. PCA over the whole r_feat ball, query
// included.  Same traversal as the feature kernel (features.cu): a warp owns 32 consecutive sorted
// points, candidate rows are staged 32 at a time in shared memory, a packed-FP32 pass builds the exact
// membership mask and each lane then adds its neighbours' un-centred FP32 moments
// (computeMeanAndCovarianceMatrix, PCL 1.8.0) one by one in ascending sorted position.  PCL adds them in
// its sorted-search order (d2, index); reproducing that would mean sorting ~2500 neighbours per point, so
// the order here is the canonical (cell key, index) one -- the oracle's `order 1` -- and differs from
// PCL's by FP32 re-association only.
struct NormRadParams {
    int n, reach, span;
    float r2, cellf, rcull2, vpx, vpy, vpz;
    uint64_t one2;
};

__global__ void __launch_bounds__(32)
normals_radius_kernel(const float4* __restrict__ s_pos, const uint32_t* __restrict__ skey, const int32_t* __restrict__ cell_start,
                      const int2* __restrict__ work, int dimx, int dimy, int dimz, NormRadParams P, float4* __restrict__ s_nrm)
{
    __shared__ __align__(16) float tile[96];
    float* sx = tile; float* sy = tile + 32; float* sz = tile + 64;
    const int lane = threadIdx.x;
    const int2 item = __ldg(work + blockIdx.x);              // up to 32 consecutive sorted points of one run (grid.cu)
    const int q = item.x + lane;
    const bool active = lane < item.y;
    float4 qp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        qp = __ldg(s_pos + q);
        const uint32_t key = __ldg(skey + q);
        const uint32_t t = key / (uint32_t)dimx;
        cx = (int)(key - t * (uint32_t)dimx);
        cz = (int)(t / (uint32_t)dimy);
        cy = (int)(t - (uint32_t)cz * (uint32_t)dimy);
    }
    const uint64_t QY = pack2(qp.y, qp.y), QZ = pack2(qp.z, qp.z);
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int cnt_nb = 0;
    const float* sx31 = sx + 31;

    unsigned remaining = __ballot_sync(0xFFFFFFFFu, active);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int minx = __shfl_sync(0xFFFFFFFFu, cx, leader);
        const int gy0 = __shfl_sync(0xFFFFFFFFu, cy, leader), gz0 = __shfl_sync(0xFFFFFFFFu, cz, leader);
        const bool member = active && ((remaining >> lane) & 1u) && cy == gy0 && cz == gz0 && (unsigned)(cx - minx) <= (unsigned)P.span;
        remaining &= ~__ballot_sync(0xFFFFFFFFu, member);
        const int maxx = __reduce_max_sync(0xFFFFFFFFu, member ? cx : minx);
        const float px = member ? qp.x : CUDART_NAN_F;
        const uint64_t QX = pack2(px, px);
        const int y0 = max(gy0 - P.reach, 0), y1 = min(gy0 + P.reach, dimy - 1);
        const int z0 = max(gz0 - P.reach, 0), z1 = min(gz0 + P.reach, dimz - 1);
        const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);

        int rb = 0, nr = 0, l = 0, row_s = 0, row_e = 0, cur = 0, end = 0;
        auto next_tile = [&](int& tb, int& te) -> bool {
            while (cur >= end) {
                if (l >= nr) {
                    if (rb >= nrows) return false;
                    row_s = 0; row_e = 0;
                    const int r = rb + lane;
                    if (r < nrows) {
                        const int zz = z0 + r / ny, yy = y0 + r % ny;
                        const int gy = max(abs(yy - gy0) - 1, 0), gz = max(abs(zz - gz0) - 1, 0);
                        const float gap2 = (float)(gy * gy + gz * gz) * P.cellf * P.cellf;
                        if (gap2 < P.rcull2) {
                            int rx = (int)(sqrtf(P.rcull2 - gap2) / P.cellf) + 1;
                            rx = min(rx, P.reach);
                            const int xa = max(minx - rx, 0), xb = min(maxx + rx, dimx - 1);
                            const int64_t base = ((int64_t)zz * dimy + yy) * dimx;
                            row_s = __ldg(cell_start + base + xa);
                            row_e = __ldg(cell_start + base + xb + 1);
                        }
                    }
                    nr = min(32, nrows - rb);
                    rb += 32;
                    l = 0;
                }
                cur = __shfl_sync(0xFFFFFFFFu, row_s, l);
                end = __shfl_sync(0xFFFFFFFFu, row_e, l);
                ++l;
            }
            tb = cur; te = end;
            cur += 32;
            return true;
        };

        int tb = 0, te = 0;
        bool have = next_tile(tb, te);
        float4 cp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
        if (have && tb + lane < te) cp = __ldg(s_pos + tb + lane);
        while (have) {
            __syncwarp();
            sx[lane] = cp.x; sy[lane] = cp.y; sz[lane] = cp.z;
            const int cnt = min(32, te - tb);
            have = next_tile(tb, te);
            cp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
            if (have && tb + lane < te) cp = __ldg(s_pos + tb + lane);
            __syncwarp();
            uint32_t mask = 0;
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 8) {
                if (k0 < cnt) {
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(sx + k0 + u);
                        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(sy + k0 + u);
                        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(sz + k0 + u);
                        float d0, d1, d2, d3;
                        unpack2(dist2_x2(QX, QY, QZ, X.x, Y.x, Z.x, P.one2), d0, d1);
                        unpack2(dist2_x2(QX, QY, QZ, X.y, Y.y, Z.y, P.one2), d2, d3);
                        if (d0 < P.r2) mask |= 0x80000000u >> (k0 + u);
                        if (d1 < P.r2) mask |= 0x80000000u >> (k0 + u + 1);
                        if (d2 < P.r2) mask |= 0x80000000u >> (k0 + u + 2);
                        if (d3 < P.r2) mask |= 0x80000000u >> (k0 + u + 3);
                    }
                }
            }
            cnt_nb += __popc(mask);
            while (mask) {
                const int m = 31 - __clz(mask);
                mask ^= 1u << m;
                const float* t = sx31 - m;
                const float x = t[0], y = t[32], z = t[64];
                accu[0] = __fadd_rn(accu[0], __fmul_rn(x, x));
                accu[1] = __fadd_rn(accu[1], __fmul_rn(x, y));
                accu[2] = __fadd_rn(accu[2], __fmul_rn(x, z));
                accu[3] = __fadd_rn(accu[3], __fmul_rn(y, y));
                accu[4] = __fadd_rn(accu[4], __fmul_rn(y, z));
                accu[5] = __fadd_rn(accu[5], __fmul_rn(z, z));
                accu[6] = __fadd_rn(accu[6], x);
                accu[7] = __fadd_rn(accu[7], y);
                accu[8] = __fadd_rn(accu[8], z);
            }
        }
    }
    if (active) s_nrm[q] = normal_from_moments(accu, cnt_nb, qp.x, qp.y, qp.z, P.vpx, P.vpy, P.vpz);
}

cudaError_t launch_normals_radius(kpl_ctx* c, int64_t n)
{
    const kpl_params& U = c->params;
    NormRadParams P;
    const double r = (double)U.radius_features;
    P.n = (int)n; P.reach = c->grid.reach_feat; P.span = U.cells_per_radius;
    P.r2 = (float)(r * r);                         // static_cast<float>(radius*radius), KdTreeFLANN::radiusSearch
    P.cellf = (float)c->grid.cell;
    P.rcull2 = (float)(r * r * (1.0 + 1e-5));
    P.vpx = U.viewpoint[0]; P.vpy = U.viewpoint[1]; P.vpz = U.viewpoint[2];
    P.one2 = 0x3F8000003F800000ull;
    if (c->nwarps_norm > 0)
        normals_radius_kernel<<<(unsigned)c->nwarps_norm, 32, 0, c->stream>>>(c->s_pos.p, c->key_b.p, c->cell_start.p, c->work_n.p,
                                                                    c->grid.dim[0], c->grid.dim[1], c->grid.dim[2], P, c->s_nrm.p);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace kpl
