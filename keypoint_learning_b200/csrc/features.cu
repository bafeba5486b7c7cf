// features.cu -- radius search + annuli x bins histogram, the dominant kernel of the path.
// Replaces KeypointLearningDetector::computePointFeatures (impl/KeypointLearning.hpp:321-376) with
// its helpers findAnnulusPair / findBinPair (src/KeypointLearning.cpp:41-92) and the FLANN radius
// search behind searchForNeighbors (hpp:334).
//
// Mapping: one warp owns 32 consecutive cell-sorted query points, one lane per query.  The warp
// walks the cell rows that can hold neighbours of any of its queries; each row is ONE contiguous
// range of the sorted arrays (grid.cu), staged 32 candidates at a time into a per-warp shared
// tile with coalesced float4 loads.  Every lane then visits the 32 staged candidates in order
// (shared-memory broadcast reads), so each query sees its neighbours in ascending sorted position
// = canonical (cell key, index) order, and accumulates its votes sequentially in FP32 into a
// lane-private histogram column hist[cell][lane] (bank == lane: conflict-free, no atomics).
// That fixed order is what makes the histogram bit-identical to the oracle.
#include "kpl_internal.h"
#include "kpl_math.cuh"

namespace kpl {

static constexpr int FEAT_WARPS = 4;

struct FeatParams {
    int n, A, B, F, reach;
    float r2, support, adim, ahalf, bdim, bhalf, cellf, rcull2;
};

__device__ __forceinline__ void key_to_cell_f(uint32_t key, int dimx, int dimy, int& cx, int& cy, int& cz)
{
    uint32_t t = key / (uint32_t)dimx;
    cx = (int)(key - t * (uint32_t)dimx);
    cz = (int)(t / (uint32_t)dimy);
    cy = (int)(t - (uint32_t)cz * (uint32_t)dimy);
}

__global__ void __launch_bounds__(FEAT_WARPS * 32)
feature_kernel(const float4* __restrict__ s_pos, const float4* __restrict__ s_nrm, const uint32_t* __restrict__ skey,
               const int32_t* __restrict__ cell_start, const uint8_t* __restrict__ s_role,
               int dimx, int dimy, int dimz, FeatParams P, float* __restrict__ feat, unsigned long long* __restrict__ counters)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = P.F * 32 + 256;
    float* hist = smem + warp * per_warp;
    float4* tpos = reinterpret_cast<float4*>(hist + P.F * 32);
    float4* tnrm = tpos + 32;

    const int q0 = (blockIdx.x * FEAT_WARPS + warp) * 32;
    if (q0 >= P.n) return;
    const int q = q0 + lane;
    bool active = q < P.n;
    if (active && s_role) active = (s_role[q] & 1) != 0;
    if (!__any_sync(0xFFFFFFFFu, active)) {
        // nothing to score in this warp: rows stay zero
        int nvalid = min(32, P.n - q0);
        for (int e = lane; e < nvalid * P.F; e += 32) feat[(int64_t)q0 * P.F + e] = 0.0f;
        return;
    }
    float4 qp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f), qn = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        qp = __ldg(s_pos + q);
        qn = __ldg(s_nrm + q);
        key_to_cell_f(__ldg(skey + q), dimx, dimy, cx, cy, cz);
    }
    const uint32_t qidx = __float_as_uint(qp.w);
    const int BIG = 0x3FFFFFFF;
    const int minx = __reduce_min_sync(0xFFFFFFFFu, active ? cx : BIG), maxx = __reduce_max_sync(0xFFFFFFFFu, active ? cx : -BIG);
    const int miny = __reduce_min_sync(0xFFFFFFFFu, active ? cy : BIG), maxy = __reduce_max_sync(0xFFFFFFFFu, active ? cy : -BIG);
    const int minz = __reduce_min_sync(0xFFFFFFFFu, active ? cz : BIG), maxz = __reduce_max_sync(0xFFFFFFFFu, active ? cz : -BIG);

    for (int f = 0; f < P.F; ++f) hist[f * 32 + lane] = 0.0f;

    const int y0 = max(miny - P.reach, 0), y1 = min(maxy + P.reach, dimy - 1);
    const int z0 = max(minz - P.reach, 0), z1 = min(maxz + P.reach, dimz - 1);
    const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
    unsigned npairs = 0, ncand = 0;

    for (int rb = 0; rb < nrows; rb += 32) {
        // each lane resolves one cell row: [s, e) in the sorted arrays, after conservative culling
        int s = 0, e = 0;
        {
            int r = rb + lane;
            if (r < nrows) {
                int zz = z0 + r / ny, yy = y0 + r % ny;
                int gy = max(max(miny - yy, yy - maxy) - 1, 0), gz = max(max(minz - zz, zz - maxz) - 1, 0);
                float gap2 = (float)(gy * gy + gz * gz) * P.cellf * P.cellf;
                if (gap2 < P.rcull2) {
                    int rx = (int)(sqrtf(P.rcull2 - gap2) / P.cellf) + 1;
                    rx = min(rx, P.reach);
                    int xa = max(minx - rx, 0), xb = min(maxx + rx, dimx - 1);
                    int64_t base = ((int64_t)zz * dimy + yy) * dimx;
                    s = __ldg(cell_start + base + xa);
                    e = __ldg(cell_start + base + xb + 1);
                }
            }
        }
        const int nr = min(32, nrows - rb);
        for (int l = 0; l < nr; ++l) {
            const int sl = __shfl_sync(0xFFFFFFFFu, s, l), el = __shfl_sync(0xFFFFFFFFu, e, l);
            for (int base = sl; base < el; base += 32) {
                const int j = base + lane;
                float4 cp, cn;
                if (j < el) {
                    cp = __ldg(s_pos + j);
                    cn = __ldg(s_nrm + j);
                    // neighbours with a non-finite normal never vote (hpp:338): poison the position
                    if (!(isfinite(cn.x) && isfinite(cn.y) && isfinite(cn.z))) cp.x = CUDART_NAN_F;
                }
                __syncwarp();
                if (j < el) { tpos[lane] = cp; tnrm[lane] = cn; }
                __syncwarp();
                const int cnt = min(32, el - base);
                ncand += cnt;
                for (int k = 0; k < cnt; ++k) {
                    const float4 c = tpos[k];
                    const float d2 = dist2(qp.x, qp.y, qp.z, c.x, c.y, c.z);
                    if (d2 < P.r2 && __float_as_uint(c.w) != qidx) {
                        const float4 nj = tnrm[k];
                        float cosine = __fsub_rn(1.0f, dot3_eigen(qn.x, qn.y, qn.z, nj.x, nj.y, nj.z));
                        const float dist = __fsqrt_rn(d2);
                        int a, ap, b, bp;
                        float wa, wb;
                        soft_bin(dist, P.adim, P.ahalf, P.A, a, ap, wa);
                        if (cosine < 0.0f) cosine = 0.0f;
                        if (cosine > 2.0f) cosine = 2.0f;
                        soft_bin(cosine, P.bdim, P.bhalf, P.B, b, bp, wb);
                        const float ua = __fsub_rn(1.0f, wa), ub = __fsub_rn(1.0f, wb);
                        float* h0 = hist + (a * P.B) * 32 + lane;
                        float* h1 = hist + (ap * P.B) * 32 + lane;
                        // the four `+=` of hpp:350-355, in source order (cells may coincide)
                        h0[b * 32] = __fadd_rn(h0[b * 32], __fmul_rn(ub, ua));
                        h0[bp * 32] = __fadd_rn(h0[bp * 32], __fmul_rn(wb, ua));
                        h1[b * 32] = __fadd_rn(h1[b * 32], __fmul_rn(ub, wa));
                        h1[bp * 32] = __fadd_rn(h1[bp * 32], __fmul_rn(wb, wa));
                        npairs++;
                    }
                }
            }
        }
    }
    __syncwarp();
    // per-annulus L2 normalisation (hpp:360-365): sequential sum of squares, IEEE sqrt and divide
    if (active) {
        for (int a = 0; a < P.A; ++a) {
            float* h = hist + (a * P.B) * 32 + lane;
            float ss = 0.0f;
            for (int b = 0; b < P.B; ++b) ss = __fadd_rn(ss, __fmul_rn(h[b * 32], h[b * 32]));
            const float norm = __fsqrt_rn(ss);
            if (norm > 0.0f)
                for (int b = 0; b < P.B; ++b) h[b * 32] = __fdiv_rn(h[b * 32], norm);
        }
    }
    __syncwarp();
    // coalesced store of the warp's 32 rows (row-major, sorted order)
    {
        const int nvalid = min(32, P.n - q0);
        const int total = nvalid * P.F;
        float* dst = feat + (int64_t)q0 * P.F;
        for (int e = lane; e < total; e += 32) {
            int row = e / P.F, f = e - row * P.F;
            dst[e] = hist[f * 32 + row];
        }
    }
    npairs = __reduce_add_sync(0xFFFFFFFFu, npairs);
    if (lane == 0) {
        atomicAdd(counters + 0, (unsigned long long)npairs);
        atomicAdd(counters + 1, (unsigned long long)ncand * 32ull);
    }
}

cudaError_t launch_features(kpl_ctx* c, int64_t n, bool use_role)
{
    const kpl_params& U = c->params;
    FeatParams P;
    P.n = (int)n; P.A = U.n_annulus; P.B = U.n_bins; P.F = P.A * P.B; P.reach = c->grid.reach_feat;
    const double r = (double)U.radius_features;
    P.r2 = (float)(r * r);                       // static_cast<float>(radius*radius), KdTreeFLANN::radiusSearch
    P.support = (float)r;                        // findAnnulusPair(.., (float)search_radius_, ..) hpp:345
    P.adim = P.support / (float)P.A;             // src/KeypointLearning.cpp:43
    P.ahalf = P.adim / 2.0f;                     // :52
    P.bdim = 2.0f / (float)P.B;                  // :75
    P.bhalf = P.bdim / 2.0f;                     // :84
    P.cellf = (float)c->grid.cell;
    P.rcull2 = (float)(r * r * (1.0 + 1e-5));
    cudaError_t e;
    if ((e = ensure(c->feat, (size_t)n * P.F))) return e;
    size_t smem = (size_t)FEAT_WARPS * (P.F * 32 + 256) * sizeof(float);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        if ((e = cudaFuncSetAttribute(feature_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
        configured = smem;
    }
    int warps = (int)((n + 31) / 32);
    int blocks = (warps + FEAT_WARPS - 1) / FEAT_WARPS;
    feature_kernel<<<blocks, FEAT_WARPS * 32, smem, c->stream>>>(c->s_pos.p, c->s_nrm.p, c->key_b.p, c->cell_start.p,
                                                                 use_role ? c->s_role.p : nullptr,
                                                                 c->grid.dim[0], c->grid.dim[1], c->grid.dim[2], P, c->feat.p, c->counters.p);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace kpl
