// features.cu -- radius search + annuli x bins histogram, the dominant kernel of the path.
// Replaces KeypointLearningDetector::computePointFeatures (impl/KeypointLearning.hpp:321-376) with
// its helpers findAnnulusPair / findBinPair (src/KeypointLearning.cpp:41-92) and the FLANN radius
// search behind searchForNeighbors (hpp:334).
//
// Mapping: one warp (= one block) owns 32 consecutive queries of the query order (grid.cu: build_lists, a
// Hilbert curve over sub-cells, so the 32 queries are a compact patch of the surface), one lane per query, and
// the warp's queries share a tight candidate box.  The warp walks the cell rows that can hold neighbours
// of any of its queries; each row is ONE contiguous range of the sorted arrays, staged 32 candidates at a
// time into a shared SoA tile with coalesced float4 loads (the next tile's loads are in flight while the
// current one is consumed).  A tile is consumed in two phases:
//   1. membership: every lane evaluates the exact FLANN predicate d2 < r2 for all 32 candidates --
//      broadcast 128-bit shared loads, four candidates per step in packed FP32 (FADD2/FMUL2/FFMA2) --
//      into a 32-bit mask;
//   2. votes: every lane walks the set bits of ITS mask in ascending candidate order, two neighbours per iteration
//      with the FP32 arithmetic of both packed, so the expensive vote runs only for real neighbours (tiles some
//      lane takes nearly whole are stepped through in order instead, every lane storing only its own votes).
// Each query therefore sees its neighbours in ascending sorted position = canonical
// (cell key, index) order and accumulates its votes sequentially in FP32 into a lane-private
// histogram column hist[cell][lane] (bank == lane: conflict-free, no atomics).  That fixed order
// is what makes the histogram bit-identical to the oracle.  The tail normalises the row per annulus and,
// in kpl_detect*, evaluates the forest straight out of the histogram column (forest.cuh).
//
// Arithmetic: IEEE binary32 RN without contraction, as the reference evaluates it.  The FAST
// variant replaces the IEEE square root and the divisions by the two run constants (annulus and
// bin width) with short FMA-corrected sequences; they are used only after selftest_kernel has
// proved them bit-identical to __fsqrt_rn / __fdiv_rn for every input mantissa on this device.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include "kpl_internal.h"
#include "kpl_math.cuh"
#include "forest.cuh"

namespace kpl {

// One warp per block: its shared memory is released the moment it finishes (4 / 2 / 1 warps per block measured
// 219.5 / -- / 204.0 ms on the 10 M-point scene, profiles/r1e_sweep_warps_per_block.txt).  __launch_bounds__(32) also
// gives ptxas its best allocation (round 1: 64 registers without spills; bounds of 128 threads: 64 with spills,
// 180.0 ms; (32, 28): 71 registers, 181.9 ms; (32): 178.0 ms.  The round-2 kernel takes 72 registers, which still
// fits the 28 blocks per SM that shared memory allows).
static constexpr int FEAT_WARPS = 1;
// Largest mask of a tile from which the in-order loop (16 iterations of ~120 instructions) beats the bit walk
// (ceil(max / 2) iterations of ~137)
static constexpr int DENSE_TILE = 29;

struct FeatParams {
    int n, A, B, F, reach;
    int unbias;            // 1 - bits(2^23): see soft_bin_x2
    int group_extent;      // cells a group of queries may span beyond its first, per axis
    int dense_tile;        // DENSE_TILE
    int recip_normalize;   // row.normalize() as Eigen 3.2.x: multiply by 1/norm instead of dividing (kpl_params.eigen32_normalize)
    float r2, support, adim, ahalf, ainv, bdim, bhalf, binv, cellf, rcull2, mhalf;
    uint64_t one2;   // (1.0f, 1.0f), opaque to the compiler: see dist2_x2
    // packed (v, v) copies of the run constants for the two-votes-per-iteration loop; n* = negated.  The bin constants
    // of that loop are those of the HALF cosine (bdim/2, bhalf/2, 2*binv: exact scalings, see vote_pair)
    uint64_t adim2, nadim2, ainv2, ahalf2, bdim2, nbdim2, binv2, bhalf2;
};

// Forest evaluation fused into the tail of the feature kernel (nodes == nullptr: not fused): the
// normalised row is scored straight out of the lane's shared-memory histogram column, so the
// A*B*4-byte row never travels to HBM and back.
struct FusedForest {
    const PackedNode* nodes;
    const int32_t* roots;
    int ntrees;
    float* s_score;   // sorted order
    float* score;     // original order
    uint8_t* fragile; // original order: 1 = some split on the point's walks was decided within 1e-5 (forest.cuh)
};

__device__ __forceinline__ void key_to_cell_f(uint32_t key, int dimx, int dimy, int& cx, int& cy, int& cz)
{
    uint32_t t = key / (uint32_t)dimx;
    cx = (int)(key - t * (uint32_t)dimx);
    cz = (int)(t / (uint32_t)dimy);
    cy = (int)(t - (uint32_t)cz * (uint32_t)dimy);
}

// ---- verified-fast arithmetic --------------------------------------------------------------------
// The FMA-corrected rsqrt square root is bit-identical to the IEEE one on [2^-100, 2^40) (selftest_kernel checks every
// float of that range on the device; below 2^-102 the correction term turns denormal and the last bit goes wrong,
// tools/probe_fast_sqrt.cu).  Smaller d2 -- duplicate points (0), denormal-range distances -- need no branch: with the
// rsqrt argument clamped to 2^-100 the sequence returns 0 for 0 and something below 2^-50 otherwise, like the true root,
// and both vanish in the only two places the distance is used (floor(dist / adim) = 0; dist - adim / 2 rounds to
// -adim / 2) as long as adim >= 2^-20, which launch_features requires of the FAST variant.
static constexpr float FAST_SQRT_LO = 7.888609052210118e-31f;   // 2^-100
static constexpr float FAST_SQRT_HI = 1.099511627776e12f;       // 2^40
static constexpr float FAST_ADIM_MIN = 9.5367431640625e-07f;    // 2^-20

__device__ __forceinline__ float fast_sqrt_core(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(x, FAST_SQRT_LO)));
    const float s = __fmul_rn(x, y), h = __fmul_rn(y, 0.5f);
    const float e = __fmaf_rn(-s, s, x);
    return __fmaf_rn(e, h, s);
}
// The caller guarantees x < FAST_SQRT_HI (x < r2, and the FAST variant is only launched when r2 <= 2^40).
template <bool FAST>
__device__ __forceinline__ float ksqrt(float x)
{
    return FAST ? fast_sqrt_core(x) : __fsqrt_rn(x);
}
__device__ __forceinline__ float fast_div_core(float x, float c, float cinv)
{
    const float q = __fmul_rn(x, cinv);
    const float r = __fmaf_rn(-q, c, x);
    return __fmaf_rn(r, cinv, q);
}
template <bool FAST>
__device__ __forceinline__ float kdiv(float x, float c, float cinv)
{
    return FAST ? fast_div_core(x, c, cinv) : __fdiv_rn(x, c);
}

// src/KeypointLearning.cpp:41-65 / :68-92 with float abs; dim = bin width, half = dim/2, inv = fl(1/dim).
// v >= 0 and RN(v/dim) <= n on this path, so `if (i == n) i--` is min(i, n-1) and the two pair clamps
// (`-1 -> 0`, `n -> i`) are a clamp of i +- 1 to [0, n-1].
template <bool FAST>
__device__ __forceinline__ void soft_bin_k(float v, float dim, float half, float inv, int nm1, int& idx, int& pair, float& w)
{
    const int i = min(__float2int_rd(kdiv<FAST>(v, dim, inv)), nm1);   // static_cast<int>(floor(v/dim)); if (i == n) i--
    const float center = __fadd_rn(__fmul_rn((float)i, dim), half);
    const float ww = kdiv<FAST>(__fsub_rn(v, center), dim, inv);
    const int p = min(max(i + ((ww > 0.0f) ? 1 : -1), 0), nm1);
    idx = i; pair = p; w = fabsf(ww);
}

// -a as RN(0 - a): one FADD2 instead of two sign flips (a == 0 never reaches the consumer: d2 == 0 takes the IEEE path)
__device__ __forceinline__ uint64_t neg2(uint64_t a) { return sub2(0ull, a); }
__device__ __forceinline__ uint64_t abs2(uint64_t a) { return a & 0x7FFFFFFF7FFFFFFFull; }
// fast_div_core for two values: q = x*cinv; r = x - q*c; q + r*cinv  (nc = (-c, -c))
__device__ __forceinline__ uint64_t fast_div_x2(uint64_t x, uint64_t nc, uint64_t cinv)
{
    const uint64_t q = mul2(x, cinv);
    const uint64_t r = fma2(q, nc, x);
    return fma2(r, cinv, q);
}
// fast_sqrt_core for two values (y = rsqrt.approx of each half)
__device__ __forceinline__ uint64_t fast_sqrt_x2(uint64_t x, float x0, float x1)
{
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fmaxf(x0, FAST_SQRT_LO)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fmaxf(x1, FAST_SQRT_LO)));
    const uint64_t y = pack2(y0, y1);
    const uint64_t s = mul2(x, y), h = mul2(y, 0x3F0000003F000000ull);
    const uint64_t e = fma2(neg2(s), s, x);
    return fma2(e, h, s);
}
// floor of two non-negative quotients without conversions: RD(q + 2^23) = 2^23 + floor(q) for 0 <= q < 2^22, so the low
// mantissa bits of the sum ARE the integer and (sum - 2^23) is its float (FADD2.RM + FADD2 instead of 2 F2I + 2 I2F)
static constexpr uint32_t FLOOR_MAGIC_BITS = 0x4B000000u;                    // 2^23
static constexpr uint64_t FLOOR_MAGIC2 = 0x4B0000004B000000ull;
__device__ __forceinline__ uint64_t add_rm2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2u(uint64_t v, uint32_t& lo, uint32_t& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
// soft_bin_k for two values at once (FAST arithmetic only): same expressions, packed where they are FP32.
// Returns i + 1 (j0, j1) and the clamped pair index (p0, p1); ww keeps its SIGN (|ww| is the weight: the callers fold the
// absolute value into the operand modifiers of the consuming FADDs, products of weights commute with it) and
// u = 1 - |ww|.  When ww == 0 the pair cell receives an exact +0, so its index is arbitrary (i + 1 here, i - 1 in
// soft_bin_k).
__device__ __forceinline__ void soft_bin_x2(uint64_t v, uint64_t dim2, uint64_t ndim2, uint64_t inv2, uint64_t half2, uint64_t one2,
                                            int nm1, int unbias, int& j0, int& j1, int& p0, int& p1, uint64_t& ww, uint64_t& u)
{
    uint32_t t0, t1;
    unpack2u(add_rm2(fast_div_x2(v, ndim2, inv2), FLOOR_MAGIC2), t0, t1);
    t0 = min(t0, FLOOR_MAGIC_BITS + (uint32_t)nm1);                                        // if (i == n) i--
    t1 = min(t1, FLOOR_MAGIC_BITS + (uint32_t)nm1);
    const uint64_t fi = sub2(pack2u(t0, t1), FLOOR_MAGIC2);                                // (float)i
    const uint64_t center = fma2(mul2(fi, dim2), one2, half2);                             // RN(RN(i*dim) + half)
    ww = fast_div_x2(sub2(v, center), ndim2, inv2);
    uint32_t w0, w1;
    unpack2u(ww, w0, w1);
    j0 = (int)t0 + unbias;        // unbias = 1 - FLOOR_MAGIC_BITS, passed at run time: a compile-time constant would be
    j1 = (int)t1 + unbias;        // re-associated into every cell address (two extra adds per read-modify-write)
    p0 = __vimin_s32_relu(j0 + 2 * ((int)w0 >> 31), nm1);                                   // clamp(i +- 1, 0, n - 1)
    p1 = __vimin_s32_relu(j1 + 2 * ((int)w1 >> 31), nm1);
    u = pack2(__fsub_rn(1.0f, fabsf(__uint_as_float(w0))), __fsub_rn(1.0f, fabsf(__uint_as_float(w1))));
}
// clamp(RN(1 - dot), 0, 2) / 2 in one instruction: RN(0.5 - 0.5*dot) is RN(1 - dot) / 2 (scaling by two commutes with
// rounding; 1 - dot is 0 or >= 2^-24, no underflow) and .sat clamps it to [0, 1]
__device__ __forceinline__ float half_cosine(float dot, float minus_half)
{
    float h;
    asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(h) : "f"(dot), "f"(minus_half));   // (one immediate per instruction)
    return h;
}

// One thread per (binade, mantissa): result[0] counts sqrt mismatches over [2^-100, 2^40), result[1]
// / result[2] division mismatches for the annulus / bin width over every mantissa of [1, 2) and [-2,-1)
// (exact power-of-two scaling extends the proof to every binade without under/overflow), result[3]
// mismatches of the packed squared distance against the scalar dist2() on values built from the mantissa.
__global__ void __launch_bounds__(256) selftest_kernel(float adim, float ainv, float bdim, float binv, uint64_t one2, unsigned* __restrict__ result)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;   // 2^23 mantissas
    if (m >= (1u << 23)) return;
    unsigned bad_s = 0, bad_a = 0, bad_b = 0, bad_p = 0;
    for (int ex = 127 - 100; ex < 127 + 40; ++ex) {
        const float x = __uint_as_float(((uint32_t)ex << 23) | m);
        bad_s += (__float_as_uint(fast_sqrt_core(x)) != __float_as_uint(__fsqrt_rn(x)));
    }
    const float x1 = __uint_as_float((127u << 23) | m);
    bad_a += (__float_as_uint(fast_div_core(x1, adim, ainv)) != __float_as_uint(__fdiv_rn(x1, adim)));
    bad_a += (__float_as_uint(fast_div_core(-x1, adim, ainv)) != __float_as_uint(__fdiv_rn(-x1, adim)));
    bad_b += (__float_as_uint(fast_div_core(x1, bdim, binv)) != __float_as_uint(__fdiv_rn(x1, bdim)));
    bad_b += (__float_as_uint(fast_div_core(-x1, bdim, binv)) != __float_as_uint(__fdiv_rn(-x1, bdim)));
    {
        uint32_t h = m * 2654435761u + 12345u;
        float v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            h = h * 1664525u + 1013904223u;
            v[t] = __uint_as_float(((h >> 31) << 31) | ((118u + ((h >> 23) & 15u)) << 23) | (h & 0x7FFFFFu));   // +-[2^-9, 2^7)
        }
        const uint64_t d = dist2_x2(pack2(v[0], v[0]), pack2(v[1], v[1]), pack2(v[2], v[2]),
                                    pack2(v[3], v[6]), pack2(v[4], v[7]), pack2(v[5], v[8]), one2);
        float d0, d1;
        unpack2(d, d0, d1);
        bad_p += (__float_as_uint(d0) != __float_as_uint(dist2(v[0], v[1], v[2], v[3], v[4], v[5])));
        bad_p += (__float_as_uint(d1) != __float_as_uint(dist2(v[0], v[1], v[2], v[6], v[7], v[8])));
        // the packed sqrt / division / soft-binning of the two-vote loop against their scalar forms
        const float xa = fabsf(v[3]), xb = x1;
        float r0, r1;
        unpack2(fast_sqrt_x2(pack2(xa, xb), xa, xb), r0, r1);
        bad_p += (__float_as_uint(r0) != __float_as_uint(fast_sqrt_core(xa))) + (__float_as_uint(r1) != __float_as_uint(fast_sqrt_core(xb)));
        unpack2(fast_div_x2(pack2(v[4], x1), pack2(-adim, -adim), pack2(ainv, ainv)), r0, r1);
        bad_p += (__float_as_uint(r0) != __float_as_uint(fast_div_core(v[4], adim, ainv))) + (__float_as_uint(r1) != __float_as_uint(fast_div_core(x1, adim, ainv)));
        {
            // the packed soft binning of the HALF cosine (floor by round-down addition, signed weights) against the scalar
            // soft binning of the cosine itself, and the saturating half cosine against the reference's clamp
            const float c0 = fminf(fabsf(v[5]), 2.0f), c1 = __fsub_rn(x1, 1.0f) * 2.0f;    // cosines in [0, 2]
            const int Bm1 = (int)(2.0f / bdim + 0.5f) - 1;
            const float hb = bdim * 0.5f;
            int j0, j1, p0, p1, si, sp; uint64_t ww, u; float sw, w0, w1, u0, u1;
            soft_bin_x2(pack2(c0 * 0.5f, c1 * 0.5f), pack2(hb, hb), pack2(-hb, -hb), pack2(binv * 2.0f, binv * 2.0f),
                        pack2(hb * 0.5f, hb * 0.5f), one2, Bm1, 1 - (int)FLOOR_MAGIC_BITS, j0, j1, p0, p1, ww, u);
            unpack2(ww, w0, w1); unpack2(u, u0, u1);
            soft_bin_k<true>(c0, bdim, bdim / 2.0f, binv, Bm1, si, sp, sw);
            bad_p += (si != j0 - 1) + (sw != 0.0f && sp != p0) + (__float_as_uint(sw) != __float_as_uint(fabsf(w0))) + (__float_as_uint(__fsub_rn(1.0f, sw)) != __float_as_uint(u0));
            soft_bin_k<true>(c1, bdim, bdim / 2.0f, binv, Bm1, si, sp, sw);
            bad_p += (si != j1 - 1) + (sw != 0.0f && sp != p1) + (__float_as_uint(sw) != __float_as_uint(fabsf(w1))) + (__float_as_uint(__fsub_rn(1.0f, sw)) != __float_as_uint(u1));
            // dots in (-4, 4) (beyond [-1, 1] the clamp acts), and the exact ends
            const float dots[4] = {__fsub_rn(x1, 1.5f) * 8.0f, v[6] * 0.03125f, m == 0 ? 1.0f : -1.0f, __fsub_rn(1.0f, v[7] * 1.52587890625e-05f)};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float ref = fminf(fmaxf(__fsub_rn(1.0f, dots[t]), 0.0f), 2.0f) * 0.5f;
                bad_p += (__float_as_uint(half_cosine(dots[t], -0.5f)) != __float_as_uint(ref));
            }
        }
    }
    bad_s = __reduce_add_sync(0xFFFFFFFFu, bad_s);
    bad_a = __reduce_add_sync(0xFFFFFFFFu, bad_a);
    bad_b = __reduce_add_sync(0xFFFFFFFFu, bad_b);
    bad_p = __reduce_add_sync(0xFFFFFFFFu, bad_p);
    if ((threadIdx.x & 31) == 0) {
        if (bad_s) atomicAdd(result + 0, bad_s);
        if (bad_a) atomicAdd(result + 1, bad_a);
        if (bad_b) atomicAdd(result + 2, bad_b);
        if (bad_p) atomicAdd(result + 3, bad_p);
    }
}

// shared-memory accesses of the vote loop by 32-bit shared address (keeps the address arithmetic to one
// add per cell and the four read-modify-writes in source order)
template <int OFF = 0>
__device__ __forceinline__ float lds_f32(unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF = 0>
__device__ __forceinline__ void sts_f32(unsigned a, float v)
{
    asm volatile("st.shared.f32 [%0+%2], %1;" :: "r"(a), "f"(v), "n"(OFF));
}

// bit walking without constants in registers: FLO and BMSK
__device__ __forceinline__ unsigned top_bit(uint32_t mask)           // index of the most significant set bit; 0xFFFFFFFF for 0
{
    unsigned r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(mask));
    return r;
}
__device__ __forceinline__ uint32_t bits_below(unsigned b)           // (1 << b) - 1 for b in [0, 32]
{
    uint32_t r;
    asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r) : "r"(b));
    return r;
}
// one float of the candidate tile: OFF = 128 * component (x, y, z, nx, ny, nz rows of the SoA tile)
template <int OFF>
__device__ __forceinline__ float lds_tile(unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}

// The four `+=` of hpp:350-355 for one neighbour, in source order (cells may coincide when a pair index
// was clamped onto the primary one, so the read-modify-writes must stay sequential).  Loading the four
// cells together and forwarding coinciding values was measured slower (more issue slots than it saves
// in shared-memory latency at 28 resident warps: 211 ms vs 204 ms on the 10 M-point scene).
// ja / jb = primary annulus / bin index + 1 (hbm = this lane's column one annulus row BELOW the histogram), ap / bp = pair
// indices; the votes carry the signs of the soft-binning quotients, |v| is the weight.
__device__ __forceinline__ void vote4(unsigned hbm, unsigned hb, unsigned row_bytes, int ja, int ap, int jb, int bp, float v00, float v01, float v10,
                                      float v11, bool on = true)
{
    const unsigned ra = hbm + (unsigned)ja * row_bytes, rp = hb + (unsigned)ap * row_bytes;
    const unsigned c00 = ra + ((unsigned)jb << 7), c01 = ra + ((unsigned)bp << 7);      // c00 / c10 lie one bin too high: OFF = -128
    const unsigned c10 = rp + ((unsigned)jb << 7), c11 = rp + ((unsigned)bp << 7);
    // `on` = false (the odd last vote of a lane, paired with itself): everything but the stores runs -- predicated stores
    // instead of a divergent branch around the four updates
    const float h00 = __fadd_rn(lds_f32<-128>(c00), fabsf(v00));
    if (on) sts_f32<-128>(c00, h00);
    const float h01 = __fadd_rn(lds_f32<0>(c01), fabsf(v01));
    if (on) sts_f32<0>(c01, h01);
    const float h10 = __fadd_rn(lds_f32<-128>(c10), fabsf(v10));
    if (on) sts_f32<-128>(c10, h10);
    const float h11 = __fadd_rn(lds_f32<0>(c11), fabsf(v11));
    if (on) sts_f32<0>(c11, h11);
}

template <bool FAST, bool FRAGILE>
__global__ void __launch_bounds__(FEAT_WARPS * 32)
feature_kernel(const float4* __restrict__ s_pos, const float4* __restrict__ s_nrm, const uint32_t* __restrict__ skey,
               const int32_t* __restrict__ cell_start,
               const int32_t* __restrict__ qlist, const int32_t* __restrict__ warp_starts, int nwarps, int nlist,
               const uint32_t* __restrict__ warp_order, int rows_by_list, int dimx, int dimy, int dimz, FeatParams P, FusedForest FF, float* __restrict__ feat,
               unsigned long long* __restrict__ counters)
{
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = P.F * 32 + 192;
    float* hist = smem + warp * per_warp;
    float* sx = hist + P.F * 32;                                  // SoA candidate tile: x, y, z, nx, ny, nz
    float* sy = sx + 32; float* sz = sy + 32; float* snx = sz + 32; float* sny = snx + 32; float* snz = sny + 32;

    // Warp w takes entries [warp_starts[w], warp_starts[w + 1]) -- at most 32 -- of the query list (sorted positions): the
    // Hilbert order of every point with a scoring role (kpl_detect*, kpl_features of a whole cloud; grid.cu: build_lists),
    // or 32 consecutive entries of the ascending list of an index subset (computePointsForTrainingFeatures, hpp:299-318:
    // no warp_starts, rows_by_list: output row = list entry).
    int w = blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= nwarps) return;
    if (warp_order) w = (int)__ldg(warp_order + w);
    const int q0 = warp_starts ? __ldg(warp_starts + w) : w * 32;
    const int nvalid = warp_starts ? __ldg(warp_starts + w + 1) - q0 : min(32, nlist - q0);
    const bool valid = lane < nvalid;
    const int q = valid ? __ldg(qlist + q0 + lane) : 0;
    bool active = valid;
    float4 qp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f), qn = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        qp = __ldg(s_pos + q);
        qn = __ldg(s_nrm + q);
        key_to_cell_f(__ldg(skey + q), dimx, dimy, cx, cy, cz);
        // a query without a finite normal is not scored (hpp:277): its row stays zero
        if (!(isfinite(qn.x) && isfinite(qn.y) && isfinite(qn.z))) { active = false; qp.x = CUDART_NAN_F; }
    }

    for (int f = 0; f < P.F; ++f) hist[f * 32 + lane] = 0.0f;
    unsigned npairs = 0, ncand = 0;
    const unsigned hb = (unsigned)__cvta_generic_to_shared(hist) + (unsigned)lane * 4u;   // this lane's histogram column
    const unsigned row_bytes = (unsigned)P.B * 128u;
    const unsigned hbm = hb - row_bytes;                   // vote4 adds (annulus + 1) rows to it
    const unsigned tile0 = (unsigned)__cvta_generic_to_shared(sx);     // slot 0 of the candidate tile's x row
    const int Am1 = P.A - 1, Bm1 = P.B - 1;
    const uint64_t QY = pack2(qp.y, qp.y), QZ = pack2(qp.z, qp.z);
    const uint64_t QNX = pack2(qn.x, qn.x), QNY = pack2(qn.y, qn.y), QNZ = pack2(qn.z, qn.z);

    // The queries of a warp are neighbours on the curve, almost always within a cell or two of each other.  They are
    // processed in groups: the lanes within P.group_extent cells of the first remaining lane whose cells span at most
    // P.group_extent + 1 per axis, so the candidate region of a pass is a tight box around the group; the rare lanes outside it (a chain of
    // queries drifting over several cells, far-apart entries of an index subset) idle for that pass (their px is NaN).
    const int GE = P.group_extent;
    unsigned remaining = __ballot_sync(0xFFFFFFFFu, active);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int lx = __shfl_sync(0xFFFFFFFFu, cx, leader), ly = __shfl_sync(0xFFFFFFFFu, cy, leader), lz = __shfl_sync(0xFFFFFFFFu, cz, leader);
        const bool near = active && ((remaining >> lane) & 1u) && (unsigned)(cx - lx + GE) <= 2u * GE && (unsigned)(cy - ly + GE) <= 2u * GE &&
                          (unsigned)(cz - lz + GE) <= 2u * GE;
        const int minx = __reduce_min_sync(0xFFFFFFFFu, near ? cx : lx), gy0 = __reduce_min_sync(0xFFFFFFFFu, near ? cy : ly);
        const int gz0 = __reduce_min_sync(0xFFFFFFFFu, near ? cz : lz);
        const bool member = near && cx - minx <= GE && cy - gy0 <= GE && cz - gz0 <= GE;        // (the leader is one: it is near itself)
        remaining &= ~__ballot_sync(0xFFFFFFFFu, member);
        const int maxx = __reduce_max_sync(0xFFFFFFFFu, member ? cx : lx);
        const int gy1 = __reduce_max_sync(0xFFFFFFFFu, member ? cy : ly), gz1 = __reduce_max_sync(0xFFFFFFFFu, member ? cz : lz);
        const float px = member ? qp.x : CUDART_NAN_F;
        const uint64_t QX = pack2(px, px);

        const int y0 = max(gy0 - P.reach, 0), y1 = min(gy1 + P.reach, dimy - 1);
        const int z0 = max(gz0 - P.reach, 0), z1 = min(gz1 + P.reach, dimz - 1);
        const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);

        // ---- warp-uniform iterator over the candidate tiles: rows in ascending (z, y), 32 points per tile
        int rb = 0, nr = 0, l = 0, row_s = 0, row_e = 0, cur = 0, end = 0;
        auto next_tile = [&](int& tb, int& te) -> bool {
            while (cur >= end) {
                if (l >= nr) {
                    if (rb >= nrows) return false;
                    // each lane resolves one cell row: [s, e) in the sorted arrays, after conservative culling
                    row_s = 0; row_e = 0;
                    const int r = rb + lane;
                    if (r < nrows) {
                        const int zz = z0 + r / ny, yy = y0 + r % ny;
                        // whole empty cells between the row and the group's cells: a lower bound of the distance
                        const int gy = max(max(gy0 - yy, yy - gy1) - 1, 0), gz = max(max(gz0 - zz, zz - gz1) - 1, 0);
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
        float4 cp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f), cn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have && tb + lane < te) { cp = __ldg(s_pos + tb + lane); cn = __ldg(s_nrm + tb + lane); }

        while (have) {
            // neighbours with a non-finite normal never vote (hpp:338): poison the position
            if (!(isfinite(cn.x) && isfinite(cn.y) && isfinite(cn.z))) cp.x = CUDART_NAN_F;
            __syncwarp();                          // previous tile fully consumed
            // candidate k of the tile goes to SLOT 31 - k, the number of its mask bit: the vote loop finds a set bit with
            // FLO and that is the slot's index
            sx[31 - lane] = cp.x; sy[31 - lane] = cp.y; sz[31 - lane] = cp.z;
            snx[31 - lane] = cn.x; sny[31 - lane] = cn.y; snz[31 - lane] = cn.z;
            const int cnt = min(32, te - tb);
            const unsigned self = (unsigned)(q - tb);   // slot of the query itself when it lies in this tile
            have = next_tile(tb, te);
            cp = make_float4(CUDART_NAN_F, 0.f, 0.f, 0.f);
            if (have && tb + lane < te) { cp = __ldg(s_pos + tb + lane); cn = __ldg(s_nrm + tb + lane); }
            __syncwarp();
            ncand += cnt;

            // phase 1: membership mask, candidate k -> bit 31-k (candidates >= cnt hold NaN positions and never
            // pass).  Four candidates per step: broadcast 128-bit loads of the SoA tile, packed FP32 math.
            uint32_t mask = 0;
#pragma unroll
            for (int s0 = 24; s0 >= 0; s0 -= 8) {                 // slots [s0, s0 + 8) = candidates [24 - s0, 32 - s0)
                if (24 - s0 < cnt) {
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(sx + s0 + u);
                        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(sy + s0 + u);
                        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(sz + s0 + u);
                        float d0, d1, d2, d3;
                        unpack2(dist2_x2(QX, QY, QZ, X.x, Y.x, Z.x, P.one2), d0, d1);
                        unpack2(dist2_x2(QX, QY, QZ, X.y, Y.y, Z.y, P.one2), d2, d3);
                        if (d0 < P.r2) mask |= 1u << (s0 + u);
                        if (d1 < P.r2) mask |= 1u << (s0 + u + 1);
                        if (d2 < P.r2) mask |= 1u << (s0 + u + 2);
                        if (d3 < P.r2) mask |= 1u << (s0 + u + 3);
                    }
                }
            }
            if (self < 32u) mask &= ~(0x80000000u >> self);   // the query is not its own neighbour (hpp:336)
            npairs += __popc(mask);

            // phase 2: votes of this lane's neighbours, ascending sorted position
            if constexpr (FAST) {
                // two neighbours per iteration: every FP32 expression of the two votes is evaluated with one
                // packed instruction (same IEEE RN operations, half the issue slots); the eight histogram
                // read-modify-writes stay scalar and in order.  An odd last vote is paired with itself and
                // its second set of updates is skipped.  (Software-pipelining the loop -- computing the next
                // pair while the current eight read-modify-writes drain -- was measured slower: 80 registers,
                // 211 ms vs 204 ms on the 10 M-point scene; capped at 72 registers 215 ms.)
                // hi_first: the candidate in the HIGH half of the packed operands is the earlier one (64-bit loads of the
                // reversed tile).  on_lo / on_hi: whether that half's four updates are stored (predicated stores).
                auto vote_pair = [&](const uint64_t X, const uint64_t Y, const uint64_t Z, const uint64_t NX, const uint64_t NY,
                                     const uint64_t NZ, const bool on_lo, const bool on_hi, auto hi_first) {
                    const uint64_t D = dist2_x2(QX, QY, QZ, X, Y, Z, P.one2);
                    // 1 - (n0*m0 + (n1*m1 + n2*m2)), hpp:341-342 with Eigen's reduction order, clamped to [0, 2]
                    // (src/KeypointLearning.cpp:70-73) -- carried as its exact half
                    const uint64_t dot = fma2(mul2(QNX, NX), P.one2, fma2(mul2(QNY, NY), P.one2, mul2(QNZ, NZ)));
                    float t0, t1, d0, d1;
                    unpack2(dot, t0, t1);
                    unpack2(D, d0, d1);
                    const uint64_t DIST = fast_sqrt_x2(D, d0, d1);                                     // hpp:345 sqrt(distances[..])
                    const uint64_t HCOS = pack2(half_cosine(t0, P.mhalf), half_cosine(t1, P.mhalf));
                    int a0, a1, ap0, ap1, b0, b1, bp0, bp1;
                    uint64_t WA, UA, WB, UB;
                    soft_bin_x2(DIST, P.adim2, P.nadim2, P.ainv2, P.ahalf2, P.one2, Am1, P.unbias, a0, a1, ap0, ap1, WA, UA);
                    soft_bin_x2(HCOS, P.bdim2, P.nbdim2, P.binv2, P.bhalf2, P.one2, Bm1, P.unbias, b0, b1, bp0, bp1, WB, UB);
                    float v00a, v00b, v01a, v01b, v10a, v10b, v11a, v11b;
                    unpack2(mul2(UB, UA), v00a, v00b);
                    unpack2(mul2(WB, UA), v01a, v01b);
                    unpack2(mul2(UB, WA), v10a, v10b);
                    unpack2(mul2(WB, WA), v11a, v11b);
                    if constexpr (decltype(hi_first)::value) {
                        vote4(hbm, hb, row_bytes, a1, ap1, b1, bp1, v00b, v01b, v10b, v11b, on_hi);
                        vote4(hbm, hb, row_bytes, a0, ap0, b0, bp0, v00a, v01a, v10a, v11a, on_lo);
                    } else {
                        vote4(hbm, hb, row_bytes, a0, ap0, b0, bp0, v00a, v01a, v10a, v11a, on_lo);
                        vote4(hbm, hb, row_bytes, a1, ap1, b1, bp1, v00b, v01b, v10b, v11b, on_hi);
                    }
                };
                // Dense tiles -- some query takes (nearly) all 32 candidates, so the bit walk below would run 15 or 16
                // iterations anyway; half of all vote iterations, a third of them in tiles EVERY query takes whole -- are
                // stepped through in order instead: no bit walking, the candidate pairs come straight out of 64-bit
                // broadcast loads, and a lane stores only the votes of its own mask bits.
                if (__reduce_max_sync(0xFFFFFFFFu, (unsigned)__popc(mask)) >= (unsigned)P.dense_tile) {
#pragma unroll 2
                    for (int sl = 30; sl >= 0; sl -= 2) {          // candidates 31 - sl - 1 (high half) and 31 - sl (low half)
                        const unsigned two_bits = mask >> sl;
                        vote_pair(*reinterpret_cast<const uint64_t*>(sx + sl), *reinterpret_cast<const uint64_t*>(sy + sl),
                                  *reinterpret_cast<const uint64_t*>(sz + sl), *reinterpret_cast<const uint64_t*>(snx + sl),
                                  *reinterpret_cast<const uint64_t*>(sny + sl), *reinterpret_cast<const uint64_t*>(snz + sl),
                                  (two_bits & 1u) != 0, (two_bits & 2u) != 0, std::true_type());
                    }
                    mask = 0;
                }
                while (mask) {
                    const unsigned m0 = top_bit(mask);             // highest bit = smallest candidate index; the bit number is the slot
                    mask &= bits_below(m0);
                    const bool two = mask != 0;
                    const unsigned m1 = min(top_bit(mask), m0);     // an odd last vote is paired with itself (FLO of 0 is 0xFFFFFFFF)
                    mask &= bits_below(m1);                          // (already empty when the last vote is odd)
                    const unsigned t0 = tile0 + 4u * m0, t1 = tile0 + 4u * m1;
                    vote_pair(pack2(lds_tile<0>(t0), lds_tile<0>(t1)), pack2(lds_tile<128>(t0), lds_tile<128>(t1)),
                              pack2(lds_tile<256>(t0), lds_tile<256>(t1)), pack2(lds_tile<384>(t0), lds_tile<384>(t1)),
                              pack2(lds_tile<512>(t0), lds_tile<512>(t1)), pack2(lds_tile<640>(t0), lds_tile<640>(t1)), true, two, std::false_type());
                }
            } else {
                while (mask) {
                    const int msb = 31 - __clz(mask);              // candidate k sits at bit 31-k: highest bit = smallest k
                    mask &= ~(1u << msb);
                    const float* t = sx + msb;                       // slot number = bit number
                    const float d2 = dist2(qp.x, qp.y, qp.z, t[0], t[32], t[64]);
                    float cosine = __fsub_rn(1.0f, dot3_eigen(qn.x, qn.y, qn.z, t[96], t[128], t[160]));   // hpp:341-342
                    const float dist = ksqrt<FAST>(d2);                                                    // hpp:345 sqrt(distances[..])
                    int a, ap, b, bp;
                    float wa, wb;
                    soft_bin_k<FAST>(dist, P.adim, P.ahalf, P.ainv, Am1, a, ap, wa);
                    cosine = fminf(fmaxf(cosine, 0.0f), 2.0f);          // src/KeypointLearning.cpp:70-73 (cosine is finite here)
                    soft_bin_k<FAST>(cosine, P.bdim, P.bhalf, P.binv, Bm1, b, bp, wb);
                    const float ua = __fsub_rn(1.0f, wa), ub = __fsub_rn(1.0f, wb);
                    const unsigned ra = hb + (unsigned)a * row_bytes, rp = hb + (unsigned)ap * row_bytes;
                    const unsigned ob = (unsigned)b << 7, op = (unsigned)bp << 7;
                    // the four `+=` of hpp:350-355, in source order (cells may coincide)
                    sts_f32(ra + ob, __fadd_rn(lds_f32(ra + ob), __fmul_rn(ub, ua)));
                    sts_f32(ra + op, __fadd_rn(lds_f32(ra + op), __fmul_rn(wb, ua)));
                    sts_f32(rp + ob, __fadd_rn(lds_f32(rp + ob), __fmul_rn(ub, wa)));
                    sts_f32(rp + op, __fadd_rn(lds_f32(rp + op), __fmul_rn(wb, wa)));
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
            if (norm > 0.0f) {
                if (P.recip_normalize) {
                    const float inv = __fdiv_rn(1.0f, norm);
                    for (int b = 0; b < P.B; ++b) h[b * 32] = __fmul_rn(h[b * 32], inv);
                } else
                    for (int b = 0; b < P.B; ++b) h[b * 32] = __fdiv_rn(h[b * 32], norm);
            }
        }
    }
    __syncwarp();
    // fused forest: score = 1 - sum/ntrees (hpp:281-287); unscored points (halo role, no finite normal) get NaN
    unsigned nfragile = 0;
    if (FF.nodes) {
        bool fragile = false;
        const float sc = active ? forest_score<FRAGILE>(hist + lane, 32, FF.nodes, FF.roots, FF.ntrees, fragile) : CUDART_NAN_F;
        if (valid) {
            const uint32_t orig = __float_as_uint(__ldg(&s_pos[q].w));
            FF.s_score[q] = sc;
            FF.score[orig] = sc;
            if (FRAGILE) FF.fragile[orig] = (active && fragile) ? 1 : 0;
        }
        if (FRAGILE) nfragile = __popc(__ballot_sync(0xFFFFFFFFu, active && fragile));
    }
    // store of the warp's rows: row of the list entry (index subsets) or of the query's sorted position
    if (feat) {
        int* sq = reinterpret_cast<int*>(sx);            // the candidate tile is free now
        __syncwarp();
        sq[lane] = rows_by_list ? q0 + lane : q;
        __syncwarp();
        const int total = nvalid * P.F;
        for (int e = lane; e < total; e += 32) {
            const int row = e / P.F, f = e - row * P.F;
            feat[(int64_t)sq[row] * P.F + f] = hist[f * 32 + row];
        }
    }
    npairs = __reduce_add_sync(0xFFFFFFFFu, npairs);
    const unsigned nscored = __popc(__ballot_sync(0xFFFFFFFFu, active));
    if (lane == 0) {
        atomicAdd(counters + 5, (unsigned long long)nscored);
        atomicAdd(counters + 0, (unsigned long long)npairs);
        atomicAdd(counters + 1, (unsigned long long)ncand * 32ull);
        if (nfragile) atomicAdd(counters + 8, (unsigned long long)nfragile);
    }
}

// Runs the exhaustive self-test once per (adim, bdim) pair and caches the verdict.
static cudaError_t fast_math_verdict(kpl_ctx* c, const FeatParams& P, bool& fast)
{
    struct Verdict { float adim, bdim; int device; bool fast; };
    static std::vector<Verdict> cache;          // process-wide: contexts on several host threads share it
    static std::mutex cache_mutex;
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        for (const Verdict& v : cache)
            if (v.adim == P.adim && v.bdim == P.bdim && v.device == c->device) { fast = v.fast && P.r2 <= FAST_SQRT_HI && P.adim >= FAST_ADIM_MIN; return cudaSuccess; }
    }
    unsigned* d_res = reinterpret_cast<unsigned*>(c->counters.p + 6);   // counters[6..7] are scratch
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_res, 0, 4 * sizeof(unsigned), c->stream))) return e;
    selftest_kernel<<<(1u << 23) / 256, 256, 0, c->stream>>>(P.adim, P.ainv, P.bdim, P.binv, P.one2, d_res);
    unsigned h[4];
    if ((e = cudaMemcpyAsync(h, d_res, sizeof h, cudaMemcpyDeviceToHost, c->stream))) return e;
    if ((e = cudaStreamSynchronize(c->stream))) return e;
    if ((e = cudaMemsetAsync(d_res, 0, 4 * sizeof(unsigned), c->stream))) return e;
    if (h[3] != 0) return cudaErrorAssert;   // packed FP32 is not IEEE RN on this device: no exact path exists
    fast = (h[0] == 0 && h[1] == 0 && h[2] == 0);
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        cache.push_back({P.adim, P.bdim, c->device, fast});
    }
    fast = fast && P.r2 <= FAST_SQRT_HI && P.adim >= FAST_ADIM_MIN;   // range of the fast square root (see FAST_SQRT_LO)
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_features(kpl_ctx* c, int64_t n, bool use_role, bool fuse_forest, bool store_rows, const int32_t* d_qlist, int64_t m_list)
{
    const kpl_params& U = c->params;
    FeatParams P;
    P.n = (int)n; P.A = U.n_annulus; P.B = U.n_bins; P.F = P.A * P.B; P.reach = c->grid.reach_feat;
    P.recip_normalize = U.eigen32_normalize != 0;
    const double r = (double)U.radius_features;
    P.r2 = (float)(r * r);                       // static_cast<float>(radius*radius), KdTreeFLANN::radiusSearch
    P.support = (float)r;                        // findAnnulusPair(.., (float)search_radius_, ..) hpp:345
    P.adim = P.support / (float)P.A;             // src/KeypointLearning.cpp:43
    P.ahalf = P.adim / 2.0f;                     // :52
    P.ainv = 1.0f / P.adim;
    P.bdim = 2.0f / (float)P.B;                  // :75
    P.bhalf = P.bdim / 2.0f;                     // :84
    P.binv = 1.0f / P.bdim;
    P.cellf = (float)c->grid.cell;
    P.rcull2 = (float)(r * r * (1.0 + 1e-5));
    P.one2 = 0x3F8000003F800000ull;
    P.mhalf = -0.5f;
    P.unbias = 1 - (int)FLOOR_MAGIC_BITS;
    P.group_extent = 3;      // 2 / 3 cells measured 151.8 / 151.1 ms on the 10 M scene, 1.77 / 1.28 ms on the 64 k-point bundled view
    P.dense_tile = DENSE_TILE;
#ifdef KPL_EXPERIMENTS
    if (const char* ev = getenv("KPL_GROUP_E")) P.group_extent = atoi(ev);
    if (const char* ev = getenv("KPL_DENSE_TILE")) P.dense_tile = atoi(ev);
#endif
    auto dup = [](float v) { uint32_t b; memcpy(&b, &v, 4); return ((uint64_t)b << 32) | b; };
    P.adim2 = dup(P.adim); P.nadim2 = dup(-P.adim); P.ainv2 = dup(P.ainv); P.ahalf2 = dup(P.ahalf);
    // the packed loop bins the HALF cosine: bin width, half width and reciprocal scaled by exact powers of two
    P.bdim2 = dup(P.bdim * 0.5f); P.nbdim2 = dup(-P.bdim * 0.5f); P.binv2 = dup(P.binv * 2.0f); P.bhalf2 = dup(P.bhalf * 0.5f);
    cudaError_t e;
    if (d_qlist && (fuse_forest || !store_rows)) return cudaErrorInvalidValue;
    if (store_rows && (e = ensure(c->feat, (size_t)(d_qlist ? m_list : n) * P.F))) return e;
    if (store_rows && !d_qlist && use_role &&
        (e = cudaMemsetAsync(c->feat.p, 0, (size_t)n * P.F * sizeof(float), c->stream))) return e;      // rows of points that are no queries
    FusedForest FF = {nullptr, nullptr, 0, nullptr, nullptr, nullptr};
    if (fuse_forest) {
        if ((e = ensure(c->s_score, n)) || (e = ensure(c->score, n)) || (e = ensure(c->fragile, n))) return e;
        FF = {c->forest.d_nodes, c->forest.d_roots, c->forest.ntrees, c->s_score.p, c->score.p, c->fragile.p};
        if (use_role) {
            // points without a scoring role are not in the query order (grid.cu): unscored = NaN (all-ones is a NaN)
            if ((e = cudaMemsetAsync(c->s_score.p, 0xFF, (size_t)n * sizeof(float), c->stream)) ||
                (e = cudaMemsetAsync(c->score.p, 0xFF, (size_t)n * sizeof(float), c->stream))) return e;
            if (U.report_fragile && (e = cudaMemsetAsync(c->fragile.p, 0, (size_t)n, c->stream))) return e;
        }
    }
    bool fast = false;
    bool try_fast = true;
#ifdef KPL_EXPERIMENTS
    try_fast = getenv("KPL_NO_FAST_MATH") == nullptr;
#endif
    if (try_fast && (e = fast_math_verdict(c, P, fast))) return e;
    c->fast_math = fast;
    const int wpb = FEAT_WARPS;
    size_t smem = (size_t)wpb * (P.F * 32 + 192) * sizeof(float);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // the near-split report (kpl_params.report_fragile) costs ~2 % of the kernel: a separate instantiation
    const bool frag = fuse_forest && U.report_fragile != 0;
    auto kern = fast ? (frag ? feature_kernel<true, true> : feature_kernel<true, false>)
                     : (frag ? feature_kernel<false, true> : feature_kernel<false, false>);
    // per device and per process state of the runtime: set it on every launch that needs it (a host-side call)
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    // the query order and its warp list were built by the caller (build_lists); an index subset brings its own list
    const int warps = d_qlist ? (int)((m_list + 31) / 32) : c->nwarps_feat;
    if (warps == 0) return cudaSuccess;
    int blocks = (warps + wpb - 1) / wpb;
    kern<<<blocks, wpb * 32, smem, c->stream>>>(c->s_pos.p, c->s_nrm.p, c->key_b.p, c->cell_start.p,
                                                d_qlist ? d_qlist : c->qorder_f, d_qlist ? nullptr : c->warp_starts_f, warps, (int)m_list,
                                                (!d_qlist && c->have_warp_order) ? c->warp_order.p : nullptr, d_qlist ? 1 : 0,
                                                c->grid.dim[0], c->grid.dim[1], c->grid.dim[2], P, FF,
                                                store_rows ? c->feat.p : nullptr, c->counters.p);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace kpl
