// kpl_math.cuh -- device arithmetic of the hot path.
//
// Contract (must match the CPU oracle bit-for-bit, and is what the reference evaluates on an
// FMA-free x86-64 / MSVC fp:precise build): IEEE binary32, round-to-nearest, no contraction.
// The translation unit is compiled with -fmad=false; where a fused operation is wanted it is
// written explicitly with __fmaf_rn.  IEEE division / square root are the default nvcc
// -prec-div=true -prec-sqrt=true expansions.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <float.h>
#include <stdint.h>

namespace kpl {

// FLANN L2_Simple<float> (result += diff*diff over x,y,z): the radius-membership expression of
// searchForNeighbors (impl/KeypointLearning.hpp:334) and tree_->radiusSearch (:213).
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Eigen fixed-size-3 reduction order: a0 + (a1 + a2).
__device__ __forceinline__ float dot3_eigen(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fmul_rn(ax, bx), __fadd_rn(__fmul_rn(ay, by), __fmul_rn(az, bz)));
}

// ---- packed FP32 (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE RN operations per issued instruction ---
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// FLANN L2_Simple for two candidates at once: ((dx*dx) + dy*dy) + dz*dz, every product and sum rounded.
// The sums are written RN(m*1 + acc) with `one` = (1.0f, 1.0f) passed at RUN time: ptxas 12.9 contracts a
// packed mul feeding a packed add into FFMA2 even for .rn operands, which would change d2 and with it
// the neighbour sets; a product by an opaque 1.0 is exact, rounds once and cannot be folded.
__device__ __forceinline__ uint64_t dist2_x2(uint64_t qx, uint64_t qy, uint64_t qz, uint64_t cx, uint64_t cy, uint64_t cz, uint64_t one)
{
    const uint64_t dx = sub2(qx, cx), dy = sub2(qy, cy), dz = sub2(qz, cz);
    const uint64_t mx = mul2(dx, dx), my = mul2(dy, dy), mz = mul2(dz, dz);
    return fma2(mz, one, fma2(my, one, mx));     // RN(RN(mx + my) + mz); addition commutes bit-exactly
}

// ---- libm-independent trig: double-precision series, rounded to float once -----------------
__device__ __forceinline__ double atan_series(double z)
{
    double z2 = __dmul_rn(z, z);
    double s = __ddiv_rn(1.0, 39.0);
#pragma unroll 1
    for (int k = 18; k >= 0; --k) s = __dsub_rn(__ddiv_rn(1.0, (double)(2 * k + 1)), __dmul_rn(z2, s));
    return __dmul_rn(z, s);
}
__device__ __forceinline__ double atan_unit(double a)
{
    if (a > 0.41421356237309503) {
        double z = __ddiv_rn(__dsub_rn(a, 1.0), __dadd_rn(a, 1.0));
        return __dadd_rn(0.78539816339744828, atan_series(z));
    }
    return atan_series(a);
}
__device__ __forceinline__ float kpl_atan2f(float y, float x)
{
    if (isnan(x) || isnan(y)) return CUDART_NAN_F;
    double ax = fabs((double)x), ay = fabs((double)y), r;
    if (ax == 0.0 && ay == 0.0) r = 0.0;
    else if (ax >= ay) r = atan_unit(__ddiv_rn(ay, ax));
    else r = __dsub_rn(1.5707963267948966, atan_unit(__ddiv_rn(ax, ay)));
    if (signbit(x)) r = __dsub_rn(3.1415926535897931, r);
    if (signbit(y)) r = -r;
    return __double2float_rn(r);
}
__device__ __forceinline__ float kpl_cosf(float xf)
{
    double x = (double)xf, x2 = __dmul_rn(x, x), s = 1.0;
#pragma unroll 1
    for (int k = 13; k >= 1; --k)
        s = __dsub_rn(1.0, __dmul_rn(__ddiv_rn(x2, (double)((2 * k - 1) * (2 * k))), s));
    return __double2float_rn(s);
}
__device__ __forceinline__ float kpl_sinf(float xf)
{
    double x = (double)xf, x2 = __dmul_rn(x, x), s = 1.0;
#pragma unroll 1
    for (int k = 13; k >= 1; --k)
        s = __dsub_rn(1.0, __dmul_rn(__ddiv_rn(x2, (double)((2 * k) * (2 * k + 1))), s));
    return __double2float_rn(__dmul_rn(x, s));
}

// ---- PCL 1.8.0 common/impl/eigen.hpp: computeRoots2 / computeRoots / eigen33 (smallest) ------
__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3])
{
    roots[0] = 0.0f;
    float d = __double2float_rn(__dsub_rn((double)__fmul_rn(b, b), __dmul_rn(4.0, (double)c)));
    if (d < 0.0f) d = 0.0f;
    float sd = __fsqrt_rn(d);
    roots[2] = __fmul_rn(0.5f, __fadd_rn(b, sd));
    roots[1] = __fmul_rn(0.5f, __fsub_rn(b, sd));
}

__device__ __forceinline__ void compute_roots(const float m[9], float roots[3])
{
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    // left-to-right evaluation of the source expressions, every product and sum rounded
    float c0 = __fmul_rn(__fmul_rn(m00, m11), m22);
    c0 = __fadd_rn(c0, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, m01), m02), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m00, m12), m12));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m11, m02), m02));
    c0 = __fsub_rn(c0, __fmul_rn(__fmul_rn(m22, m01), m01));
    float c1 = __fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m01));
    c1 = __fadd_rn(c1, __fmul_rn(m00, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m02, m02));
    c1 = __fadd_rn(c1, __fmul_rn(m11, m22));
    c1 = __fsub_rn(c1, __fmul_rn(m12, m12));
    float c2 = __fadd_rn(__fadd_rn(m00, m11), m22);
    if (fabsf(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
    const float s_inv3 = 0.3333333432674407958984375f;   // (float)(1.0/3.0)
    const float s_sqrt3 = 1.73205077648162841796875f;    // sqrtf(3.0f)
    float c2_over_3 = __fmul_rn(c2, s_inv3);
    float a_over_3 = __fmul_rn(__fsub_rn(c1, __fmul_rn(c2, c2_over_3)), s_inv3);
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float t = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, c2_over_3), c2_over_3), c1);
    float half_b = __fmul_rn(0.5f, __fadd_rn(c0, __fmul_rn(c2_over_3, t)));
    float q = __fadd_rn(__fmul_rn(half_b, half_b), __fmul_rn(__fmul_rn(a_over_3, a_over_3), a_over_3));
    if (q > 0.0f) q = 0.0f;
    float rho = __fsqrt_rn(-a_over_3);
    float theta = __fmul_rn(kpl_atan2f(__fsqrt_rn(-q), half_b), s_inv3);
    float cos_theta = kpl_cosf(theta);
    float sin_theta = kpl_sinf(theta);
    roots[0] = __fadd_rn(c2_over_3, __fmul_rn(__fmul_rn(2.0f, rho), cos_theta));
    roots[1] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fadd_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    roots[2] = __fsub_rn(c2_over_3, __fmul_rn(rho, __fsub_rn(cos_theta, __fmul_rn(s_sqrt3, sin_theta))));
    float w;
    if (roots[0] >= roots[1]) { w = roots[0]; roots[0] = roots[1]; roots[1] = w; }
    if (roots[1] >= roots[2]) {
        w = roots[1]; roots[1] = roots[2]; roots[2] = w;
        if (roots[0] >= roots[1]) { w = roots[0]; roots[0] = roots[1]; roots[1] = w; }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o)
{
    o[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    o[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    o[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

// accu[9] = raw moment sums (xx,xy,xz,yy,yz,zz,x,y,z) over cnt neighbours, in neighbour order.
// PCL 1.8.0 computeMeanAndCovarianceMatrix + solvePlaneParameters + flipNormalTowardsViewpoint.
__device__ __forceinline__ float4 normal_from_moments(float accu[9], int cnt, float px, float py, float pz,
                                                      float vpx, float vpy, float vpz)
{
    if (cnt < 3) return make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
    float fn = (float)cnt;
#pragma unroll
    for (int i = 0; i < 9; ++i) accu[i] = __fdiv_rn(accu[i], fn);
    float cov[9];
    cov[0] = __fsub_rn(accu[0], __fmul_rn(accu[6], accu[6]));
    cov[1] = __fsub_rn(accu[1], __fmul_rn(accu[6], accu[7]));
    cov[2] = __fsub_rn(accu[2], __fmul_rn(accu[6], accu[8]));
    cov[4] = __fsub_rn(accu[3], __fmul_rn(accu[7], accu[7]));
    cov[5] = __fsub_rn(accu[4], __fmul_rn(accu[7], accu[8]));
    cov[8] = __fsub_rn(accu[5], __fmul_rn(accu[8], accu[8]));
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(cov[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = __fdiv_rn(cov[i], scale);
    float roots[3];
    compute_roots(m, roots);
    float lambda = __fmul_rn(roots[0], scale);
    m[0] = __fsub_rn(m[0], roots[0]); m[4] = __fsub_rn(m[4], roots[0]); m[8] = __fsub_rn(m[8], roots[0]);
    float v1[3], v2[3], v3[3];
    cross3(m + 0, m + 3, v1);
    cross3(m + 0, m + 6, v2);
    cross3(m + 3, m + 6, v3);
    float l1 = dot3_eigen(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
    float l2 = dot3_eigen(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]);
    float l3 = dot3_eigen(v3[0], v3[1], v3[2], v3[0], v3[1], v3[2]);
    float ex, ey, ez, l;
    if (l1 >= l2 && l1 >= l3) { ex = v1[0]; ey = v1[1]; ez = v1[2]; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { ex = v2[0]; ey = v2[1]; ez = v2[2]; l = l2; }
    else { ex = v3[0]; ey = v3[1]; ez = v3[2]; l = l3; }
    float s = __fsqrt_rn(l);
    ex = __fdiv_rn(ex, s); ey = __fdiv_rn(ey, s); ez = __fdiv_rn(ez, s);
    float eig_sum = __fadd_rn(__fadd_rn(cov[0], cov[4]), cov[8]);
    float curv = (eig_sum != 0.0f) ? fabsf(__fdiv_rn(lambda, eig_sum)) : 0.0f;
    float vx = __fsub_rn(vpx, px), vy = __fsub_rn(vpy, py), vz = __fsub_rn(vpz, pz);
    float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, ex), __fmul_rn(vy, ey)), __fmul_rn(vz, ez));
    if (cos_theta < 0.0f) { ex = -ex; ey = -ey; ez = -ez; }
    return make_float4(ex, ey, ez, curv);
}

// src/KeypointLearning.cpp:41-65 / :68-92 with float abs; dim = bin width, half = dim/2.
__device__ __forceinline__ void soft_bin(float v, float dim, float half, int n, int& idx, int& pair, float& w)
{
    int i = (int)floorf(__fdiv_rn(v, dim));
    if (i == n) i--;
    float center = __fadd_rn(__fmul_rn((float)i, dim), half);
    float ww = __fdiv_rn(__fsub_rn(v, center), dim);
    int p = (ww > 0.0f) ? i + 1 : i - 1;
    if (p == -1) p = 0;
    if (p == n) p = i;
    idx = i; pair = p; w = fabsf(ww);
}

}  // namespace kpl
