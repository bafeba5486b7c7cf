// Experiments tool: is the FMA-corrected rsqrt square root of the feature kernel (features.cu: fast_sqrt_core / fast_sqrt_x2)
// bit-identical to the IEEE square root on EVERY normal float below 2^40, and what does it return for zero / denormal inputs
// when the rsqrt argument is clamped to FLT_MIN?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o tools/build/probe_fast_sqrt tools/probe_fast_sqrt.cu
#include <cstdio>
#include <cstdint>
#include <cfloat>
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fast_sqrt_x2_clamped(uint64_t x, float x0, float x1)
{
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fmaxf(x0, FLT_MIN)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fmaxf(x1, FLT_MIN)));
    const uint64_t y = pack2(y0, y1);
    const uint64_t s = mul2(x, y), h = mul2(y, 0x3F0000003F000000ull);
    const uint64_t e = fma2(sub2(0ull, s), s, x);
    return fma2(e, h, s);
}
__global__ void probe(unsigned long long* bad_per_binade, float* worst_sub)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;   // 2^23 mantissas
    for (int ex = 1; ex < 127 + 40; ex += 2) {
        const float a = __uint_as_float(((uint32_t)ex << 23) | m), b = __uint_as_float(((uint32_t)(ex + 1) << 23) | m);
        float r0, r1;
        unpack2(fast_sqrt_x2_clamped(pack2(a, b), a, b), r0, r1);
        if (__float_as_uint(r0) != __float_as_uint(__fsqrt_rn(a))) atomicAdd(bad_per_binade + ex, 1ull);
        if (ex + 1 < 127 + 40 && __float_as_uint(r1) != __float_as_uint(__fsqrt_rn(b))) atomicAdd(bad_per_binade + ex + 1, 1ull);
    }
    // zero and denormals: magnitude of what comes back (must stay far below any annulus half width)
    const float d = __uint_as_float(m);            // exponent field 0: zero / denormal
    float r0, r1;
    unpack2(fast_sqrt_x2_clamped(pack2(d, 0.0f), d, 0.0f), r0, r1);
    if (!(fabsf(r0) <= 1e-18f) || r1 != 0.0f || !(r0 == r0)) atomicAdd(bad_per_binade + 0, 1ull);
    atomicMax((int*)worst_sub, __float_as_int(fabsf(r0)));
}
int main()
{
    unsigned long long* d_bad; float* d_w;
    cudaMalloc(&d_bad, 256 * 8); cudaMemset(d_bad, 0, 256 * 8);
    cudaMalloc(&d_w, 4); cudaMemset(d_w, 0, 4);
    probe<<<(1u << 23) / 256, 256>>>(d_bad, d_w);
    unsigned long long h[256]; float w;
    cudaMemcpy(h, d_bad, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(&w, d_w, 4, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    printf("zero/denormal inputs out of bounds: %llu, largest result %g\n", h[0], w);
    unsigned long long tot = 0; int first = -1, last = -1;
    for (int ex = 1; ex < 127 + 40; ++ex) if (h[ex]) { tot += h[ex]; if (first < 0) first = ex; last = ex; printf("binade 2^%d: %llu mismatches\n", ex - 127, h[ex]); }
    printf("total mismatches over [2^-126, 2^40): %llu (binades %d..%d)\n", tot, first - 127, last - 127);
    return 0;
}
