// Micro-benchmark: issue/pipe rate of packed FP32 (FADD2/FMUL2/FFMA2) against scalar FADD/FMUL/FFMA on
// sm_100a, plus a bit-exactness check of the packed forms against the scalar IEEE RN intrinsics.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pk(float a, float b){ uint64_t r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(uint64_t v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v));}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b){ uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b){ uint64_t r; asm volatile("sub.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b){ uint64_t r; asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){ uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}

template <int MODE>
__global__ void __launch_bounds__(256) rate(float seed, uint64_t one2, float* out, int iters)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint64_t p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), p4 = pk(a1, a2), p5 = pk(a3, a4), p6 = pk(a5, a6), p7 = pk(a7, a0);
    const float m = 1.0000001f, c = 1e-7f;
    const uint64_t M = pk(m, m), C = pk(c, c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
                             a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c); }
            if (MODE == 1) { a0 = __fadd_rn(a0, c); a1 = __fadd_rn(a1, c); a2 = __fadd_rn(a2, c); a3 = __fadd_rn(a3, c);
                             a4 = __fadd_rn(a4, c); a5 = __fadd_rn(a5, c); a6 = __fadd_rn(a6, c); a7 = __fadd_rn(a7, c); }
            if (MODE == 2) { a0 = __fmul_rn(a0, m); a1 = __fmul_rn(a1, m); a2 = __fmul_rn(a2, m); a3 = __fmul_rn(a3, m);
                             a4 = __fmul_rn(a4, m); a5 = __fmul_rn(a5, m); a6 = __fmul_rn(a6, m); a7 = __fmul_rn(a7, m); }
            if (MODE == 3) { p0 = fma2(p0, M, C); p1 = fma2(p1, M, C); p2 = fma2(p2, M, C); p3 = fma2(p3, M, C);
                             p4 = fma2(p4, M, C); p5 = fma2(p5, M, C); p6 = fma2(p6, M, C); p7 = fma2(p7, M, C); }
            if (MODE == 4) { p0 = add2(p0, C); p1 = add2(p1, C); p2 = add2(p2, C); p3 = add2(p3, C);
                             p4 = add2(p4, C); p5 = add2(p5, C); p6 = add2(p6, C); p7 = add2(p7, C); }
            if (MODE == 5) { p0 = mul2(p0, M); p1 = mul2(p1, M); p2 = mul2(p2, M); p3 = mul2(p3, M);
                             p4 = mul2(p4, M); p5 = mul2(p5, M); p6 = mul2(p6, M); p7 = mul2(p7, M); }
            // MODE 6: packed FP mixed with integer ALU work (do they dual-issue across pipes?)
            if (MODE == 6) { p0 = fma2(p0, M, C); p1 = fma2(p1, M, C); p2 = fma2(p2, M, C); p3 = fma2(p3, M, C);
                             a0 = __int_as_float(__float_as_int(a0) * 3 + 1); a1 = __int_as_float(__float_as_int(a1) ^ (__float_as_int(a0) >> 3));
                             a2 = __int_as_float(__float_as_int(a2) + __float_as_int(a1)); a3 = __int_as_float(__float_as_int(a3) ^ (__float_as_int(a2) << 1)); }
            if (MODE == 7) { a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
                             a0 = __int_as_float(__float_as_int(a0) * 3 + 1); a1 = __int_as_float(__float_as_int(a1) ^ (__float_as_int(a0) >> 3));
                             a2 = __int_as_float(__float_as_int(a2) + __float_as_int(a1)); a3 = __int_as_float(__float_as_int(a3) ^ (__float_as_int(a2) << 1)); }
        }
    }
    float x0, x1; float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    upk(p0, x0, x1); s += x0 + x1; upk(p1, x0, x1); s += x0 + x1; upk(p2, x0, x1); s += x0 + x1; upk(p3, x0, x1); s += x0 + x1;
    upk(p4, x0, x1); s += x0 + x1; upk(p5, x0, x1); s += x0 + x1; upk(p6, x0, x1); s += x0 + x1; upk(p7, x0, x1); s += x0 + x1;
    if (s == 12345.678f) out[0] = s;
}

__global__ void exact(const float* a, const float* b, const float* c, int n, unsigned* bad)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    uint64_t A = pk(a[2 * i], a[2 * i + 1]), B = pk(b[2 * i], b[2 * i + 1]), Cc = pk(c[2 * i], c[2 * i + 1]);
    float r0, r1; unsigned e = 0;
    upk(add2(A, B), r0, r1); e += (__float_as_uint(r0) != __float_as_uint(__fadd_rn(a[2 * i], b[2 * i]))) + (__float_as_uint(r1) != __float_as_uint(__fadd_rn(a[2 * i + 1], b[2 * i + 1])));
    upk(sub2(A, B), r0, r1); e += (__float_as_uint(r0) != __float_as_uint(__fsub_rn(a[2 * i], b[2 * i]))) + (__float_as_uint(r1) != __float_as_uint(__fsub_rn(a[2 * i + 1], b[2 * i + 1])));
    upk(mul2(A, B), r0, r1); e += (__float_as_uint(r0) != __float_as_uint(__fmul_rn(a[2 * i], b[2 * i]))) + (__float_as_uint(r1) != __float_as_uint(__fmul_rn(a[2 * i + 1], b[2 * i + 1])));
    upk(fma2(A, B, Cc), r0, r1); e += (__float_as_uint(r0) != __float_as_uint(__fmaf_rn(a[2 * i], b[2 * i], c[2 * i]))) + (__float_as_uint(r1) != __float_as_uint(__fmaf_rn(a[2 * i + 1], b[2 * i + 1], c[2 * i + 1])));
    if (e) atomicAdd(bad, e);
}

template <int MODE> void run(const char* name, int ops_per_inst)
{
    float* out; cudaMalloc(&out, 4);
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate<MODE><<<blocks, 256>>>(1.0f, 0x3f8000003f800000ull, out, 16);
    cudaEventRecord(e0);
    rate<MODE><<<blocks, 256>>>(1.0f, 0x3f8000003f800000ull, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)blocks * 256 / 32 * iters * 64;   // warp instructions of the measured kind (modes 6/7: 32 FP + 32 int)
    printf("%-28s %8.3f ms  %7.2f Gwarp-inst/s  (%d flop-lanes/inst)  err=%s\n", name, ms, inst / ms * 1e-6, ops_per_inst, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    run<0>("FFMA  scalar", 1); run<1>("FADD  scalar", 1); run<2>("FMUL  scalar", 1);
    run<3>("FFMA2 packed", 2); run<4>("FADD2 packed", 2); run<5>("FMUL2 packed", 2);
    run<6>("FFMA2 + int (32+32)", 2); run<7>("FFMA + int (32+32)", 1);
    const int n = 1 << 22;
    float *a, *b, *c; unsigned* bad;
    cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4); cudaMallocManaged(&c, n * 4); cudaMallocManaged(&bad, 4);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; uint32_t ex = 100 + (s >> 8) % 56; uint32_t m = (s * 2654435761u) & 0x7FFFFFu; uint32_t sg = (s >> 3) & 1u;
                       uint32_t bits = (sg << 31) | (ex << 23) | m; float f; memcpy(&f, &bits, 4); return f; };
    for (int i = 0; i < n; ++i) { a[i] = rnd(); b[i] = rnd(); c[i] = rnd(); }
    *bad = 0;
    exact<<<(n / 2 + 255) / 256, 256>>>(a, b, c, n, bad);
    cudaDeviceSynchronize();
    printf("packed-vs-scalar RN mismatches over %d lanes x 4 ops: %u (%s)\n", n, *bad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
