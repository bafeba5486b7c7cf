// organized.cu -- normals of an ORGANIZED cloud: pcl::IntegralImageNormalEstimation with SIMPLE_3D_GRADIENT and
// setNormalSmoothingSize(5.0), the branch KeypointLearningDetector::initCompute takes when no normals were set and the
// surface is organized (impl/KeypointLearning.hpp:138-145).
//
// [3P-recalled] PCL 1.8.0 features/impl/integral_image_normal.hpp (computeFeature: depth-change map, two-pass distance
// map; computeFeatureFull with BORDER_POLICY_IGNORE and no depth-dependent smoothing; computePointNormal) and
// features/impl/integral_image2D.hpp (IntegralImage2D<float,3>: sums in double, non-finite elements skipped).  The
// estimator is PCL code that is absent from this build, so this is a restatement from the pinned version; the tests hold
// it bit for bit to an independent CPU restatement of the same PCL sources (tests/test_gpu_parity.py).
//
// The two distance-map passes and the integral image are recurrences with a fixed evaluation order; they are evaluated
// as WAVEFRONTS by one thread block (every element with the same number of the recurrence's longest dependency chain is
// independent), which reproduces the sequential results bit for bit -- including PCL's reads of the neighbouring row at
// the first / last column.  An organized cloud is an image (<= a few M pixels): these passes are microseconds to a few
// milliseconds and never on the hot path of the detector.
#include <cmath>
#include "kpl_internal.h"
#include "kpl_math.cuh"

namespace kpl {

__global__ void __launch_bounds__(256) ii_change_kernel(const float4* __restrict__ xyz, int W, int H, float factor, unsigned char* __restrict__ change)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)(W - 1) * (H - 1)) return;
    const int ri = (int)(t / (W - 1)), ci = (int)(t - (int64_t)ri * (W - 1));
    const int64_t index = (int64_t)ri * W + ci;
    const float depth = __ldg(&xyz[index].z), depthR = __ldg(&xyz[index + 1].z), depthD = __ldg(&xyz[index + W].z);
    const float lim = __fmul_rn(__fmul_rn(factor, __fadd_rn(fabsf(depth), 1.0f)), 2.0f);
    // only zeros are ever written: the order of the writes does not matter
    if (fabsf(__fsub_rn(depth, depthR)) > lim || !isfinite(depth) || !isfinite(depthR)) { change[index] = 0; change[index + 1] = 0; }
    if (fabsf(__fsub_rn(depth, depthD)) > lim || !isfinite(depth) || !isfinite(depthD)) { change[index] = 0; change[index + W] = 0; }
}

__global__ void __launch_bounds__(256) ii_dist_init_kernel(const unsigned char* __restrict__ change, int64_t n, float far_value, float* __restrict__ dist)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dist[i] = change[i] == 0 ? 0.0f : far_value;
}

// Both passes of the distance map, one block.  Forward: element (ri, ci), ri >= 1, ci >= 1, reads (ri-1, ci-1..ci+1) and
// (ri, ci-1): all of them carry a smaller 2*ri + ci.  Backward: the mirror image.
__global__ void __launch_bounds__(1024) ii_distance_kernel(float* __restrict__ dist, int W, int H)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int t = 3; t <= 2 * (H - 1) + (W - 1); ++t) {
        const int rlo = max(1, (t - (W - 1) + 1) / 2), rhi = min(H - 1, (t - 1) / 2);
        for (int ri = rlo + tid; ri <= rhi; ri += nt) {
            const int ci = t - 2 * ri;
            const float* prev = dist + (int64_t)(ri - 1) * W;
            float* cur = dist + (int64_t)ri * W;
            const float upLeft = __fadd_rn(prev[ci - 1], 1.4f), up = __fadd_rn(prev[ci], 1.0f), upRight = __fadd_rn(prev[ci + 1], 1.4f);
            const float left = __fadd_rn(cur[ci - 1], 1.0f);
            const float m = fminf(fminf(upLeft, up), fminf(left, upRight));
            if (m < cur[ci]) cur[ci] = m;
        }
        __syncthreads();
    }
    for (int t = 0; t <= 2 * (H - 2) + (W - 2); ++t) {
        // mirrored coordinates r' = H-2-ri, c' = W-2-ci, wavefront 2*r' + c'
        const int rlo = max(0, (t - (W - 2) + 1) / 2), rhi = min(H - 2, t / 2);
        for (int rp = rlo + tid; rp <= rhi; rp += nt) {
            const int cp = t - 2 * rp;
            const int ri = H - 2 - rp, ci = W - 2 - cp;
            const float* next = dist + (int64_t)(ri + 1) * W;
            float* cur = dist + (int64_t)ri * W;
            // ci == 0 reads next[-1], the last element of the current row, as PCL does
            const float lowerLeft = __fadd_rn(next[ci - 1], 1.4f), lower = __fadd_rn(next[ci], 1.0f), lowerRight = __fadd_rn(next[ci + 1], 1.4f);
            const float right = __fadd_rn(cur[ci + 1], 1.0f);
            const float m = fminf(fminf(lowerLeft, lower), fminf(right, lowerRight));
            if (m < cur[ci]) cur[ci] = m;
        }
        __syncthreads();
    }
}

// IntegralImage2D<float,3>::computeIntegralImages, first order: I[r+1][c+1] = ((I[r][c+1] + I[r+1][c]) - I[r][c]) + v in
// double, v skipped when the element's float sum is not finite.  Wavefront over r + c, one block.  I: (H+1) x (W+1) x 3.
__global__ void __launch_bounds__(1024) ii_integral_kernel(const float4* __restrict__ xyz, int W, int H, double* __restrict__ I)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int S = W + 1;
    for (int64_t i = tid; i < (int64_t)S * 3; i += nt) I[i] = 0.0;                               // row 0
    for (int r = tid; r <= H; r += nt) { double* p = I + (int64_t)r * S * 3; p[0] = p[1] = p[2] = 0.0; }   // column 0
    __syncthreads();
    for (int t = 0; t <= (H - 1) + (W - 1); ++t) {
        const int rlo = max(0, t - (W - 1)), rhi = min(H - 1, t);
        for (int r = rlo + tid; r <= rhi; r += nt) {
            const int c = t - r;
            const float4 e = __ldg(xyz + (int64_t)r * W + c);
            const bool fin = isfinite(__fadd_rn(e.x, __fadd_rn(e.y, e.z)));
            const float ev[3] = {e.x, e.y, e.z};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double v = __dsub_rn(__dadd_rn(I[((int64_t)r * S + (c + 1)) * 3 + a], I[((int64_t)(r + 1) * S + c) * 3 + a]), I[((int64_t)r * S + c) * 3 + a]);
                if (fin) v = __dadd_rn(v, (double)ev[a]);
                I[((int64_t)(r + 1) * S + (c + 1)) * 3 + a] = v;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double ii_sum(const double* __restrict__ I, int S, int sx, int sy, int w, int h, int a)
{
    // getFirstOrderSum: (lower_right + upper_left - upper_right - lower_left), left to right
    return __dsub_rn(__dsub_rn(__dadd_rn(I[((int64_t)(sy + h) * S + sx + w) * 3 + a], I[((int64_t)sy * S + sx) * 3 + a]),
                               I[((int64_t)sy * S + sx + w) * 3 + a]),
                     I[((int64_t)(sy + h) * S + sx) * 3 + a]);
}

__global__ void __launch_bounds__(256) ii_normals_kernel(const float4* __restrict__ xyz, const float* __restrict__ dist, const double* __restrict__ I,
                                                         int W, int H, float smoothing_size, float vpx, float vpy, float vpz, float4* __restrict__ out)
{
    const int64_t index = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (index >= (int64_t)W * H) return;
    const float nanf_ = CUDART_NAN_F;
    float4 res = make_float4(nanf_, nanf_, nanf_, nanf_);
    const int ri = (int)(index / W), ci = (int)(index - (int64_t)ri * W);
    const int border = (int)smoothing_size;
    const float4 p = __ldg(xyz + index);
    if (ri >= border && ri < H - border && ci >= border && ci < W - border && isfinite(p.z)) {
        const float smoothing = fminf(dist[index], smoothing_size);
        if (smoothing > 2.0f) {
            const int rw = (int)smoothing, rh = rw, rw2 = rw / 2, rh2 = rh / 2, S = W + 1;
            double gx[3], gy[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                gx[a] = __dsub_rn(ii_sum(I, S, ci + rw2, ri - rh2, 1, rh, a), ii_sum(I, S, ci - rw2, ri - rh2, 1, rh, a));
                gy[a] = __dsub_rn(ii_sum(I, S, ci - rw2, ri + rh2, rw, 1, a), ii_sum(I, S, ci - rw2, ri - rh2, rw, 1, a));
            }
            // gradient_y.cross(gradient_x)
            const double nx_ = __dsub_rn(__dmul_rn(gy[1], gx[2]), __dmul_rn(gy[2], gx[1]));
            const double ny_ = __dsub_rn(__dmul_rn(gy[2], gx[0]), __dmul_rn(gy[0], gx[2]));
            const double nz_ = __dsub_rn(__dmul_rn(gy[0], gx[1]), __dmul_rn(gy[1], gx[0]));
            const double len = __dadd_rn(__dmul_rn(nx_, nx_), __dadd_rn(__dmul_rn(ny_, ny_), __dmul_rn(nz_, nz_)));
            if (len != 0.0) {
                const double s = __dsqrt_rn(len);
                float nx = __double2float_rn(__ddiv_rn(nx_, s)), ny = __double2float_rn(__ddiv_rn(ny_, s)), nz = __double2float_rn(__ddiv_rn(nz_, s));
                // flipNormalTowardsViewpoint
                const float vx = __fsub_rn(vpx, p.x), vy = __fsub_rn(vpy, p.y), vz = __fsub_rn(vpz, p.z);
                const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
                if (cos_theta < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
                res = make_float4(nx, ny, nz, nanf_);           // curvature = bad_point for this method
            }
        }
    }
    out[index] = res;
}

// d_xyz: width*height float4 (row-major, NaN = no measurement) -> d_out: (nx, ny, nz, curvature) per pixel
cudaError_t launch_normals_integral_image(kpl_ctx* c, const float4* d_xyz, int W, int H, float smoothing_size, float4* d_out)
{
    const int64_t n = (int64_t)W * H;
    cudaError_t e;
    DevBuf<uint8_t>& change = c->s_state;
    if ((e = ensure(change, (size_t)n)) || (e = ensure(c->s_score, (size_t)n)) ||
        (e = ensure(c->scratch_f, (size_t)(W + 1) * (H + 1) * 6 + 16)))
        return e;
    float* dist = c->s_score.p;
    double* I = reinterpret_cast<double*>(c->scratch_f.p);
    const float factor = 20.0f * 0.001f;                       // max_depth_change_factor_ (constructor default; the reference never sets it)
    if ((e = cudaMemsetAsync(change.p, 255, (size_t)n, c->stream))) return e;
    if (W > 1 && H > 1) {
        const int64_t m = (int64_t)(W - 1) * (H - 1);
        ii_change_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(d_xyz, W, H, factor, change.p);
    }
    ii_dist_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(change.p, n, (float)(W + H), dist);
    ii_distance_kernel<<<1, 1024, 0, c->stream>>>(dist, W, H);
    ii_integral_kernel<<<1, 1024, 0, c->stream>>>(d_xyz, W, H, I);
    const kpl_params& P = c->params;
    ii_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_xyz, dist, I, W, H, smoothing_size, P.viewpoint[0], P.viewpoint[1],
                                                                         P.viewpoint[2], d_out);
    c->launches += 5;
    return cudaGetLastError();
}

}  // namespace kpl
