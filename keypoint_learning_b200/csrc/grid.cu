// grid.cu -- uniform-grid spatial hash: the replacement of pcl::search::KdTree / FLANN
// KDTreeSingleIndex on the detection path (reference call sites impl/KeypointLearning.hpp:213,334
// and main_test_detector.cpp:166-169).
//
// Layout in HBM after build_grid():
//   s_pos[i]   float4  cell-sorted position, .w = bit pattern of the original point index
//   s_nrm[i]   float4  cell-sorted normal (when given), .w = curvature / unused
//   key_b[i]   u32     canonical cell key (cz*dimy + cy)*dimx + cx of sorted point i
//   idx_b[i]   u32     original index of sorted point i
//   cell_start[k], k in [0, ncells]   first sorted position whose key is >= k  (monotone, so any
//              run of cells along x is ONE contiguous range [cell_start[k0], cell_start[k1+1]))
// The sort is a stable LSD radix sort (cub::DeviceRadixSort over the significant key bits), so
// inside a cell points stay in ascending original index: sorted position == canonical
// (cell key, index) order, which is the order feature votes are accumulated in.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/reverse_iterator.h>
#include <thrust/iterator/counting_iterator.h>
#include "kpl_internal.h"

namespace kpl {

__device__ __forceinline__ uint32_t enc_float(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// bbox[0..2] = min (encoded), bbox[3..5] = max (encoded), bbox[6] = non-finite flag
__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ xyz, int64_t n, uint32_t* __restrict__ bbox)
{
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    uint32_t bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = __ldg(xyz + i);
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { bad = 1; continue; }
        uint32_t e[3] = {enc_float(p.x), enc_float(p.y), enc_float(p.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], e[a]); hi[a] = max(hi[a], e[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]);
        hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
    }
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(bbox + a, lo[a]); atomicMax(bbox + 3 + a, hi[a]); }
        if (bad) atomicOr(bbox + 6, 1u);
    }
}

__global__ void bbox_init_kernel(uint32_t* bbox)
{
    int t = threadIdx.x;
    if (t < 3) bbox[t] = 0xFFFFFFFFu;
    else if (t < 8) bbox[t] = 0u;
}

cudaError_t launch_bbox(kpl_ctx* c, const float4* xyz, int64_t n, float* d_bbox)
{
    bbox_init_kernel<<<1, 32, 0, c->stream>>>((uint32_t*)d_bbox);
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    if (blocks < 1) blocks = 1;
    bbox_kernel<<<blocks, 256, 0, c->stream>>>(xyz, n, (uint32_t*)d_bbox);
    c->launches += 2;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) cell_key_kernel(const float4* __restrict__ xyz, int64_t n, GridDesc g,
                                                       uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                                       uint32_t* __restrict__ bbox_flags)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(xyz + i);
    float v[3] = {p.x, p.y, p.z};
    int cc[3];
    bool out = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double q = floor(__ddiv_rn(__dsub_rn((double)v[a], g.org[a]), g.cell)) - (double)g.off[a];
        if (!(q >= 0.0)) { out = true; q = 0.0; }
        if (q > (double)(g.dim[a] - 1)) { out = true; q = (double)(g.dim[a] - 1); }
        cc[a] = (int)q;
    }
    if (out) atomicOr(bbox_flags + 7, 1u);
    keys[i] = (uint32_t)(((int64_t)cc[2] * g.dim[1] + cc[1]) * g.dim[0] + cc[0]);
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* __restrict__ p, int64_t n, int32_t v)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void __launch_bounds__(256) reorder_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm,
                                                      const uint8_t* __restrict__ role,
                                                      const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sidx, int64_t n,
                                                      float4* __restrict__ s_pos, float4* __restrict__ s_nrm, uint8_t* __restrict__ s_role,
                                                      int32_t* __restrict__ cell_start)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o = sidx[i];
    float4 p = __ldg(xyz + o);
    p.w = __uint_as_float(o);
    s_pos[i] = p;
    if (nrm) s_nrm[i] = __ldg(nrm + o);
    if (role) s_role[i] = role[o];
    uint32_t k = skey[i];
    if (i == 0 || skey[i - 1] != k) cell_start[k] = (int32_t)i;
}

struct MinOp {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a < b ? a : b; }
};

cudaError_t build_grid(kpl_ctx* c, const float4* xyz, const float4* nrm, const uint8_t* role, int64_t n)
{
    const GridDesc& g = c->grid;
    cudaError_t e;
    if ((e = ensure(c->key_a, n)) || (e = ensure(c->key_b, n)) || (e = ensure(c->idx_a, n)) || (e = ensure(c->idx_b, n)) ||
        (e = ensure(c->s_pos, n)) || (e = ensure(c->s_nrm, n)) || (e = ensure(c->s_role, n)) ||
        (e = ensure(c->cell_start, (size_t)g.ncells + 1)))
        return e;
    int end_bit = 1;
    while (end_bit < 32 && (1ll << end_bit) < g.ncells) end_bit++;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, c->key_a.p, c->key_b.p, c->idx_a.p, c->idx_b.p, (int)n, 0, end_bit, c->stream);
    typedef thrust::reverse_iterator<int32_t*> rev_t;
    rev_t rbegin(c->cell_start.p + g.ncells + 1);
    cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, rbegin, rbegin, MinOp(), (int)(g.ncells + 1), c->stream);
    if ((e = ensure(c->cub_tmp, std::max(sort_bytes, scan_bytes)))) return e;

    int blocks = (int)((n + 255) / 256);
    cell_key_kernel<<<blocks, 256, 0, c->stream>>>(xyz, n, g, c->key_a.p, c->idx_a.p, (uint32_t*)c->d_bbox);
    size_t tmp = c->cub_tmp.cap;
    if ((e = cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->key_a.p, c->key_b.p, c->idx_a.p, c->idx_b.p, (int)n, 0, end_bit, c->stream)))
        return e;
    int fblocks = (int)std::min<int64_t>((g.ncells + 1 + 255) / 256, 148 * 16);
    fill_i32_kernel<<<fblocks, 256, 0, c->stream>>>(c->cell_start.p, g.ncells + 1, (int32_t)n);
    reorder_kernel<<<blocks, 256, 0, c->stream>>>(xyz, nrm, role, c->key_b.p, c->idx_b.p, n, c->s_pos.p, c->s_nrm.p, c->s_role.p, c->cell_start.p);
    tmp = c->cub_tmp.cap;
    if ((e = cub::DeviceScan::InclusiveScan(c->cub_tmp.p, tmp, rbegin, rbegin, MinOp(), (int)(g.ncells + 1), c->stream))) return e;
    c->launches += 3 + ((end_bit + 7) / 8) + 2 + 2;  // key, fill, reorder + radix passes (+histogram) + scan
    return cudaGetLastError();
}

// ---- keypoint compaction: ascending original index (keypoints_indices_, hpp:252-253) ----------
cudaError_t launch_compact(kpl_ctx* c, int64_t n, int32_t* d_kp_idx_out)
{
    size_t bytes = 0;
    thrust::counting_iterator<int32_t> it(0);
    int32_t* d_cnt = (int32_t*)(c->counters.p + 3);
    cub::DeviceSelect::Flagged(nullptr, bytes, it, c->flag.p, d_kp_idx_out, d_cnt, (int)n, c->stream);
    cudaError_t e;
    if ((e = ensure(c->cub_tmp, bytes))) return e;
    bytes = c->cub_tmp.cap;
    if ((e = cub::DeviceSelect::Flagged(c->cub_tmp.p, bytes, it, c->flag.p, d_kp_idx_out, d_cnt, (int)n, c->stream))) return e;
    c->launches += 2;
    return cudaGetLastError();
}

// ---- small reorder helpers (sorted order <-> original order) ---------------------------------
__global__ void __launch_bounds__(256) unsort_rows_kernel(const float* __restrict__ rows, const uint32_t* __restrict__ sidx,
                                                          int64_t n, int width, float* __restrict__ out)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = n * width;
    if (t >= total) return;
    int64_t i = t / width;
    int f = (int)(t - i * width);
    out[(int64_t)sidx[i] * width + f] = rows[t];
}
cudaError_t launch_unsort_rows(kpl_ctx* c, const float* rows, int64_t n, int width, float* out)
{
    int64_t total = n * width;
    unsort_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(rows, c->idx_b.p, n, width, out);
    c->launches++;
    return cudaGetLastError();
}
__global__ void __launch_bounds__(256) unsort_normals_kernel(const float4* __restrict__ s_nrm, const uint32_t* __restrict__ sidx,
                                                             int64_t n, float4* __restrict__ out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[sidx[i]] = s_nrm[i];
}
cudaError_t launch_unsort_normals(kpl_ctx* c, int64_t n, float4* out)
{
    unsort_normals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->s_nrm.p, c->idx_b.p, n, out);
    c->launches++;
    return cudaGetLastError();
}
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ rows, const int32_t* __restrict__ indices,
                                                          int64_t m, int width, float* __restrict__ out)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= m * width) return;
    int64_t k = t / width;
    int f = (int)(t - k * width);
    out[t] = rows[(int64_t)indices[k] * width + f];
}
cudaError_t launch_gather_rows(kpl_ctx* c, const float* rows, const int32_t* indices, int64_t m, int width, float* out)
{
    int64_t total = m * width;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(rows, indices, m, width, out);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace kpl
