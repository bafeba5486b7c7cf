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
#include <cmath>
#include <cstdlib>
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

// view of point i: the last v with view_offsets[v] <= i
__device__ __forceinline__ int view_of_point(const int64_t* __restrict__ view_offsets, int nviews, int64_t i)
{
    int lo = 0, hi = nviews - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(view_offsets + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Per-view bounding boxes of a batch: bbox[8 * v + ..] laid out like the single-cloud box.
__global__ void __launch_bounds__(256) bbox_views_kernel(const float4* __restrict__ xyz, int64_t n, const int64_t* __restrict__ view_offsets,
                                                         int nviews, uint32_t* __restrict__ bbox)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool in = i < n;
    int v = -1;
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};      // neutral for lanes beyond n
    bool bad = false;
    if (in) {
        v = view_of_point(view_offsets, nviews, i);
        const float4 p = __ldg(xyz + i);
        bad = !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
        lo[0] = hi[0] = enc_float(p.x); lo[1] = hi[1] = enc_float(p.y); lo[2] = hi[2] = enc_float(p.z);
    }
    const unsigned inmask = __ballot_sync(0xFFFFFFFFu, in);
    if (!inmask) return;                                       // the whole warp lies beyond n
    // the lanes of a warp almost always share one view: one set of atomics per warp then
    const int v0 = __shfl_sync(0xFFFFFFFFu, v, __ffs(inmask) - 1);
    if (__all_sync(0xFFFFFFFFu, !in || (v == v0 && !bad))) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]); hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]); }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { atomicMin(bbox + 8 * v0 + a, lo[a]); atomicMax(bbox + 8 * v0 + 3 + a, hi[a]); }
        }
    } else if (in) {
        if (bad) atomicOr(bbox + 8 * v + 6, 1u);
        else {
#pragma unroll
            for (int a = 0; a < 3; ++a) { atomicMin(bbox + 8 * v + a, lo[a]); atomicMax(bbox + 8 * v + 3 + a, hi[a]); }
        }
    }
}

__global__ void bbox_views_init_kernel(uint32_t* bbox, int nviews)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nviews * 8) bbox[t] = ((t & 7) < 3) ? 0xFFFFFFFFu : 0u;
}

cudaError_t launch_bbox_views(kpl_ctx* c, const float4* xyz, int64_t n, const int64_t* d_view_offsets, int nviews, uint32_t* d_bbox)
{
    bbox_views_init_kernel<<<(nviews * 8 + 255) / 256, 256, 0, c->stream>>>(d_bbox, nviews);
    bbox_views_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(xyz, n, d_view_offsets, nviews, d_bbox);
    c->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_bbox_init(kpl_ctx* c)
{
    bbox_init_kernel<<<1, 32, 0, c->stream>>>((uint32_t*)c->d_bbox);
    c->launches++;
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
    double org[3] = {g.org[0], g.org[1], g.org[2]};
    int lo[3] = {0, 0, 0}, hi[3] = {g.dim[0] - 1, g.dim[1] - 1, g.dim[2] - 1};
    int zoff = 0;
    if (g.views) {                                   // batch: the point's cell in its own view grid, stacked along z
        const ViewDesc* V = g.views + view_of_point(g.view_offsets, g.nviews, i);
        org[0] = V->org[0]; org[1] = V->org[1]; org[2] = V->org[2];
        zoff = V->zoff; lo[2] = V->zoff; hi[2] = V->zoff + V->dimz - 1;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double q = floor(__ddiv_rn(__dsub_rn((double)v[a], org[a]), g.cell)) - (double)g.off[a] + (a == 2 ? (double)zoff : 0.0);
        if (!(q >= (double)lo[a])) { out = true; q = (double)lo[a]; }
        if (q > (double)hi[a]) { out = true; q = (double)hi[a]; }
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

// ---- warp work list of the normal kernels ---------------------------------------------------------------
// A warp works best when its 32 queries share one cell row and span few cells in x: then they form ONE
// group and no lane idles while another group is processed.  Consecutive sorted points do not have that
// property (a closed surface crosses a cell row in several separate places), so the queries are cut into
// runs -- maximal stretches of one row spanning at most span + 1 cells -- and every warp gets up to 32
// consecutive points of one run: work[w] = (first sorted position, count).  One thread walks one row.
template <bool FILL>
__global__ void __launch_bounds__(128) run_list_kernel(const int32_t* __restrict__ cell_start, int dimx, int64_t nrows, int span,
                                                       int32_t* __restrict__ row_warps, const int32_t* __restrict__ row_offset,
                                                       int2* __restrict__ work)
{
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int32_t* cs = cell_start + row * dimx;
    int warps = 0;
    int out = FILL ? row_offset[row] : 0;
    int run_x = -1, run_s = 0, run_e = 0;
    auto close_run = [&]() {
        for (int s = run_s; s < run_e; s += 32) {
            if (FILL) work[out++] = make_int2(s, min(32, run_e - s));
            ++warps;
        }
    };
    int prev = __ldg(cs);
    for (int x = 0; x < dimx; ++x) {
        const int next = __ldg(cs + x + 1);
        if (next > prev) {                                   // cell x holds points [prev, next)
            if (run_x < 0 || x - run_x > span) {
                if (run_x >= 0) close_run();
                run_x = x; run_s = prev;
            }
            run_e = next;
        }
        prev = next;
    }
    if (run_x >= 0) close_run();
    if (!FILL) row_warps[row] = warps;
}

// ---- query order of the feature kernel -----------------------------------------------------------------------
// A warp of the feature kernel takes 32 queries and walks the union of their neighbourhoods; every lane idles while
// the others vote for candidates it does not take.  The tighter the 32 queries sit together, the more alike their
// neighbour sets are -- and ANY assignment of queries to warps gives the same results, because each query accumulates
// its own votes in canonical candidate order.  So the queries are ordered along a Hilbert curve over a lattice of
// cell / 2^sub sub-cells (sub = 2: 4 x 4 x 4 per grid cell) and every warp takes 32 CONSECUTIVE queries of that order:
// nearly all warps are full (runs of a cell row filled 0.90 of the lanes), and the 32 queries of a warp are a
// compact patch of the surface (model on the 10 M-point scene, tools/sim_query_order.py: 0.69 -> 0.78 of the vote
// slots used).  Where the curve jumps -- two consecutive queries more than one cell apart: it left the surface and
// came back elsewhere -- the list is cut, so that a warp never holds two far-apart clusters (it would walk both
// neighbourhoods one after the other with most lanes idle, and such long-running warps set the duration of a
// single-wave launch).  Points without a scoring role are left out of the feature kernel's list; the cooperative k-NN
// normal kernel works through the same order over ALL points (normals.cu).
__device__ __forceinline__ uint64_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits)
{
    // J. Skilling, "Programming the Hilbert curve" (2004): axes -> transposed index, then interleave
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    uint64_t h = 0;
    for (int b = bits - 1; b >= 0; --b)
        h = (h << 3) | (uint64_t)((((X[0] >> b) & 1u) << 2) | (((X[1] >> b) & 1u) << 1) | ((X[2] >> b) & 1u));
    return h;
}

__global__ void __launch_bounds__(256) curve_key_kernel(const float4* __restrict__ s_pos, const uint32_t* __restrict__ skey, int64_t n, GridDesc g,
                                                        int sub, int shift, int bits,
                                                        uint64_t* __restrict__ keys, int32_t* __restrict__ pos, unsigned long long* __restrict__ d_n)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i == 0) *d_n = (unsigned long long)n;
    if (i >= n) return;
    const uint32_t k = __ldg(skey + i);
    const uint32_t t = k / (uint32_t)g.dim[0];
    int cc[3];
    cc[0] = (int)(k - t * (uint32_t)g.dim[0]);
    cc[2] = (int)(t / (uint32_t)g.dim[1]);
    cc[1] = (int)(t - (uint32_t)cc[2] * (uint32_t)g.dim[1]);
    double org[3] = {g.org[0], g.org[1], g.org[2]};
    int zoff = 0;
    if (g.views) {
        const int v = __ldg(g.layer_view + cc[2]);
        if (v >= 0) { const ViewDesc* V = g.views + v; org[0] = V->org[0]; org[1] = V->org[1]; org[2] = V->org[2]; zoff = V->zoff; }
    }
    const float4 p = __ldg(s_pos + i);
    const float v3[3] = {p.x, p.y, p.z};
    uint32_t f[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // position inside the cell in [0, 1): only the ORDER of the queries depends on it, never a result
        const double fr = ((double)v3[a] - org[a]) / g.cell - (double)(cc[a] + g.off[a] - (a == 2 ? zoff : 0));
        const int sc = min(max((int)(fr * (double)(1 << sub)), 0), (1 << sub) - 1);
        f[a] = (((uint32_t)cc[a] << sub) | (uint32_t)sc) >> shift;
    }
    keys[i] = hilbert3(f[0], f[1], f[2], bits);
    pos[i] = (int32_t)i;
}
__global__ void widen_count_kernel(const int32_t* __restrict__ in, unsigned long long* __restrict__ out) { *out = (unsigned long long)*in; }

// head[i] = i when entry i of the query order starts a segment (first entry, or more than one cell from its predecessor in
// some axis), else 0: an inclusive max-scan turns it into the segment start of every entry
__global__ void __launch_bounds__(256) segment_head_kernel(const int32_t* __restrict__ qorder, const unsigned long long* __restrict__ d_nq, int64_t n,
                                                           const uint32_t* __restrict__ skey, int dimx, int dimy, int jump, int32_t* __restrict__ head)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int h = 0;
    if (i > 0 && i < (int64_t)*d_nq) {
        const uint32_t ka = __ldg(skey + __ldg(qorder + i - 1)), kb = __ldg(skey + __ldg(qorder + i));
        const uint32_t ta = ka / (uint32_t)dimx, tb = kb / (uint32_t)dimx;
        const int dx = (int)(ka - ta * (uint32_t)dimx) - (int)(kb - tb * (uint32_t)dimx);
        const int za = (int)(ta / (uint32_t)dimy), zb = (int)(tb / (uint32_t)dimy);
        const int dy = (int)(ta - (uint32_t)za * (uint32_t)dimy) - (int)(tb - (uint32_t)zb * (uint32_t)dimy);
        if (abs(dx) > jump || abs(dy) > jump || abs(za - zb) > jump) h = (int)i;
    }
    head[i] = h;
}
// flag[i] = 1 where a warp starts: every 32nd entry of a segment
__global__ void __launch_bounds__(256) warp_start_kernel(const int32_t* __restrict__ segstart, const unsigned long long* __restrict__ d_nq, int64_t n,
                                                         uint8_t* __restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i < (int64_t)*d_nq && (((int)i - segstart[i]) & 31) == 0) ? 1 : 0;
}
__global__ void terminate_starts_kernel(int32_t* __restrict__ starts, const int32_t* __restrict__ d_nwarps, const unsigned long long* __restrict__ d_nq)
{
    starts[*d_nwarps] = (int32_t)*d_nq;
}
struct MaxOp {
    __host__ __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

// Cost estimate of warp w of the query order = candidate points in the box the kernel will stage for it.
__global__ void __launch_bounds__(128) warp_cost_kernel(const int32_t* __restrict__ qorder, const int32_t* __restrict__ starts, int nwarps,
                                                        const uint32_t* __restrict__ skey, const int32_t* __restrict__ cell_start,
                                                        int dimx, int dimy, int dimz, int reach, float cellf, float rcull2,
                                                        uint32_t* __restrict__ cost, uint32_t* __restrict__ order)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwarps) return;
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {-1, -1, -1};
    for (int l = __ldg(starts + w); l < __ldg(starts + w + 1); ++l) {
        const uint32_t k = __ldg(skey + __ldg(qorder + l));
        const uint32_t t = k / (uint32_t)dimx;
        const int cz = (int)(t / (uint32_t)dimy);
        const int cc[3] = {(int)(k - t * (uint32_t)dimx), (int)(t - (uint32_t)cz * (uint32_t)dimy), cz};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], cc[a]); hi[a] = max(hi[a], cc[a]); }
    }
    unsigned c = 0;
    if (hi[0] >= 0)
        for (int zz = max(lo[2] - reach, 0); zz <= min(hi[2] + reach, dimz - 1); ++zz)
            for (int yy = max(lo[1] - reach, 0); yy <= min(hi[1] + reach, dimy - 1); ++yy) {
                const int gy = max(max(lo[1] - yy, yy - hi[1]) - 1, 0), gz = max(max(lo[2] - zz, zz - hi[2]) - 1, 0);
                const float gap2 = (float)(gy * gy + gz * gz) * cellf * cellf;
                if (!(gap2 < rcull2)) continue;
                const int rx = min((int)(sqrtf(rcull2 - gap2) / cellf) + 1, reach);
                const int64_t base = ((int64_t)zz * dimy + yy) * dimx;
                c += (unsigned)(__ldg(cell_start + base + min(hi[0] + rx, dimx - 1) + 1) - __ldg(cell_start + base + max(lo[0] - rx, 0)));
            }
    cost[w] = c;
    order[w] = (uint32_t)w;
}

// role flag of entry i of a query order
__global__ void __launch_bounds__(256) role_flag_kernel(const int32_t* __restrict__ qorder, const uint8_t* __restrict__ s_role, int64_t n,
                                                        uint8_t* __restrict__ flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = s_role[qorder[i]] & 1;
}

// Warp list of a query order of *d_nq entries (n = capacity): segments cut at the jumps of the curve, 32 entries per warp
// inside a segment.  starts gets *d_nwarps + 1 entries.  seg (n ints) and flag (n bytes) are scratch.
static cudaError_t make_warp_list(kpl_ctx* c, const int32_t* qorder, const unsigned long long* d_nq, int64_t n, int jump,
                                  int32_t* seg, uint8_t* flag, int32_t* starts, int32_t* d_nwarps)
{
    const GridDesc& g = c->grid;
    const unsigned pblocks = (unsigned)((n + 255) / 256);
    cudaError_t e;
    segment_head_kernel<<<pblocks, 256, 0, c->stream>>>(qorder, d_nq, n, c->key_b.p, g.dim[0], g.dim[1], jump, seg);
    size_t tmp = c->cub_tmp.cap;
    if ((e = cub::DeviceScan::InclusiveScan(c->cub_tmp.p, tmp, seg, seg, MaxOp(), (int)n, c->stream))) return e;
    warp_start_kernel<<<pblocks, 256, 0, c->stream>>>(seg, d_nq, n, flag);
    tmp = c->cub_tmp.cap;
    if ((e = cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, thrust::counting_iterator<int32_t>(0), flag, starts, d_nwarps, (int)n, c->stream))) return e;
    terminate_starts_kernel<<<1, 1, 0, c->stream>>>(starts, d_nwarps, d_nq);
    c->launches += 5;
    return cudaGetLastError();
}

// Builds, for the grid in place, the lists of the normal kernels and of the feature kernel with ONE host synchronisation
// for their sizes:
//   span_n >= 0        run list of the radius-normal kernel (c->work_n, c->nwarps_norm);
//   curve_normals      Hilbert order of ALL points + warp list for the cooperative k-NN kernel (c->qorder_all,
//                      c->warp_starts_n, c->nwarps_norm);
//   want_features      query order + warp list of the feature kernel (c->qorder_f / c->warp_starts_f, c->nwarps_feat): the
//                      points with a scoring role, in the same Hilbert order (without roles it IS the list of all points).
// The feature warps additionally get a launch order (c->warp_order), most expensive first, when their number is below
// longest_first_below: one warp of the feature kernel runs for milliseconds (32 queries x thousands of neighbours), so a
// launch of only a few waves -- one slab of an 8-GPU job is ~9 -- idles ~half a warp-time per SM slot at its end unless the
// cheap warps are the ones left for the tail.  Blocks retire independently, so the order of the launch never changes a
// result.
cudaError_t build_lists(kpl_ctx* c, int64_t n, int span_n, bool curve_normals, bool want_features, bool use_role, int64_t longest_first_below)
{
    const GridDesc& g = c->grid;
    const int64_t nrows = (int64_t)g.dim[1] * g.dim[2];
    const unsigned rblocks = (unsigned)((nrows + 127) / 128);
    const unsigned pblocks = (unsigned)((n + 255) / 256);
    cudaError_t e;
    c->nwarps_norm = c->nwarps_feat = 0;
    c->have_warp_order = false;
    c->qorder_f = nullptr; c->warp_starts_f = nullptr;
    if (n <= 0) return cudaSuccess;
    const bool curve = curve_normals || want_features;
    int32_t total_n = 0, total_f = 0, total_c = 0;
    size_t bytes = 0, need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)nrows + 1, c->stream);
    need = bytes;
    if (curve) {
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 64, c->stream);
        need = std::max(need, bytes);
        cub::DeviceScan::InclusiveScan(nullptr, bytes, (int32_t*)nullptr, (int32_t*)nullptr, MaxOp(), (int)n, c->stream);
        need = std::max(need, bytes);
        cub::DeviceSelect::Flagged(nullptr, bytes, thrust::counting_iterator<int32_t>(0), (uint8_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, c->stream);
        need = std::max(need, bytes);
        cub::DeviceSelect::Flagged(nullptr, bytes, (int32_t*)nullptr, (uint8_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, c->stream);
        need = std::max(need, bytes);
    }
    if ((e = ensure(c->cub_tmp, need))) return e;
    // ---- sizes
    if (span_n >= 0) {
        if ((e = ensure(c->row_warps_n, (size_t)nrows + 1)) || (e = ensure(c->row_offset_n, (size_t)nrows + 1))) return e;
        run_list_kernel<false><<<rblocks, 128, 0, c->stream>>>(c->cell_start.p, g.dim[0], nrows, span_n, c->row_warps_n.p, nullptr, nullptr);
        if ((e = cudaMemsetAsync(c->row_warps_n.p + nrows, 0, sizeof(int32_t), c->stream))) return e;
        size_t tmp = c->cub_tmp.cap;
        if ((e = cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->row_warps_n.p, c->row_offset_n.p, (int)nrows + 1, c->stream))) return e;
        if ((e = cudaMemcpyAsync(&total_n, c->row_offset_n.p + nrows, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream))) return e;
        c->launches += 3;
    }
    // counters: [11] all points [12] warps of the all-points list [13] queries with a scoring role [14] their warps
    unsigned long long* d_nall = c->counters.p + 11;
    int32_t* d_nwarps_all = reinterpret_cast<int32_t*>(c->counters.p + 12);
    unsigned long long* d_nq = c->counters.p + 13;
    int32_t* d_nwarps_q = reinterpret_cast<int32_t*>(c->counters.p + 14);
    if (curve) {
        if ((e = ensure(c->ckey_a, (size_t)n)) || (e = ensure(c->ckey_b, (size_t)n)) || (e = ensure(c->qorder_a, (size_t)n)) ||
            (e = ensure(c->qorder_all, (size_t)n)) || (e = ensure(c->warp_starts_n, (size_t)n + 2)))
            return e;
        // lattice: 2^sub sub-cells per cell and axis while the largest axis fits 21 bits (3 x 21 = 63-bit curve index);
        // grids beyond 2^21 cells along one axis drop low bits instead
        const int maxdim = std::max(g.dim[0], std::max(g.dim[1], g.dim[2]));
        int sub = 2, shift = 0, jump = 1;
#ifdef KPL_EXPERIMENTS
        if (const char* ev = getenv("KPL_SUB")) sub = atoi(ev);
        if (const char* ev = getenv("KPL_JUMP")) jump = atoi(ev);
#endif
        while (sub > 0 && ((int64_t)maxdim << sub) > (1ll << 21)) --sub;
        while ((((int64_t)maxdim << sub) >> shift) > (1ll << 21)) ++shift;
        int bits = 1;
        while ((1ll << bits) < (((int64_t)maxdim << sub) >> shift) + 1) ++bits;
        if ((e = cudaMemsetAsync(d_nall, 0, 4 * sizeof(unsigned long long), c->stream))) return e;
        curve_key_kernel<<<pblocks, 256, 0, c->stream>>>(c->s_pos.p, c->key_b.p, n, g, sub, shift, bits, c->ckey_a.p, c->qorder_a.p, d_nall);
        const int end_bit = 3 * bits;
        size_t tmp = c->cub_tmp.cap;
        if ((e = cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->ckey_a.p, c->ckey_b.p, c->qorder_a.p, c->qorder_all.p, (int)n, 0, end_bit, c->stream)))
            return e;
        c->launches += 2 + (end_bit + 7) / 8;
        // (the sort buffers are free now: scratch of the warp lists)
        int32_t* seg = c->qorder_a.p;
        uint8_t* flag = reinterpret_cast<uint8_t*>(c->ckey_a.p);
        if (curve_normals || !use_role) {
            if ((e = make_warp_list(c, c->qorder_all.p, d_nall, n, jump, seg, flag, c->warp_starts_n.p, d_nwarps_all))) return e;
            if ((e = cudaMemcpyAsync(&total_c, d_nwarps_all, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream))) return e;
        }
        if (want_features && use_role) {
            // the queries with a scoring role, in the same order
            if ((e = ensure(c->qorder_role, (size_t)n)) || (e = ensure(c->warp_starts_role, (size_t)n + 2))) return e;
            role_flag_kernel<<<pblocks, 256, 0, c->stream>>>(c->qorder_all.p, c->s_role.p, n, flag);
            int32_t* d_sel = reinterpret_cast<int32_t*>(c->counters.p + 15);
            tmp = c->cub_tmp.cap;
            if ((e = cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, c->qorder_all.p, flag, c->qorder_role.p, d_sel, (int)n, c->stream))) return e;
            widen_count_kernel<<<1, 1, 0, c->stream>>>(d_sel, d_nq);
            c->launches += 3;
            if ((e = make_warp_list(c, c->qorder_role.p, d_nq, n, jump, seg, flag, c->warp_starts_role.p, d_nwarps_q))) return e;
            if ((e = cudaMemcpyAsync(&total_f, d_nwarps_q, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream))) return e;
        }
    }
    if (span_n >= 0 || curve) {
        if ((e = cudaStreamSynchronize(c->stream))) return e;
        c->syncs++;
    }
    // ---- the run list itself
    if (span_n >= 0 && total_n > 0) {
        if ((e = ensure(c->work_n, (size_t)total_n + 1))) return e;
        run_list_kernel<true><<<rblocks, 128, 0, c->stream>>>(c->cell_start.p, g.dim[0], nrows, span_n, nullptr, c->row_offset_n.p, c->work_n.p);
        c->nwarps_norm = total_n;
        c->launches++;
    }
    if (curve_normals) c->nwarps_norm = total_c;
    if (want_features) {
        if (use_role) { c->qorder_f = c->qorder_role.p; c->warp_starts_f = c->warp_starts_role.p; c->nwarps_feat = total_f; }
        else { c->qorder_f = c->qorder_all.p; c->warp_starts_f = c->warp_starts_n.p; c->nwarps_feat = total_c; }
    }
    if (want_features && c->nwarps_feat > 1 && c->nwarps_feat < longest_first_below) {
        const int nwarps = c->nwarps_feat;
        if ((e = ensure(c->scratch_i, 3 * (size_t)nwarps + 16)) || (e = ensure(c->warp_order, (size_t)nwarps + 1))) return e;
        uint32_t* cost_a = (uint32_t*)c->scratch_i.p;
        uint32_t* cost_b = cost_a + nwarps;
        uint32_t* ord_a = cost_b + nwarps;
        bytes = 0;
        cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, cost_a, cost_b, ord_a, c->warp_order.p, nwarps, 0, 32, c->stream);
        if ((e = ensure(c->cub_tmp, bytes))) return e;
        const double r = (double)c->params.radius_features;
        warp_cost_kernel<<<(nwarps + 127) / 128, 128, 0, c->stream>>>(c->qorder_f, c->warp_starts_f, nwarps, c->key_b.p, c->cell_start.p, g.dim[0], g.dim[1],
                                                                       g.dim[2], g.reach_feat, (float)g.cell, (float)(r * r * (1.0 + 1e-5)), cost_a, ord_a);
        bytes = c->cub_tmp.cap;
        if ((e = cub::DeviceRadixSort::SortPairsDescending(c->cub_tmp.p, bytes, cost_a, cost_b, ord_a, c->warp_order.p, nwarps, 0, 32, c->stream))) return e;
        c->have_warp_order = true;
        c->launches += 1 + 5;
    }
    return cudaGetLastError();
}

// ---- occupied cells (kpl_normals sizes its k-NN grid from the data) ---------------------------------------
__global__ void __launch_bounds__(256) count_heads_kernel(const uint32_t* __restrict__ skey, int64_t n, unsigned long long* __restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool head = i < n && (i == 0 || skey[i - 1] != skey[i]);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, head);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)__popc(m));
}
cudaError_t launch_count_occupied_cells(kpl_ctx* c, int64_t n, unsigned long long* d_out)
{
    count_heads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->key_b.p, n, d_out);
    c->launches++;
    return cudaGetLastError();
}

// ---- query lists (computePointsForTrainingFeatures on an index subset, hpp:299-318) -----------------------
// indices[m] (original order) -> qlist[m]: their sorted positions in ascending order, perm[m]: the caller's row of
// each list entry.  Only n + O(m) words of scratch: nothing of size n x F is ever allocated for a subset.
__global__ void __launch_bounds__(256) inverse_perm_kernel(const uint32_t* __restrict__ sidx, int64_t n, int32_t* __restrict__ inv)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) inv[sidx[i]] = (int32_t)i;
}
__global__ void __launch_bounds__(256) query_pos_kernel(const int32_t* __restrict__ indices, int64_t m, const int32_t* __restrict__ inv,
                                                        uint32_t* __restrict__ qpos, uint32_t* __restrict__ row)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < m) { qpos[k] = (uint32_t)inv[indices[k]]; row[k] = (uint32_t)k; }
}
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ rows, const uint32_t* __restrict__ perm, int64_t m, int width,
                                                           float* __restrict__ out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= m * width) return;
    const int64_t j = t / width;
    const int f = (int)(t - j * width);
    out[(int64_t)perm[j] * width + f] = rows[t];
}
// d_indices: m original indices (device).  Leaves qlist in c->qlist[0..m) and perm in c->qlist[m..2m).
cudaError_t build_query_list(kpl_ctx* c, int64_t n, const int32_t* d_indices, int64_t m)
{
    cudaError_t e;
    if ((e = ensure(c->scratch_i, (size_t)n + 2 * (size_t)m + 16)) || (e = ensure(c->qlist, 2 * (size_t)m + 16))) return e;
    int32_t* inv = c->scratch_i.p;
    uint32_t* qpos = (uint32_t*)(c->scratch_i.p + n);
    uint32_t* row = qpos + m;
    uint32_t* qsorted = (uint32_t*)c->qlist.p;
    uint32_t* perm = qsorted + m;
    int end_bit = 1;
    while (end_bit < 32 && (1ll << end_bit) < n) end_bit++;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, qpos, qsorted, row, perm, (int)m, 0, end_bit, c->stream);
    if ((e = ensure(c->cub_tmp, bytes))) return e;
    inverse_perm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->idx_b.p, n, inv);
    query_pos_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(d_indices, m, inv, qpos, row);
    bytes = c->cub_tmp.cap;
    if ((e = cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, bytes, qpos, qsorted, row, perm, (int)m, 0, end_bit, c->stream))) return e;
    c->launches += 2 + (end_bit + 7) / 8 + 1;
    return cudaGetLastError();
}
cudaError_t launch_scatter_rows(kpl_ctx* c, const float* d_rows, int64_t m, int width, float* d_out)
{
    const int64_t total = m * width;
    scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_rows, (const uint32_t*)c->qlist.p + m, m, width, d_out);
    c->launches++;
    return cudaGetLastError();
}

// ---- batch of views: per-view keypoint ranges of the ascending concatenated index list ---------------------
// kp_offsets[v] = number of keypoints with concatenated index < view_offsets[v]; kp_idx is made view-local.
__global__ void __launch_bounds__(256) view_ranges_kernel(const int32_t* __restrict__ kp_idx, const int32_t* __restrict__ d_count,
                                                          const int64_t* __restrict__ view_offsets, int nviews, int64_t* __restrict__ kp_offsets)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nviews) return;
    const int cnt = *d_count;
    const int64_t bound = view_offsets[v];
    int lo = 0, hi = cnt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)kp_idx[mid] < bound) lo = mid + 1; else hi = mid;
    }
    kp_offsets[v] = lo;
}
__global__ void __launch_bounds__(256) view_localise_kernel(int32_t* __restrict__ kp_idx, const int32_t* __restrict__ d_count,
                                                            const int64_t* __restrict__ view_offsets, int nviews)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *d_count) return;
    const int32_t g = kp_idx[i];
    kp_idx[i] = (int32_t)(g - view_offsets[view_of_point(view_offsets, nviews, g)]);
}
cudaError_t launch_view_ranges(kpl_ctx* c, int64_t n, int32_t* d_kp_idx, const int64_t* d_view_offsets, int nviews, int64_t* d_kp_offsets)
{
    const int32_t* d_cnt = (const int32_t*)(c->counters.p + 3);
    view_ranges_kernel<<<(nviews + 1 + 255) / 256, 256, 0, c->stream>>>(d_kp_idx, d_cnt, d_view_offsets, nviews, d_kp_offsets);
    view_localise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_kp_idx, d_cnt, d_view_offsets, nviews);
    c->launches += 2;
    return cudaGetLastError();
}

// ---- the keypoint cloud: (x, y, z, response) of every keypoint (PointXYZI of hpp:246-253) --------------------
__global__ void __launch_bounds__(256) gather_keypoints_kernel(const float4* __restrict__ xyz, const float* __restrict__ score,
                                                               const int32_t* __restrict__ kp_idx, int64_t nkp, float4* __restrict__ out)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nkp) return;
    const int32_t i = kp_idx[k];
    float4 p = __ldg(xyz + i);
    p.w = score[i];
    out[k] = p;
}
cudaError_t launch_gather_keypoints(kpl_ctx* c, const float4* d_xyz, const int32_t* d_kp_idx, int64_t nkp, float4* d_out)
{
    gather_keypoints_kernel<<<(unsigned)((nkp + 255) / 256), 256, 0, c->stream>>>(d_xyz, c->score.p, d_kp_idx, nkp, d_out);
    c->launches++;
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

// ---- pcl::UniformSampling (src/main_test_detector.cpp:145-157) ----------------------------------
// One survivor per leaf-sized voxel, ijk = floor(p * (1/leaf)) in FP32.  [3P-recalled] PCL 1.8.0
// filters/impl/uniform_sampling.hpp keeps the point with the smaller
//     (p.getVector4fMap() - ijk.cast<float>()).squaredNorm()
// i.e. it measures the distance to the voxel INDEX vector taken as a point (not to the voxel centre), over the four
// floats (x, y, z, 1) - (i, j, k, 0), summed as an SSE2 packet reduction (d0 + d2) + (d1 + d3); a later point wins only
// if strictly closer, so ties go to the lower index.  centre = true selects the voxel centre (ijk + 0.5) * leaf instead
// (kpl_params.uniform_sampling_centre).  Survivors are reported in ascending original index (PCL emits them in
// unordered_map order, which is platform dependent).
// Same machinery as the neighbour grid: voxel keys, a stable radix sort, one scan of each run of equal keys.
__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ xyz, int64_t n, float inv_leaf, float leaf, bool centre,
                                                        int mbx, int mby, int mbz, unsigned long long dx, unsigned long long dy,
                                                        unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx, float* __restrict__ dist)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(xyz + i);
    const float fx = floorf(__fmul_rn(p.x, inv_leaf)), fy = floorf(__fmul_rn(p.y, inv_leaf)), fz = floorf(__fmul_rn(p.z, inv_leaf));
    const unsigned long long ix = (unsigned long long)((long long)fx - mbx), iy = (unsigned long long)((long long)fy - mby),
                             iz = (unsigned long long)((long long)fz - mbz);
    keys[i] = (iz * dy + iy) * dx + ix;
    idx[i] = (uint32_t)i;
    if (centre) {
        const float cx = __fmul_rn(__fadd_rn(fx, 0.5f), leaf), cy = __fmul_rn(__fadd_rn(fy, 0.5f), leaf), cz = __fmul_rn(__fadd_rn(fz, 0.5f), leaf);
        const float ex = __fsub_rn(cx, p.x), ey = __fsub_rn(cy, p.y), ez = __fsub_rn(cz, p.z);
        dist[i] = __fadd_rn(__fmul_rn(ex, ex), __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ez, ez)));
    } else {
        const float ex = __fsub_rn(p.x, fx), ey = __fsub_rn(p.y, fy), ez = __fsub_rn(p.z, fz);      // fourth component: 1 - 0
        dist[i] = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ez, ez)), __fadd_rn(__fmul_rn(ey, ey), 1.0f));
    }
}

__global__ void __launch_bounds__(256) voxel_pick_kernel(const unsigned long long* __restrict__ skeys, const uint32_t* __restrict__ sidx,
                                                         const float* __restrict__ dist, int64_t n, uint8_t* __restrict__ flag)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = skeys[i];
    if (i > 0 && skeys[i - 1] == k) return;                 // not the head of its run
    uint32_t best = sidx[i];
    float bd = dist[best];
    for (int64_t j = i + 1; j < n && skeys[j] == k; ++j) {  // the sort is stable: ascending original index inside a run
        const uint32_t o = sidx[j];
        const float d = dist[o];
        if (d < bd) { bd = d; best = o; }
    }
    flag[best] = 1;
}

cudaError_t uniform_sample(kpl_ctx* c, const float4* xyz, int64_t n, float leaf, const float mn[3], const float mx[3],
                           int32_t* d_idx_out, std::string& err)
{
    const float inv_leaf = 1.0f / leaf;
    long long mb[3], db[3];
    for (int a = 0; a < 3; ++a) {
        mb[a] = (long long)floorf(mn[a] * inv_leaf);
        db[a] = (long long)floorf(mx[a] * inv_leaf) - mb[a] + 1;
        if (mb[a] < -2000000000ll || mb[a] > 2000000000ll || db[a] < 1) { err = "leaf size too small for this cloud"; return cudaErrorInvalidValue; }
    }
    if ((double)db[0] * (double)db[1] * (double)db[2] > 1.8e19) { err = "leaf size too small for this cloud (more than 2^64 voxels)"; return cudaErrorInvalidValue; }
    cudaError_t e;
    DevBuf<uint8_t>& tmp = c->cub_tmp;
    // scratch: keys a/b (u64), idx a/b (u32), dist (f32), flags (u8) carved from scratch buffers
    if ((e = ensure(c->scratch_f, (size_t)n * 5 + 16)) || (e = ensure(c->scratch_i, (size_t)n * 2 + 16)) || (e = ensure(c->flag, (size_t)n))) return e;
    unsigned long long* key_a = reinterpret_cast<unsigned long long*>(c->scratch_f.p);
    unsigned long long* key_b = key_a + n;
    float* dist = c->scratch_f.p + 4 * (size_t)n + 8;
    uint32_t* idx_a = reinterpret_cast<uint32_t*>(c->scratch_i.p);
    uint32_t* idx_b = idx_a + n;
    int end_bit = 1;
    const double nvox = (double)db[0] * (double)db[1] * (double)db[2];
    while (end_bit < 64 && std::ldexp(1.0, end_bit) < nvox) end_bit++;
    size_t sort_bytes = 0, sel_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, key_a, key_b, idx_a, idx_b, (int)n, 0, end_bit, c->stream);
    thrust::counting_iterator<int32_t> it(0);
    int32_t* d_cnt = (int32_t*)(c->counters.p + 3);
    cub::DeviceSelect::Flagged(nullptr, sel_bytes, it, c->flag.p, d_idx_out, d_cnt, (int)n, c->stream);
    if ((e = ensure(tmp, std::max(sort_bytes, sel_bytes)))) return e;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    voxel_key_kernel<<<blocks, 256, 0, c->stream>>>(xyz, n, inv_leaf, leaf, c->params.uniform_sampling_centre != 0, (int)mb[0], (int)mb[1], (int)mb[2],
                                                    (unsigned long long)db[0], (unsigned long long)db[1], key_a, idx_a, dist);
    size_t bytes = tmp.cap;
    if ((e = cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_a, key_b, idx_a, idx_b, (int)n, 0, end_bit, c->stream))) return e;
    if ((e = cudaMemsetAsync(c->flag.p, 0, (size_t)n, c->stream))) return e;
    voxel_pick_kernel<<<blocks, 256, 0, c->stream>>>(key_b, idx_b, dist, n, c->flag.p);
    bytes = tmp.cap;
    if ((e = cub::DeviceSelect::Flagged(tmp.p, bytes, it, c->flag.p, d_idx_out, d_cnt, (int)n, c->stream))) return e;
    c->launches += 4 + (end_bit + 7) / 8 + 1;
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
}  // namespace kpl
