// kpl_internal.h -- shared declarations of the CUDA implementation behind include/kpl.h.
// Built for sm_100a only, with -fmad=false: every float expression below is evaluated with
// separately rounded multiplies and adds, which is the arithmetic contract of the hot path
// (FLANN L2_Simple / Eigen / the reference's own helpers are FMA-free on the authors' platform).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/kpl.h"

namespace kpl {

// Canonical uniform grid.  cell = r_feat*(1+2^-20)/cells_per_radius; coordinates are computed in
// double so that two points closer than r_feat are never more than cells_per_radius cells apart.
// A batch of independent views (kpl_detect_batch) shares ONE grid: every view keeps the cells of its own
// canonical grid (origin = its bounding-box minimum) and the views are stacked along z with empty padding
// layers between them, so no kernel ever pairs points of two views and the (cell key, index) order inside a
// view is the one of its stand-alone run.
struct ViewDesc {
    double org[3];        // the view's own grid origin
    int32_t zoff, dimz;   // first stacked z layer / number of layers of the view
};
struct GridDesc {
    double org[3];
    double cell;
    int32_t dim[3];
    int32_t off[3];       // cell offset of a forced (slab) grid, 0 otherwise
    int32_t reach_feat;   // cells to search for radius_features
    int32_t reach_nms;    // cells to search for radius_nms
    int64_t ncells;
    // batch of views (nullptr / 0 for a single cloud)
    const ViewDesc* views;          // device, nviews entries
    const int32_t* layer_view;      // device, dim[2] entries: view of a stacked z layer, -1 = padding
    const int64_t* view_offsets;    // device, nviews + 1 entries: first point of each view
    int32_t nviews;
    // slab of a larger cloud: the x faces of the local grid that are NOT faces of the global grid.  A k-NN
    // search that would have to look beyond such a face cannot be exact (normals.cu)
    int32_t interior_lo, interior_hi;
    int32_t guard_cells;  // points in the outermost guard_cells columns next to an interior face are expected to be clipped
    int32_t owned_lo, owned_hi;   // local columns holding every point with a scoring role (0,0: not stated)
};

// One forest node, 8 bytes: thr_or_value + packed(child_block_offset << 10 | var); var == 1023 => leaf.
// Nodes are grouped in 32-byte blocks {node, left child, right child, unused}: see pack_forest (forest.cu).
struct __align__(8) PackedNode {
    float thr;
    uint32_t packed;
};
static constexpr uint32_t KPL_LEAF_VAR = 1023u;

struct Forest {
    int32_t ntrees = 0, nnodes = 0, var_count = 0, max_depth = 0;
    int32_t max_var = -1;                 // largest variable index any split reads (checked against annuli*bins per call)
    std::vector<int32_t> roots;           // host copy (block index of each root)
    PackedNode* d_nodes = nullptr;
    int32_t* d_roots = nullptr;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
};

struct HostForestArrays {
    std::vector<int32_t> roots, var, left, right;
    std::vector<float> thr, value;
    int32_t var_count = 0;
};

// forest_yaml.cpp
int parse_forest_yaml(const char* path, HostForestArrays& out, std::string& err);

}  // namespace kpl

struct kpl_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    kpl_params params;
    kpl::Forest forest;
    kpl::GridDesc grid;
    kpl_timings timings;
    kpl_stats stats;
    std::string err;
    int launches = 0;

    // device buffers (grown on demand, never shrunk)
    kpl::DevBuf<float4> in_xyz, in_nrm;          // staged host input (host API only)
    kpl::DevBuf<uint8_t> in_role, s_role;
    kpl::DevBuf<uint32_t> key_a, key_b, idx_a, idx_b;
    kpl::DevBuf<uint8_t> cub_tmp;
    kpl::DevBuf<int32_t> cell_start;
    kpl::DevBuf<int32_t> row_warps_n, row_offset_n;   // normal kernels' work list: warps per cell row and their prefix sum
    int nwarps_norm = 0, nwarps_feat = 0;         // sizes of work_n / warp_starts for the grid in place
    kpl::DevBuf<int2> work_n;                     // (first sorted position, count <= 32) per warp of the normal kernels
    // feature kernel: queries (sorted positions) in Hilbert order of a sub-cell lattice, 32 consecutive ones per warp
    // (grid.cu: build_lists); warp w takes entries [warp_starts[w], warp_starts[w + 1]) of the order.  _all / _n: every
    // point (cooperative k-NN normals; the feature kernel too when no roles are given), _role: the points with a scoring role
    kpl::DevBuf<uint64_t> ckey_a, ckey_b;
    kpl::DevBuf<int32_t> qorder_a, qorder_all, warp_starts_n, qorder_role, warp_starts_role;
    const int32_t* qorder_f = nullptr;            // the feature kernel's list for the grid in place (one of the above)
    const int32_t* warp_starts_f = nullptr;
    kpl::DevBuf<uint32_t> warp_order;             // launch order of the warps, most expensive first (few-wave launches)
    bool have_warp_order = false;
    kpl::DevBuf<float4> s_pos, s_nrm;            // cell-sorted positions (w = original index bits) / normals
    kpl::DevBuf<float> feat;                     // n x F, sorted order
    kpl::DevBuf<float> s_score, score;           // sorted order / original order
    kpl::DevBuf<uint8_t> flag;                   // keypoint flag, original order
    kpl::DevBuf<uint8_t> fragile;                // near-split flag of the forest walk, original order (forest.cuh)
    kpl::DevBuf<kpl::ViewDesc> views;            // kpl_detect_batch: per-view grids
    kpl::DevBuf<int32_t> layer_view;
    kpl::DevBuf<int64_t> view_offsets;
    kpl::DevBuf<int32_t> qlist;                  // kpl_features with an index subset: sorted positions of the queries
    kpl::DevBuf<uint8_t> s_state;                // draws-remove NMS state, sorted order
    kpl::DevBuf<int32_t> kp_idx;
    kpl::DevBuf<float> scratch_f;                // fetch / reorder scratch
    kpl::DevBuf<int32_t> scratch_i;
    // [0] feature pairs [1] candidate pairs [2] above th [3] n_kp [4] unscored [5] scored [6] scratch [7] near threshold
    // [8] fragile points [9] clipped k-NN searches (slab edge) [10] scratch [11..15] sizes of the query lists (grid.cu: build_lists)
    kpl::DevBuf<unsigned long long> counters;
    static constexpr int NCOUNTERS = 16;
    int syncs = 0;                               // cudaStreamSynchronize calls of the call in flight
    const float4* cur_xyz = nullptr;             // original-order inputs of the call in flight (device)
    const float4* cur_nrm = nullptr;
    float* d_bbox = nullptr;                     // 6 ordered-int encoded floats + flags
    int64_t last_n = 0;
    int last_F = 0;
    bool last_has_normals = false, last_has_features = false, last_has_fragile = false;
    bool fast_math = false;                      // feature kernel variant chosen by the arithmetic self-test
    bool keep_intermediates = false;             // kpl_set_keep_intermediates: materialise feature rows in kpl_detect*
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace kpl {

// ---- launch wrappers implemented in the .cu files; all enqueue on ctx->stream and return cudaError_t
cudaError_t launch_bbox(kpl_ctx* c, const float4* xyz, int64_t n, float* d_bbox /* 8 uint32 */);
cudaError_t build_grid(kpl_ctx* c, const float4* xyz, const float4* nrm_or_null, const uint8_t* role_or_null, int64_t n);
cudaError_t launch_normals_knn(kpl_ctx* c, int64_t n);
cudaError_t launch_normals_radius(kpl_ctx* c, int64_t n);
cudaError_t launch_flip_normals(kpl_ctx* c, int64_t n);
cudaError_t launch_check_normals(kpl_ctx* c, int64_t n, bool use_role);
bool normals_knn_uses_work_list(const kpl_params& P);
cudaError_t launch_bbox_init(kpl_ctx* c);
cudaError_t launch_gather_keypoints(kpl_ctx* c, const float4* d_xyz, const int32_t* d_kp_idx, int64_t nkp, float4* d_out);
cudaError_t launch_normals_integral_image(kpl_ctx* c, const float4* d_xyz, int W, int H, float smoothing_size, float4* d_out);
cudaError_t build_lists(kpl_ctx* c, int64_t n, int span_n, bool curve_normals, bool want_features, bool use_role, int64_t longest_first_below);
cudaError_t launch_count_occupied_cells(kpl_ctx* c, int64_t n, unsigned long long* d_out);
cudaError_t build_query_list(kpl_ctx* c, int64_t n, const int32_t* d_indices, int64_t m);
cudaError_t launch_scatter_rows(kpl_ctx* c, const float* d_rows, int64_t m, int width, float* d_out);
cudaError_t launch_view_ranges(kpl_ctx* c, int64_t n, int32_t* d_kp_idx, const int64_t* d_view_offsets, int nviews, int64_t* d_kp_offsets);
cudaError_t launch_bbox_views(kpl_ctx* c, const float4* xyz, int64_t n, const int64_t* d_view_offsets, int nviews, uint32_t* d_bbox);
// qlist != nullptr: features of the m_list listed sorted positions only (rows in list order, c->feat holds m_list x F)
cudaError_t launch_features(kpl_ctx* c, int64_t n, bool use_role, bool fuse_forest, bool store_rows, const int32_t* d_qlist = nullptr, int64_t m_list = 0);
#ifdef KPL_EXPERIMENTS
cudaError_t launch_forest(kpl_ctx* c, int64_t n, bool use_role);
#endif
cudaError_t launch_nms(kpl_ctx* c, int64_t n, bool use_role);
cudaError_t launch_nms_draws(kpl_ctx* c, int64_t n, bool use_role);
cudaError_t launch_compact(kpl_ctx* c, int64_t n, int32_t* d_kp_idx_out);

cudaError_t uniform_sample(kpl_ctx* c, const float4* xyz, int64_t n, float leaf, const float mn[3], const float mx[3],
                           int32_t* d_idx_out, std::string& err);
cudaError_t launch_nearest(kpl_ctx* c, const float4* d_queries, int64_t m, int32_t* d_idx, float* d_d2);
cudaError_t launch_radius_stats(kpl_ctx* c, int64_t n, double radius, int32_t* d_counts, unsigned long long* d_hash);
cudaError_t launch_radius_lists(kpl_ctx* c, int64_t n, double radius, const int32_t* d_queries, int64_t m,
                                const int64_t* d_offsets, int32_t* d_indices);
cudaError_t launch_unsort_rows(kpl_ctx* c, const float* d_sorted_rows, int64_t n, int width, float* d_out_orig_order);
cudaError_t launch_unsort_normals(kpl_ctx* c, int64_t n, float4* d_out_orig_order);

// detection phases (capi.cu), shared with the slab-sharded driver (shard.cu); all return kpl_status
int detect_check(kpl_ctx* ctx, int64_t n, bool sharded);
int detect_begin(kpl_ctx* ctx);
int detect_grid_phase(kpl_ctx* ctx, const float4* d_xyz, const float4* d_nrm, const uint8_t* d_role, int64_t n);
int detect_score_phase(kpl_ctx* ctx, bool normals_given, bool use_role, int64_t n);
int detect_nms_phase(kpl_ctx* ctx, bool use_role, int64_t n, int32_t* d_kp_out);
int detect_finish(kpl_ctx* ctx, int64_t n, int64_t* n_kp_out);

template <typename T>
cudaError_t ensure(DevBuf<T>& b, size_t n)
{
    if (n <= b.cap) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&b.p, want * sizeof(T));
    if (e == cudaSuccess) b.cap = want;
    return e;
}

}  // namespace kpl
