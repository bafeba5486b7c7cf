// forest.cu -- random-forest evaluation: replaces cv::ml::RTrees::predict(feat, result, PREDICT_SUM)
// and the score line of KeypointLearningDetector::runForest (impl/KeypointLearning.hpp:281-287).
//
// Nodes are re-laid out on the host in pre-order, 8 bytes each (PackedNode): the left child is the
// next node, the right child is node + right_offset, so a visit is one 8-byte load.  One thread
// evaluates one point; its feature row sits in shared memory as sf[var][thread] (bank == lane, no
// conflicts for the data-dependent var), and four trees are walked concurrently per thread so four
// independent node loads are in flight (the node arrays live in L2 / L1).
#include <algorithm>
#include "kpl_internal.h"
#include "kpl_math.cuh"
#include "forest.cuh"

namespace kpl {

static constexpr int FOREST_THREADS = 128;

__global__ void __launch_bounds__(FOREST_THREADS)
forest_kernel(const float* __restrict__ feat, const PackedNode* __restrict__ nodes, const int32_t* __restrict__ roots,
              int ntrees, int F, int n, const uint8_t* __restrict__ s_role, const float4* __restrict__ s_pos,
              const float4* __restrict__ s_nrm,
              float* __restrict__ s_score, float* __restrict__ score, unsigned long long* __restrict__ counters)
{
    extern __shared__ __align__(16) float sf[];
    const int tid = threadIdx.x;
    const int base = blockIdx.x * FOREST_THREADS;
    const int rows = min(FOREST_THREADS, n - base);
    const float* src = feat + (int64_t)base * F;
    for (int e = tid; e < rows * F; e += FOREST_THREADS) {
        int row = e / F, f = e - row * F;
        sf[f * FOREST_THREADS + row] = __ldg(src + e);
    }
    __syncthreads();
    const int i = base + tid;
    if (i >= n) return;
    const uint32_t orig = __float_as_uint(__ldg(s_pos + i).w);
    const float4 qn = __ldg(s_nrm + i);
    // unscored: halo role, or no finite normal (the reference skips the point, hpp:277)
    if ((s_role && !(s_role[i] & 1)) || !(isfinite(qn.x) && isfinite(qn.y) && isfinite(qn.z))) {
        s_score[i] = CUDART_NAN_F;
        score[orig] = CUDART_NAN_F;
        return;
    }
    const float sc = forest_score(sf + tid, FOREST_THREADS, nodes, roots, ntrees);
    s_score[i] = sc;
    score[orig] = sc;
    (void)counters;
}

cudaError_t launch_forest(kpl_ctx* c, int64_t n, bool use_role)
{
    const int F = c->params.n_annulus * c->params.n_bins;
    cudaError_t e;
    if ((e = ensure(c->s_score, n)) || (e = ensure(c->score, n))) return e;
    size_t smem = (size_t)F * FOREST_THREADS * sizeof(float);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(forest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return e;
    int blocks = (int)((n + FOREST_THREADS - 1) / FOREST_THREADS);
    forest_kernel<<<blocks, FOREST_THREADS, smem, c->stream>>>(c->feat.p, c->forest.d_nodes, c->forest.d_roots, c->forest.ntrees, F, (int)n,
                                                               use_role ? c->s_role.p : nullptr, c->s_pos.p, c->s_nrm.p, c->s_score.p, c->score.p,
                                                               c->counters.p);
    c->launches++;
    return cudaGetLastError();
}

// Host side: arbitrary (roots,var,thr,left,right,value) arrays -> pre-order PackedNode array.
int pack_forest(const HostForestArrays& in, std::vector<PackedNode>& nodes, std::vector<int32_t>& roots, int& max_depth, std::string& err)
{
    const int32_t nn = (int32_t)in.var.size();
    nodes.clear(); roots.clear(); max_depth = 0;
    nodes.reserve(nn);
    std::vector<std::pair<int32_t, int32_t>> stack;   // (source node, packed index of parent waiting for its right child or -1)
    std::vector<int32_t> depth_stack;
    for (size_t t = 0; t < in.roots.size(); ++t) {
        roots.push_back((int32_t)nodes.size());
        stack.clear(); depth_stack.clear();
        stack.push_back({in.roots[t], -1}); depth_stack.push_back(0);
        while (!stack.empty()) {
            auto [src, parent] = stack.back(); stack.pop_back();
            int d = depth_stack.back(); depth_stack.pop_back();
            if (src < 0 || src >= nn) { err = "forest: node index out of range"; return KPL_E_FOREST; }
            if ((int64_t)nodes.size() > (int64_t)nn) { err = "forest: cyclic node graph"; return KPL_E_FOREST; }
            int32_t me = (int32_t)nodes.size();
            if (parent >= 0) {
                int64_t off = (int64_t)me - parent;
                if (off >= (1 << 22)) { err = "forest: subtree larger than 2^22 nodes"; return KPL_E_FOREST; }
                nodes[parent].packed |= (uint32_t)off << 10;
            }
            max_depth = std::max(max_depth, d);
            PackedNode pn;
            if (in.var[src] < 0) { pn.thr = in.value[src]; pn.packed = KPL_LEAF_VAR; nodes.push_back(pn); }
            else {
                if (in.var[src] >= (int32_t)KPL_LEAF_VAR) { err = "forest: variable index >= 1023"; return KPL_E_FOREST; }
                pn.thr = in.thr[src]; pn.packed = (uint32_t)in.var[src];
                nodes.push_back(pn);
                // right is pushed first so that the left subtree is emitted immediately after `me`
                stack.push_back({in.right[src], me}); depth_stack.push_back(d + 1);
                stack.push_back({in.left[src], -1}); depth_stack.push_back(d + 1);
            }
        }
    }
    return KPL_OK;
}

}  // namespace kpl
