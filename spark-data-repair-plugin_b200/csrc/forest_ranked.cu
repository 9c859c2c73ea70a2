// dr_forest_predict_ranked: rank-coded forest inference for all-discrete models (see the header).
//
// Compared with the generic float64 kernel: a node is one 32-bit word instead of 12 bytes, a cell's
// features are one byte each instead of eight, so a CTA of 256 cells needs ~25 KB of features and the
// rest of shared memory holds bigger forest chunks at 2 CTAs (16 warps) per SM.  Each thread walks FOUR
// trees at a time for a fixed number of levels (leaves self-loop), which gives the scheduler four
// independent dependent-load chains per thread and removes all data-dependent branches; the four
// leaf values are then added to the sequence's float64 accumulator in tree order, so the margins are
// bit-identical to the generic kernel and to the oracle.
#include "common.cuh"

namespace {

constexpr int T = 256;
constexpr int kChunkNodes = 8192;   // 32 KB of node words
constexpr int kChunkLeaves = 4352;  // 34 KB of float64 leaf values

struct RankedParams {
    dr_forest_ranked f;
    int32_t* tile;
    int n_cols;
    const int32_t* cells;
    int64_t n_cells;
    int target_col;
    double* out_margin;
    int feat_stride;  // bytes per thread row of the feature tile (multiple of 4, odd number of words)
};

// One level: two shared-memory loads (feature rank, next node word) and six ALU ops.  Leaves point
// at themselves (left = right = own index), so a finished tree just re-reads its leaf word.
__device__ __forceinline__ uint32_t step_node(const uint32_t* __restrict__ nodes, int root, uint32_t w,
                                              const uint8_t* __restrict__ my_feat) {
    const uint32_t r = my_feat[w >> 20];            // rank of feature (w >> 21) in its NaN-left / NaN-right variant
    const uint32_t thr = (w >> 12) & 0xFFu;
    const uint32_t child = (r < thr) ? ((w >> 6) & 0x3Fu) : (w & 0x3Fu);
    return nodes[root + (int)child];
}

__global__ void __launch_bounds__(T, 2) k_forest_predict_ranked(const __grid_constant__ RankedParams p) {
    extern __shared__ unsigned char smem_raw[];
    const dr_forest_ranked& F = p.f;
    double* s_leaf = reinterpret_cast<double*>(smem_raw);                            // kChunkLeaves
    uint32_t* s_node = reinterpret_cast<uint32_t*>(s_leaf + kChunkLeaves);           // kChunkNodes
    uint8_t* s_feat = reinterpret_cast<uint8_t*>(s_node + kChunkNodes);              // T * feat_stride
    const int t = threadIdx.x;
    uint8_t* my_feat = s_feat + (size_t)t * p.feat_stride;
    const int depth = F.max_depth;

    for (int64_t base = (int64_t)blockIdx.x * T; base < p.n_cells; base += (int64_t)gridDim.x * T) {
        const int64_t i = base + t;
        const bool live = i < p.n_cells;
        const int64_t row = live ? p.cells[i] : 0;
        __syncthreads();
        {
            const int32_t* trow = p.tile + row * p.n_cols;
            for (int f = 0; f < F.n_feat; ++f) {
                uint8_t r = 255;  // NaN
                if (live) {
                    const int lo = F.rank_lut_off[f], hi = F.rank_lut_off[f + 1];
                    const int k = lo + trow[F.feat_col[f]] + 1;
                    if (k >= lo && k < hi) r = __ldg(F.rank_lut + k);
                }
                // ranks are stored +1 (1..254); NaN becomes 0 in the "NaN goes left" variant and 255
                // in the "NaN goes right" one, so `rank < thr` needs no special case
                my_feat[2 * f + 0] = r == 255 ? 255 : r;   // variant read by nodes whose NaN goes right
                my_feat[2 * f + 1] = r == 255 ? 0 : r;     // variant read by nodes whose NaN goes left
            }
        }
        double best = 0.0, margin0 = 0.0;
        int best_s = 0;
        for (int s = 0; s < F.n_seq; ++s) {
            double acc = F.baseline[s];
            const int t_end = F.seq_tree_off[s + 1];
            int tr = F.seq_tree_off[s];
            while (tr < t_end) {
                const int n0 = F.tree_node_off[tr], l0 = F.tree_leaf_off[tr];
                int tr_hi = tr + 1;
                while (tr_hi < t_end && F.tree_node_off[tr_hi + 1] - n0 <= kChunkNodes &&
                       F.tree_leaf_off[tr_hi + 1] - l0 <= kChunkLeaves)
                    ++tr_hi;
                const int n1 = F.tree_node_off[tr_hi], l1 = F.tree_leaf_off[tr_hi];
                __syncthreads();
                for (int k = t; k < n1 - n0; k += T) s_node[k] = F.node_word[n0 + k];
                for (int k = t; k < l1 - l0; k += T) s_leaf[k] = F.leaf_value[l0 + k];
                __syncthreads();
                int q = tr;
                for (; q + 4 <= tr_hi; q += 4) {
                    const int r0 = F.tree_node_off[q] - n0, r1 = F.tree_node_off[q + 1] - n0;
                    const int r2 = F.tree_node_off[q + 2] - n0, r3 = F.tree_node_off[q + 3] - n0;
                    uint32_t w0 = s_node[r0], w1 = s_node[r1], w2 = s_node[r2], w3 = s_node[r3];
                    for (int d = 0; d < depth; ++d) {
                        w0 = step_node(s_node, r0, w0, my_feat);
                        w1 = step_node(s_node, r1, w1, my_feat);
                        w2 = step_node(s_node, r2, w2, my_feat);
                        w3 = step_node(s_node, r3, w3, my_feat);
                    }
                    acc += s_leaf[F.tree_leaf_off[q] - l0 + (int)((w0 >> 12) & 0xFFu)];
                    acc += s_leaf[F.tree_leaf_off[q + 1] - l0 + (int)((w1 >> 12) & 0xFFu)];
                    acc += s_leaf[F.tree_leaf_off[q + 2] - l0 + (int)((w2 >> 12) & 0xFFu)];
                    acc += s_leaf[F.tree_leaf_off[q + 3] - l0 + (int)((w3 >> 12) & 0xFFu)];
                }
                for (; q < tr_hi; ++q) {
                    const int r0 = F.tree_node_off[q] - n0;
                    uint32_t w0 = s_node[r0];
                    for (int d = 0; d < depth; ++d) w0 = step_node(s_node, r0, w0, my_feat);
                    acc += s_leaf[F.tree_leaf_off[q] - l0 + (int)((w0 >> 12) & 0xFFu)];
                }
                tr = tr_hi;
            }
            if (live && p.out_margin) p.out_margin[i * F.n_seq + s] = acc;
            if (s == 0) { best = acc; best_s = 0; margin0 = acc; }
            else if (acc > best) { best = acc; best_s = s; }
        }
        if (live) {
            const int cls = F.n_seq == 1 ? (margin0 > 0.0 ? 1 : 0) : best_s;
            p.tile[row * p.n_cols + p.target_col] = cls < F.n_classes ? F.class_code[cls] : -1;
        }
    }
}

}  // namespace

extern "C" int dr_forest_predict_ranked(dr_ctx* ctx, const dr_forest_ranked* forest, int32_t* tile, int n_cols,
                                        const int32_t* cells, int64_t n_cells, int target_col, double* out_margin,
                                        void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_cells <= 0) return DR_OK;
    DR_REQUIRE(ctx, forest && cells && tile, "null pointer");
    const dr_forest_ranked& f = *forest;
    DR_REQUIRE(ctx, f.n_seq >= 1 && f.n_feat >= 0 && f.n_feat <= 2047, "bad forest sizes");
    DR_REQUIRE(ctx, f.seq_tree_off && f.tree_node_off && f.tree_leaf_off && f.baseline && f.feat_col &&
                        f.rank_lut_off && f.class_code, "null forest array");
    DR_REQUIRE(ctx, f.max_depth >= 0 && f.max_depth <= 63, "bad max_depth");
    DR_REQUIRE(ctx, target_col >= 0 && target_col < n_cols, "bad target column");
    DR_REQUIRE(ctx, f.n_seq == 1 ? f.n_classes >= 2 : f.n_classes == f.n_seq, "class count mismatch");
    RankedParams p;
    p.f = f;
    p.tile = tile;
    p.n_cols = n_cols;
    p.cells = cells;
    p.n_cells = n_cells;
    p.target_col = target_col;
    p.out_margin = out_margin;
    int words = (2 * f.n_feat + 3) / 4;
    if (words < 1) words = 1;
    if ((words & 1) == 0) ++words;  // odd word stride: consecutive threads land on different banks
    p.feat_stride = words * 4;
    const size_t smem = (size_t)kChunkLeaves * 8 + (size_t)kChunkNodes * 4 + (size_t)T * p.feat_stride;
    if (smem > 200 * 1024)
        return dr_fail(ctx, DR_ERR_UNSUPPORTED, "ranked forest with %d features exceeds shared memory", f.n_feat);
    DR_CUDA(ctx, cudaFuncSetAttribute(k_forest_predict_ranked, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int grid = dr_grid_for(ctx, n_cells, T, per_sm);
    k_forest_predict_ranked<<<grid, T, smem, (cudaStream_t)stream>>>(p);
    DR_LAUNCHED(ctx);
    return DR_OK;
}
