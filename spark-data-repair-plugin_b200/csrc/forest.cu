// dr_forest_predict: fused feature-gather -> encode -> tree traversal -> argmax -> in-place fill for
// one target attribute (model.py:1095-1135).
//
// One thread per cell.  The cell's encoded feature vector lives in shared memory, feature-major
// (feat[f * T + t]) so that lanes reading the same feature hit 32 different banks.  The forest is
// streamed through shared memory in chunks of whole trees, sequence (class) by sequence; every
// thread walks every tree of the chunk for its own cell and keeps ONE float64 accumulator -- the
// margin of the current sequence, summed in tree order exactly like the oracle -- plus the running
// arg-max.  This kernel is latency/issue bound (about 300 * C * depth dependent shared-memory
// look-ups per cell), not HBM bound; DESIGN.md reports its node-visit rate next to the HBM figure.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kChunkNodes = 6144;  // 6144 * 12 B = 72 KB of nodes per chunk
constexpr uint32_t kLeaf = 0xFFFu;

struct ForestParams {
    dr_forest f;
    int32_t* tile;
    int n_cols;
    double* ctile;
    int n_ccols;
    const int32_t* cells;
    int64_t n_cells;
    int target_col;
    double* out_margin;
};

template <int T>
__global__ void __launch_bounds__(T) k_forest_predict(const __grid_constant__ ForestParams p) {
    extern __shared__ double smem_d[];
    const dr_forest& F = p.f;
    double* feat = smem_d;                                               // n_feat * T
    double* s_thr = smem_d + (size_t)F.n_feat * T;                       // kChunkNodes
    uint32_t* s_meta = reinterpret_cast<uint32_t*>(s_thr + kChunkNodes);  // kChunkNodes
    const int t = threadIdx.x;

    for (int64_t base = (int64_t)blockIdx.x * T; base < p.n_cells; base += (int64_t)gridDim.x * T) {
        const int64_t i = base + t;
        const bool live = i < p.n_cells;
        const int64_t row = live ? p.cells[i] : 0;
        __syncthreads();  // previous iteration is done with feat / node chunks
        if (live) {
            const int32_t* trow = p.tile + row * p.n_cols;
            for (int f = 0; f < F.n_feat; ++f) {
                const int col = F.feat_col[f];
                double x;
                if (col >= 0) {
                    const int code = trow[col];
                    const int lo = F.enc_lut_off[f], hi = F.enc_lut_off[f + 1];
                    const int k = lo + code + 1;
                    x = (k >= lo && k < hi) ? __ldg(F.enc_lut + k) : (double)NAN;  // unseen code -> NaN
                } else {
                    x = p.ctile[row * p.n_ccols + (-col - 1)];
                }
                feat[f * T + t] = x;
            }
        }
        double best = 0.0;
        int best_s = 0;
        double margin0 = 0.0;
        for (int s = 0; s < F.n_seq; ++s) {
            double acc = F.baseline[s];
            const int t_begin = F.seq_tree_off[s], t_end = F.seq_tree_off[s + 1];
            int tr = t_begin;
            while (tr < t_end) {
                // chunk = as many whole trees as fit
                const int n0 = F.tree_node_off[tr];
                int tr_hi = tr + 1;
                while (tr_hi < t_end && F.tree_node_off[tr_hi + 1] - n0 <= kChunkNodes) ++tr_hi;
                const int n1 = F.tree_node_off[tr_hi];
                __syncthreads();
                for (int k = t; k < n1 - n0; k += T) {
                    s_thr[k] = F.node_thr[n0 + k];
                    s_meta[k] = F.node_meta[n0 + k];
                }
                __syncthreads();
                if (live) {
                    for (int q = tr; q < tr_hi; ++q) {
                        const int root = F.tree_node_off[q] - n0;
                        int node = 0;
                        uint32_t m = s_meta[root];
                        while ((m & kLeaf) != kLeaf) {
                            const double x = feat[(m & kLeaf) * T + t];
                            const double thr = s_thr[root + node];
                            const bool go_left = (x != x) ? ((m >> 12) & 1u) : (x <= thr);
                            node = go_left ? ((m >> 13) & 0x1FFu) : ((m >> 22) & 0x1FFu);
                            m = s_meta[root + node];
                        }
                        acc += s_thr[root + node];
                    }
                }
                tr = tr_hi;
            }
            if (live && p.out_margin) p.out_margin[i * F.n_seq + s] = acc;
            if (s == 0) { best = acc; best_s = 0; margin0 = acc; }
            else if (acc > best) { best = acc; best_s = s; }
        }
        if (live) {
            if (F.kind == 0) {
                const int cls = F.n_seq == 1 ? (margin0 > 0.0 ? 1 : 0) : best_s;
                p.tile[row * p.n_cols + p.target_col] = cls < F.n_classes ? F.class_code[cls] : -1;
            } else {
                const double v = F.integral ? rint(margin0) : margin0;
                p.ctile[row * p.n_ccols + p.target_col] = v;
            }
        }
    }
}

template <int T>
int launch(dr_ctx* ctx, const ForestParams& p, cudaStream_t st) {
    const size_t smem = (size_t)p.f.n_feat * T * sizeof(double) + (size_t)kChunkNodes * 12;
    DR_CUDA(ctx, cudaFuncSetAttribute(k_forest_predict<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / T) per_sm = 2048 / T;
    const int grid = dr_grid_for(ctx, p.n_cells, T, per_sm);
    k_forest_predict<T><<<grid, T, smem, st>>>(p);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

}  // namespace

extern "C" int dr_forest_predict(dr_ctx* ctx, const dr_forest* forest, int32_t* tile, int n_cols, double* ctile,
                                 int n_ccols, const int32_t* cells, int64_t n_cells, int target_col,
                                 double* out_margin, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_cells <= 0) return DR_OK;
    DR_REQUIRE(ctx, forest && cells, "null pointer");
    const dr_forest& f = *forest;
    DR_REQUIRE(ctx, f.n_seq >= 1 && f.n_trees >= 0 && f.n_nodes >= 0 && f.n_feat >= 0, "bad forest sizes");
    DR_REQUIRE(ctx, f.seq_tree_off && f.tree_node_off && f.baseline && f.feat_col && f.enc_lut_off,
               "null forest array");
    DR_REQUIRE(ctx, f.n_nodes == 0 || (f.node_thr && f.node_meta), "null node arrays");
    DR_REQUIRE(ctx, f.n_feat < 4095, "too many encoded features (max 4094)");
    DR_REQUIRE(ctx, tile != nullptr || n_cols == 0, "null tile");
    if (f.kind == 0) {
        DR_REQUIRE(ctx, tile && target_col >= 0 && target_col < n_cols, "bad target column");
        DR_REQUIRE(ctx, f.class_code && f.n_classes >= 1, "classifier needs class codes");
        DR_REQUIRE(ctx, f.n_seq == 1 ? f.n_classes >= 2 : f.n_classes == f.n_seq, "class count mismatch");
    } else {
        DR_REQUIRE(ctx, ctile && target_col >= 0 && target_col < n_ccols, "bad continuous target column");
        DR_REQUIRE(ctx, f.n_seq == 1, "regressor has one sequence");
    }
    ForestParams p;
    p.f = f;
    p.tile = tile;
    p.n_cols = n_cols;
    p.ctile = ctile;
    p.n_ccols = n_ccols;
    p.cells = cells;
    p.n_cells = n_cells;
    p.target_col = target_col;
    p.out_margin = out_margin;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t budget = 150 * 1024;
    if ((size_t)f.n_feat * 128 * 8 <= budget) return launch<128>(ctx, p, st);
    if ((size_t)f.n_feat * 64 * 8 <= budget) return launch<64>(ctx, p, st);
    if ((size_t)f.n_feat * 32 * 8 <= budget) return launch<32>(ctx, p, st);
    return dr_fail(ctx, DR_ERR_UNSUPPORTED, "forest with %lld encoded features exceeds the shared-memory budget",
                   (long long)f.n_feat);
}
