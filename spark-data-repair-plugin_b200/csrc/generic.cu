// (1) dr_combine_counts: the local half of the sharded path's exchange step.  Every rank packs its count
//     tensors into one int64 buffer, ONE all-gather moves them over NVLink, and this kernel folds the G
//     copies segment by segment (SUM for histograms / pair tables / cell counts, MIN / MAX for the
//     per-key bounds of the denial constraints, OR for presence bits) -- one collective per exchange
//     whatever mix of reductions the pass needs (SURVEY.md 8e, 8b `dr_allreduce_counts`).
// (2) General two-tuple denial constraints (ErrorDetectorApi.scala:189-244 evaluates any parsed
//     predicate list as `EXISTS t2`): the answer for a row only depends on its PROJECTION onto the
//     attributes the constraint names, so the device (a) marks which projections occur
//     (dr_key_presence: mixed-radix key -> bit, test-then-set), (b) decides every DISTINCT projection
//     against the distinct projections of its equality group (dr_dc_exists), (c) flags the rows whose
//     projection was decided "violating" (dr_key_flag).  Cost: two streaming passes + |distinct|^2 per
//     group, instead of the reference's self semi-join over rows.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct CombineParams {
    int64_t seg_off[DR_MAX_SEGMENTS + 1];
    int32_t seg_op[DR_MAX_SEGMENTS];
    int n_seg;
};

__global__ void __launch_bounds__(kThreads) k_combine_counts(const int64_t* __restrict__ gathered, int world,
                                                             int64_t n, const __grid_constant__ CombineParams p,
                                                             int64_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int s = 0;
        while (s + 1 < p.n_seg && i >= p.seg_off[s + 1]) ++s;
        const int op = p.seg_op[s];
        int64_t acc = gathered[i];
        for (int g = 1; g < world; ++g) {
            const int64_t v = gathered[(int64_t)g * n + i];
            switch (op) {
                case DR_RED_SUM: acc += v; break;
                case DR_RED_MIN: acc = v < acc ? v : acc; break;
                case DR_RED_MAX: acc = v > acc ? v : acc; break;
                default: acc |= v; break;
            }
        }
        out[i] = acc;
    }
}

struct KeyParams {
    const int32_t* cols[DR_MAX_COLS];
    int64_t strides[DR_MAX_COLS];
    int n_keys;
};

__device__ __forceinline__ int64_t row_key(const KeyParams& k, int64_t r) {
    int64_t key = 0;
    for (int i = 0; i < k.n_keys; ++i) key += (int64_t)(__ldcs(k.cols[i] + r) + 1) * k.strides[i];
    return key;
}

// after the first few thousand rows nearly every projection has been seen: the test is a cached load,
// the atomic is rare
__global__ void __launch_bounds__(kThreads) k_key_presence(const __grid_constant__ KeyParams k, int64_t n_rows,
                                                           int64_t key_space, uint32_t* __restrict__ bits) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const int64_t key = row_key(k, r);
        if (key < 0 || key >= key_space) continue;
        const uint32_t bit = 1u << (key & 31);
        uint32_t* w = bits + (key >> 5);
        if (!(__ldcg(w) & bit)) atomicOr(w, bit);
    }
}

__global__ void __launch_bounds__(kThreads) k_key_flag(const __grid_constant__ KeyParams k, int64_t n_rows,
                                                       int64_t key_space, const uint32_t* __restrict__ viol,
                                                       uint32_t* __restrict__ bm) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n_rows + 31) & ~(int64_t)31;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += stride) {
        bool bit = false;
        if (r < n_rows) {
            const int64_t key = row_key(k, r);
            if (key >= 0 && key < key_space) bit = (__ldg(viol + (key >> 5)) >> (key & 31)) & 1u;
        }
        const unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0 && w != 0) atomicOr(bm + (r >> 5), w);
    }
}

struct ExistsParams {
    const int32_t* left[DR_MAX_PREDS];   // rank of t_i's left-hand value under predicate p (-1 = NULL)
    const int32_t* right[DR_MAX_PREDS];  // rank of t_j's right-hand value under predicate p (-1 = NULL)
    int32_t sign[DR_MAX_PREDS];
    int n_preds;
};

// SQL semantics of DenialConstraints.scala:66-79: EQ is `<=>`, IQ is NOT(<=>), LT / GT are false on NULL
__device__ __forceinline__ bool pred_holds(int sign, int l, int r) {
    switch (sign) {
        case DR_OP_EQ: return l == r;
        case DR_OP_IQ: return l != r;
        case DR_OP_LT: return l >= 0 && r >= 0 && l < r;
        default: return l >= 0 && r >= 0 && l > r;
    }
}

// one thread per distinct projection i: does some projection j of the same equality group satisfy
// every predicate?  (group_begin / group_end: the range of i's group in the group-sorted order)
__global__ void __launch_bounds__(kThreads) k_dc_exists(const __grid_constant__ ExistsParams p, int64_t n,
                                                        const int32_t* __restrict__ group_begin,
                                                        const int32_t* __restrict__ group_end,
                                                        uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l[DR_MAX_PREDS];
#pragma unroll
    for (int q = 0; q < DR_MAX_PREDS; ++q) l[q] = q < p.n_preds ? p.left[q][i] : 0;
    bool found = false;
    for (int j = group_begin[i]; j < group_end[i] && !found; ++j) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < DR_MAX_PREDS; ++q)
            if (q < p.n_preds && ok) ok = pred_holds(p.sign[q], l[q], __ldg(p.right[q] + j));
        found = ok;
    }
    out[i] = found ? 1 : 0;
}

int fill_keys(dr_ctx* ctx, KeyParams& k, const int32_t* const* cols, const int64_t* strides, int n_keys) {
    DR_REQUIRE(ctx, cols && strides && n_keys >= 1 && n_keys <= DR_MAX_COLS, "bad key columns");
    memset(&k, 0, sizeof(k));
    k.n_keys = n_keys;
    for (int i = 0; i < n_keys; ++i) {
        DR_REQUIRE(ctx, cols[i] != nullptr, "null key column");
        k.cols[i] = cols[i];
        k.strides[i] = strides[i];
    }
    return DR_OK;
}

}  // namespace

extern "C" {

int dr_combine_counts(dr_ctx* ctx, const int64_t* gathered, int world, int64_t n, const int64_t* seg_off,
                      const int32_t* seg_op, int n_seg, int64_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, gathered && out && seg_off && seg_op, "null pointer");
    DR_REQUIRE(ctx, world >= 1 && n_seg >= 1 && n_seg <= DR_MAX_SEGMENTS, "bad world size / segment count");
    DR_REQUIRE(ctx, seg_off[0] == 0 && seg_off[n_seg] == n, "segments must cover [0, n)");
    CombineParams p;
    memset(&p, 0, sizeof(p));
    p.n_seg = n_seg;
    for (int s = 0; s < n_seg; ++s) {
        DR_REQUIRE(ctx, seg_off[s] <= seg_off[s + 1], "segment offsets must ascend");
        DR_REQUIRE(ctx, seg_op[s] >= DR_RED_SUM && seg_op[s] <= DR_RED_OR, "bad reduction");
        p.seg_off[s] = seg_off[s];
        p.seg_op[s] = seg_op[s];
    }
    p.seg_off[n_seg] = n;
    k_combine_counts<<<dr_grid_for(ctx, n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(gathered, world, n, p,
                                                                                               out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_key_presence(dr_ctx* ctx, const int32_t* const* cols, const int64_t* strides, int n_keys, int64_t n_rows,
                    int64_t key_space, uint32_t* bits, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_rows <= 0) return DR_OK;
    DR_REQUIRE(ctx, bits && key_space >= 1, "bad presence table");
    KeyParams k;
    int rc = fill_keys(ctx, k, cols, strides, n_keys);
    if (rc) return rc;
    k_key_presence<<<dr_grid_for(ctx, n_rows, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(k, n_rows, key_space,
                                                                                                  bits);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_key_flag(dr_ctx* ctx, const int32_t* const* cols, const int64_t* strides, int n_keys, int64_t n_rows,
                int64_t key_space, const uint32_t* viol_bits, uint32_t* row_bitmap, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_rows <= 0) return DR_OK;
    DR_REQUIRE(ctx, viol_bits && row_bitmap && key_space >= 1, "bad arguments");
    KeyParams k;
    int rc = fill_keys(ctx, k, cols, strides, n_keys);
    if (rc) return rc;
    k_key_flag<<<dr_grid_for(ctx, n_rows, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(k, n_rows, key_space,
                                                                                              viol_bits, row_bitmap);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_dc_exists(dr_ctx* ctx, const int32_t* const* left, const int32_t* const* right, const int32_t* sign,
                 int n_preds, int64_t n, const int32_t* group_begin, const int32_t* group_end, uint8_t* out,
                 void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, left && right && sign && group_begin && group_end && out, "null pointer");
    DR_REQUIRE(ctx, n_preds >= 1 && n_preds <= DR_MAX_PREDS, "n_preds must be in [1, 8]");
    DR_REQUIRE(ctx, n < (int64_t)INT32_MAX, "too many distinct projections");
    ExistsParams p;
    memset(&p, 0, sizeof(p));
    p.n_preds = n_preds;
    for (int q = 0; q < n_preds; ++q) {
        DR_REQUIRE(ctx, left[q] && right[q], "null predicate column");
        DR_REQUIRE(ctx, sign[q] >= DR_OP_EQ && sign[q] <= DR_OP_GT, "bad predicate sign");
        p.left[q] = left[q];
        p.right[q] = right[q];
        p.sign[q] = sign[q];
    }
    const int grid = (int)((n + kThreads - 1) / kThreads);
    k_dc_exists<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, n, group_begin, group_end, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

}  // extern "C"
