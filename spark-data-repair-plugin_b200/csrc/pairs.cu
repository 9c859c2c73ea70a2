// Attribute-pair statistics: distinct-pair presence bits and exact
// co-occurrence counts.  Both stage a 256-row tile of the referenced columns in shared memory as
// 16-bit slots (code + 1), then every thread walks the pair list for its row.
//   presence: test-then-atomicOr on a bit table (after the first few tiles nearly every test hits)
//   counts:   shared-memory int32 atomics per (CTA, table), flushed to global int64 at the end
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kTile = 512;

struct PairParams {
    const int32_t* cols[DR_MAX_COLS];
    int32_t dom[DR_MAX_COLS];
    int32_t used[DR_MAX_COLS];  // columns referenced by this launch's pairs (staged in this order)
    int n_used;
    const int32_t* px;  // device, indices into `used` order
    const int32_t* py;
    const int32_t* off;  // device int32[n_pairs + 1], word/entry offsets relative to the launch
    const int32_t* skip;      // counts only: device, entry (cx) of pair q at skip[skip_off[q] + cx], or nullptr
    const int32_t* skip_off;  // device int32[n_pairs], offsets into `skip`
    int32_t skip_off0;        // skip offset of the launch's first pair
    int32_t skip_len;         // skip entries of the launch's pairs
    int n_pairs;
    int64_t n_rows;
    int64_t block_rows;
    int64_t n_blocks;
};

__device__ __forceinline__ int64_t block_start(const PairParams& p, int64_t j) {
    if (p.block_rows * p.n_blocks >= p.n_rows || p.n_blocks <= 1) return j * p.block_rows;
    return (j * (p.n_rows - p.block_rows)) / (p.n_blocks - 1);
}

// per-pair metadata staged in shared memory once per CTA (one 128-bit broadcast load per pair and row
// instead of three dependent global loads)
struct __align__(16) PairMeta {
    int32_t x_off, y_off;  // element offsets of the pair's columns in the staged code tile
    int32_t base_ny;       // first table word / entry of the pair | (dom_y + 1) << 16
    int32_t skip;          // offset of the pair's entries in the staged skip LUT
};

template <bool kCount>
__global__ void __launch_bounds__(kThreads) k_pairs(const __grid_constant__ PairParams p, int table_words,
                                                    int skip_words, void* __restrict__ out_raw) {
    extern __shared__ uint32_t smem[];
    uint32_t* tab = smem;                                                   // table_words
    PairMeta* meta = reinterpret_cast<PairMeta*>(smem + ((table_words + 3) & ~3));   // n_pairs
    uint16_t* skip = reinterpret_cast<uint16_t*>(meta + p.n_pairs);        // skip_words * 2 entries
    uint16_t* codes = skip + 2 * skip_words;                               // n_used * kTile
    for (int i = threadIdx.x; i < table_words; i += kThreads) tab[i] = 0;
    for (int q = threadIdx.x; q < p.n_pairs; q += kThreads) {
        const int x = p.px[q], y = p.py[q];
        PairMeta m;
        m.x_off = x * kTile;
        m.y_off = y * kTile;
        m.base_ny = p.off[q] | ((p.dom[p.used[y]] + 1) << 16);
        m.skip = kCount && p.skip != nullptr ? p.skip_off[q] - p.skip_off0 : 0;
        meta[q] = m;
    }
    if (kCount && p.skip != nullptr)
        for (int i = threadIdx.x; i < 2 * skip_words; i += kThreads)
            skip[i] = i < p.skip_len ? (uint16_t)p.skip[p.skip_off0 + i] : (uint16_t)0xFFFF;
    __syncthreads();
    const bool use_skip = kCount && p.skip != nullptr;
    const int64_t units_per_block = (p.block_rows + kTile - 1) / kTile;
    const int64_t n_units = units_per_block * p.n_blocks;
    for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int64_t j = u / units_per_block, k = u % units_per_block;
        const int64_t first = block_start(p, j) + k * kTile;
        int64_t last = block_start(p, j) + p.block_rows;
        if (last > p.n_rows) last = p.n_rows;
        const int64_t r = first + threadIdx.x;
        const bool live = r < last;
        // stage the row's codes, eight columns at a time with all loads issued before any is used
        constexpr int kBatch = 8;
        const int64_t rr = live ? r : first;  // always a readable row
        for (int c0 = 0; c0 < p.n_used; c0 += kBatch) {
            int v[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int c = c0 + j < p.n_used ? c0 + j : p.n_used - 1;
                v[j] = __ldcs(p.cols[p.used[c]] + rr);
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int c = c0 + j;
                if (c < p.n_used) {
                    const int dom = p.dom[p.used[c]];
                    const unsigned u = (unsigned)(v[j] + 1);
                    codes[c * kTile + threadIdx.x] = (uint16_t)(live ? (u > (unsigned)dom ? dom : (int)u) : 0);
                }
            }
        }
        // each thread only reads back its own column of `codes`: no barrier needed
        if (live) {
            for (int q = 0; q < p.n_pairs; ++q) {
                const PairMeta m = meta[q];
                const int cx = codes[m.x_off + threadIdx.x], cy = codes[m.y_off + threadIdx.x];
                const int base = m.base_ny & 0xFFFF;
                const int idx = cx * (int)((unsigned)m.base_ny >> 16) + cy;
                if (kCount) {
                    // the most frequent partner of cx is not counted: the host restores it from the column
                    // histogram (dr_cooc_skip) -- on correlated pairs that removes nearly every atomic
                    if (!use_skip || skip[m.skip + cx] != cy) atomicAdd(&tab[base + idx], 1u);
                } else {
                    const uint32_t bit = 1u << (idx & 31);
                    uint32_t* w = &tab[base + (idx >> 5)];
                    if (!(*w & bit)) atomicOr(w, bit);
                }
            }
        }
    }
    __syncthreads();
    if (kCount) {
        unsigned long long* out = reinterpret_cast<unsigned long long*>(out_raw);
        for (int i = threadIdx.x; i < table_words; i += kThreads)
            if (tab[i]) atomicAdd(out + i, (unsigned long long)tab[i]);
    } else {
        uint32_t* out = reinterpret_cast<uint32_t*>(out_raw);
        for (int i = threadIdx.x; i < table_words; i += kThreads)
            if (tab[i]) atomicOr(out + i, tab[i]);
    }
}

// Launches k_pairs over consecutive pair ranges whose tables fit the shared-memory budget.
template <bool kCount>
int run_pairs(dr_ctx* ctx, const int32_t* const* cols, const int32_t* dom, int n_cols, const int32_t* px,
              const int32_t* py, int n_pairs, const int64_t* off, int64_t n_rows, int64_t block_rows,
              int64_t n_blocks, void* out, cudaStream_t st, const int32_t* skip = nullptr,
              const int64_t* skip_off = nullptr) {
    DR_REQUIRE(ctx, cols && dom && px && py && off && out, "null pointer");
    DR_REQUIRE(ctx, n_cols >= 1 && n_cols <= DR_MAX_COLS, "n_cols must be in [1, 64]");
    DR_REQUIRE(ctx, n_rows < (int64_t)INT32_MAX, "n_rows must be < 2^31 per shard");
    if (n_pairs <= 0 || n_rows <= 0 || n_blocks <= 0 || block_rows <= 0) return DR_OK;
    for (int q = 0; q < n_pairs; ++q) {
        DR_REQUIRE(ctx, px[q] >= 0 && px[q] < n_cols && py[q] >= 0 && py[q] < n_cols && px[q] != py[q], "bad pair");
        const int64_t need = kCount ? (int64_t)(dom[px[q]] + 1) * (dom[py[q]] + 1)
                                    : ((int64_t)(dom[px[q]] + 1) * (dom[py[q]] + 1) + 31) / 32;
        DR_REQUIRE(ctx, off[q + 1] - off[q] >= need, "pair table offsets too small");
        DR_REQUIRE(ctx, dom[px[q]] < 65535 && dom[py[q]] < 65535, "domain too large for 16-bit staging");
    }
    // device copies of the per-launch pair lists live in the context scratch buffer
    int rc = dr_ensure_scratch(ctx, sizeof(int32_t) * (size_t)(4 * n_pairs + 4) + 256);
    if (rc) return rc;
    static thread_local int32_t* h = nullptr;
    static thread_local size_t h_cap = 0;
    if (h_cap < (size_t)(4 * n_pairs + 4)) {
        free(h);
        h_cap = (size_t)(4 * n_pairs + 4);
        h = (int32_t*)malloc(sizeof(int32_t) * h_cap);
        if (!h) { h_cap = 0; return dr_fail(ctx, DR_ERR_INVALID, "out of host memory"); }
    }
    int q0 = 0;
    while (q0 < n_pairs) {
        PairParams p;
        memset(&p, 0, sizeof(p));
        int slot_of[DR_MAX_COLS];
        for (int i = 0; i < n_cols; ++i) slot_of[i] = -1;
        int q1 = q0;
        int64_t words = 0;
        auto skip_entries = [&](int qa, int qb) { return skip_off ? (int64_t)(skip_off[qb] - skip_off[qa]) : 0; };
        while (q1 < n_pairs) {
            const int extra = (slot_of[px[q1]] < 0) + (slot_of[py[q1]] < 0);
            const int64_t words_after = off[q1 + 1] - off[q0];
            const size_t bytes_after = (size_t)(words_after + 4) * 4 + (size_t)(p.n_used + extra) * kTile * 2 +
                                       (size_t)(q1 + 1 - q0) * 16 + (size_t)(skip_entries(q0, q1 + 1) + 2) * 2;
            // 65535 table words: the pair metadata keeps the pair's base in 16 bits
            if (q1 > q0 && (bytes_after > (size_t)200 * 1024 || words_after > 65535)) break;
            if (slot_of[px[q1]] < 0) { slot_of[px[q1]] = p.n_used; p.used[p.n_used++] = px[q1]; }
            if (slot_of[py[q1]] < 0) { slot_of[py[q1]] = p.n_used; p.used[p.n_used++] = py[q1]; }
            words = words_after;
            ++q1;
        }
        const int np = q1 - q0;
        const int skip_words = (int)((skip_entries(q0, q1) + 1) / 2);
        const size_t smem = (size_t)((words + 3) & ~(int64_t)3) * 4 + (size_t)np * 16 + (size_t)skip_words * 4 +
                            (size_t)p.n_used * kTile * 2;
        if (smem > 220 * 1024 || words > 65535)
            return dr_fail(ctx, DR_ERR_UNSUPPORTED, "pair table of %lld words does not fit in shared memory",
                           (long long)words);
        for (int q = 0; q < np; ++q) {
            h[q] = slot_of[px[q0 + q]];
            h[np + q] = slot_of[py[q0 + q]];
            h[2 * np + q] = (int32_t)(off[q0 + q] - off[q0]);
            h[3 * np + 1 + q] = skip_off ? (int32_t)skip_off[q0 + q] : 0;
        }
        h[3 * np] = (int32_t)(off[q1] - off[q0]);
        int32_t* d = (int32_t*)ctx->scratch;
        // the scratch copy must be consumed before the next range overwrites it
        DR_CUDA(ctx, cudaStreamSynchronize(st));
        DR_CUDA(ctx, cudaMemcpyAsync(d, h, sizeof(int32_t) * (size_t)(4 * np + 1), cudaMemcpyHostToDevice, st));
        for (int i = 0; i < n_cols; ++i) { p.cols[i] = cols[i]; p.dom[i] = dom[i]; }
        p.px = d;
        p.py = d + np;
        p.off = d + 2 * np;
        p.skip = skip;
        p.skip_off = d + 3 * np + 1;
        p.skip_off0 = skip_off ? (int32_t)skip_off[q0] : 0;
        p.skip_len = (int32_t)skip_entries(q0, q1);
        p.n_pairs = np;
        p.n_rows = n_rows;
        p.block_rows = block_rows;
        p.n_blocks = n_blocks;
        DR_CUDA(ctx, cudaFuncSetAttribute(k_pairs<kCount>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t units = ((block_rows + kTile - 1) / kTile) * n_blocks;
        const int per_sm = smem <= 110 * 1024 ? 2 : 1;
        int grid = ctx->sm_count * per_sm;
        if ((int64_t)grid > units) grid = (int)units;
        char* outp = (char*)out + (kCount ? sizeof(int64_t) : sizeof(uint32_t)) * (size_t)off[q0];
        k_pairs<kCount><<<grid, kThreads, smem, st>>>(p, (int)words, skip_words, outp);
        DR_LAUNCHED(ctx);
        DR_CUDA(ctx, cudaStreamSynchronize(st));  // h / scratch are reused by the next range
        q0 = q1;
    }
    return DR_OK;
}

}  // namespace

extern "C" {

int dr_pair_presence(dr_ctx* ctx, const int32_t* const* cols, const int32_t* dom, int n_cols, const int32_t* px,
                     const int32_t* py, int n_pairs, const int64_t* bit_off, int64_t n_rows, int64_t block_rows,
                     int64_t n_blocks, uint32_t* bits, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    return run_pairs<false>(ctx, cols, dom, n_cols, px, py, n_pairs, bit_off, n_rows, block_rows, n_blocks, bits,
                            (cudaStream_t)stream);
}

int dr_cooc(dr_ctx* ctx, const int32_t* const* cols, const int32_t* dom, int n_cols, const int32_t* px,
            const int32_t* py, int n_pairs, const int64_t* tab_off, int64_t n_rows, int64_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    const int64_t block_rows = 1 << 16;
    const int64_t n_blocks = (n_rows + block_rows - 1) / block_rows;
    return run_pairs<true>(ctx, cols, dom, n_cols, px, py, n_pairs, tab_off, n_rows, block_rows, n_blocks, out,
                           (cudaStream_t)stream);
}


int dr_cooc_skip(dr_ctx* ctx, const int32_t* const* cols, const int32_t* dom, int n_cols, const int32_t* px,
                 const int32_t* py, int n_pairs, const int64_t* tab_off, int64_t n_rows, const int32_t* skip,
                 const int64_t* skip_off, int64_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, (skip == nullptr) == (skip_off == nullptr), "skip and skip_off go together");
    if (skip_off)
        for (int q = 0; q < n_pairs; ++q)
            DR_REQUIRE(ctx, skip_off[q + 1] - skip_off[q] >= dom[px[q]] + 1 && skip_off[q] < INT32_MAX,
                       "skip table offsets too small");
    const int64_t block_rows = 1 << 16;
    const int64_t n_blocks = (n_rows + block_rows - 1) / block_rows;
    return run_pairs<true>(ctx, cols, dom, n_cols, px, py, n_pairs, tab_off, n_rows, block_rows, n_blocks, out,
                           (cudaStream_t)stream, skip, skip_off);
}

}  // extern "C"
