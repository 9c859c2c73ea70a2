// Cell-set plumbing: bitmap algebra, ordered compaction (bitmap -> ascending row list), masked
// row gathers into the row-major dirty tile, cell gathers and sorted lookup.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 8;

__global__ void __launch_bounds__(kThreads) k_bitmap_or(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src,
                                                        int64_t n_words, int andnot) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride)
        dst[i] = andnot ? (dst[i] & ~src[i]) : (dst[i] | src[i]);
}

__device__ __forceinline__ uint32_t tail_mask(int64_t word, int64_t n_rows) {
    const int64_t first = word << 5;
    if (first + 32 <= n_rows) return 0xffffffffu;
    if (first >= n_rows) return 0u;
    return (1u << (int)(n_rows - first)) - 1u;
}

// ---- ordered compaction -------------------------------------------------------------------------
constexpr int kWordsPerBlock = 1024;  // 32 Ki rows per block

__device__ __forceinline__ void block_popc(const uint32_t* __restrict__ bm, int64_t n_rows, int64_t n_words,
                                           unsigned long long* __restrict__ counts) {
    __shared__ unsigned int warp_sum[kThreads / 32];
    const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock;
    unsigned int c = 0;
    for (int i = threadIdx.x; i < kWordsPerBlock; i += kThreads) {
        const int64_t w = w0 + i;
        if (w < n_words) c += __popc(bm[w] & tail_mask(w, n_rows));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int i = 0; i < kThreads / 32; ++i) t += warp_sum[i];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kThreads) k_block_popc(const uint32_t* __restrict__ bm, int64_t n_rows,
                                                         int64_t n_words, unsigned long long* __restrict__ counts) {
    block_popc(bm, n_rows, n_words, counts);
}

// exclusive scan of the per-block counts, one CTA; counts[n_blocks] receives the total
__device__ __forceinline__ void scan_counts(unsigned long long* __restrict__ counts, int n_blocks) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n_blocks ? counts[i] : 0ull;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;  // inclusive
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? warp_tot[warp - 1] : 0ull) + (x - v);
        if (i < n_blocks) counts[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n_blocks] = carry;
}

__global__ void __launch_bounds__(1024) k_scan_counts(unsigned long long* __restrict__ counts, int n_blocks) {
    scan_counts(counts, n_blocks);
}

__device__ __forceinline__ void write_rows(const uint32_t* __restrict__ bm, int64_t n_rows, int64_t n_words,
                                           const unsigned long long* __restrict__ offsets,
                                           int32_t* __restrict__ out, int64_t capacity) {
    __shared__ unsigned int warp_sum[kThreads / 32];
    __shared__ unsigned long long block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) block_base = offsets[blockIdx.x];
    __syncthreads();
    const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock;
    for (int step = 0; step < kWordsPerBlock; step += kThreads) {
        const int64_t w = w0 + step + threadIdx.x;
        uint32_t bits = w < n_words ? (bm[w] & tail_mask(w, n_rows)) : 0u;
        const unsigned int c = __popc(bits);
        unsigned int x = c;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        unsigned int before = x - c;
        unsigned int total = 0;
        for (int i = 0; i < kThreads / 32; ++i) {
            if (i < warp) before += warp_sum[i];
            total += warp_sum[i];
        }
        unsigned long long pos = block_base + before;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if ((int64_t)pos < capacity) out[pos] = (int32_t)((w << 5) + b);
            ++pos;
        }
        __syncthreads();
        if (threadIdx.x == 0) block_base += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) k_write_rows(const uint32_t* __restrict__ bm, int64_t n_rows,
                                                         int64_t n_words,
                                                         const unsigned long long* __restrict__ offsets,
                                                         int32_t* __restrict__ out, int64_t capacity) {
    write_rows(bm, n_rows, n_words, offsets, out, capacity);
}

__global__ void __launch_bounds__(kThreads) k_bitmap_gather(const uint32_t* __restrict__ src,
                                                            const int32_t* __restrict__ rows, int64_t n,
                                                            uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool bit = false;
        if (i < n) {
            const int r = rows[i];
            bit = (__ldg(src + (r >> 5)) >> (r & 31)) & 1u;
        }
        const unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) out[i >> 5] = w;
    }
}

__global__ void __launch_bounds__(kThreads) k_bitmap_clear_rows(uint32_t* __restrict__ bm,
                                                                const int32_t* __restrict__ rows,
                                                                const uint8_t* __restrict__ flags, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!flags[i]) continue;
        const int r = rows[i];
        atomicAnd(bm + (r >> 5), ~(1u << (r & 31)));  // several cleared rows may share a word
    }
}

// ---- masked row gather into the row-major tile ---------------------------------------------------
struct GatherParams {
    const void* cols[DR_MAX_COLS];
    const uint32_t* bitmaps[DR_MAX_COLS];
    int n_cols;
};

// A warp gathers 32 rows: lane = row while reading (each column read is a 32-row gather from one
// column array), then the 32 x K tile is written out contiguously (coalesced) through shared memory.
// Columns are taken eight at a time with all sixteen loads (mask word + value) issued before any is
// used -- the gather is latency bound otherwise.  Optionally the NULL bitmap of every tile column is
// produced on the way (one ballot per column and row group), which saves a pass over the tile.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_gather_rows_masked(const __grid_constant__ GatherParams p,
                                                                 const int32_t* __restrict__ rows, int64_t n,
                                                                 T* __restrict__ out, T null_value,
                                                                 uint32_t* __restrict__ null_out,
                                                                 int64_t null_words_per_col) {
    extern __shared__ unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int K = p.n_cols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* my = tile + (size_t)warp * 32 * (K + 1);
    const int64_t n_groups = (n + 31) >> 5;
    constexpr int kBatch = 8;
    for (int64_t grp = (int64_t)blockIdx.x * (kThreads / 32) + warp; grp < n_groups;
         grp += (int64_t)gridDim.x * (kThreads / 32)) {
        const int64_t i = (grp << 5) + lane;
        const int r = i < n ? rows[i] : -1;
        const int rr = r < 0 ? 0 : r;  // always a readable row
        for (int c0 = 0; c0 < K; c0 += kBatch) {
            T v[kBatch];
            uint32_t mw[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int c = c0 + j < K ? c0 + j : K - 1;
                const uint32_t* bm = p.bitmaps[c];
                mw[j] = bm != nullptr ? __ldg(bm + (rr >> 5)) : 0u;
                v[j] = reinterpret_cast<const T*>(p.cols[c])[rr];
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int c = c0 + j;
                if (c < K) {  // warp uniform
                    T x = v[j];
                    if (r < 0 || ((mw[j] >> (rr & 31)) & 1u)) x = null_value;
                    my[lane * (K + 1) + c] = x;
                    if (null_out != nullptr) {
                        bool isnull;
                        if constexpr (sizeof(T) == 8) isnull = x != x; else isnull = x < 0;
                        const unsigned bits = __ballot_sync(0xffffffffu, isnull && r >= 0);
                        if (lane == 0) null_out[(int64_t)c * null_words_per_col + grp] = bits;
                    }
                }
            }
        }
        __syncwarp();
        const int64_t first = grp << 5;
        const int64_t cnt = (n - first < 32 ? n - first : 32) * K;
        for (int64_t j = lane; j < cnt; j += 32) out[first * K + j] = my[(j / K) * (K + 1) + (j % K)];
        __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_tile_null_bitmap(const T* __restrict__ tile, int64_t n, int K,
                                                               int col, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool bit = false;
        if (i < n) {
            const T v = tile[i * K + col];
            if constexpr (sizeof(T) == 8) bit = v != v; else bit = v < 0;
        }
        const unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) out[i >> 5] = w;
    }
}

// NULL bitmaps of ALL columns of the row-major tile in one pass: a warp takes 32 rows (lane = row,
// each lane walks its own 4*K-byte row, which stays L1-resident), one ballot per column.
__global__ void __launch_bounds__(kThreads) k_tile_null_bitmaps(const int32_t* __restrict__ tile, int64_t n, int K,
                                                                int64_t words_per_col, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t n_groups = (n + 31) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < n_groups; grp += warps) {
        const int64_t i = (grp << 5) + lane;
        const int32_t* row = tile + (i < n ? i : 0) * K;
        for (int c = 0; c < K; ++c) {
            const bool bit = i < n && row[c] < 0;
            const unsigned w = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) out[(int64_t)c * words_per_col + grp] = w;
        }
    }
}

// Ingest: codes of small dictionaries cross PCIe as one byte each (255 = NULL) and are widened to
// the int32 layout on the device: 16 codes per thread, 128-bit load -> four 128-bit stores.
__global__ void __launch_bounds__(kThreads) k_widen_u8(const uint8_t* __restrict__ src, int64_t n,
                                                       int32_t* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n16 = n >> 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int4* out = reinterpret_cast<int4*>(dst) + i * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int4 o;
            const uint32_t b0 = w[j] & 0xFF, b1 = (w[j] >> 8) & 0xFF, b2 = (w[j] >> 16) & 0xFF, b3 = w[j] >> 24;
            o.x = b0 == 255 ? -1 : (int)b0;
            o.y = b1 == 255 ? -1 : (int)b1;
            o.z = b2 == 255 ? -1 : (int)b2;
            o.w = b3 == 255 ? -1 : (int)b3;
            __stcs(out + j, o);
        }
    }
    for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t b = src[i];
        dst[i] = b == 255 ? -1 : (int)b;
    }
}

// bit i = repaired[i] IS NULL OR NOT(current[i] <=> repaired[i])  (model.py:1401), on codes
__global__ void __launch_bounds__(kThreads) k_changed_bitmap(const int32_t* __restrict__ cur,
                                                             const int32_t* __restrict__ rep, int64_t n,
                                                             uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool bit = false;
        if (i < n) {
            const int r = rep[i];
            bit = r < 0 || r != cur[i];
        }
        const unsigned w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) out[i >> 5] = w;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_gather(const T* __restrict__ col, const int32_t* __restrict__ rows,
                                                     int64_t n, int64_t row_stride, int64_t col_off,
                                                     T* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = col[(int64_t)rows[i] * row_stride + col_off];
}

__global__ void __launch_bounds__(kThreads) k_lookup_sorted(const int32_t* __restrict__ sorted, int64_t n_sorted,
                                                            const int32_t* __restrict__ keys, int64_t n,
                                                            int32_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int key = keys[i];
        int64_t lo = 0, hi = n_sorted;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(sorted + mid) < key) lo = mid + 1; else hi = mid;
        }
        out[i] = (lo < n_sorted && __ldg(sorted + lo) == key) ? (int32_t)lo : -1;
    }
}

__global__ void __launch_bounds__(kThreads) k_tile_fill(int32_t* __restrict__ tile, int K, int col,
                                                        const int32_t* __restrict__ cells, int64_t n, int32_t value) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        tile[(int64_t)cells[i] * K + col] = value;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_scatter(T* __restrict__ col, const int32_t* __restrict__ rows,
                                                      const T* __restrict__ vals, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) col[rows[i]] = vals[i];
}

// tile[cells[i]][y_col] = lut[tile[cells[i]][x_col] + 1]: the functional-dependency model X -> Y
__global__ void __launch_bounds__(kThreads) k_tile_lut_fill(int32_t* __restrict__ tile, int K, int x_col, int y_col,
                                                            const int32_t* __restrict__ cells, int64_t n,
                                                            const int32_t* __restrict__ lut, int32_t lut_size) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t* row = tile + (int64_t)cells[i] * K;
        const int32_t x = row[x_col];
        row[y_col] = (x >= -1 && x + 1 < lut_size) ? __ldg(lut + x + 1) : -1;
    }
}

// per key x (NULL and masked cells skipped): min / max of the non-NULL, unmasked y codes
__global__ void __launch_bounds__(kThreads) k_fd_map_build(const int32_t* __restrict__ x_col,
                                                           const uint32_t* __restrict__ x_mask,
                                                           const int32_t* __restrict__ y_col,
                                                           const uint32_t* __restrict__ y_mask, int64_t n_rows,
                                                           int32_t dom_x, int32_t* __restrict__ lo,
                                                           int32_t* __restrict__ hi) {
    extern __shared__ int32_t s_tab[];  // [2][dom_x] when it fits, else straight to global
    const bool use_smem = dom_x <= 8192;
    if (use_smem) {
        for (int i = threadIdx.x; i < dom_x; i += blockDim.x) { s_tab[i] = INT32_MAX; s_tab[dom_x + i] = INT32_MIN; }
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const int32_t x = x_col[r], y = y_col[r];
        if (x < 0 || y < 0 || x >= dom_x) continue;
        if (x_mask && (x_mask[r >> 5] >> (r & 31)) & 1u) continue;
        if (y_mask && (y_mask[r >> 5] >> (r & 31)) & 1u) continue;
        if (use_smem) {
            if (y < s_tab[x]) atomicMin(&s_tab[x], y);
            if (y > s_tab[dom_x + x]) atomicMax(&s_tab[dom_x + x], y);
        } else {
            if (y < lo[x]) atomicMin(&lo[x], y);
            if (y > hi[x]) atomicMax(&hi[x], y);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < dom_x; i += blockDim.x) {
            if (s_tab[i] != INT32_MAX) atomicMin(&lo[i], s_tab[i]);
            if (s_tab[dom_x + i] != INT32_MIN) atomicMax(&hi[i], s_tab[dom_x + i]);
        }
    }
}

// popcount of many bitmaps in one launch: blockIdx.y = bitmap, blocks of a row stride over its words
struct ManyBitmaps {
    const uint32_t* bm[DR_MAX_COUNT_MANY];
};
__global__ void __launch_bounds__(kThreads) k_popc_many(const __grid_constant__ ManyBitmaps p, int64_t n_rows,
                                                        unsigned long long* __restrict__ totals) {
    const uint32_t* bm = p.bm[blockIdx.y];
    const int64_t n_words = (n_rows + 31) >> 5;
    unsigned long long local = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = bm[w];
        if (w == n_words - 1 && (n_rows & 31)) v &= (1u << (n_rows & 31)) - 1u;  // bits past the last row
        local += __popc(v);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(totals + blockIdx.y, local);
}

// ordered compaction of many bitmaps in three launches: blockIdx.y = bitmap
struct ManyRows {
    int32_t* out[DR_MAX_COUNT_MANY];
    long long capacity[DR_MAX_COUNT_MANY];
};
__global__ void __launch_bounds__(kThreads) k_block_popc_many(const __grid_constant__ ManyBitmaps p, int64_t n_rows,
                                                              int64_t n_words, int n_blocks,
                                                              unsigned long long* __restrict__ counts) {
    block_popc(p.bm[blockIdx.y], n_rows, n_words, counts + (size_t)blockIdx.y * (n_blocks + 1));
}
__global__ void __launch_bounds__(1024) k_scan_counts_many(unsigned long long* __restrict__ counts, int n_blocks) {
    scan_counts(counts + (size_t)blockIdx.x * (n_blocks + 1), n_blocks);
}
__global__ void __launch_bounds__(kThreads) k_write_rows_many(const __grid_constant__ ManyBitmaps p,
                                                              const __grid_constant__ ManyRows o, int64_t n_rows,
                                                              int64_t n_words, int n_blocks,
                                                              const unsigned long long* __restrict__ counts) {
    if (o.capacity[blockIdx.y] <= 0) return;
    write_rows(p.bm[blockIdx.y], n_rows, n_words, counts + (size_t)blockIdx.y * (n_blocks + 1), o.out[blockIdx.y],
               o.capacity[blockIdx.y]);
}

inline int grid_rows(const dr_ctx* ctx, int64_t n) { return dr_grid_for(ctx, n, kThreads, kCtasPerSm); }

}  // namespace

extern "C" {

int dr_bitmap_or(dr_ctx* ctx, uint32_t* dst, const uint32_t* src, int64_t n_rows, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, dst && src, "null pointer");
    const int64_t n_words = (n_rows + 31) >> 5;
    if (n_words <= 0) return DR_OK;
    k_bitmap_or<<<grid_rows(ctx, n_words), kThreads, 0, (cudaStream_t)stream>>>(dst, src, n_words, 0);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_bitmap_andnot(dr_ctx* ctx, uint32_t* dst, const uint32_t* src, int64_t n_rows, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, dst && src, "null pointer");
    const int64_t n_words = (n_rows + 31) >> 5;
    if (n_words <= 0) return DR_OK;
    k_bitmap_or<<<grid_rows(ctx, n_words), kThreads, 0, (cudaStream_t)stream>>>(dst, src, n_words, 1);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

static int compaction_counts(dr_ctx* ctx, const uint32_t* bitmap, int64_t n_rows, int* n_blocks_out,
                             cudaStream_t st) {
    const int64_t n_words = (n_rows + 31) >> 5;
    const int64_t n_blocks = (n_words + kWordsPerBlock - 1) / kWordsPerBlock;
    DR_REQUIRE(ctx, n_blocks < (1 << 30), "bitmap too large");
    int rc = dr_ensure_scratch(ctx, sizeof(unsigned long long) * (size_t)(n_blocks + 1));
    if (rc) return rc;
    unsigned long long* counts = (unsigned long long*)ctx->scratch;
    k_block_popc<<<(int)n_blocks, kThreads, 0, st>>>(bitmap, n_rows, n_words, counts);
    DR_LAUNCHED(ctx);
    k_scan_counts<<<1, 1024, 0, st>>>(counts, (int)n_blocks);
    DR_LAUNCHED(ctx);
    *n_blocks_out = (int)n_blocks;
    return DR_OK;
}

int dr_bitmap_count(dr_ctx* ctx, const uint32_t* bitmap, int64_t n_rows, int64_t* out_count, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, bitmap && out_count, "null pointer");
    *out_count = 0;
    if (n_rows <= 0) return DR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int n_blocks = 0;
    int rc = compaction_counts(ctx, bitmap, n_rows, &n_blocks, st);
    if (rc) return rc;
    unsigned long long* counts = (unsigned long long*)ctx->scratch;
    DR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, counts + n_blocks, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                 st));
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    *out_count = (int64_t) * (unsigned long long*)ctx->pinned;
    return DR_OK;
}

int dr_bitmap_count_many(dr_ctx* ctx, const uint32_t* const* bitmaps, int n_bitmaps, int64_t n_rows,
                         int64_t* out_counts, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_bitmaps <= 0) return DR_OK;
    DR_REQUIRE(ctx, bitmaps && out_counts, "null pointer");
    DR_REQUIRE(ctx, n_bitmaps <= DR_MAX_COUNT_MANY, "too many bitmaps for one call");
    for (int i = 0; i < n_bitmaps; ++i) out_counts[i] = 0;
    if (n_rows <= 0) return DR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ManyBitmaps p;
    for (int i = 0; i < n_bitmaps; ++i) {
        DR_REQUIRE(ctx, bitmaps[i] != nullptr, "null bitmap");
        p.bm[i] = bitmaps[i];
    }
    unsigned long long* totals = (unsigned long long*)ctx->scratch;  // (clobbers the compaction offsets)
    DR_CUDA(ctx, cudaMemsetAsync(totals, 0, sizeof(unsigned long long) * n_bitmaps, st));
    const int64_t n_words = (n_rows + 31) >> 5;
    int per_row = (int)((n_words + kThreads * 8 - 1) / (kThreads * 8));
    const int cap = ctx->sm_count * 8 / n_bitmaps + 1;
    if (per_row > cap) per_row = cap;
    if (per_row < 1) per_row = 1;
    k_popc_many<<<dim3(per_row, n_bitmaps), kThreads, 0, st>>>(p, n_rows, totals);
    DR_LAUNCHED(ctx);
    DR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, totals, sizeof(unsigned long long) * n_bitmaps, cudaMemcpyDeviceToHost,
                                 st));
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < n_bitmaps; ++i) out_counts[i] = (int64_t)((unsigned long long*)ctx->pinned)[i];
    return DR_OK;
}

int dr_bitmap_to_rows_async(dr_ctx* ctx, const uint32_t* bitmap, int64_t n_rows, int32_t* out_rows, int64_t count,
                            void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, bitmap && (out_rows || count == 0), "null pointer");
    if (n_rows <= 0 || count <= 0) return DR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int n_blocks = 0;
    int rc = compaction_counts(ctx, bitmap, n_rows, &n_blocks, st);
    if (rc) return rc;
    k_write_rows<<<n_blocks, kThreads, 0, st>>>(bitmap, n_rows, (n_rows + 31) >> 5,
                                                (const unsigned long long*)ctx->scratch, out_rows, count);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_bitmaps_to_rows_many(dr_ctx* ctx, const uint32_t* const* bitmaps, int n_bitmaps, int64_t n_rows,
                            int32_t* const* out_rows, const int64_t* counts, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_bitmaps <= 0 || n_rows <= 0) return DR_OK;
    DR_REQUIRE(ctx, bitmaps && out_rows && counts, "null pointer");
    DR_REQUIRE(ctx, n_bitmaps <= DR_MAX_COUNT_MANY, "too many bitmaps for one call");
    const int64_t n_words = (n_rows + 31) >> 5;
    const int64_t n_blocks = (n_words + kWordsPerBlock - 1) / kWordsPerBlock;
    DR_REQUIRE(ctx, n_blocks < 65536 * 16, "bitmap too large");
    ManyBitmaps p;
    ManyRows o;
    bool any = false;
    for (int i = 0; i < n_bitmaps; ++i) {
        DR_REQUIRE(ctx, bitmaps[i] != nullptr && counts[i] >= 0 && (counts[i] == 0 || out_rows[i]), "null bitmap / list");
        p.bm[i] = bitmaps[i];
        o.out[i] = out_rows[i];
        o.capacity[i] = counts[i];
        any |= counts[i] > 0;
    }
    if (!any) return DR_OK;
    int rc = dr_ensure_scratch(ctx, sizeof(unsigned long long) * (size_t)(n_blocks + 1) * (size_t)n_bitmaps);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* cnt = (unsigned long long*)ctx->scratch;
    k_block_popc_many<<<dim3((unsigned)n_blocks, (unsigned)n_bitmaps), kThreads, 0, st>>>(p, n_rows, n_words,
                                                                                       (int)n_blocks, cnt);
    DR_LAUNCHED(ctx);
    k_scan_counts_many<<<n_bitmaps, 1024, 0, st>>>(cnt, (int)n_blocks);
    DR_LAUNCHED(ctx);
    k_write_rows_many<<<dim3((unsigned)n_blocks, (unsigned)n_bitmaps), kThreads, 0, st>>>(p, o, n_rows, n_words,
                                                                                       (int)n_blocks, cnt);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_bitmap_to_rows(dr_ctx* ctx, const uint32_t* bitmap, int64_t n_rows, int32_t* out_rows, int64_t capacity,
                      int64_t* out_count, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, bitmap && out_count && (out_rows || capacity == 0), "null pointer");
    *out_count = 0;
    if (n_rows <= 0) return DR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int n_blocks = 0;
    int rc = compaction_counts(ctx, bitmap, n_rows, &n_blocks, st);
    if (rc) return rc;
    unsigned long long* counts = (unsigned long long*)ctx->scratch;
    const int64_t n_words = (n_rows + 31) >> 5;
    if (capacity > 0) {
        k_write_rows<<<n_blocks, kThreads, 0, st>>>(bitmap, n_rows, n_words, counts, out_rows, capacity);
        DR_LAUNCHED(ctx);
    }
    DR_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, counts + n_blocks, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                 st));
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    *out_count = (int64_t) * (unsigned long long*)ctx->pinned;
    if (*out_count > capacity) return dr_fail(ctx, DR_ERR_INVALID, "dr_bitmap_to_rows: capacity too small");
    return DR_OK;
}

int dr_bitmap_rows_after_count(dr_ctx* ctx, const uint32_t* bitmap, int64_t n_rows, int32_t* out_rows,
                               int64_t capacity, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, bitmap && (out_rows || capacity == 0), "null pointer");
    if (n_rows <= 0 || capacity <= 0) return DR_OK;
    const int64_t n_words = (n_rows + 31) >> 5;
    const int64_t n_blocks = (n_words + kWordsPerBlock - 1) / kWordsPerBlock;
    DR_REQUIRE(ctx, ctx->scratch_bytes >= sizeof(unsigned long long) * (size_t)(n_blocks + 1),
               "dr_bitmap_rows_after_count must follow dr_bitmap_count on the same bitmap");
    k_write_rows<<<(int)n_blocks, kThreads, 0, (cudaStream_t)stream>>>(
        bitmap, n_rows, n_words, (const unsigned long long*)ctx->scratch, out_rows, capacity);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_null_bitmaps(dr_ctx* ctx, const int32_t* tile, int64_t n, int n_cols, int64_t words_per_col,
                         uint32_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0 || n_cols <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && out && words_per_col >= (n + 31) / 32, "bad arguments");
    k_tile_null_bitmaps<<<dr_grid_for(ctx, (n + 31) / 32, kThreads / 32, kCtasPerSm), kThreads, 0,
                          (cudaStream_t)stream>>>(tile, n, n_cols, words_per_col, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_widen_u8(dr_ctx* ctx, const uint8_t* src, int64_t n, int32_t* dst, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, src && dst, "null pointer");
    DR_REQUIRE(ctx, ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "buffers must be 16-byte aligned");
    k_widen_u8<<<dr_grid_for(ctx, (n + 15) / 16, kThreads, kCtasPerSm), kThreads, 0, (cudaStream_t)stream>>>(src, n,
                                                                                                            dst);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_changed_bitmap(dr_ctx* ctx, const int32_t* current, const int32_t* repaired, int64_t n, uint32_t* out,
                      void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, current && repaired && out, "null pointer");
    k_changed_bitmap<<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(current, repaired, n, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_bitmap_gather(dr_ctx* ctx, const uint32_t* src, const int32_t* rows, int64_t n, uint32_t* out,
                     void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, src && rows && out, "null pointer");
    k_bitmap_gather<<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(src, rows, n, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_bitmap_clear_rows(dr_ctx* ctx, uint32_t* bitmap, const int32_t* rows, const uint8_t* flags, int64_t n,
                         void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, bitmap && rows && flags, "null pointer");
    k_bitmap_clear_rows<<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(bitmap, rows, flags, n);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

}  // extern "C"

template <typename T>
static int gather_rows_masked(dr_ctx* ctx, const T* const* cols, uint32_t* const* bitmaps, int n_cols,
                              const int32_t* rows, int64_t n, T* out, T null_value, uint32_t* null_out,
                              int64_t null_words_per_col, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0 || n_cols == 0) return DR_OK;
    DR_REQUIRE(ctx, cols && rows && out, "null pointer");
    DR_REQUIRE(ctx, null_out == nullptr || null_words_per_col >= (n + 31) / 32, "null bitmap rows too short");
    DR_REQUIRE(ctx, n_cols >= 1 && n_cols <= DR_MAX_COLS, "n_cols must be in [1, 64]");
    GatherParams p;
    memset(&p, 0, sizeof(p));
    p.n_cols = n_cols;
    for (int c = 0; c < n_cols; ++c) {
        DR_REQUIRE(ctx, cols[c] != nullptr, "null column pointer");
        p.cols[c] = cols[c];
        p.bitmaps[c] = bitmaps ? bitmaps[c] : nullptr;
    }
    const size_t smem = (size_t)(kThreads / 32) * 32 * (n_cols + 1) * sizeof(T);
    DR_CUDA(ctx, cudaFuncSetAttribute(k_gather_rows_masked<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int grid = dr_grid_for(ctx, (n + 31) / 32, kThreads / 32, 2);
    k_gather_rows_masked<T><<<grid, kThreads, smem, (cudaStream_t)stream>>>(p, rows, n, out, null_value, null_out,
                                                                            null_words_per_col);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

extern "C" {

int dr_gather_rows_masked(dr_ctx* ctx, const int32_t* const* cols, uint32_t* const* bitmaps, int n_cols,
                          const int32_t* rows, int64_t n, int32_t* out, void* stream) {
    return gather_rows_masked<int32_t>(ctx, cols, bitmaps, n_cols, rows, n, out, -1, nullptr, 0, stream);
}

int dr_gather_rows_masked_nulls(dr_ctx* ctx, const int32_t* const* cols, uint32_t* const* bitmaps, int n_cols,
                                const int32_t* rows, int64_t n, int32_t* out, uint32_t* null_out,
                                int64_t null_words_per_col, void* stream) {
    return gather_rows_masked<int32_t>(ctx, cols, bitmaps, n_cols, rows, n, out, -1, null_out, null_words_per_col,
                                       stream);
}

int dr_gather_rows_masked_f64(dr_ctx* ctx, const double* const* cols, uint32_t* const* bitmaps, int n_cols,
                              const int32_t* rows, int64_t n, double* out, void* stream) {
    return gather_rows_masked<double>(ctx, cols, bitmaps, n_cols, rows, n, out, (double)NAN, nullptr, 0, stream);
}

int dr_tile_null_bitmap(dr_ctx* ctx, const int32_t* tile, int64_t n, int n_cols, int col, uint32_t* out,
                        void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && out && col >= 0 && col < n_cols, "bad tile column");
    k_tile_null_bitmap<int32_t><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(tile, n, n_cols, col, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_null_bitmap_f64(dr_ctx* ctx, const double* tile, int64_t n, int n_cols, int col, uint32_t* out,
                            void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && out && col >= 0 && col < n_cols, "bad tile column");
    k_tile_null_bitmap<double><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(tile, n, n_cols, col, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_gather_i32(dr_ctx* ctx, const int32_t* col, const int32_t* rows, int64_t n, int32_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, col && rows && out, "null pointer");
    k_gather<int32_t><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(col, rows, n, 1, 0, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_gather_f64(dr_ctx* ctx, const double* col, const int32_t* rows, int64_t n, double* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, col && rows && out, "null pointer");
    k_gather<double><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(col, rows, n, 1, 0, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_gather_i32(dr_ctx* ctx, const int32_t* tile, int n_cols, int col, const int32_t* drows, int64_t n,
                       int32_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && drows && out && col >= 0 && col < n_cols, "bad tile column");
    k_gather<int32_t><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(tile, drows, n, n_cols, col, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_gather_f64(dr_ctx* ctx, const double* tile, int n_cols, int col, const int32_t* drows, int64_t n,
                       double* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && drows && out && col >= 0 && col < n_cols, "bad tile column");
    k_gather<double><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(tile, drows, n, n_cols, col, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_lookup_sorted(dr_ctx* ctx, const int32_t* sorted, int64_t n_sorted, const int32_t* keys, int64_t n,
                     int32_t* out, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, (sorted || n_sorted == 0) && keys && out, "null pointer");
    k_lookup_sorted<<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(sorted, n_sorted, keys, n, out);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_fill_i32(dr_ctx* ctx, int32_t* tile, int n_cols, int col, const int32_t* cells, int64_t n_cells,
                     int32_t value, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_cells <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && cells && col >= 0 && col < n_cols, "bad tile column");
    k_tile_fill<<<grid_rows(ctx, n_cells), kThreads, 0, (cudaStream_t)stream>>>(tile, n_cols, col, cells, n_cells,
                                                                               value);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_scatter_i32(dr_ctx* ctx, int32_t* col, const int32_t* rows, const int32_t* vals, int64_t n, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, col && rows && vals, "null pointer");
    k_scatter<int32_t><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(col, rows, vals, n);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_scatter_f64(dr_ctx* ctx, double* col, const int32_t* rows, const double* vals, int64_t n, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n <= 0) return DR_OK;
    DR_REQUIRE(ctx, col && rows && vals, "null pointer");
    k_scatter<double><<<grid_rows(ctx, n), kThreads, 0, (cudaStream_t)stream>>>(col, rows, vals, n);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_tile_lut_fill(dr_ctx* ctx, int32_t* tile, int n_cols, int x_col, int y_col, const int32_t* cells,
                     int64_t n_cells, const int32_t* lut, int32_t lut_size, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_cells <= 0) return DR_OK;
    DR_REQUIRE(ctx, tile && cells && lut && lut_size >= 1, "null pointer");
    DR_REQUIRE(ctx, x_col >= 0 && x_col < n_cols && y_col >= 0 && y_col < n_cols && x_col != y_col, "bad tile column");
    k_tile_lut_fill<<<grid_rows(ctx, n_cells), kThreads, 0, (cudaStream_t)stream>>>(tile, n_cols, x_col, y_col, cells,
                                                                                   n_cells, lut, lut_size);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

int dr_fd_map_build(dr_ctx* ctx, const int32_t* x_col, const uint32_t* x_mask, const int32_t* y_col,
                    const uint32_t* y_mask, int64_t n_rows, int32_t dom_x, int32_t* lo, int32_t* hi, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_rows <= 0 || dom_x <= 0) return DR_OK;
    DR_REQUIRE(ctx, x_col && y_col && lo && hi, "null pointer");
    const size_t smem = dom_x <= 8192 ? (size_t)dom_x * 2 * sizeof(int32_t) : 0;
    if (smem > 48 * 1024)
        DR_CUDA(ctx, cudaFuncSetAttribute(k_fd_map_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fd_map_build<<<grid_rows(ctx, n_rows), kThreads, smem, (cudaStream_t)stream>>>(x_col, x_mask, y_col, y_mask,
                                                                                     n_rows, dom_x, lo, hi);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

}  // extern "C"
