// dr_scan_hist: NULL-cell scan fused with per-column histograms.
//
// One pass over the columns; HBM-bound: 4 bytes per cell in, 1 bit per cell out.
//
// Shared-memory atomics cost ~2 cycles per lane on this part, an order of magnitude more than
// the HBM stream can afford at 32 columns.  The kernel therefore needs NO atomics in its hot loop:
//   * every warp owns a disjoint subset of the columns of its CTA (column c -> warp c % 8), so
//     bins are never shared between warps;
//   * inside a warp every lane owns a private copy of each bin, laid out bin*32 + lane, i.e. lane
//     L only ever touches shared-memory bank L: plain LDS/IADD/STS, conflict-free by construction.
// A CTA holds sum(dom+1) * 128 bytes of bins; columns are packed into groups (gridDim.y) that fit
// the shared-memory budget.  Each lane streams 4 consecutive rows per 128-bit load; the NULL bits
// of 128 rows are assembled with three xor-shuffles and OR-ed into the column bitmap (each word is
// owned by exactly one warp).  At the end the lane-private bins are reduced and added to the
// global int64 histogram with one atomic per (CTA, bin).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTileRows = 128;             // rows per warp-wide 128-bit load
constexpr int kTilesPerChunk = 4;          // tiles per chunk; two chunks (current + prefetched) in registers
constexpr int kChunkRows = kTileRows * kTilesPerChunk;

struct ScanParams {
    const int32_t* cols[DR_MAX_COLS];
    uint32_t* bitmaps[DR_MAX_COLS];
    int32_t dom[DR_MAX_COLS];
    int32_t slot_off[DR_MAX_COLS];   // offset into the global histogram
    int32_t local_off[DR_MAX_COLS];  // offset into the group's shared-memory bins
    int32_t group_start[DR_MAX_COLS + 1];
    int32_t group_slots[DR_MAX_COLS];
    int64_t n_rows;
    int64_t* hist;
};

// lane-private bin update: `b` already points at this lane's copy of the column's slot 0
__device__ __forceinline__ void count_one(int32_t* b, unsigned dom, int code) {
    const unsigned s = min((unsigned)(code + 1), dom);  // clamp: memory safety for out-of-range codes
    b[s * 32] += 1;
}

__device__ __forceinline__ void store_null_bits(uint32_t* __restrict__ bm, int64_t tile_base, unsigned nib,
                                                int lane) {
    unsigned w = nib << (4 * (lane & 7));
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    // RED.OR: fire-and-forget, so the streaming loop never waits on a read-modify-write of the bitmap
    if ((lane & 7) == 0 && w != 0) atomicOr(bm + (tile_base >> 5) + (lane >> 3), w);
}

__device__ __forceinline__ void load_chunk(int4 (&v)[kTilesPerChunk], const int32_t* __restrict__ col,
                                           int64_t base, int lane) {
#pragma unroll
    for (int t = 0; t < kTilesPerChunk; ++t)
        v[t] = __ldcs(reinterpret_cast<const int4*>(col + base + t * kTileRows) + lane);
}

// Counts the four codes of one 128-bit load.  The four bin addresses may coincide (same lane, same
// column), so instead of four dependent read-modify-writes the four bins are LOADED together, each
// new value gets +1 for itself and +1 for every earlier equal code of the quad, and the four stores
// go out in order -- the last store to an address carries the full count.  One shared-memory latency
// per quad instead of four.
__device__ __forceinline__ void count_quad(int32_t* b, unsigned dom, const int4 q) {
    const unsigned s0 = min((unsigned)(q.x + 1), dom), s1 = min((unsigned)(q.y + 1), dom);
    const unsigned s2 = min((unsigned)(q.z + 1), dom), s3 = min((unsigned)(q.w + 1), dom);
    const int v0 = b[s0 * 32], v1 = b[s1 * 32], v2 = b[s2 * 32], v3 = b[s3 * 32];
    b[s0 * 32] = v0 + 1;
    b[s1 * 32] = v1 + 1 + (s1 == s0);
    b[s2 * 32] = v2 + 1 + (s2 == s0) + (s2 == s1);
    b[s3 * 32] = v3 + 1 + (s3 == s0) + (s3 == s1) + (s3 == s2);
}

__device__ __forceinline__ void count_chunk(const int4 (&v)[kTilesPerChunk], int32_t* b, unsigned dom,
                                            uint32_t* __restrict__ bm, int64_t base, int lane) {
#pragma unroll
    for (int t = 0; t < kTilesPerChunk; ++t) {
        const int4 q = v[t];
        count_quad(b, dom, q);
        if (bm != nullptr) {
            const unsigned nib = ((unsigned)q.x >> 31) | (((unsigned)q.y >> 31) << 1) |
                                 (((unsigned)q.z >> 31) << 2) | (((unsigned)q.w >> 31) << 3);
            store_null_bits(bm, base + t * kTileRows, nib, lane);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_scan_hist(const __grid_constant__ ScanParams p) {
    extern __shared__ int32_t bins[];
    const int g = blockIdx.y;
    const int c_begin = p.group_start[g], c_end = p.group_start[g + 1];
    const int n_slots = p.group_slots[g];
    for (int i = threadIdx.x; i < n_slots * 32; i += kThreads) bins[i] = 0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_rows = p.n_rows;
    const int64_t n_full = n_rows / kChunkRows;  // chunks that need no bounds checks
    const int64_t n_chunks = (n_rows + kChunkRows - 1) / kChunkRows;

    for (int c = c_begin + warp; c < c_end; c += kWarps) {
        const int32_t* __restrict__ col = p.cols[c];
        uint32_t* __restrict__ bm = p.bitmaps[c];
        const int loc = p.local_off[c];
        const unsigned dom = (unsigned)p.dom[c];
        int32_t* const b = bins + loc * 32 + lane;
        const bool vec_ok = ((uintptr_t)col & 15) == 0;
        int64_t chunk = blockIdx.x;
        if (vec_ok) {
            // software pipeline, ping-pong register buffers: the next chunk's four 128-bit loads are in
            // flight while this one is counted
            int4 bufA[kTilesPerChunk], bufB[kTilesPerChunk];
            if (chunk < n_full) load_chunk(bufA, col, chunk * kChunkRows, lane);
            while (chunk < n_full) {
                int64_t next = chunk + gridDim.x;
                if (next < n_full) load_chunk(bufB, col, next * kChunkRows, lane);
                count_chunk(bufA, b, dom, bm, chunk * kChunkRows, lane);
                chunk = next;
                if (chunk >= n_full) break;
                next = chunk + gridDim.x;
                if (next < n_full) load_chunk(bufA, col, next * kChunkRows, lane);
                count_chunk(bufB, b, dom, bm, chunk * kChunkRows, lane);
                chunk = next;
            }
        }
        // ragged tail (and unaligned columns): bounds-checked scalar loads
        for (; chunk < n_chunks; chunk += gridDim.x) {
            const int64_t base = chunk * kChunkRows;
            for (int t = 0; t < kTilesPerChunk; ++t) {
                unsigned nib = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t r = base + t * kTileRows + lane * 4 + j;
                    if (r < n_rows) {
                        const int code = col[r];
                        count_one(b, dom, code);
                        nib |= (unsigned)(code < 0) << j;
                    }
                }
                if (bm != nullptr) store_null_bits(bm, base + t * kTileRows, nib, lane);
            }
        }
        // reduce the lane-private copies of this column's bins; lane L sums bin (s0 + L), reading
        // the 32 copies in a rotated order so that the 32 lanes hit 32 different banks
        __syncwarp();
        for (int s0 = 0; s0 <= (int)dom; s0 += 32) {
            const int s = s0 + lane;
            if (s <= (int)dom) {
                long long total = 0;
#pragma unroll 8
                for (int j = 0; j < 32; ++j) total += bins[(loc + s) * 32 + ((j + lane) & 31)];
                if (total != 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.hist + p.slot_off[c] + s),
                              (unsigned long long)total);
            }
        }
    }
}

// Fallback for domains too large for lane-private bins: global int64 atomics (low contention when
// the domain is large) + the same bitmap logic, one thread per row.
__global__ void k_scan_hist_global(const int32_t* __restrict__ col, uint32_t* __restrict__ bm, int dom,
                                   int64_t n_rows, int64_t* __restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n_rows + 31) & ~(int64_t)31;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_pad; r += stride) {
        bool is_null = false;
        if (r < n_rows) {
            const int code = col[r];
            unsigned s = (unsigned)(code + 1);
            s = s > (unsigned)dom ? (unsigned)dom : s;
            atomicAdd(reinterpret_cast<unsigned long long*>(hist + s), 1ull);
            is_null = code < 0;
        }
        const unsigned w = __ballot_sync(0xffffffffu, is_null);
        if (bm != nullptr && lane == 0 && w != 0) atomicOr(bm + (r >> 5), w);
    }
}

}  // namespace

extern "C" int dr_scan_hist(dr_ctx* ctx, const int32_t* const* cols, const int32_t* dom, int n_cols,
                            int64_t n_rows, uint32_t* const* bitmaps, int64_t* hist, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    DR_REQUIRE(ctx, cols && dom && hist, "null pointer");
    DR_REQUIRE(ctx, n_cols >= 0 && n_cols <= DR_MAX_COLS, "n_cols must be in [0, 64]");
    DR_REQUIRE(ctx, n_rows >= 0 && n_rows < (int64_t)INT32_MAX, "n_rows must be < 2^31 per shard");
    if (n_cols == 0 || n_rows == 0) return DR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DR_CUDA(ctx, cudaSetDevice(ctx->device));

    constexpr int kBudgetSlots = 880;     // 880 * 128 B = 110 KB -> two CTAs per SM (2 * 111 KB <= 228 KB)
    constexpr int kMaxGroupSlots = 1760;  // 220 KB, one CTA per SM
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.n_rows = n_rows;
    p.hist = hist;
    // columns small enough for lane-private bins; the rest take the global-atomic fallback
    int total_slots = 0, n_small = 0;
    for (int i = 0; i < n_cols; ++i) {
        DR_REQUIRE(ctx, dom[i] >= 0, "negative domain size");
        DR_REQUIRE(ctx, cols[i] != nullptr, "null column pointer");
        if (dom[i] + 1 <= kMaxGroupSlots) { total_slots += dom[i] + 1; ++n_small; }
    }
    // balanced packing: as few groups as the budget allows, about the same number of columns each
    int want_groups = (total_slots + kBudgetSlots - 1) / kBudgetSlots;
    if (want_groups < 1) want_groups = 1;
    const int cols_per_group = (n_small + want_groups - 1) / want_groups;
    int n_packed = 0, n_groups = 0, cur_slots = 0, cur_cols = 0, off = 0;
    int max_group_slots = 0;
    p.group_start[0] = 0;
    for (int i = 0; i < n_cols; ++i) {
        const int slots = dom[i] + 1;
        if (slots > kMaxGroupSlots) {  // large-domain fallback, launched separately
            const int threads = 256;
            const int grid = dr_grid_for(ctx, n_rows, threads, 8);
            k_scan_hist_global<<<grid, threads, 0, st>>>(cols[i], bitmaps ? bitmaps[i] : nullptr, dom[i], n_rows,
                                                         hist + off);
            DR_LAUNCHED(ctx);
            off += slots;
            continue;
        }
        const int limit = slots > kBudgetSlots ? kMaxGroupSlots : kBudgetSlots;
        if (cur_cols > 0 && (cur_slots + slots > limit || cur_cols >= cols_per_group)) {
            p.group_slots[n_groups] = cur_slots;
            max_group_slots = cur_slots > max_group_slots ? cur_slots : max_group_slots;
            p.group_start[++n_groups] = n_packed;
            cur_slots = 0;
            cur_cols = 0;
        }
        p.cols[n_packed] = cols[i];
        p.bitmaps[n_packed] = bitmaps ? bitmaps[i] : nullptr;
        p.dom[n_packed] = dom[i];
        p.slot_off[n_packed] = off;
        p.local_off[n_packed] = cur_slots;
        cur_slots += slots;
        ++cur_cols;
        off += slots;
        ++n_packed;
    }
    if (n_packed == 0) return DR_OK;
    p.group_slots[n_groups] = cur_slots;
    max_group_slots = cur_slots > max_group_slots ? cur_slots : max_group_slots;
    p.group_start[++n_groups] = n_packed;

    const size_t smem = (size_t)max_group_slots * 32 * sizeof(int32_t);
    DR_CUDA(ctx, cudaFuncSetAttribute(k_scan_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int64_t n_chunks = (n_rows + kChunkRows - 1) / kChunkRows;
    int grid_x = (ctx->sm_count * ctas_per_sm + n_groups - 1) / n_groups;
    if ((int64_t)grid_x > n_chunks) grid_x = (int)n_chunks;
    if (grid_x < 1) grid_x = 1;
    dim3 grid(grid_x, n_groups);
    k_scan_hist<<<grid, kThreads, smem, st>>>(p);
    DR_LAUNCHED(ctx);
    return DR_OK;
}
