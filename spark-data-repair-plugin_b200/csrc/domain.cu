// dr_domain_score: weak-label cell-domain analysis for one target attribute (SURVEY.md 8a row a9,
// RepairApi.scala:479-675).  The naive-Bayes term of a (correlated value, candidate) pair depends on
// three counts only, so a first kernel evaluates exp(log(.) + log(.)) once per table entry
// (<= 81 x 81 per correlated attribute) into a score table; the per-cell kernel -- one thread per noisy
// cell -- then only adds table entries, in the same order as before: identical doubles, no
// transcendental math on the 10^7-cell path.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kMaxCorr = 8;

struct DomainParams {
    const int32_t* corr[kMaxCorr];
    const int64_t* cooc[kMaxCorr];
    int32_t dom_c[kMaxCorr];
    int64_t tau[kMaxCorr];
    int n_corr;
};

struct ScoreTables {
    double* score[kMaxCorr];   // [dom_c][dom_t]: the term of (value v, candidate n), or -1 when cnt <= tau
    int32_t* any[kMaxCorr];    // [dom_c]: does value v keep at least one candidate
};

__global__ void __launch_bounds__(kThreads) k_domain_prepare(const __grid_constant__ DomainParams p,
                                                             const __grid_constant__ ScoreTables t, int dom_t,
                                                             const int64_t* __restrict__ hist_t, double n_total) {
    const int j = blockIdx.y;
    const int64_t n_entries = (int64_t)p.dom_c[j] * dom_t;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
        const int v = (int)(e / dom_t), n = (int)(e % dom_t);
        const long long cnt = p.cooc[j][(int64_t)(v + 1) * (dom_t + 1) + 1 + n];
        const long long cnt_n = hist_t[n + 1];
        double s = -1.0;
        if (cnt > p.tau[j]) {
            atomicOr(t.any[j] + v, 1);
            if (cnt_n > 0) {
                const double c = fmax((double)cnt - 1.0, 0.1);
                s = exp(log((double)cnt_n / n_total) + log(c / (double)cnt_n));
            }
        }
        t.score[j][e] = s;
    }
}

__global__ void __launch_bounds__(kThreads) k_domain_score(const __grid_constant__ DomainParams p,
                                                           const __grid_constant__ ScoreTables t,
                                                           const int32_t* __restrict__ rows, int64_t n_cells,
                                                           const int32_t* __restrict__ target, int dom_t,
                                                           const int64_t* __restrict__ hist_t, double beta,
                                                           int32_t* __restrict__ out_top1,
                                                           double* __restrict__ out_prob,
                                                           uint8_t* __restrict__ out_weak) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
        const int r = rows[i];
        const int cur = target[r];
        // rows of the score tables selected by this cell's correlated values
        int vals[kMaxCorr];
#pragma unroll
        for (int j = 0; j < kMaxCorr; ++j) vals[j] = j < p.n_corr ? p.corr[j][r] : -1;
        const double* tab[kMaxCorr];
        int start = 0;
#pragma unroll
        for (int j = 0; j < kMaxCorr; ++j) {
            tab[j] = nullptr;
            if (j < p.n_corr) {
                const int v = vals[j];
                const bool any = v >= 0 && v < p.dom_c[j] && __ldg(t.any[j] + v) != 0;
                if (any) tab[j] = t.score[j] + (int64_t)v * dom_t;
                // IF(ISNOTNULL(l.domain), CONCAT(l.domain, r.d), r.d): a missing r.d nulls the prefix
                if (!any) start = j + 1;
            }
        }
        double denom = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double best = -1.0;
            int best_n = -1;
            for (int n = 0; n < dom_t; ++n) {
                if (__ldg(hist_t + n + 1) <= 0) continue;
                double score = 0.0;
                bool hit = false;
#pragma unroll
                for (int j = 0; j < kMaxCorr; ++j) {
                    if (j >= start && j < p.n_corr) {
                        const double s = __ldg(tab[j] + n);
                        if (s >= 0.0) { score += s; hit = true; }
                    }
                }
                if (!hit) continue;
                if (pass == 0) {
                    denom += score;
                } else {
                    const double prob = score / denom;
                    if (prob > beta && prob > best) { best = prob; best_n = n; }
                }
            }
            if (pass == 1) {
                out_top1[i] = best_n;
                out_prob[i] = best_n >= 0 ? best : 0.0;
                out_weak[i] = (best_n >= 0 && cur == best_n) ? 1 : 0;
            }
        }
    }
}

// ---- all targets in three launches (dr_domain_prune) ------------------------------------------------
// A cell's verdict only depends on its correlated values, so the top-1 candidate is computed once per
// COMBINATION of correlated values (<= 81^2 with the default two correlated attributes) by the very
// loop k_domain_score runs per cell -- identical doubles -- and the pass over the cells is a table
// look-up driven by the error bitmap itself: no row lists, no per-cell arrays.
struct PruneTarget {
    DomainParams p;
    ScoreTables t;
    const int32_t* target;
    uint32_t* bitmap;
    const int64_t* hist_t;
    int32_t* combo_top1;     // [prod (dom_c[j] + 1)], index = sum (v_j + 1) * stride_j, slot 0 = NULL
    int64_t n_combos;
    int32_t stride[kMaxCorr];
    int dom_t;
};

__global__ void __launch_bounds__(kThreads) k_domain_prepare_many(const PruneTarget* __restrict__ targets,
                                                                  double n_total) {
    const PruneTarget& T = targets[blockIdx.y];
    for (int j = 0; j < T.p.n_corr; ++j) {
        const int64_t n_entries = (int64_t)T.p.dom_c[j] * T.dom_t;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
            const int v = (int)(e / T.dom_t), n = (int)(e % T.dom_t);
            const long long cnt = T.p.cooc[j][(int64_t)(v + 1) * (T.dom_t + 1) + 1 + n];
            const long long cnt_n = T.hist_t[n + 1];
            double s = -1.0;
            if (cnt > T.p.tau[j]) {
                atomicOr(T.t.any[j] + v, 1);
                if (cnt_n > 0) {
                    const double c = fmax((double)cnt - 1.0, 0.1);
                    s = exp(log((double)cnt_n / n_total) + log(c / (double)cnt_n));
                }
            }
            T.t.score[j][e] = s;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_domain_combos(const PruneTarget* __restrict__ targets, double beta) {
    const PruneTarget& T = targets[blockIdx.y];
    const int dom_t = T.dom_t;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < T.n_combos; q += stride) {
        const double* tab[kMaxCorr];
        int start = 0;
        for (int j = 0; j < kMaxCorr; ++j) {
            tab[j] = nullptr;
            if (j < T.p.n_corr) {
                const int v = (int)((q / T.stride[j]) % (T.p.dom_c[j] + 1)) - 1;
                const bool any = v >= 0 && T.t.any[j][v] != 0;
                if (any) tab[j] = T.t.score[j] + (int64_t)v * dom_t;
                if (!any) start = j + 1;
            }
        }
        double denom = 0.0;
        int top1 = -1;
        for (int pass = 0; pass < 2; ++pass) {
            double best = -1.0;
            int best_n = -1;
            for (int n = 0; n < dom_t; ++n) {
                if (T.hist_t[n + 1] <= 0) continue;
                double score = 0.0;
                bool hit = false;
                for (int j = 0; j < kMaxCorr; ++j) {
                    if (j >= start && j < T.p.n_corr) {
                        const double s = tab[j][n];
                        if (s >= 0.0) { score += s; hit = true; }
                    }
                }
                if (!hit) continue;
                if (pass == 0) {
                    denom += score;
                } else {
                    const double prob = score / denom;
                    if (prob > beta && prob > best) { best = prob; best_n = n; }
                }
            }
            top1 = best_n;
        }
        T.combo_top1[q] = top1;
    }
}

// one thread per bitmap word; the few set bits of a word are looked up one after the other
__global__ void __launch_bounds__(256) k_domain_prune_bits(const PruneTarget* __restrict__ targets, int64_t n_rows,
                                                           int64_t* __restrict__ removed) {
    const PruneTarget& T = targets[blockIdx.y];
    const int64_t n_words = (n_rows + 31) >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int mine = 0;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        uint32_t bits = T.bitmap[w];
        if (!bits) continue;
        uint32_t keep = bits;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int64_t r = (w << 5) + b;
            if (r >= n_rows) break;
            const int cur = T.target[r];
            if (cur < 0) continue;   // a NULL cell is never its own top-1 candidate
            int64_t q = 0;
            for (int j = 0; j < T.p.n_corr; ++j) {
                const int v = T.p.corr[j][r];
                q += (int64_t)((v >= 0 && v < T.p.dom_c[j]) ? v + 1 : 0) * T.stride[j];
            }
            if (__ldg(T.combo_top1 + q) == cur) { keep &= ~(1u << b); ++mine; }
        }
        if (keep != T.bitmap[w]) T.bitmap[w] = keep;
    }
    // block total -> one atomic
    __shared__ int s_sum[8];
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < 8; ++i) tot += s_sum[i];
        if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(removed + blockIdx.y), (unsigned long long)tot);
    }
}

}  // namespace

extern "C" int dr_domain_score(dr_ctx* ctx, const int32_t* rows, int64_t n_cells, const int32_t* target,
                               int32_t dom_t, const int32_t* const* corr, const int32_t* dom_c,
                               const int64_t* const* cooc, int n_corr, const int64_t* hist_t, const int64_t* tau,
                               int64_t n_total_rows, double beta, int32_t* out_top1, double* out_prob,
                               uint8_t* out_weak, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_cells <= 0) return DR_OK;
    DR_REQUIRE(ctx, rows && target && hist_t && out_top1 && out_prob && out_weak, "null pointer");
    DR_REQUIRE(ctx, n_corr >= 0 && n_corr <= kMaxCorr, "n_corr must be in [0, 8]");
    DR_REQUIRE(ctx, n_corr == 0 || (corr && dom_c && cooc && tau), "null pointer");
    DR_REQUIRE(ctx, n_total_rows > 0 && dom_t >= 0, "bad sizes");
    DomainParams p;
    memset(&p, 0, sizeof(p));
    p.n_corr = n_corr;
    for (int j = 0; j < n_corr; ++j) {
        DR_REQUIRE(ctx, corr[j] && cooc[j], "null correlated column / table");
        p.corr[j] = corr[j];
        p.cooc[j] = cooc[j];
        p.dom_c[j] = dom_c[j];
        p.tau[j] = tau[j];
    }
    // score tables + "any candidate" flags live in the context scratch buffer
    ScoreTables t;
    memset(&t, 0, sizeof(t));
    size_t bytes = 0;
    for (int j = 0; j < n_corr; ++j) bytes += (size_t)dom_c[j] * dom_t * sizeof(double);
    const size_t flags_at = bytes;
    for (int j = 0; j < n_corr; ++j) bytes += (size_t)dom_c[j] * sizeof(int32_t);
    int rc = dr_ensure_scratch(ctx, bytes + 16);
    if (rc) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    size_t at = 0, fat = flags_at;
    int max_dom = 0;
    for (int j = 0; j < n_corr; ++j) {
        t.score[j] = reinterpret_cast<double*>(base + at);
        t.any[j] = reinterpret_cast<int32_t*>(base + fat);
        at += (size_t)dom_c[j] * dom_t * sizeof(double);
        fat += (size_t)dom_c[j] * sizeof(int32_t);
        if (dom_c[j] > max_dom) max_dom = dom_c[j];
    }
    if (n_corr > 0 && dom_t > 0 && max_dom > 0) {
        DR_CUDA(ctx, cudaMemsetAsync(base + flags_at, 0, bytes - flags_at, (cudaStream_t)stream));
        const int64_t entries = (int64_t)max_dom * dom_t;
        dim3 grid((unsigned)((entries + kThreads - 1) / kThreads), (unsigned)n_corr);
        k_domain_prepare<<<grid, kThreads, 0, (cudaStream_t)stream>>>(p, t, dom_t, hist_t, (double)n_total_rows);
        DR_LAUNCHED(ctx);
    }
    k_domain_score<<<dr_grid_for(ctx, n_cells, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        p, t, rows, n_cells, target, dom_t, hist_t, beta, out_top1, out_prob, out_weak);
    DR_LAUNCHED(ctx);
    return DR_OK;
}

extern "C" int dr_domain_prune(dr_ctx* ctx, const dr_domain_target* targets, int n_targets, int64_t n_rows,
                               int64_t n_total_rows, double beta, int64_t* removed, void* stream) {
    if (!ctx) return DR_ERR_INVALID;
    if (n_targets <= 0 || n_rows <= 0) return DR_OK;
    DR_REQUIRE(ctx, targets && removed && n_total_rows > 0, "null pointer");
    // scratch layout: [PruneTarget x n_targets | per target: score tables, any flags, combo table]
    std::vector<PruneTarget> host((size_t)n_targets);
    size_t at = ((size_t)n_targets * sizeof(PruneTarget) + 255) / 256 * 256;
    int64_t max_entries = 1, max_combos = 1;
    std::vector<size_t> off_score((size_t)n_targets * kMaxCorr), off_any((size_t)n_targets * kMaxCorr),
        off_combo((size_t)n_targets);
    size_t zero_from = 0, zero_to = 0;
    for (int i = 0; i < n_targets; ++i) {
        const dr_domain_target& d = targets[i];
        DR_REQUIRE(ctx, d.target && d.bitmap && d.hist_t && d.dom_t >= 0 && d.n_corr >= 1 && d.n_corr <= kMaxCorr,
                   "bad target");
        PruneTarget& T = host[(size_t)i];
        memset(&T, 0, sizeof(T));
        T.target = d.target; T.bitmap = d.bitmap; T.hist_t = d.hist_t; T.dom_t = d.dom_t; T.p.n_corr = d.n_corr;
        int64_t combos = 1;
        for (int j = 0; j < d.n_corr; ++j) {
            DR_REQUIRE(ctx, d.corr[j] && d.cooc[j] && d.dom_c[j] >= 0, "null correlated column / table");
            T.p.corr[j] = d.corr[j]; T.p.cooc[j] = d.cooc[j]; T.p.dom_c[j] = d.dom_c[j]; T.p.tau[j] = d.tau[j];
            T.stride[j] = (int32_t)combos;
            combos *= d.dom_c[j] + 1;
            if (combos > (1 << 20))
                return dr_fail(ctx, DR_ERR_UNSUPPORTED, "more than 2^20 combinations of correlated values (%lld)",
                               (long long)combos);
            off_score[(size_t)i * kMaxCorr + j] = at;
            at += ((size_t)d.dom_c[j] * d.dom_t * sizeof(double) + 255) / 256 * 256;
            if ((int64_t)d.dom_c[j] * d.dom_t > max_entries) max_entries = (int64_t)d.dom_c[j] * d.dom_t;
        }
        T.n_combos = combos;
        if (combos > max_combos) max_combos = combos;
        off_combo[(size_t)i] = at;
        at += ((size_t)combos * sizeof(int32_t) + 255) / 256 * 256;
    }
    zero_from = at;   // the "any candidate" flags of all targets sit together: one memset
    for (int i = 0; i < n_targets; ++i)
        for (int j = 0; j < targets[i].n_corr; ++j) {
            off_any[(size_t)i * kMaxCorr + j] = at;
            at += ((size_t)targets[i].dom_c[j] * sizeof(int32_t) + 255) / 256 * 256;
        }
    zero_to = at;
    int rc = dr_ensure_scratch(ctx, at + 256);
    if (rc) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->scratch);
    for (int i = 0; i < n_targets; ++i) {
        PruneTarget& T = host[(size_t)i];
        for (int j = 0; j < T.p.n_corr; ++j) {
            T.t.score[j] = reinterpret_cast<double*>(base + off_score[(size_t)i * kMaxCorr + j]);
            T.t.any[j] = reinterpret_cast<int32_t*>(base + off_any[(size_t)i * kMaxCorr + j]);
        }
        T.combo_top1 = reinterpret_cast<int32_t*>(base + off_combo[(size_t)i]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    DR_CUDA(ctx, cudaMemcpyAsync(base, host.data(), host.size() * sizeof(PruneTarget), cudaMemcpyHostToDevice, st));
    if (zero_to > zero_from) DR_CUDA(ctx, cudaMemsetAsync(base + zero_from, 0, zero_to - zero_from, st));
    const PruneTarget* dev = reinterpret_cast<const PruneTarget*>(base);
    {
        dim3 grid((unsigned)((max_entries + kThreads - 1) / kThreads), (unsigned)n_targets);
        k_domain_prepare_many<<<grid, kThreads, 0, st>>>(dev, (double)n_total_rows);
        DR_LAUNCHED(ctx);
    }
    {
        dim3 grid((unsigned)((max_combos + kThreads - 1) / kThreads), (unsigned)n_targets);
        k_domain_combos<<<grid, kThreads, 0, st>>>(dev, beta);
        DR_LAUNCHED(ctx);
    }
    {
        const int64_t n_words = (n_rows + 31) >> 5;
        int gx = dr_grid_for(ctx, n_words, 256 * 4, 8);
        if (gx * n_targets < ctx->sm_count * 8) gx = dr_grid_for(ctx, n_words, 256, 8);
        dim3 grid((unsigned)gx, (unsigned)n_targets);
        k_domain_prune_bits<<<grid, 256, 0, st>>>(dev, n_rows, removed);
        DR_LAUNCHED(ctx);
    }
    // (the host array was copied into the stream's staging area by cudaMemcpyAsync: pageable source)
    DR_CUDA(ctx, cudaStreamSynchronize(st));
    return DR_OK;
}
