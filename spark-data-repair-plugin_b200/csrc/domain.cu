// dr_domain_score: weak-label cell-domain analysis for one target attribute (SURVEY.md 8a row a9,
// RepairApi.scala:479-675).  The naive-Bayes term of a (correlated value, candidate) pair depends on
// three counts only, so a first kernel evaluates exp(log(.) + log(.)) once per table entry
// (<= 81 x 81 per correlated attribute) into a score table; the per-cell kernel -- one thread per noisy
// cell -- then only adds table entries, in the same order as before: identical doubles, no
// transcendental math on the 10^7-cell path.
#include <math.h>

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
