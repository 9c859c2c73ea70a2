// dr_domain_score: weak-label cell-domain analysis for one target attribute (SURVEY.md 8a row a9,
// RepairApi.scala:479-675).  One thread per noisy cell; the count tables (<= 81 x 81 int64 with the
// default discrete threshold) stay L1/L2 resident, laid out so that the candidate index is the
// fastest-varying one.
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

__global__ void __launch_bounds__(kThreads) k_domain_score(const __grid_constant__ DomainParams p,
                                                           const int32_t* __restrict__ rows, int64_t n_cells,
                                                           const int32_t* __restrict__ target, int dom_t,
                                                           const int64_t* __restrict__ hist_t, double n_total,
                                                           double beta, int32_t* __restrict__ out_top1,
                                                           double* __restrict__ out_prob,
                                                           uint8_t* __restrict__ out_weak) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += stride) {
        const int r = rows[i];
        const int cur = target[r];
        // rows of the count tables selected by this cell's correlated values
        const int64_t* tab[kMaxCorr];
        int start = 0;
        for (int j = 0; j < p.n_corr; ++j) {
            const int v = p.corr[j][r];
            tab[j] = nullptr;
            bool any = false;
            if (v >= 0 && v < p.dom_c[j]) {
                const int64_t* t = p.cooc[j] + (int64_t)(v + 1) * (dom_t + 1) + 1;  // entry [v][n], n >= 0
                for (int n = 0; n < dom_t && !any; ++n) any = __ldg(t + n) > p.tau[j];
                if (any) tab[j] = t;
            }
            // IF(ISNOTNULL(l.domain), CONCAT(l.domain, r.d), r.d): a missing r.d nulls the prefix
            if (!any) start = j + 1;
        }
        double denom = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            double best = -1.0;
            int best_n = -1;
            for (int n = 0; n < dom_t; ++n) {
                const long long cnt_n = __ldg(hist_t + n + 1);
                if (cnt_n <= 0) continue;
                double score = 0.0;
                bool hit = false;
                for (int j = start; j < p.n_corr; ++j) {
                    const long long cnt = __ldg(tab[j] + n);
                    if (cnt > p.tau[j]) {
                        const double c = fmax((double)cnt - 1.0, 0.1);
                        score += exp(log((double)cnt_n / n_total) + log(c / (double)cnt_n));
                        hit = true;
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
    k_domain_score<<<dr_grid_for(ctx, n_cells, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        p, rows, n_cells, target, dom_t, hist_t, (double)n_total_rows, beta, out_top1, out_prob, out_weak);
    DR_LAUNCHED(ctx);
    return DR_OK;
}
