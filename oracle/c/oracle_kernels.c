/*
 * Plain-C restatement of the oracle's three heavy loops (TEST INFRASTRUCTURE, see
 * oracle/__init__.py): used by bench.py's cpu_baseline / --impl reference legs so that the CPU
 * side is timed with compiled, multi-threaded (OpenMP) code instead of NumPy dispatch overhead,
 * and by tests/test_oracle_c.py, which checks every function against the NumPy oracle.
 *
 *  o_scan_hist       NULL counts + per-column histograms   (ErrorDetectorApi.scala:128-157,
 *                                                           RepairApi.scala:231-273 singles)
 *  o_cooc            pair co-occurrence counts             (RepairApi.scala:231-273 pairs)
 *  o_forest_margins  flat-forest evaluation, margins summed in tree order
 *                                                          (model.py:1118-1133 -> LightGBM predict)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs ask for the host's cores */
void o_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* cols: K pointers to int32[n]; hist: int64[sum(dom+1)], slot 0 = NULL */
void o_scan_hist(const int32_t* const* cols, const int32_t* dom, int k, int64_t n, int64_t* hist) {
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (k + 1));
    off[0] = 0;
    for (int c = 0; c < k; ++c) off[c + 1] = off[c] + dom[c] + 1;
    memset(hist, 0, sizeof(int64_t) * off[k]);
#pragma omp parallel
    {
        int64_t* local = (int64_t*)calloc(off[k], sizeof(int64_t));
#pragma omp for schedule(static) collapse(1)
        for (int64_t blk = 0; blk < (n + 65535) / 65536; ++blk) {
            const int64_t lo = blk * 65536, hi = lo + 65536 < n ? lo + 65536 : n;
            for (int c = 0; c < k; ++c) {
                const int32_t* col = cols[c];
                int64_t* h = local + off[c];
                for (int64_t r = lo; r < hi; ++r) h[col[r] + 1]++;
            }
        }
#pragma omp critical
        for (int64_t i = 0; i < off[k]; ++i) hist[i] += local[i];
        free(local);
    }
    free(off);
}

/* out: int64, pair q at tab_off[q], entry (cx+1)*(dom[y]+1)+(cy+1) */
void o_cooc(const int32_t* const* cols, const int32_t* dom, const int32_t* px, const int32_t* py, int n_pairs,
            const int64_t* tab_off, int64_t n, int64_t* out) {
    memset(out, 0, sizeof(int64_t) * tab_off[n_pairs]);
#pragma omp parallel
    {
        int64_t* local = (int64_t*)calloc(tab_off[n_pairs], sizeof(int64_t));
#pragma omp for schedule(static)
        for (int64_t blk = 0; blk < (n + 16383) / 16384; ++blk) {
            const int64_t lo = blk * 16384, hi = lo + 16384 < n ? lo + 16384 : n;
            for (int q = 0; q < n_pairs; ++q) {
                const int32_t *x = cols[px[q]], *y = cols[py[q]];
                const int64_t ny = dom[py[q]] + 1;
                int64_t* t = local + tab_off[q];
                for (int64_t r = lo; r < hi; ++r) t[(int64_t)(x[r] + 1) * ny + (y[r] + 1)]++;
            }
        }
#pragma omp critical
        for (int64_t i = 0; i < tab_off[n_pairs]; ++i) out[i] += local[i];
        free(local);
    }
}

/* X: float64[n][n_feat] row-major, NaN = missing; raw: float64[n][n_seq] */
void o_forest_margins(int64_t n, int n_feat, const double* X, int n_seq, const double* baseline, int n_trees,
                      const int32_t* tree_seq, const int64_t* tree_offset, const int32_t* feature,
                      const double* threshold, const uint8_t* missing_left, const int32_t* left,
                      const int32_t* right, const double* value, double* raw) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double* x = X + i * n_feat;
        double* r = raw + i * n_seq;
        for (int s = 0; s < n_seq; ++s) r[s] = baseline[s];
        for (int t = 0; t < n_trees; ++t) {
            const int64_t base = tree_offset[t];
            int64_t node = 0;
            while (feature[base + node] >= 0) {
                const double v = x[feature[base + node]];
                const int go_left = isnan(v) ? missing_left[base + node] == 1 : v <= threshold[base + node];
                node = go_left ? left[base + node] : right[base + node];
            }
            r[tree_seq[t]] += value[base + node];
        }
    }
}
