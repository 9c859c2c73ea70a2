// Shared helpers for libb200repair.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "b200repair.h"

struct dr_ctx {
    int device;
    int sm_count;
    int64_t launches;
    char err[512];
    // small scratch: block counters for compaction, host-readable results
    void* scratch;        // device
    size_t scratch_bytes;
    void* pinned;         // host pinned, 4 KB
};

static inline int dr_fail(dr_ctx* ctx, int code, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}
template <typename A, typename... Args>
static inline int dr_fail(dr_ctx* ctx, int code, const char* fmt, A a, Args... args) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, args...);
    return code;
}

#define DR_CUDA(ctx, expr)                                                                     \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s failed: %s (%s:%d)", #expr,          \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                             \
            return DR_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define DR_REQUIRE(ctx, cond, msg)                                                             \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            snprintf((ctx)->err, sizeof((ctx)->err), "invalid argument: %s (%s)", msg, #cond); \
            return DR_ERR_INVALID;                                                             \
        }                                                                                      \
    } while (0)

// Count a launch and surface launch-configuration errors immediately.
#define DR_LAUNCHED(ctx)                 \
    do {                                 \
        (ctx)->launches++;               \
        DR_CUDA(ctx, cudaGetLastError()); \
    } while (0)

static inline int dr_grid_for(const dr_ctx* ctx, int64_t work_items, int per_block, int ctas_per_sm) {
    int64_t need = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)ctx->sm_count * ctas_per_sm;  // persistent: a multiple of the SM count
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static inline int dr_ensure_scratch(dr_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return DR_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    DR_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return DR_OK;
}

static __device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
