// Context management for libb200repair.so.
#include "common.cuh"

extern "C" {

int dr_abi_version(void) { return 4; }

int dr_ctx_create(int device, dr_ctx** out) {
    if (!out) return DR_ERR_INVALID;
    *out = nullptr;
    dr_ctx* ctx = new dr_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        // keep the ctx alive so the caller can read the message
        snprintf(ctx->err, sizeof(ctx->err), "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
        *out = ctx;
        return DR_ERR_CUDA;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
        *out = ctx;
        return DR_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    DR_CUDA(ctx, cudaMallocHost(&ctx->pinned, 4096));
    return dr_ensure_scratch(ctx, 1u << 20);
}

int dr_ctx_destroy(dr_ctx* ctx) {
    if (!ctx) return DR_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    delete ctx;
    return DR_OK;
}

const char* dr_last_error(const dr_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int64_t dr_launch_count(const dr_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
