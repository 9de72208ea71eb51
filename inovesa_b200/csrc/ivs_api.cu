// ivs_api.cu -- context, memory and transfer entry points of the C ABI (include/ivs.h).
//
// This is what takes the place of OCLH (src/CL/OpenCLHandler.cpp): one device, one in-order
// stream, thin wrappers with status codes instead of cl::Error exceptions.
#include <cstdio>
#include <cstring>

#include "ivs_common.cuh"

namespace {
std::string g_create_error;
}

ivs_status ivs_fail(ivs_ctx* ctx, ivs_status code, const std::string& msg) {
    if (ctx) ctx->error = msg; else g_create_error = msg;
    return code;
}

ivs_status ivs_check_cuda(ivs_ctx* ctx, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return ivs_fail(ctx, IVS_ERR_CUDA, msg);
}

extern "C" {

const char* ivs_version(void) { return "ivs 0.1 (B200 VFP step; CUDA sm_100a)"; }

int ivs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

ivs_status ivs_ctx_create(int device, int arith, ivs_ctx** out) {
    if (!out) return ivs_fail(nullptr, IVS_ERR_INVALID_ARGUMENT, "ivs_ctx_create: out is NULL");
    *out = nullptr;
    if (arith != IVS_ARITH_FMA && arith != IVS_ARITH_STRICT)
        return ivs_fail(nullptr, IVS_ERR_INVALID_ARGUMENT, "ivs_ctx_create: unknown arithmetic mode");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return ivs_fail(nullptr, IVS_ERR_NO_DEVICE,
                        "ivs_ctx_create: no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n)
        return ivs_fail(nullptr, IVS_ERR_INVALID_ARGUMENT, "ivs_ctx_create: device index out of range");
    ivs_ctx* ctx = new ivs_ctx;
    ctx->device = device;
    ctx->arith = arith;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete ctx; return ivs_check_cuda(nullptr, e, "cudaSetDevice"); }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return ivs_check_cuda(nullptr, e, "cudaGetDeviceProperties"); }
    if (prop.major != 10) {
        delete ctx;
        return ivs_fail(nullptr, IVS_ERR_UNSUPPORTED, "ivs_ctx_create: kernels are built for sm_100a (Blackwell B200) only");
    }
    ctx->sm_count = prop.multiProcessorCount;
    // the context's stream is the most urgent one: side streams of a simulation (tracked particles) only
    // get the SMs it leaves idle
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) prio_lo = prio_hi = 0;
    if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) { delete ctx; return ivs_check_cuda(nullptr, e, "cudaStreamCreate"); }
    if ((e = cudaMalloc((void**)&ctx->scratch, 4096)) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return ivs_check_cuda(nullptr, e, "cudaMalloc"); }
    *out = ctx;
    return IVS_OK;
}

ivs_status ivs_ctx_destroy(ivs_ctx* ctx) {
    if (!ctx) return IVS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 16; i++) if (ctx->twiddle[i]) cudaFree(ctx->twiddle[i]);
    cudaFree(ctx->scratch);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return IVS_OK;
}

const char* ivs_last_error(const ivs_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }
void* ivs_ctx_stream(ivs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t ivs_ctx_launch_count(const ivs_ctx* ctx) { return ctx ? ctx->launches : 0; }

ivs_status ivs_malloc(ivs_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_ptr != nullptr, "ivs_malloc: d_ptr is NULL");
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));
    IVS_CUDA(ctx, cudaMalloc(d_ptr, bytes ? bytes : 1));
    return IVS_OK;
}

ivs_status ivs_free(ivs_ctx* ctx, void* d_ptr) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    if (!d_ptr) return IVS_OK;
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    IVS_CUDA(ctx, cudaFree(d_ptr));
    return IVS_OK;
}

ivs_status ivs_upload(ivs_ctx* ctx, void* d_dst, const void* host, size_t bytes) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, (d_dst && host) || bytes == 0, "ivs_upload: null pointer");
    IVS_CUDA(ctx, cudaMemcpyAsync(d_dst, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return IVS_OK;
}

ivs_status ivs_download(ivs_ctx* ctx, void* host, const void* d_src, size_t bytes) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, (host && d_src) || bytes == 0, "ivs_download: null pointer");
    IVS_CUDA(ctx, cudaMemcpyAsync(host, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IVS_OK;
}

ivs_status ivs_copy(ivs_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, (d_dst && d_src) || bytes == 0, "ivs_copy: null pointer");
    IVS_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return IVS_OK;
}

ivs_status ivs_memset_zero(ivs_ctx* ctx, void* d_dst, size_t bytes) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_dst || bytes == 0, "ivs_memset_zero: null pointer");
    IVS_CUDA(ctx, cudaMemsetAsync(d_dst, 0, bytes, ctx->stream));
    return IVS_OK;
}

ivs_status ivs_host_alloc(ivs_ctx* ctx, size_t bytes, void** host_ptr) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, host_ptr != nullptr, "ivs_host_alloc: host_ptr is NULL");
    IVS_CUDA(ctx, cudaMallocHost(host_ptr, bytes ? bytes : 1));
    return IVS_OK;
}

ivs_status ivs_host_free(ivs_ctx* ctx, void* host_ptr) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    if (host_ptr) IVS_CUDA(ctx, cudaFreeHost(host_ptr));
    return IVS_OK;
}

ivs_status ivs_flush(ivs_ctx* ctx) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    // work is submitted eagerly; cudaStreamQuery nudges the driver like clFlush
    cudaError_t e = cudaStreamQuery(ctx->stream);
    if (e != cudaSuccess && e != cudaErrorNotReady) return ivs_check_cuda(ctx, e, "cudaStreamQuery");
    return IVS_OK;
}

ivs_status ivs_sync(ivs_ctx* ctx) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IVS_OK;
}

}  // extern "C"
