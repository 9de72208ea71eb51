// ivs_reduce.cu -- PhaseSpace projections, integral, normalisation and moments on the device.
//
// The reference computes these as sequential fp32 sums on the host (src/PS/PhaseSpace.cpp:
// 229-340), pulling the whole grid over the bus for the Y projection, normalize() and the
// moments.  Here they are warp-shuffle / block tree reductions.  The products
// are rounded exactly like the reference's (no contraction); the summation ORDER is a fixed
// tree instead of left-to-right, so results agree to fp32 round-off of the sum (~1e-7
// relative), deterministically (no atomics).
#include "ivs_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in every thread; blockDim.x multiple of 32, <= 1024
__device__ float block_sum(float v, float* smem /* 32 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    float t = (lane < nw) ? smem[lane] : 0.0f;
    return warp_sum(t);
}
__device__ double block_sum_d(double v, double* smem /* 32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double t = (lane < nw) ? smem[lane] : 0.0;
    return warp_sum_d(t);
}

// PhaseSpace::updateXProjection: proj[x] = sum_y f[x][y]*ws[y]; one warp per row, lanes stride
// over y (coalesced), 4 independent partial sums per lane for ILP, then a shuffle tree.
__global__ void __launch_bounds__(256)
project_x_kernel(const float* __restrict__ grid, uint32_t n, const float* __restrict__ ws,
                 float* __restrict__ proj, uint32_t rows_total) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= rows_total) return;
    const float* row = grid + (size_t)warp * n;
    float acc = 0.0f;
    for (uint32_t y = lane; y < n; y += 32) acc = __fadd_rn(acc, __fmul_rn(row[y], ws[y]));
    acc = warp_sum(acc);
    if (lane == 0) proj[warp] = acc;
}

// PhaseSpace::updateYProjection: proj[y] = sum_x f[x][y]*ws[x]; CTA = 32 columns x 32 row
// lanes: every warp walks rows x = warp, warp+32, ... reading 32 consecutive y (coalesced),
// partial sums are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(1024)
project_y_kernel(const float* __restrict__ grid, uint32_t n, const float* __restrict__ ws,
                 float* __restrict__ proj) {
    __shared__ float part[32][33];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t y = blockIdx.x * 32 + lane;
    const float* g = grid + (size_t)blockIdx.y * n * n;
    float acc = 0.0f;
    if (y < n) {
        for (uint32_t x = warp; x < n; x += 32) acc = __fadd_rn(acc, __fmul_rn(g[(size_t)x * n + y], ws[x]));
    }
    part[warp][lane] = acc;
    __syncthreads();
    // transpose-read: warp w sums column w over the 32 row lanes
    float v = part[lane][warp];
    v = warp_sum(v);
    const uint32_t yo = blockIdx.x * 32 + warp;
    if (lane == 0 && yo < n) proj[(size_t)blockIdx.y * n + yo] = v;
}

// PhaseSpace::integrate: one CTA per member
__global__ void __launch_bounds__(256)
integrate_kernel(const float* __restrict__ proj, uint32_t n, const float* __restrict__ ws,
                 float* __restrict__ integral) {
    __shared__ float sm[32];
    const float* p = proj + (size_t)blockIdx.x * n;
    float acc = 0.0f;
    for (uint32_t x = threadIdx.x; x < n; x += blockDim.x) acc = __fadd_rn(acc, __fmul_rn(p[x], ws[x]));
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) integral[blockIdx.x] = acc;
}

// PhaseSpace::normalize: data *= filling_set/filling (the quotient is formed first, as the
// reference's `_data[n][x][y] *= _filling_set[n]/_filling[n]`)
__global__ void __launch_bounds__(256)
normalize_kernel(float* __restrict__ grid, size_t cells_per_member, float filling_set,
                 const float* __restrict__ integral, const float* __restrict__ filling_sets) {
    // filling_sets: one target per member (the bunches of a train carry their charge fractions), else filling_set
    const float s = __fdiv_rn(filling_sets ? filling_sets[blockIdx.y] : filling_set, integral[blockIdx.y]);
    float* g = grid + (size_t)blockIdx.y * cells_per_member;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells_per_member;
         i += (size_t)gridDim.x * blockDim.x) {
        g[i] = __fmul_rn(g[i], s);
    }
}

// PhaseSpace::average + variance on one projection per CTA:
// mean = (sum proj*q) * delta/filling ; var = (sum proj*(q-mean)^2 [double]) * delta/filling
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ proj, uint32_t n, float amin, float delta,
               const float* __restrict__ integral, float* __restrict__ out3) {
    __shared__ float sm[32];
    __shared__ double smd[32];
    const float* p = proj + (size_t)blockIdx.x * n;
    const float filling = integral[blockIdx.x];
    float acc = 0.0f;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float q = __fadd_rn(amin, __fmul_rn((float)i, delta));  // Ruler::at(i)
        acc = __fadd_rn(acc, __fmul_rn(p[i], q));
    }
    acc = block_sum(acc, sm);
    const float scale = __fdiv_rn(delta, filling);
    const float mean = __fmul_rn(acc, scale);
    double vacc = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float q = __fadd_rn(amin, __fmul_rn((float)i, delta));
        const double d = (double)__fsub_rn(q, mean);  // std::pow(float,int) works in double
        vacc += (double)p[i] * (d * d);
    }
    vacc = block_sum_d(vacc, smd);
    if (threadIdx.x == 0) {
        const float var = __fmul_rn((float)vacc, scale);
        out3[blockIdx.x * 3 + 0] = mean;
        out3[blockIdx.x * 3 + 1] = var;
        out3[blockIdx.x * 3 + 2] = sqrtf(var);
    }
}

}  // namespace

// member-batched entry points shared with the simulation object (ivs_step.cu)
ivs_status ivs_project_x_batched(ivs_ctx* ctx, const float* grid, uint32_t n, uint32_t members,
                                 const float* ws, float* proj) {
    const uint32_t rows = n * members;
    project_x_kernel<<<ivs_div_up(rows, 8), 256, 0, ctx->stream>>>(grid, n, ws, proj, rows);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
ivs_status ivs_project_y_batched(ivs_ctx* ctx, const float* grid, uint32_t n, uint32_t members,
                                 const float* ws, float* proj) {
    project_y_kernel<<<dim3(ivs_div_up(n, 32), members), 1024, 0, ctx->stream>>>(grid, n, ws, proj);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
ivs_status ivs_integrate_batched(ivs_ctx* ctx, const float* proj, uint32_t n, uint32_t members,
                                 const float* ws, float* integral) {
    integrate_kernel<<<members, 256, 0, ctx->stream>>>(proj, n, ws, integral);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
ivs_status ivs_normalize_batched(ivs_ctx* ctx, float* grid, uint32_t n, uint32_t members,
                                 float filling_set, const float* integral, const float* filling_sets) {
    const size_t cells = (size_t)n * n;
    const uint32_t bx = (uint32_t)((cells + 255) / 256 < 1184 ? (cells + 255) / 256 : 1184);
    normalize_kernel<<<dim3(bx, members), 256, 0, ctx->stream>>>(grid, cells, filling_set, integral, filling_sets);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
ivs_status ivs_moments_batched(ivs_ctx* ctx, const float* proj, uint32_t n, uint32_t members,
                               float amin, float delta, const float* integral, float* out3) {
    moments_kernel<<<members, 256, 0, ctx->stream>>>(proj, n, amin, delta, integral, out3);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

extern "C" {

ivs_status ivs_project_x(ivs_ctx* ctx, const void* d_grid, uint32_t n, const void* d_ws,
                         void* d_proj) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_grid && d_ws && d_proj && n >= 2, "ivs_project_x: null pointer or n < 2");
    return ivs_project_x_batched(ctx, (const float*)d_grid, n, 1, (const float*)d_ws, (float*)d_proj);
}

ivs_status ivs_project_y(ivs_ctx* ctx, const void* d_grid, uint32_t n, const void* d_ws,
                         void* d_proj) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_grid && d_ws && d_proj && n >= 2, "ivs_project_y: null pointer or n < 2");
    return ivs_project_y_batched(ctx, (const float*)d_grid, n, 1, (const float*)d_ws, (float*)d_proj);
}

ivs_status ivs_integrate(ivs_ctx* ctx, const void* d_proj, uint32_t n, const void* d_ws,
                         void* d_integral) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_proj && d_ws && d_integral && n >= 2, "ivs_integrate: null pointer or n < 2");
    return ivs_integrate_batched(ctx, (const float*)d_proj, n, 1, (const float*)d_ws, (float*)d_integral);
}

ivs_status ivs_normalize(ivs_ctx* ctx, void* d_grid, uint32_t n, float filling_set,
                         const void* d_integral) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_grid && d_integral && n >= 2, "ivs_normalize: null pointer or n < 2");
    return ivs_normalize_batched(ctx, (float*)d_grid, n, 1, filling_set, (const float*)d_integral, nullptr);
}

ivs_status ivs_moments(ivs_ctx* ctx, const void* d_proj, uint32_t n, float amin, float delta,
                       const void* d_integral, void* d_out3) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_proj && d_integral && d_out3 && n >= 2, "ivs_moments: null pointer or n < 2");
    return ivs_moments_batched(ctx, (const float*)d_proj, n, 1, amin, delta, (const float*)d_integral,
                               (float*)d_out3);
}

}  // extern "C"
