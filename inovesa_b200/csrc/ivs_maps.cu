// ivs_maps.cu -- the reference's per-map kernels, one launch per map ("unfused" path).
//
// These are the 1:1 replacements of the OpenCL kernels K1-K6 of SURVEY.md section 2a; the
// arithmetic follows the reference's CPU loops (the parity oracle), not its OpenCL variants.
// They accept any grid size and any displacement (every tap is bounds checked), and are what
// the fused step kernel (ivs_step.cu) is verified against on the device.
#include <algorithm>

#include "ivs_common.cuh"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- KickMap, Axis::y
// src/SM/KickMap.cpp:243-266.  One CTA per row x: the row's taps are computed once (thread 0)
// with the arithmetic of updateSM(), the row is staged in shared memory (coalesced, read once
// from L2/HBM), every thread produces outputs y = tid, tid+blockDim, ...
template <bool FMA>
__global__ void __launch_bounds__(kThreads)
kick_y_kernel(const float* __restrict__ in, float* __restrict__ out, uint32_t n,
              const float* __restrict__ offset, uint32_t offset_stride, uint32_t it) {
    extern __shared__ float row[];
    __shared__ ivs_taps taps;
    const uint32_t x = blockIdx.x;  // global row = member*n + row
    if (threadIdx.x == 0) taps = ivs_kick_taps(offset[(x / n) * offset_stride + (x % n)], n, it);
    const float* src = in + (size_t)x * n;
    for (uint32_t y = threadIdx.x; y < n; y += blockDim.x) row[y] = src[y];
    __syncthreads();
    const int first = taps.first;
    const float w0 = taps.w[0], w1 = taps.w[1], w2 = taps.w[2], w3 = taps.w[3];
    float* dst = out + (size_t)x * n;
    for (uint32_t y = threadIdx.x; y < n; y += blockDim.x) {
        const int p = (int)y + first;
        float v = 0.0f;
        if ((uint32_t)(p) < n) v = madd<FMA>(row[p], w0, v);
        if ((uint32_t)(p + 1) < n) v = madd<FMA>(row[p + 1], w1, v);
        if ((uint32_t)(p + 2) < n) v = madd<FMA>(row[p + 2], w2, v);
        if ((uint32_t)(p + 3) < n) v = madd<FMA>(row[p + 3], w3, v);
        dst[y] = v;
    }
}

// ---------------------------------------------------------------- KickMap, Axis::x
// src/SM/KickMap.cpp:221-242.  Thread = column y (coalesced along y), CTA = 128 columns x a
// strip of rows; the per-column taps are computed once per thread and reused down the strip, and the
// four source rows of consecutive output rows overlap in three: they slide through registers, one load
// per output row (rows outside the grid are skipped taps, as in the reference's loop).
template <bool FMA>
__global__ void __launch_bounds__(kThreads)
kick_x_kernel(const float* __restrict__ in_all, float* __restrict__ out_all, uint32_t n,
              const float* __restrict__ offset, uint32_t offset_stride, uint32_t it, uint32_t strip) {
    const uint32_t y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= n) return;
    const float* in = in_all + (size_t)blockIdx.z * n * n;
    float* out = out_all + (size_t)blockIdx.z * n * n;
    const ivs_taps t = ivs_kick_taps(offset[blockIdx.z * offset_stride + y], n, it);
    const uint32_t x0 = blockIdx.y * strip;
    const uint32_t x1 = min(x0 + strip, n);
    auto row = [&](int p) { return (uint32_t)p < n ? __ldg(in + (size_t)p * n + y) : 0.0f; };
    int p = (int)x0 + t.first;
    float r0 = row(p), r1 = row(p + 1), r2 = row(p + 2), r3;
#pragma unroll 4
    for (uint32_t x = x0; x < x1; x++, p++) {
        r3 = row(p + 3);
        float v = 0.0f;
        if ((uint32_t)(p) < n) v = madd<FMA>(r0, t.w[0], v);
        if ((uint32_t)(p + 1) < n) v = madd<FMA>(r1, t.w[1], v);
        if ((uint32_t)(p + 2) < n) v = madd<FMA>(r2, t.w[2], v);
        if ((uint32_t)(p + 3) < n) v = madd<FMA>(r3, t.w[3], v);
        out[(size_t)x * n + y] = v;
        r0 = r1; r1 = r2; r2 = r3;
    }
}

// ---------------------------------------------------------------- FokkerPlanckMap
// src/SM/FokkerPlanckMap.cpp:197-211 with the stencil as 5 dense weights on y-2..y+2 (the
// reference's 3 or 4 taps in ascending order; the padding weights are exactly 0).
template <bool FMA>
__global__ void __launch_bounds__(kThreads)
fokker_planck_kernel(const float* __restrict__ in, float* __restrict__ out, uint32_t n,
                     const float* __restrict__ w5) {
    extern __shared__ float row[];
    const uint32_t x = blockIdx.x;  // global row = member*n + row
    const float* src = in + (size_t)x * n;
    for (uint32_t y = threadIdx.x; y < n; y += blockDim.x) row[y] = src[y];
    __syncthreads();
    float* dst = out + (size_t)x * n;
    for (uint32_t y = threadIdx.x; y < n; y += blockDim.x) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int p = (int)y + k - 2;
            const float w = w5[y * 5 + k];
            if (w != 0.0f && (uint32_t)p < n) v = madd<FMA>(row[p], w, v);
        }
        dst[y] = v;
    }
}

// ---------------------------------------------------------------- generic gather
// src/SM/SourceMap.cpp:98-107 (kernel applySM1D :121-139)
struct hi_t {
    uint32_t index;
    float weight;
};
template <bool FMA>
__global__ void __launch_bounds__(kThreads)
source_map_kernel(const float* __restrict__ in, float* __restrict__ out, uint32_t ncells,
                  const hi_t* __restrict__ sm, uint32_t ip) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    float v = 0.0f;
    for (uint32_t j = 0; j < ip; j++) {
        const hi_t h = sm[(size_t)i * ip + j];
        v = madd<FMA>(in[h.index], h.weight, v);
    }
    out[i] = v;
}

// ---------------------------------------------------------------- RotationMap
// src/SM/RotationMap.cpp:190-260 (genHInfo) + :142-174 (apply): source coordinates and the
// it x it Lagrange weights are recomputed per cell with the reference's fp32 expression order;
// the table variant of the reference produces the same numbers, so there is no table here
// (the reference's table costs 128 B/cell of traffic).
struct rot_params {
    float cos_dt, sin_dt;
    float min0, min1, d0, d1, zb0, zb1;
    uint32_t n, it;
    int clamp;
};

template <bool FMA>
__global__ void __launch_bounds__(kThreads)
rotate_kernel(const float* __restrict__ in, float* __restrict__ out, rot_params p) {
    const uint32_t y0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t x0 = blockIdx.y;
    const uint32_t n = p.n, it = p.it;
    if (y0 >= n) return;
    // Ruler::at(i) = min + i*delta (inc/PS/Ruler.hpp:47-49)
    const float ax = IVS_ADD(p.min0, IVS_MUL((float)x0, p.d0));
    const float ay = IVS_ADD(p.min1, IVS_MUL((float)y0, p.d1));
    float x1r = IVS_ADD(IVS_DIV(IVS_SUB(IVS_MUL(p.cos_dt, ax), IVS_MUL(p.sin_dt, ay)), p.d0), p.zb0);
    float y1r = IVS_ADD(IVS_DIV(IVS_ADD(IVS_MUL(p.sin_dt, ax), IVS_MUL(p.cos_dt, ay)), p.d1), p.zb1);
    const float xi = truncf(x1r), yi = truncf(y1r);
    const float xf = IVS_SUB(x1r, xi), yf = IVS_SUB(y1r, yi);
    const uint32_t x1 = (uint32_t)(long long)xi, y1 = (uint32_t)(long long)yi;
    float v = 0.0f;
    if (x1 < n && y1 < n) {
        float icq[4], icp[4];
        ivs_calc_coefficients(xf, it, icq);
        ivs_calc_coefficients(yf, it, icp);
        const uint32_t lo = (it - 1) / 2;
        float ceil = 1.17549435e-38f /* FLT_MIN */, flor = 3.402823466e+38f /* FLT_MAX */;
        for (uint32_t i1 = 0; i1 < it; i1++) {
            const uint32_t i0 = x1 + i1 - lo;
            for (uint32_t j1 = 0; j1 < it; j1++) {
                const uint32_t j0 = y1 + j1 - lo;
                const bool ok = (i0 < n && j0 < n);
                // out-of-grid taps are stored as {index 0, weight 0} (:247-249)
                const float src = in[ok ? (size_t)i0 * n + j0 : 0];
                const float w = ok ? IVS_MUL(icq[i1], icp[j1]) : 0.0f;
                v = madd<FMA>(src, w, v);
                if (p.clamp && i1 >= 1 && i1 <= 2 && j1 >= 1 && j1 <= 2) {
                    ceil = (ceil < src) ? src : ceil;  // std::max(ceil, src)
                    flor = (src < flor) ? src : flor;  // std::min(flor, src)
                }
            }
        }
        if (p.clamp) {
            const float m = (v < ceil) ? v : ceil;  // std::min(ceil, v)
            v = (m < flor) ? flor : m;              // std::max(m, flor)
        }
    } else if (p.clamp) {
        // all 16 taps are {0,0}: value 0 clamped to [in[0], in[0]] (:161-171)
        const float src = in[0];
        float ceil = 1.17549435e-38f, flor = 3.402823466e+38f;
        ceil = (ceil < src) ? src : ceil;
        flor = (src < flor) ? src : flor;
        const float m = (0.0f < ceil) ? 0.0f : ceil;
        v = (m < flor) ? flor : m;
    }
    out[(size_t)x0 * n + y0] = v;
}

// ---------------------------------------------------------------- RotationMap, tiled
// The same arithmetic with the source cells served from shared memory.  A CTA owns an output tile
// of kRotTX rows x kRotTY columns.  The map is affine, so the source coordinates of the tile are
// bounded by those of its four corners: that bounding box (plus the stencil, clipped to the grid,
// columns aligned to 16 bytes) is staged by the TMA engine, one cp.async.bulk per row completing
// on an mbarrier.  For the rotation of a simulation step (angle = 2 pi / steps) the box is a few
// cells larger than the tile, so the grid is read about once (8 B per cell); at any angle it stays
// below (kRotTX + kRotTY + 8)^2 cells, which the launcher sizes the shared memory for.
// Taps outside the grid are {index 0, weight 0} in the reference's table (:247-249): their source
// value is in[0], read once per thread, so every sum (and the clamp) sees the same operands.
constexpr int kRotTX = 32, kRotTY = 128;
constexpr int kRotGY = 4;  // row groups of a CTA: thread (c, g) forms the rows g, g + kRotGY, ... of column c

__device__ __forceinline__ uint32_t rot_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// CLAMP: the clamp of the cubic map (src/SM/RotationMap.cpp:161-171) as a template parameter, so that the
// unclamped map does not pay for the min / max of the four inner source cells
template <bool FMA, bool CLAMP>
__global__ void __launch_bounds__(kRotTY * kRotGY)
rotate_tile_kernel(const float* __restrict__ in, float* __restrict__ out, rot_params p, uint32_t cap_rows,
                   uint32_t pitch) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int box[5];  // first row, rows, first column (multiple of 4), columns (multiple of 4), box not clipped by the grid
    const uint32_t n = p.n, it = p.it;
    const uint32_t tx0 = blockIdx.y * kRotTX, ty0 = blockIdx.x * kRotTY;
    const uint32_t lo = (it - 1) / 2;
    auto source = [&](uint32_t x0, uint32_t y0, float& x1r, float& y1r) {
        // Ruler::at(i) = min + i*delta (inc/PS/Ruler.hpp:47-49); src/SM/RotationMap.cpp:203-214
        const float ax = IVS_ADD(p.min0, IVS_MUL((float)x0, p.d0));
        const float ay = IVS_ADD(p.min1, IVS_MUL((float)y0, p.d1));
        x1r = IVS_ADD(IVS_DIV(IVS_SUB(IVS_MUL(p.cos_dt, ax), IVS_MUL(p.sin_dt, ay)), p.d0), p.zb0);
        y1r = IVS_ADD(IVS_DIV(IVS_ADD(IVS_MUL(p.sin_dt, ax), IVS_MUL(p.cos_dt, ay)), p.d1), p.zb1);
    };
    const uint32_t tlin = threadIdx.y * kRotTY + threadIdx.x;
    if (tlin == 0) {
        const uint32_t xe = min(tx0 + kRotTX, n) - 1, ye = min(ty0 + kRotTY, n) - 1;
        float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
        const uint32_t cx[2] = {tx0, xe}, cy[2] = {ty0, ye};
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                float xr, yr;
                source(cx[a], cy[b], xr, yr);
                xmin = fminf(xmin, xr); xmax = fmaxf(xmax, xr);
                ymin = fminf(ymin, yr); ymax = fmaxf(ymax, yr);
            }
        // one cell of slack for the rounding of the corner coordinates, then the stencil
        const float big = 1.0e9f;
        int r0 = (int)floorf(fmaxf(fminf(xmin, big), -big)) - 1 - (int)lo;
        int r1 = (int)floorf(fmaxf(fminf(xmax, big), -big)) + 1 + (int)(it - 1 - lo);
        int c0 = (int)floorf(fmaxf(fminf(ymin, big), -big)) - 1 - (int)lo;
        int c1 = (int)floorf(fmaxf(fminf(ymax, big), -big)) + 1 + (int)(it - 1 - lo);
        // (one cell of slack on every side of the corners' stencils: unclipped means every cell's are inside)
        const bool unclipped = r0 >= 0 && c0 >= 0 && r1 <= (int)n - 1 && c1 <= (int)n - 1;
        r0 = max(r0, 0); c0 = max(c0, 0) & ~3;
        r1 = min(r1, (int)n - 1); c1 = min(c1 | 3, (int)n - 1);
        int rows = r1 - r0 + 1, cols = c1 - c0 + 1;
        if (rows <= 0 || cols <= 0) rows = cols = 0;
        if ((uint32_t)rows > cap_rows || (uint32_t)cols > pitch || (cols & 3)) rows = -1;  // cannot be staged: gather from global
        box[0] = r0; box[1] = rows; box[2] = c0; box[3] = cols; box[4] = (unclipped && rows > 0) ? 1 : 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rot_smem_u32(&bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (rows > 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rot_smem_u32(&bar)),
                         "r"((uint32_t)rows * (uint32_t)cols * 4u) : "memory");
    }
    __syncthreads();
    const int r0 = box[0], rows = box[1], c0 = box[2], cols = box[3];
    if (rows > 0) {
        for (int r = tlin; r < rows; r += kRotTY * kRotGY) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             rot_smem_u32(tile + (size_t)r * pitch)),
                         "l"(in + (size_t)(r0 + r) * n + c0), "r"((uint32_t)cols * 4u), "r"(rot_smem_u32(&bar))
                         : "memory");
        }
    }
    const float in0 = __ldg(in);
    const uint32_t y0 = ty0 + threadIdx.x;
    if (rows > 0) {
        uint32_t ok;
        do {
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                         : "=r"(ok) : "r"(rot_smem_u32(&bar)), "r"(0u) : "memory");
        } while (!ok);
    }
    if (y0 >= n) return;
    const bool staged = rows >= 0;
    // every tap of every cell of the tile inside the staged box and the grid (the box was clipped to the
    // grid only if it reached beyond it): the common case, no test per tap
    const bool interior = rows > 0 && box[4] != 0;
    for (uint32_t x0 = tx0 + threadIdx.y; x0 < min(tx0 + kRotTX, n); x0 += kRotGY) {
        float x1r, y1r;
        source(x0, y0, x1r, y1r);
        const float xi = truncf(x1r), yi = truncf(y1r);
        const float xf = IVS_SUB(x1r, xi), yf = IVS_SUB(y1r, yi);
        const uint32_t x1 = (uint32_t)(long long)xi, y1 = (uint32_t)(long long)yi;
        float v = 0.0f;
        if (interior && it == 4) {
            // cubic interpolation, all 16 taps valid: straight-line code (same operations, same order)
            float icq[4], icp[4];
            ivs_calc_coefficients(xf, 4, icq);
            ivs_calc_coefficients(yf, 4, icp);
            const float* t0 = tile + (size_t)((int)x1 - 1 - r0) * pitch + ((int)y1 - 1 - c0);
            float ceil = 1.17549435e-38f /* FLT_MIN */, flor = 3.402823466e+38f /* FLT_MAX */;
#pragma unroll
            for (int i1 = 0; i1 < 4; i1++) {
#pragma unroll
                for (int j1 = 0; j1 < 4; j1++) {
                    const float src = t0[(size_t)i1 * pitch + j1];
                    v = madd<FMA>(src, IVS_MUL(icq[i1], icp[j1]), v);
                    if (CLAMP && i1 >= 1 && i1 <= 2 && j1 >= 1 && j1 <= 2) {
                        ceil = (ceil < src) ? src : ceil;  // std::max(ceil, src)
                        flor = (src < flor) ? src : flor;  // std::min(flor, src)
                    }
                }
            }
            if (CLAMP) {
                const float m = (v < ceil) ? v : ceil;  // std::min(ceil, v)
                v = (m < flor) ? flor : m;              // std::max(m, flor)
            }
        } else if (x1 < n && y1 < n) {
            float icq[4], icp[4];
            ivs_calc_coefficients(xf, it, icq);
            ivs_calc_coefficients(yf, it, icp);
            float ceil = 1.17549435e-38f /* FLT_MIN */, flor = 3.402823466e+38f /* FLT_MAX */;
            for (uint32_t i1 = 0; i1 < it; i1++) {
                const uint32_t i0 = x1 + i1 - lo;
                for (uint32_t j1 = 0; j1 < it; j1++) {
                    const uint32_t j0 = y1 + j1 - lo;
                    const bool ok = (i0 < n && j0 < n);
                    float src = in0;
                    if (ok) src = staged ? tile[(size_t)((int)i0 - r0) * pitch + ((int)j0 - c0)] : in[(size_t)i0 * n + j0];
                    const float w = ok ? IVS_MUL(icq[i1], icp[j1]) : 0.0f;
                    v = madd<FMA>(src, w, v);
                    if (CLAMP && i1 >= 1 && i1 <= 2 && j1 >= 1 && j1 <= 2) {
                        ceil = (ceil < src) ? src : ceil;  // std::max(ceil, src)
                        flor = (src < flor) ? src : flor;  // std::min(flor, src)
                    }
                }
            }
            if (CLAMP) {
                const float m = (v < ceil) ? v : ceil;  // std::min(ceil, v)
                v = (m < flor) ? flor : m;              // std::max(m, flor)
            }
        } else if (CLAMP) {
            // all 16 taps are {0,0}: value 0 clamped to [in[0], in[0]] (:161-171)
            float ceil = 1.17549435e-38f, flor = 3.402823466e+38f;
            ceil = (ceil < in0) ? in0 : ceil;
            flor = (in0 < flor) ? in0 : flor;
            const float m = (0.0f < ceil) ? 0.0f : ceil;
            v = (m < flor) ? flor : m;
        }
        out[(size_t)x0 * n + y0] = v;
    }
}

template <typename K, typename... Args>
ivs_status launch(ivs_ctx* ctx, K kernel_fma, K kernel_strict, dim3 grid, dim3 block, size_t smem,
                  Args... args) {
    if (ctx->arith == IVS_ARITH_FMA) kernel_fma<<<grid, block, smem, ctx->stream>>>(args...);
    else kernel_strict<<<grid, block, smem, ctx->stream>>>(args...);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

}  // namespace

// member-batched launchers shared with the simulation object (ivs_step.cu)
ivs_status ivs_kick_y_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n,
                              uint32_t members, const float* offset, uint32_t offset_stride,
                              uint32_t it) {
    const size_t smem = n * sizeof(float);
    if (smem > 48 * 1024) {
        IVS_CUDA(ctx, cudaFuncSetAttribute(kick_y_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(kick_y_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
    }
    return launch(ctx, kick_y_kernel<true>, kick_y_kernel<false>, dim3(n * members), dim3(kThreads),
                  smem, in, out, n, offset, offset_stride, it);
}
ivs_status ivs_kick_x_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n,
                              uint32_t members, const float* offset, uint32_t offset_stride,
                              uint32_t it) {
    const uint32_t bx = 128;
    // rows per CTA: long strips amortise the three extra source rows, short ones fill the machine at small N
    uint32_t strip = 64;
    while (strip > 8 && (uint64_t)ivs_div_up(n, bx) * ivs_div_up(n, strip) * members < 2u * (uint32_t)ctx->sm_count) strip /= 2;
    return launch(ctx, kick_x_kernel<true>, kick_x_kernel<false>,
                  dim3(ivs_div_up(n, bx), ivs_div_up(n, strip), members), dim3(bx), 0, in, out, n,
                  offset, offset_stride, it, strip);
}
ivs_status ivs_fokker_planck_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n,
                                     uint32_t members, const float* w5) {
    const size_t smem = n * sizeof(float);
    if (smem > 48 * 1024) {
        IVS_CUDA(ctx, cudaFuncSetAttribute(fokker_planck_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fokker_planck_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
    }
    return launch(ctx, fokker_planck_kernel<true>, fokker_planck_kernel<false>, dim3(n * members),
                  dim3(kThreads), smem, in, out, n, w5);
}

extern "C" {

ivs_status ivs_kick_y(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n, const void* d_offset,
                      uint32_t it) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_in && d_out && d_offset && n >= 2, "ivs_kick_y: null pointer or n < 2");
    IVS_REQUIRE(ctx, it >= 1 && it <= 4, "ivs_kick_y: interpolation type must be 1..4");
    IVS_REQUIRE(ctx, d_in != d_out, "ivs_kick_y: in-place application is not supported");
    return ivs_kick_y_batched(ctx, (const float*)d_in, (float*)d_out, n, 1, (const float*)d_offset, 0, it);
}

ivs_status ivs_kick_x(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n, const void* d_offset,
                      uint32_t it) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_in && d_out && d_offset && n >= 2, "ivs_kick_x: null pointer or n < 2");
    IVS_REQUIRE(ctx, it >= 1 && it <= 4, "ivs_kick_x: interpolation type must be 1..4");
    IVS_REQUIRE(ctx, d_in != d_out, "ivs_kick_x: in-place application is not supported");
    return ivs_kick_x_batched(ctx, (const float*)d_in, (float*)d_out, n, 1, (const float*)d_offset, 0, it);
}

ivs_status ivs_fokker_planck(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n,
                             const void* d_weights) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_in && d_out && d_weights && n >= 4, "ivs_fokker_planck: null pointer or n < 4");
    IVS_REQUIRE(ctx, d_in != d_out, "ivs_fokker_planck: in-place application is not supported");
    return ivs_fokker_planck_batched(ctx, (const float*)d_in, (float*)d_out, n, 1, (const float*)d_weights);
}

ivs_status ivs_apply_source_map(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t ncells,
                                const void* d_hinfo, uint32_t ip) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_in && d_out && d_hinfo && ncells > 0, "ivs_apply_source_map: null pointer");
    return launch(ctx, source_map_kernel<true>, source_map_kernel<false>,
                  dim3(ivs_div_up(ncells, kThreads)), dim3(kThreads), 0, (const float*)d_in,
                  (float*)d_out, ncells, (const hi_t*)d_hinfo, ip);
}

ivs_status ivs_identity(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n) {
    return ivs_copy(ctx, d_out, d_in, (size_t)n * n * sizeof(float));
}

ivs_status ivs_rotate(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n, float angle,
                      float qmin, float qmax, float pmin, float pmax, uint32_t it, int clamp) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_in && d_out && n >= 2, "ivs_rotate: null pointer or n < 2");
    IVS_REQUIRE(ctx, it >= 1 && it <= 4, "ivs_rotate: interpolation type must be 1..4");
    IVS_REQUIRE(ctx, qmin < qmax && pmin < pmax, "Tried to set up Ruler with min >= max.");
    // src/SM/RotationMap.cpp:83-87
    IVS_REQUIRE(ctx, !(clamp && it != 4),
                "Clamping only supported with cubic interpolation and rotmapsize > 0.");
    rot_params p;
    p.cos_dt = cosf(-angle);  // src/SM/RotationMap.cpp:24-25, std::cos(float)
    p.sin_dt = sinf(-angle);
    p.n = n; p.it = it; p.clamp = clamp;
    p.min0 = qmin; p.min1 = pmin;
    p.d0 = (qmax - qmin) / (float)(n - 1);  // inc/PS/Ruler.hpp:38
    p.d1 = (pmax - pmin) / (float)(n - 1);
    p.zb0 = ((qmin + qmax) / (qmin - qmax) + 1) * (float)(n - 1) / 2;  // inc/PS/Ruler.hpp:40
    p.zb1 = ((pmin + pmax) / (pmin - pmax) + 1) * (float)(n - 1) / 2;
    if (n % 4 != 0)  // rows are staged in 16-byte units
        return launch(ctx, rotate_kernel<true>, rotate_kernel<false>, dim3(ivs_div_up(n, 128), n),
                      dim3(128), 0, (const float*)d_in, (float*)d_out, p);
    // shared memory for the bounding box of a rotated tile (+ stencil, slack and column alignment)
    const float ac = fabsf(p.cos_dt), as = fabsf(p.sin_dt);
    // the two axes may have different cell sizes: source rows per output row / column
    const float rx = p.d1 / p.d0, ry = p.d0 / p.d1;
    uint32_t cap_rows = (uint32_t)ceilf(kRotTX * ac + kRotTY * as * rx) + 8;
    uint32_t pitch = ((uint32_t)ceilf(kRotTY * ac + kRotTX * as * ry) + 8 + 6 + 3) & ~3u;
    cap_rows = std::min<uint32_t>(cap_rows, n);
    pitch = std::min<uint32_t>(pitch, n);
    size_t smem = (size_t)cap_rows * pitch * sizeof(float);
    if (smem > (size_t)kIvsMaxOptinSmem - 1024) {  // very anisotropic axes: keep what fits, the kernel gathers the rest
        cap_rows = (uint32_t)(((size_t)kIvsMaxOptinSmem - 1024) / ((size_t)pitch * sizeof(float)));
        smem = (size_t)cap_rows * pitch * sizeof(float);
    }
    // once per context: the attribute belongs to the kernel (two API calls of 15 us each otherwise)
    const bool opted_in = (ctx->optin & 1u) != 0;
    if (smem > 40 * 1024 && !opted_in) {
        IVS_CUDA(ctx, cudaFuncSetAttribute(rotate_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(rotate_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(rotate_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(rotate_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        ctx->optin |= 1u;
    }
    const dim3 grid(ivs_div_up(n, kRotTY), ivs_div_up(n, kRotTX)), block(kRotTY, kRotGY);
    if (clamp)
        return launch(ctx, rotate_tile_kernel<true, true>, rotate_tile_kernel<false, true>, grid, block, smem,
                      (const float*)d_in, (float*)d_out, p, cap_rows, pitch);
    return launch(ctx, rotate_tile_kernel<true, false>, rotate_tile_kernel<false, false>, grid, block, smem,
                  (const float*)d_in, (float*)d_out, p, cap_rows, pitch);
}

}  // extern "C"
