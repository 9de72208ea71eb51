// ivs_train.cu -- wake potential of a TRAIN of bunches (PhaseSpace::nb > 1).
// The reference pads the profiles of all bunches into one array of nmax bins, bunch b at
// bucket[b] * spacing_bins, transforms, multiplies by the impedance, transforms back and reads the wake
// potential of bunch b at the same offset (src/PS/ElectricField.cpp:186-263, src/main.cpp:248-321): a
// circular convolution of nmax = spaced_bins points (millions at a storage ring's bunch spacing).  The
// impedance is static, so the same numbers are
//     W_b[x] = scaling * sum_b' sum_x' rho_b'[x'] * w[(bucket_b - bucket_b') * spacing_bins + x - x']
// with w the unit wake function (one inverse transform on the host, ivs_train_wake_windows in ivs_host.h):
// nb^2 windows of 2 n - 1 lags and a Toeplitz product per pair of bunches, accumulated in double like the
// fused step kernels' (ivs_stream8.cu); no transform per step, no padded array on the device.
#include "ivs_common.cuh"

namespace {

constexpr int kTrainTX = 64;      // rows x per block
constexpr int kTrainTY = 8;       // threads per row: thread (tx, ty) sums the source bins x' = ty (mod kTrainTY)
constexpr int kTrainChunk = 512;  // source bins x' per shared-memory pass

__global__ void __launch_bounds__(kTrainTX * kTrainTY)
train_wake_kernel(const float* __restrict__ profiles, uint32_t n, uint32_t nb, const double* __restrict__ windows,
                  const float* __restrict__ scaling, float scal, float* __restrict__ wake) {
    __shared__ double rho[kTrainChunk];
    __shared__ double win[kTrainChunk + kTrainTX];
    __shared__ double red[kTrainTY][kTrainTX];
    const uint32_t b = blockIdx.y, x0 = blockIdx.x * kTrainTX, tx = threadIdx.x, ty = threadIdx.y;
    const uint32_t tlin = ty * kTrainTX + tx;
    // fixed order: source bunch, then source bin, ascending, as two interleaved partial sums per thread (the
    // multiply-adds of one are dependent); the kTrainTY threads of a row are added in order at the end
    double a0 = 0.0, a1 = 0.0;
    for (uint32_t c = 0; c < nb; c++) {
        const double* w = windows + ((size_t)b * nb + c) * 2 * n;
        const float* rc = profiles + (size_t)c * n;
        for (uint32_t s0 = 0; s0 < n; s0 += kTrainChunk) {
            // lag index of (x0, s0 + kTrainChunk - 1): the lowest one this pass reads
            const int64_t base = (int64_t)x0 - (int64_t)s0 - (kTrainChunk - 1) + (int64_t)(n - 1);
            for (uint32_t k = tlin; k < kTrainChunk; k += kTrainTX * kTrainTY) rho[k] = s0 + k < n ? (double)rc[s0 + k] : 0.0;
            for (uint32_t j = tlin; j < kTrainChunk + kTrainTX - 1; j += kTrainTX * kTrainTY) {
                const int64_t i = base + j;
                win[j] = (i >= 0 && i < (int64_t)(2 * n - 1)) ? w[i] : 0.0;
            }
            __syncthreads();
            const double* wt = win + tx + (kTrainChunk - 1);
#pragma unroll 4
            for (uint32_t k = ty; k < kTrainChunk; k += 2 * kTrainTY) {
                a0 = fma(rho[k], wt[-(int)k], a0);
                a1 = fma(rho[k + kTrainTY], wt[-(int)(k + kTrainTY)], a1);
            }
            __syncthreads();
        }
    }
    red[ty][tx] = a0 + a1;
    __syncthreads();
    const uint32_t x = x0 + tx;
    if (ty == 0 && x < n) {
        double acc = 0.0;
#pragma unroll
        for (int t = 0; t < kTrainTY; t++) acc += red[t][tx];
        wake[(size_t)b * n + x] = __fmul_rn(scaling ? scaling[b] : scal, (float)acc);  // src/PS/ElectricField.cpp:236-241
    }
}

}  // namespace

ivs_status ivs_train_wake_launch(ivs_ctx* ctx, const float* profiles, uint32_t n, uint32_t nb, const double* windows,
                                 const float* scaling, float scal, float* wake) {
    train_wake_kernel<<<dim3(ivs_div_up(n, kTrainTX), nb), dim3(kTrainTX, kTrainTY), 0, ctx->stream>>>(profiles, n, nb, windows,
                                                                                      scaling, scal, wake);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

extern "C" {

ivs_status ivs_train_wake_potential(ivs_ctx* ctx, const void* d_profiles, uint32_t n, uint32_t nb,
                                    const void* d_windows, float wakescaling, void* d_wake) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_profiles && d_windows && d_wake, "ivs_train_wake_potential: null pointer");
    IVS_REQUIRE(ctx, n >= 2 && nb >= 1 && nb <= 65535, "ivs_train_wake_potential: n >= 2, 1 <= nb <= 65535");
    return ivs_train_wake_launch(ctx, (const float*)d_profiles, n, nb, (const double*)d_windows, nullptr, wakescaling,
                                 (float*)d_wake);
}

}  // extern "C"
