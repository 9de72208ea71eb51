// ivs_fused.cu -- the whole simulation step as ONE kernel (sm_100a).
//
//   wake potential -> wake kick (y) -> RF kick (y) -> drift (x) -> Fokker-Planck (y) -> X projection
//   (src/main.cpp:948-1053; maps src/SM/KickMap.cpp:221-266, src/SM/FokkerPlanckMap.cpp:197-211,
//    projection src/PS/PhaseSpace.cpp:313-321, wake src/PS/ElectricField.cpp:186-253)
//
// Data flow per CTA (one tile = TX consecutive rows x of one ensemble member, all N columns y):
//   0. the tile plus the rows the drift will need from below/above (halo HL/HR, derived from
//      the drift offsets) is staged in shared memory by the TMA engine: one cp.async.bulk per
//      row (8 KB at N = 2048) completing on an mbarrier; every row keeps PAD zero columns on
//      both sides and out-of-grid rows are zero, which implements "taps outside the grid are
//      dropped" (src/SM/KickMap.cpp:231-237,255-261) without a branch per tap.
//   1. while the copies fly, every warp forms the wake potential of its rows as the Toeplitz
//      product  W[x] = scaling * sum_x' rho[x'] w[x-x']  of last step's bunch profile with the
//      precomputed wake function (same operator as pad -> r2c -> xZ -> c2r), in
//      double, and turns it into interpolation taps with the arithmetic of KickMap::updateSM.
//   2. wake kick and RF kick: row-local 4-tap passes, one warp per row, in place.
//   3. drift: column-local 4-tap pass down the rows, one thread per column, in place.
//   4. Fokker-Planck stencil along y, Simpson-weighted row sums (the X projection for the next
//      step) and the coalesced store of the TX finished rows.
// The grid is read once and written once: 8 B per cell-update; halo rows are re-read by the
// neighbouring tile out of L2.  Intermediate grids t2/t3 of the reference never exist.
// Every tap accumulation is madd<FMA>() in the reference's order, so the result is bit-identical
// to the unfused kernel sequence (ivs_maps.cu) -- tested.
#include <algorithm>
#include <vector>

#include "ivs_internal.h"
#include "ivs_sim_struct.h"

namespace {

constexpr int kPad = 16;   // zero columns on each side of a staged row (floats)
constexpr int kChunk = 8;  // outputs per lane and chunk in the row passes

struct fused_args {
    const float* in;
    float* out;
    const float* rho_in;
    float* rho_out;
    float* wake_out;
    const double* wakefn;
    const float* wakescaling;
    const int* rf_first;
    const float4* rf_w;
    const int* dr_first;
    const float4* dr_w;
    const int* fp_base;
    const float4* fp_w;
    const float* fp_w5;
    const float* ws;
    int* status;
    uint32_t n, it, hl, hr, pitch;
    int has_wake;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// TMA engine, 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}

template <bool FMA>
__device__ __forceinline__ float tap4(float a0, float a1, float a2, float a3, const float4 w) {
    // value = 0; value += in*w, four times in the reference's order.  The first add (0 + a*w)
    // only decides the sign of a zero result; it is folded away.
    float v = __fmul_rn(a0, w.x);
    v = madd<FMA>(a1, w.y, v);
    v = madd<FMA>(a2, w.z, v);
    v = madd<FMA>(a3, w.w, v);
    return v;
}

// one 4-tap pass along a staged row, in place, by one warp.  `row` points at column 0; columns
// [-kPad, n+kPad) are addressable and the pads hold zeros.
template <bool FMA>
__device__ __forceinline__ void row_pass(float* row, uint32_t n, int first, const float4 w, uint32_t lane) {
    const float* src = row + first;
    float prev[kChunk];
    int prev_base = -1;
    for (uint32_t c = 0; c < n; c += 32 * kChunk) {
        float cur[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; k++) {
            const uint32_t y = c + lane + 32 * k;
            cur[k] = 0.0f;
            if (y < n) cur[k] = tap4<FMA>(src[y], src[y + 1], src[y + 2], src[y + 3], w);
        }
        __syncwarp();  // every lane has read what it needs from the previous chunk's range
        if (prev_base >= 0) {
#pragma unroll
            for (int k = 0; k < kChunk; k++) row[prev_base + lane + 32 * k] = prev[k];
        }
#pragma unroll
        for (int k = 0; k < kChunk; k++) prev[k] = cur[k];
        prev_base = (int)c;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kChunk; k++) {
        const uint32_t y = prev_base + lane + 32 * k;
        if (y < n) row[y] = prev[k];
    }
    __syncwarp();
}

template <int TX, bool FMA>
__global__ void __launch_bounds__(1024, 1) fused_step_generic_kernel(const fused_args g) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t n = g.n, pitch = g.pitch, hl = g.hl;
    const uint32_t R = TX + g.hl + g.hr;
    const uint32_t T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const uint32_t member = blockIdx.y;
    const uint32_t x0 = blockIdx.x * TX;
    const int xs = (int)x0 - (int)hl;  // grid row of staged row 0
    const uint32_t rows_here = min((uint32_t)TX, n - x0);

    float* tile = smem;                               // R x pitch
    float* rho = tile + (size_t)R * pitch;            // n
    int* wfirst = (int*)(rho + n);                    // R
    float4* ww = (float4*)(wfirst + ((R + 3) & ~3u)); // R
    float* red = (float*)(ww + R);                    // TX x W

    const float* gin = g.in + (size_t)member * n * n;

    // ---- 0. stage the rows through the TMA engine
    if (tid == 0) {
        mbar_init(&bar, 1);
        uint32_t valid = 0;
        for (uint32_t r = 0; r < R; r++) valid += ((uint32_t)(xs + (int)r) < n);
        mbar_expect_tx(&bar, valid * n * (uint32_t)sizeof(float));
    }
    __syncthreads();
    if (warp == 0) {
        for (uint32_t r = lane; r < R; r += 32) {
            const int x = xs + (int)r;
            if ((uint32_t)x < n) bulk_g2s(tile + (size_t)r * pitch + kPad, gin + (size_t)x * n, n * sizeof(float), &bar);
        }
    }
    // zero pads (and whole rows outside the grid) with ordinary stores: disjoint from the bulk copies
    for (uint32_t r = warp; r < R; r += W) {
        float* row = tile + (size_t)r * pitch;
        const int x = xs + (int)r;
        if ((uint32_t)x < n) {
            if (lane < kPad) row[lane] = 0.0f; else row[n + kPad + (lane - kPad)] = 0.0f;
        } else {
            for (uint32_t i = lane; i < pitch; i += 32) row[i] = 0.0f;
        }
    }

    // ---- 1. wake potential of the staged rows (Toeplitz product with last step's profile)
    if (g.has_wake) {
        const float* rin = g.rho_in + (size_t)member * n;
        for (uint32_t i = tid; i < n; i += T) rho[i] = rin[i];
        __syncthreads();
        const float scal = g.wakescaling[member];
        for (uint32_t r = warp; r < R; r += W) {
            const int x = xs + (int)r;
            if ((uint32_t)x >= n) continue;
            const double* wf = g.wakefn + (x + (int)n - 1);  // wf[-x'] = w[x - x']
            double acc = 0.0;
            for (uint32_t xp = lane; xp < n; xp += 32) acc = __fma_rn((double)rho[xp], __ldg(wf - (int)xp), acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                // wakepotential[x] = _wakescaling * padded[x]  (src/PS/ElectricField.cpp:236-241)
                const float wk = __fmul_rn(scal, (float)acc);
                ivs_taps t = ivs_kick_taps(wk, n, g.it);
                if (t.first < -kPad || t.first > kPad - 4) {  // beyond the staged pad: flag, neutralise
                    atomicExch(g.status, 1);
                    t.first = -1;
                    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
                }
                wfirst[r] = t.first;
                ww[r] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
                if (r >= hl && r < hl + rows_here) g.wake_out[(size_t)member * n + x] = wk;
            }
        }
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    // ---- 2. wake kick, then RF kick: row-local, one warp per staged row, in place
    for (uint32_t r = warp; r < R; r += W) {
        const int x = xs + (int)r;
        if ((uint32_t)x >= n) continue;
        float* row = tile + (size_t)r * pitch + kPad;
        if (g.has_wake) row_pass<FMA>(row, n, wfirst[r], ww[r], lane);
        row_pass<FMA>(row, n, __ldg(&g.rf_first[x]), __ldg(&g.rf_w[x]), lane);
    }
    __syncthreads();

    // ---- 3. drift: column-local, rows hl .. hl+TX-1 are overwritten with the drifted values
    for (uint32_t y = tid; y < n; y += T) {
        const int first = __ldg(&g.dr_first[y]);
        const float4 w = __ldg(&g.dr_w[y]);
        const float* col = tile + (size_t)((int)hl + first) * pitch + kPad + y;
        float v[TX + 3];
#pragma unroll
        for (int k = 0; k < TX + 3; k++) v[k] = col[(size_t)k * pitch];
        float* dst = tile + (size_t)hl * pitch + kPad + y;
#pragma unroll
        for (int i = 0; i < TX; i++) dst[(size_t)i * pitch] = tap4<FMA>(v[i], v[i + 1], v[i + 2], v[i + 3], w);
    }
    __syncthreads();

    // ---- 4. Fokker-Planck along y, projection partial sums, store
    float acc[TX];
#pragma unroll
    for (int i = 0; i < TX; i++) acc[i] = 0.0f;
    float* gout = g.out + ((size_t)member * n + x0) * n;
    for (uint32_t y = tid; y < n; y += T) {
        const int b = __ldg(&g.fp_base[y]);
        const float4 w = __ldg(&g.fp_w[y]);
        const float wsy = __ldg(&g.ws[y]);
        const float* src = tile + (size_t)hl * pitch + kPad + y + b;
#pragma unroll
        for (int i = 0; i < TX; i++) {
            const float* p = src + (size_t)i * pitch;
            const float o = tap4<FMA>(p[0], p[1], p[2], p[3], w);
            if ((uint32_t)i < rows_here) gout[(size_t)i * n + y] = o;
            acc[i] = __fadd_rn(acc[i], __fmul_rn(o, wsy));  // src/PS/PhaseSpace.cpp:316-319 (products rounded)
        }
    }
#pragma unroll
    for (int i = 0; i < TX; i++) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[i * W + warp] = v;
    }
    __syncthreads();
    if (tid < rows_here) {
        float v = 0.0f;
        for (uint32_t k = 0; k < W; k++) v = __fadd_rn(v, red[tid * W + k]);
        g.rho_out[(size_t)member * n + x0 + tid] = v;
    }
}


// ------------------------------------------------------------------------------------------
// Vectorised variant for grid sizes that are multiples of 256 (256 / 512 / 1024 / 2048):
// the same data flow, but every lane owns 4 consecutive columns (one float4), so a 4-tap pass
// costs 2-3 LDS.128 + 1 STS.128 per 4 outputs instead of 16 LDS + 4 STS, the Fokker-Planck pass
// stores with STG.128, and the projection partial sums are combined with a transposing butterfly
// (one shuffle per value instead of five).  Arithmetic per cell is unchanged.

template <bool FMA, int RR>
__device__ __forceinline__ void row_pass_v4_body(float* row, uint32_t nq, int a, const float4 w, uint32_t lane) {
    const float4* src = reinterpret_cast<const float4*>(row + a);
    float4* dst = reinterpret_cast<float4*>(row);
    float4 prev0 = make_float4(0, 0, 0, 0), prev1 = prev0;
    int prev_q = -1;
    for (uint32_t q0 = 0; q0 < nq; q0 += 64) {
        float4 cur[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t q = q0 + lane + 32 * u;
            float in[12];
            const float4 A = src[q], B = src[q + 1];
            in[0] = A.x; in[1] = A.y; in[2] = A.z; in[3] = A.w;
            in[4] = B.x; in[5] = B.y; in[6] = B.z; in[7] = B.w;
            if (RR >= 2) {
                const float4 C = src[q + 2];
                in[8] = C.x; in[9] = C.y; in[10] = C.z; in[11] = C.w;
            }
            cur[u].x = tap4<FMA>(in[RR + 0], in[RR + 1], in[RR + 2], in[RR + 3], w);
            cur[u].y = tap4<FMA>(in[RR + 1], in[RR + 2], in[RR + 3], in[RR + 4], w);
            cur[u].z = tap4<FMA>(in[RR + 2], in[RR + 3], in[RR + 4], in[RR + 5], w);
            cur[u].w = tap4<FMA>(in[RR + 3], in[RR + 4], in[RR + 5], in[RR + 6], w);
        }
        __syncwarp();  // all lanes have read what they need from the previous chunk's range
        if (prev_q >= 0) {
            dst[prev_q + lane] = prev0;
            dst[prev_q + lane + 32] = prev1;
        }
        prev0 = cur[0]; prev1 = cur[1];
        prev_q = (int)q0;
    }
    __syncwarp();
    dst[prev_q + lane] = prev0;
    dst[prev_q + lane + 32] = prev1;
    __syncwarp();
}

template <bool FMA>
__device__ __forceinline__ void row_pass_v4(float* row, uint32_t nq, int first, const float4 w, uint32_t lane) {
    const int a = first & ~3;  // floor to a multiple of 4 (two's complement: also for negatives)
    switch (first - a) {       // warp-uniform
    case 0: row_pass_v4_body<FMA, 0>(row, nq, a, w, lane); break;
    case 1: row_pass_v4_body<FMA, 1>(row, nq, a, w, lane); break;
    case 2: row_pass_v4_body<FMA, 2>(row, nq, a, w, lane); break;
    default: row_pass_v4_body<FMA, 3>(row, nq, a, w, lane); break;
    }
}

// sums NV per-lane partials over the 32 lanes of a warp; afterwards every lane holds the total
// of row index row_of_lane<NV>(lane) in v[0].  NV shuffles instead of 5*NV, fixed order.
template <int NV>
__device__ __forceinline__ void warp_rows_reduce(float (&v)[NV], uint32_t lane) {
    int off = 16;
#pragma unroll
    for (int m = NV; m > 1; m >>= 1) {
        const int half = m >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
        off >>= 1;
    }
    for (; off > 0; off >>= 1) v[0] = __fadd_rn(v[0], __shfl_xor_sync(0xffffffffu, v[0], off));
}
template <int NV>
__device__ __forceinline__ uint32_t row_of_lane(uint32_t lane) {
    uint32_t row = 0;
    int off = 16;
#pragma unroll
    for (int m = NV; m > 1; m >>= 1) {
        row = row * 2 + ((lane & off) ? 1u : 0u);
        off >>= 1;
    }
    return row;
}

template <int TX, int RG, bool FMA>
__global__ void __launch_bounds__(1024, 1) fused_step_kernel(const fused_args g) {
    constexpr int NV = TX / RG;  // rows per thread in the last stage
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t n = g.n, nq = g.n >> 2, pitch = g.pitch, hl = g.hl;
    const uint32_t R = TX + g.hl + g.hr;
    const uint32_t T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    const uint32_t member = blockIdx.y;
    const uint32_t x0 = blockIdx.x * TX;
    const int xs = (int)x0 - (int)hl;
    const uint32_t rows_here = min((uint32_t)TX, n - x0);

    float* tile = smem;                                      // R x pitch floats
    double* rho = reinterpret_cast<double*>(tile + (size_t)R * pitch);  // n doubles
    int* wfirst = reinterpret_cast<int*>(rho + n);           // R
    float4* ww = reinterpret_cast<float4*>(wfirst + ((R + 3) & ~3u));  // R
    float* red = reinterpret_cast<float*>(ww + R);           // TX x W/RG

    const float* gin = g.in + (size_t)member * n * n;

    // ---- 0. stage the rows through the TMA engine (one bulk copy per row)
    if (tid == 0) {
        mbar_init(&bar, 1);
        uint32_t valid = 0;
        for (uint32_t r = 0; r < R; r++) valid += ((uint32_t)(xs + (int)r) < n);
        mbar_expect_tx(&bar, valid * n * (uint32_t)sizeof(float));
    }
    __syncthreads();
    if (warp == 0) {
        for (uint32_t r = lane; r < R; r += 32) {
            const int x = xs + (int)r;
            if ((uint32_t)x < n) bulk_g2s(tile + (size_t)r * pitch + kPad, gin + (size_t)x * n, n * sizeof(float), &bar);
        }
    }
    for (uint32_t r = warp; r < R; r += W) {
        float* row = tile + (size_t)r * pitch;
        const int x = xs + (int)r;
        if ((uint32_t)x < n) {
            if (lane < kPad) row[lane] = 0.0f; else row[n + kPad + (lane - kPad)] = 0.0f;
        } else {
            for (uint32_t i = lane; i < pitch; i += 32) row[i] = 0.0f;
        }
    }

    // ---- 1. wake potential of the staged rows: Toeplitz product in double
    if (g.has_wake) {
        const float* rin = g.rho_in + (size_t)member * n;
        for (uint32_t i = tid; i < n; i += T) rho[i] = (double)rin[i];
        __syncthreads();
        const float scal = g.wakescaling[member];
        for (uint32_t r = warp; r < R; r += W) {
            const int x = xs + (int)r;
            if ((uint32_t)x >= n) continue;
            const double* wf = g.wakefn + (x + (int)n - 1) - (int)lane;  // wf[-32 j] = w[x - lane - 32 j]
            const double* rl = rho + lane;
            double acc0 = 0.0, acc1 = 0.0;
            for (uint32_t j = 0; j < n; j += 64) {
                acc0 = __fma_rn(rl[j], __ldg(wf - (int)j), acc0);
                acc1 = __fma_rn(rl[j + 32], __ldg(wf - (int)j - 32), acc1);
            }
            double acc = acc0 + acc1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const float wk = __fmul_rn(scal, (float)acc);  // src/PS/ElectricField.cpp:236-241
                ivs_taps t = ivs_kick_taps(wk, n, g.it);
                if (t.first < -kPad || t.first > kPad - 4) {
                    atomicExch(g.status, 1);
                    t.first = -1;
                    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
                }
                wfirst[r] = t.first;
                ww[r] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
                if (r >= hl && r < hl + rows_here) g.wake_out[(size_t)member * n + x] = wk;
            }
        }
    }
    mbar_wait(&bar, 0);
    __syncthreads();

    // ---- 2. wake kick, then RF kick: one warp per staged row, in place, float4 per lane
    for (uint32_t r = warp; r < R; r += W) {
        const int x = xs + (int)r;
        if ((uint32_t)x >= n) continue;
        float* row = tile + (size_t)r * pitch + kPad;
        if (g.has_wake) row_pass_v4<FMA>(row, nq, wfirst[r], ww[r], lane);
        row_pass_v4<FMA>(row, nq, __ldg(&g.rf_first[x]), __ldg(&g.rf_w[x]), lane);
    }
    __syncthreads();

    // ---- 3. drift: column-local sliding window down the rows, in place
    for (uint32_t y = tid; y < n; y += T) {
        const int first = __ldg(&g.dr_first[y]);
        const float4 w = __ldg(&g.dr_w[y]);
        const float* col = tile + (size_t)((int)hl + first) * pitch + kPad + y;
        float v[TX + 3];
#pragma unroll
        for (int k = 0; k < TX + 3; k++) v[k] = col[(size_t)k * pitch];
        float* dst = tile + (size_t)hl * pitch + kPad + y;
#pragma unroll
        for (int i = 0; i < TX; i++) dst[(size_t)i * pitch] = tap4<FMA>(v[i], v[i + 1], v[i + 2], v[i + 3], w);
    }
    __syncthreads();

    // ---- 4. Fokker-Planck (5 dense weights on y-2..y+2), projection partials, STG.128
    {
        const uint32_t q = tid % nq, grp = tid / nq;  // T == nq * RG
        const uint32_t y = q << 2;
        float w5[4][5];
        {
            const float4* wp = reinterpret_cast<const float4*>(g.fp_w5 + (size_t)y * 5);
            float t[20];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float4 v = __ldg(wp + k);
                t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int k = 0; k < 5; k++) w5[j][k] = t[j * 5 + k];
        }
        const float4 ws4 = __ldg(reinterpret_cast<const float4*>(g.ws) + q);
        float acc[NV];
        float4* gout = reinterpret_cast<float4*>(g.out + ((size_t)member * n + x0 + grp * NV) * n) + q;
        const float4* src = reinterpret_cast<const float4*>(tile + (size_t)(hl + grp * NV) * pitch + kPad) + q;
        const uint32_t pitch4 = pitch >> 2;
#pragma unroll
        for (int ii = 0; ii < NV; ii++) {
            const float4* rp = src + ii * (int)pitch4;
            const float4 A = rp[-1], B = rp[0], C = rp[1];
            const float in[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float v = __fmul_rn(in[2 + j], w5[j][0]);
                v = madd<FMA>(in[3 + j], w5[j][1], v);
                v = madd<FMA>(in[4 + j], w5[j][2], v);
                v = madd<FMA>(in[5 + j], w5[j][3], v);
                v = madd<FMA>(in[6 + j], w5[j][4], v);
                o[j] = v;
            }
            if (grp * NV + ii < rows_here) gout[(size_t)ii * nq] = make_float4(o[0], o[1], o[2], o[3]);
            float p = __fmul_rn(o[0], ws4.x);                      // src/PS/PhaseSpace.cpp:316-319
            p = __fadd_rn(p, __fmul_rn(o[1], ws4.y));
            p = __fadd_rn(p, __fmul_rn(o[2], ws4.z));
            p = __fadd_rn(p, __fmul_rn(o[3], ws4.w));
            acc[ii] = p;
        }
        warp_rows_reduce<NV>(acc, lane);
        constexpr int kGroupLanes = 32 / NV;  // lanes that end up with the same row total
        if ((lane & (kGroupLanes - 1)) == 0) {
            const uint32_t wq = (nq >> 5);    // warps per row group
            red[(grp * NV + row_of_lane<NV>(lane)) * wq + (warp % wq)] = acc[0];
        }
    }
    __syncthreads();
    if (tid < rows_here) {
        const uint32_t wq = (nq >> 5);
        float v = 0.0f;
        for (uint32_t k = 0; k < wq; k++) v = __fadd_rn(v, red[tid * wq + k]);
        g.rho_out[(size_t)member * n + x0 + tid] = v;
    }
}


// ------------------------------------------------------------------------------------------
// Streaming variant (grid sizes that are multiples of 128): a CTA owns a strip of consecutive
// rows of one member and pushes them through a warp-specialised pipeline over a ring of row
// slots in shared memory -- no block-wide phases, every row is kicked once, copies overlap math:
//
//   producer warp   : one TMA bulk copy per row into the next free slot          (full[slot])
//   kick warps      : wake potential of the row (Toeplitz, double), wake kick + RF kick in
//                     place, one warp per row, float4 per lane                   (kicked[slot])
//   drift/FP warps  : n/4 threads, each owns 4 columns for the whole strip (taps in registers):
//                     drift from the ring into a scratch row, Fokker-Planck from the scratch row
//                     straight to global memory (STG.128) + projection partials  (free[slot])
//
// Synchronisation is by mbarriers only (plus one named barrier per output row inside the
// drift/FP group).  Arithmetic per cell is identical to the tile kernels.

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct stream_args {
    fused_args f;
    uint32_t strip;     // rows per strip
    uint32_t kslots;    // ring slots (power of two)
    uint32_t ndf;       // drift/FP warps (= n/128)
    uint32_t nkick;     // kick warps
};

template <bool FMA, bool REBALANCE>
__global__ void __launch_bounds__(1024, 1) stream_step_kernel(const stream_args sa) {
    const fused_args& g = sa.f;
    extern __shared__ __align__(128) float smem[];
    const uint32_t n = g.n, nq = g.n >> 2, pitch = g.pitch, hl = g.hl;
    const uint32_t H = g.hl + g.hr;
    const uint32_t K = sa.kslots, kmask = K - 1;
    const uint32_t klog = 31 - __clz(K);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t member = blockIdx.y;
    const uint32_t xa = blockIdx.x * sa.strip;
    const uint32_t S = min(sa.strip, n - xa);  // output rows of this strip
    const uint32_t RIN = S + H;                // input rows streamed through the ring
    const int xs = (int)xa - (int)hl;          // grid row of input row 0

    float* ring = smem;                                                   // K x pitch
    float* scratch = ring + (size_t)K * pitch;                            // 2 x pitch
    double* rho = reinterpret_cast<double*>(scratch + 2 * (size_t)pitch); // n
    uint64_t* full = reinterpret_cast<uint64_t*>(rho + n);                // K
    uint64_t* kicked = full + K;                                          // K
    uint64_t* freeb = kicked + K;                                         // K
    float* red = reinterpret_cast<float*>(freeb + K);                     // 2 x 32

    const float* gin = g.in + (size_t)member * n * n;

    // ---- set-up by everybody: barriers, zero pads, last step's profile in double
    if (tid == 0) {
        for (uint32_t s = 0; s < K; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&kicked[s], 1);
            mbar_init(&freeb[s], 1);
        }
    }
    for (uint32_t r = warp; r < K + 2; r += (blockDim.x >> 5)) {
        float* row = ring + (size_t)r * pitch;  // the two scratch rows follow the ring
        if (lane < kPad) row[lane] = 0.0f; else row[n + kPad + (lane - kPad)] = 0.0f;
    }
    if (g.has_wake) {
        const float* rin = g.rho_in + (size_t)member * n;
        for (uint32_t i = tid; i < n; i += blockDim.x) rho[i] = (double)rin[i];
    }
    __syncthreads();

    if (warp < sa.ndf) {
        // =============================== drift + Fokker-Planck group ===========================
        if (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
        const uint32_t ndf_threads = sa.ndf * 32;
        const uint32_t q = tid;            // this thread's quad of columns, y = 4q .. 4q+3
        const uint32_t y = q << 2;
        // drift taps of the 4 columns
        const int4 df = __ldg(reinterpret_cast<const int4*>(g.dr_first) + q);
        const float4 dw0 = __ldg(&g.dr_w[y]), dw1 = __ldg(&g.dr_w[y + 1]);
        const float4 dw2 = __ldg(&g.dr_w[y + 2]), dw3 = __ldg(&g.dr_w[y + 3]);
        const bool uniform = (df.x == df.y) && (df.y == df.z) && (df.z == df.w);
        // Fokker-Planck weights (5 per column) and Simpson weights
        float w5[4][5];
        {
            const float4* wp = reinterpret_cast<const float4*>(g.fp_w5 + (size_t)y * 5);
            float t[20];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float4 v = __ldg(wp + k);
                t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int k = 0; k < 5; k++) w5[j][k] = t[j * 5 + k];
        }
        const float4 ws4 = __ldg(reinterpret_cast<const float4*>(g.ws) + q);
        float4* gout = reinterpret_cast<float4*>(g.out + ((size_t)member * n + xa) * n) + q;
        const uint32_t pitch4 = pitch >> 2;
        const float4* ring4 = reinterpret_cast<const float4*>(ring + kPad) + q;

        // the first output needs input rows 0..H kicked; afterwards one new row per output
        for (uint32_t r = 0; r < H; r++) mbar_wait(&kicked[r & kmask], (r >> klog) & 1);
        for (uint32_t i = 0; i < S; i++) {
            {
                const uint32_t r = i + H;
                mbar_wait(&kicked[r & kmask], (r >> klog) & 1);
            }
            // ---- drift (src/SM/KickMap.cpp:221-242): column-local 4-tap along x
            float4 d;
            if (uniform) {
                const uint32_t rb = i + hl + (uint32_t)df.x;  // >= i by construction of hl
                const float4 A0 = ring4[(size_t)((rb) & kmask) * pitch4];
                const float4 A1 = ring4[(size_t)((rb + 1) & kmask) * pitch4];
                const float4 A2 = ring4[(size_t)((rb + 2) & kmask) * pitch4];
                const float4 A3 = ring4[(size_t)((rb + 3) & kmask) * pitch4];
                d.x = tap4<FMA>(A0.x, A1.x, A2.x, A3.x, dw0);
                d.y = tap4<FMA>(A0.y, A1.y, A2.y, A3.y, dw1);
                d.z = tap4<FMA>(A0.z, A1.z, A2.z, A3.z, dw2);
                d.w = tap4<FMA>(A0.w, A1.w, A2.w, A3.w, dw3);
            } else {
                const float* ringc = ring + kPad + y;
                const int fs[4] = {df.x, df.y, df.z, df.w};
                const float4 dws[4] = {dw0, dw1, dw2, dw3};
                float o[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t rb = i + hl + (uint32_t)fs[c];
                    const float a0 = ringc[(size_t)((rb) & kmask) * pitch + c];
                    const float a1 = ringc[(size_t)((rb + 1) & kmask) * pitch + c];
                    const float a2 = ringc[(size_t)((rb + 2) & kmask) * pitch + c];
                    const float a3 = ringc[(size_t)((rb + 3) & kmask) * pitch + c];
                    o[c] = tap4<FMA>(a0, a1, a2, a3, dws[c]);
                }
                d = make_float4(o[0], o[1], o[2], o[3]);
            }
            float* srow = scratch + (size_t)(i & 1) * pitch + kPad;
            reinterpret_cast<float4*>(srow)[q] = d;
            named_bar_sync(1, ndf_threads);
            // input row i is not needed any more: hand its slot back to the producer
            if (tid == 0) mbar_arrive(&freeb[i & kmask]);
            // ---- Fokker-Planck (src/SM/FokkerPlanckMap.cpp:197-211), store, projection partial
            const float4* s4 = reinterpret_cast<const float4*>(srow) + q;
            const float4 A = s4[-1], B = s4[0], C = s4[1];
            const float in[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float v = __fmul_rn(in[2 + j], w5[j][0]);
                v = madd<FMA>(in[3 + j], w5[j][1], v);
                v = madd<FMA>(in[4 + j], w5[j][2], v);
                v = madd<FMA>(in[5 + j], w5[j][3], v);
                v = madd<FMA>(in[6 + j], w5[j][4], v);
                o[j] = v;
            }
            gout[(size_t)i * nq] = make_float4(o[0], o[1], o[2], o[3]);
            float p = __fmul_rn(o[0], ws4.x);  // src/PS/PhaseSpace.cpp:316-319 (products rounded)
            p = __fadd_rn(p, __fmul_rn(o[1], ws4.y));
            p = __fadd_rn(p, __fmul_rn(o[2], ws4.z));
            p = __fadd_rn(p, __fmul_rn(o[3], ws4.w));
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, o2));
            if (lane == 0) red[(i & 1) * 32 + warp] = p;
            // the previous row's warp partials are complete (they were written before this
            // row's barrier): one thread adds them in fixed order
            if (tid == 32 % ndf_threads && i > 0) {
                float v = 0.0f;
                for (uint32_t k = 0; k < sa.ndf; k++) v = __fadd_rn(v, red[((i - 1) & 1) * 32 + k]);
                g.rho_out[(size_t)member * n + xa + i - 1] = v;
            }
        }
        named_bar_sync(1, ndf_threads);
        if (tid == 0) {
            float v = 0.0f;
            for (uint32_t k = 0; k < sa.ndf; k++) v = __fadd_rn(v, red[((S - 1) & 1) * 32 + k]);
            g.rho_out[(size_t)member * n + xa + S - 1] = v;
        }
    } else {
        if (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == sa.ndf) {
            // =============================== producer warp ====================================
            for (uint32_t r = 0; r < RIN; r++) {
                const uint32_t s = r & kmask;
                if (r >= K) mbar_wait(&freeb[s], ((r >> klog) - 1) & 1);
                const int x = xs + (int)r;
                float* dst = ring + (size_t)s * pitch + kPad;
                if ((uint32_t)x < n) {
                    if (lane == 0) {
                        fence_proxy_async();  // earlier generic-proxy accesses to the slot are done
                        mbar_expect_tx(&full[s], n * (uint32_t)sizeof(float));
                        bulk_g2s(dst, gin + (size_t)x * n, n * sizeof(float), &full[s]);
                    }
                } else {
                    // row outside the grid: zeros ("taps outside the grid are dropped")
                    for (uint32_t i = lane; i < nq; i += 32) reinterpret_cast<float4*>(dst)[i] = make_float4(0, 0, 0, 0);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                }
                __syncwarp();
            }
        } else {
            // =============================== kick warps =======================================
            const uint32_t kw = warp - sa.ndf - 1;
            const float scal = g.has_wake ? g.wakescaling[member] : 0.0f;
            for (uint32_t r = kw; r < RIN; r += sa.nkick) {
                const uint32_t s = r & kmask;
                const int x = xs + (int)r;
                int wfirst = -1;
                float4 wwt = make_float4(0, 0, 0, 0);
                if ((uint32_t)x < n && g.has_wake) {
                    // wake potential of row x while its copy is still in flight
                    const double* wf = g.wakefn + (x + (int)n - 1) - (int)lane;
                    const double* rl = rho + lane;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
                    for (uint32_t j = 0; j < n; j += 128) {
                        a0 = __fma_rn(rl[j], __ldg(wf - (int)j), a0);
                        a1 = __fma_rn(rl[j + 32], __ldg(wf - (int)j - 32), a1);
                        a2 = __fma_rn(rl[j + 64], __ldg(wf - (int)j - 64), a2);
                        a3 = __fma_rn(rl[j + 96], __ldg(wf - (int)j - 96), a3);
                    }
                    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    const float wk = __fmul_rn(scal, (float)acc);  // src/PS/ElectricField.cpp:236-241
                    ivs_taps t = ivs_kick_taps(wk, n, g.it);
                    if (t.first < -kPad || t.first > kPad - 4) {
                        if (lane == 0) atomicExch(g.status, 1);
                        t.first = -1;
                        t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
                    }
                    wfirst = t.first;
                    wwt = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
                    if (lane == 0 && r >= hl && r < hl + S) g.wake_out[(size_t)member * n + x] = wk;
                }
                mbar_wait(&full[s], (r >> klog) & 1);
                if ((uint32_t)x < n) {
                    float* row = ring + (size_t)s * pitch + kPad;
                    if (g.has_wake) row_pass_v4<FMA>(row, nq, wfirst, wwt, lane);
                    row_pass_v4<FMA>(row, nq, __ldg(&g.rf_first[x]), __ldg(&g.rf_w[x]), lane);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&kicked[s]);
            }
        }
    }
}

struct launch_cfg {
    int tx;
    uint32_t threads;
    size_t smem;
};

size_t smem_bytes(uint32_t n, uint32_t R, int tx, uint32_t threads, bool vec) {
    const uint32_t pitch = n + 2 * kPad;
    size_t b = (size_t)R * pitch * 4 + (size_t)n * (vec ? 8 : 4);
    b += ((R + 3) & ~3u) * 4 + (size_t)R * 16 + (size_t)tx * (threads / 32) * 4;
    return b;
}

constexpr int kRG = 2;  // row groups of the vectorised kernel's last stage (threads = n/4 * kRG)

template <int TX>
ivs_status launch_vec(ivs_ctx* ctx, const fused_args& a, dim3 grid, uint32_t threads, size_t smem) {
    if (ctx->arith == IVS_ARITH_FMA) fused_step_kernel<TX, kRG, true><<<grid, threads, smem, ctx->stream>>>(a);
    else fused_step_kernel<TX, kRG, false><<<grid, threads, smem, ctx->stream>>>(a);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
template <int TX>
ivs_status set_smem_attr(ivs_ctx* ctx, bool vec, size_t smem) {
    if (vec) {
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<TX, kRG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<TX, kRG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_generic_kernel<TX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_generic_kernel<TX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    return IVS_OK;
}

template <int TX>
ivs_status launch_tx(ivs_ctx* ctx, const fused_args& a, dim3 grid, uint32_t threads, size_t smem) {
    if (ctx->arith == IVS_ARITH_FMA) fused_step_generic_kernel<TX, true><<<grid, threads, smem, ctx->stream>>>(a);
    else fused_step_generic_kernel<TX, false><<<grid, threads, smem, ctx->stream>>>(a);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}

}  // namespace

ivs_status ivs_fused_plan_create(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan) {
    const uint32_t n = s->n;
    if (n % 32 != 0)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused step needs a grid size that is a multiple of 32: pass IVS_SIM_UNFUSED");
    if (s->np && s->desc.fptrack == 2)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED,
                        "FPTracking::approximation2 reads the drifted grid, which the fused step never "
                        "materialises: pass IVS_SIM_UNFUSED");
    // static tap tables from the offsets, with the arithmetic of KickMap::updateSM
    std::vector<float> rf(n), dr(n), fpw((size_t)n * 5);
    if (ivs_download(ctx, rf.data(), s->rf, n * 4) || ivs_download(ctx, dr.data(), s->drift, n * 4) ||
        ivs_download(ctx, fpw.data(), s->fpw, (size_t)n * 5 * 4))
        return IVS_ERR_CUDA;
    std::vector<int> rf_first(n), dr_first(n), fp_base(n);
    std::vector<float4> rf_w(n), dr_w(n), fp_w(n);
    int dmin = 0, dmax = -3;
    bool any = false;
    for (uint32_t i = 0; i < n; i++) {
        ivs_taps t = ivs_kick_taps(rf[i], n, s->it);
        if (t.first < -kPad || t.first > kPad - 4)
            return ivs_fail(ctx, IVS_ERR_HALO, "RF kick displacement exceeds the fused kernel's staged pad: pass IVS_SIM_UNFUSED");
        rf_first[i] = t.first;
        rf_w[i] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
        ivs_taps d = ivs_kick_taps(dr[i], n, s->it);
        const bool zero = (d.w[0] == 0 && d.w[1] == 0 && d.w[2] == 0 && d.w[3] == 0);
        if (zero) d.first = -1;
        dr_first[i] = d.first;
        dr_w[i] = make_float4(d.w[0], d.w[1], d.w[2], d.w[3]);
        if (!zero) {
            dmin = any ? std::min(dmin, d.first) : d.first;
            dmax = any ? std::max(dmax, d.first) : d.first;
            any = true;
        }
        // Fokker-Planck: 4 consecutive taps starting at y-2 (rows below the centre, 3-point
        // stencil, identity) or y-1 (rows from the centre on), cf. src/SM/FokkerPlanckMap.cpp:71-108
        const float* w5 = &fpw[(size_t)i * 5];
        if (w5[4] == 0.0f) { fp_base[i] = -2; fp_w[i] = make_float4(w5[0], w5[1], w5[2], w5[3]); }
        else if (w5[0] == 0.0f) { fp_base[i] = -1; fp_w[i] = make_float4(w5[1], w5[2], w5[3], w5[4]); }
        else return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "5-point Fokker-Planck rows are not supported by the fused step");
    }
    if (!any) { dmin = -1; dmax = -1; }
    plan->halo_lo = (uint32_t)std::max(0, -std::min(dmin, -1));
    plan->halo_hi = (uint32_t)std::max(0, std::max(dmax, -1) + 3);
    // rows with all-zero drift weights read the window starting at -1: keep that in range
    plan->halo_lo = std::max(plan->halo_lo, 1u);
    plan->halo_hi = std::max(plan->halo_hi, 2u);
    plan->pad_y = kPad;
    // ---- streaming pipeline (default for grid sizes 256/512/1024/2048)
    plan->stream = false;
    if (n % 256 == 0 && n <= 2048 && !(s->desc.flags & IVS_SIM_TILED)) {
        const uint32_t H = plan->halo_lo + plan->halo_hi;
        const uint32_t ndf = n / 128;
        const uint32_t warps = (ndf >= 16) ? 32 : std::max<uint32_t>(8, 2 * ndf);
        const uint32_t threads_s = warps * 32;
        uint32_t K = 16;
        while (K < H + 4) K <<= 1;
        const uint32_t pitch = n + 2 * kPad;
        const size_t sb = (size_t)(K + 2) * pitch * 4 + (size_t)n * 8 + (size_t)3 * K * 8 + 2 * 32 * 4;
        if (sb <= 227 * 1024) {
            const uint32_t per_sm = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(227 * 1024 / (sb + 1024)), 1024 / threads_s));
            // strips per member: minimise rounds x (rows through the pipeline per CTA)
            uint64_t best = ~0ull;
            uint32_t bestP = 1;
            for (uint32_t P = 1; P <= n / 8; P++) {
                const uint32_t S0 = ivs_div_up(n, P);
                if ((uint64_t)S0 * (P - 1) >= n) continue;  // last strip would be empty
                const uint64_t units = (uint64_t)P * s->members;
                const uint64_t slots = (uint64_t)ctx->sm_count * per_sm;
                const uint64_t rounds = (units + slots - 1) / slots;
                const uint64_t cost = rounds * (S0 + H + 10);
                if (cost < best) { best = cost; bestP = P; }
            }
            plan->stream = true;
            plan->strip = ivs_div_up(n, bestP);
            plan->kslots = K;
            plan->ndf = ndf;
            plan->nkick = warps - ndf - 1;
            plan->threads = threads_s;
            plan->smem = sb;
            plan->tx = 0;
        }
    }
    // tile height: minimise (rounds of CTAs over the machine) x (rows staged per CTA)
    const uint32_t halo = plan->halo_lo + plan->halo_hi;
    const int candidates[3] = {16, 8, 4};
    const bool vec = (n % 256 == 0) && (n / 4) * kRG <= 1024;
    plan->vec = vec;
    int tx = 0;
    uint32_t threads = 0;
    size_t smem = 0;
    uint64_t best = ~0ull;
    for (int c : candidates) {
        if ((uint32_t)c > n) continue;
        const uint32_t R = c + halo;
        uint32_t th = std::min<uint32_t>(1024, std::max<uint32_t>(128, ((R * n / 32) / 32) * 32));
        th = std::min<uint32_t>(th, std::max<uint32_t>(128, n / 2));
        if (vec) th = (n / 4) * kRG;
        const size_t sb = smem_bytes(n, R, c, th, vec);
        if (sb > 227 * 1024) continue;
        // 64 registers per thread (launch bounds) -> at most 1024 resident threads per SM
        const uint32_t per_sm = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(227 * 1024 / (sb + 1024)), 1024 / th));
        const uint64_t ctas = (uint64_t)ivs_div_up(n, c) * s->members;
        const uint64_t rounds = (ctas + (uint64_t)ctx->sm_count * per_sm - 1) / ((uint64_t)ctx->sm_count * per_sm);
        const uint64_t cost = rounds * R * std::max<uint32_t>(1, n / th);
        if (cost < best) { best = cost; tx = c; threads = th; smem = sb; }
    }
    if (plan->stream) { tx = 16; threads = plan->threads; smem = plan->smem; }
    if (tx == 0)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "drift displacement needs more halo rows than fit in shared memory: pass IVS_SIM_UNFUSED");
    if (!plan->stream) { plan->tx = tx; plan->threads = threads; plan->smem = smem; }
    plan->launches_per_step = 1;
    auto up = [&](void** dst, const void* src, size_t bytes) -> ivs_status {
        IVS_CUDA(ctx, cudaMalloc(dst, bytes));
        IVS_CUDA(ctx, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return IVS_OK;
    };
    ivs_status st;
    if ((st = up(&plan->rf_first, rf_first.data(), n * 4)) || (st = up(&plan->rf_w, rf_w.data(), n * 16)) ||
        (st = up(&plan->dr_first, dr_first.data(), n * 4)) || (st = up(&plan->dr_w, dr_w.data(), n * 16)) ||
        (st = up(&plan->fp_base, fp_base.data(), n * 4)) || (st = up(&plan->fp_w, fp_w.data(), n * 16)))
        return st;
    // opt in to the large dynamic shared memory once (not per launch)
    if (plan->stream) {
        IVS_CUDA(ctx, cudaFuncSetAttribute(stream_step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(stream_step_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(stream_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(stream_step_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem));
        plan->ready = true;
        return IVS_OK;
    }
    switch (tx) {
    case 16: st = set_smem_attr<16>(ctx, vec, smem); break;
    case 8: st = set_smem_attr<8>(ctx, vec, smem); break;
    default: st = set_smem_attr<4>(ctx, vec, smem); break;
    }
    if (st != IVS_OK) return st;
    plan->ready = true;
    return IVS_OK;
}

void ivs_fused_plan_destroy(ivs_fused_plan* plan) {
    void** bufs[] = {&plan->rf_first, &plan->rf_w, &plan->dr_first, &plan->dr_w, &plan->fp_base, &plan->fp_w};
    for (void** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
    }
    plan->ready = false;
}

ivs_status ivs_fused_enqueue_step(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s) {
    if (!plan->ready) return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused plan missing");
    fused_args a;
    a.in = s->grid[s->cur];
    a.out = s->grid[s->cur ^ 1];
    a.rho_in = s->profile[s->pcur];
    a.rho_out = s->profile[s->pcur ^ 1];
    a.wake_out = s->wake;
    a.wakefn = s->wakefunction;
    a.wakescaling = s->wakescaling;
    a.rf_first = (const int*)plan->rf_first; a.rf_w = (const float4*)plan->rf_w;
    a.dr_first = (const int*)plan->dr_first; a.dr_w = (const float4*)plan->dr_w;
    a.fp_base = (const int*)plan->fp_base; a.fp_w = (const float4*)plan->fp_w;
    a.fp_w5 = s->fpw;
    a.ws = s->ws;
    a.status = s->status;
    a.n = s->n; a.it = s->it; a.hl = plan->halo_lo; a.hr = plan->halo_hi; a.pitch = s->n + 2 * kPad;
    a.has_wake = s->has_wake ? 1 : 0;
    ivs_status st;
    if (plan->stream) {
        stream_args sa;
        sa.f = a;
        sa.strip = plan->strip; sa.kslots = plan->kslots; sa.ndf = plan->ndf; sa.nkick = plan->nkick;
        const dim3 sgrid(ivs_div_up(s->n, plan->strip), s->members);
        const bool fma = ctx->arith == IVS_ARITH_FMA;
        const bool reb = false;  // setmaxnreg rebalancing: ptxas keeps the 64-register allocation, not used yet
        if (fma && reb) stream_step_kernel<true, true><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        else if (fma) stream_step_kernel<true, false><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        else if (reb) stream_step_kernel<false, true><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        else stream_step_kernel<false, false><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        IVS_LAUNCHED(ctx);
        st = IVS_OK;
        goto launched;
    }
    {
    const dim3 grid(ivs_div_up(s->n, plan->tx), s->members);
    if (plan->vec) {
        switch (plan->tx) {
        case 16: st = launch_vec<16>(ctx, a, grid, plan->threads, plan->smem); break;
        case 8: st = launch_vec<8>(ctx, a, grid, plan->threads, plan->smem); break;
        default: st = launch_vec<4>(ctx, a, grid, plan->threads, plan->smem); break;
        }
    } else {
        switch (plan->tx) {
        case 16: st = launch_tx<16>(ctx, a, grid, plan->threads, plan->smem); break;
        case 8: st = launch_tx<8>(ctx, a, grid, plan->threads, plan->smem); break;
        default: st = launch_tx<4>(ctx, a, grid, plan->threads, plan->smem); break;
        }
    }
    }
launched:
    if (st != IVS_OK) return st;
    s->cur ^= 1;
    s->pcur ^= 1;
    if (s->np) {
        st = ivs_track_step_launch(ctx, s->particles, s->np, s->members, s->n, s->has_wake ? s->wake : nullptr,
                                   s->rf, s->drift, s->has_fp ? s->fpw : nullptr, nullptr, s->desc.fptrack,
                                   s->desc.e1, s->desc.dp, s->desc.seed, s->steps_done);
    }
    return st;
}
