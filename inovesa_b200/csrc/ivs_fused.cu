// ivs_fused.cu -- the whole simulation step as ONE kernel (sm_100a).
//
//   wake potential -> wake kick (y) -> RF kick (y) -> drift (x) -> Fokker-Planck (y) -> X projection
//   (src/main.cpp:948-1053; maps src/SM/KickMap.cpp:221-266, src/SM/FokkerPlanckMap.cpp:197-211,
//    projection src/PS/PhaseSpace.cpp:313-321, wake src/PS/ElectricField.cpp:186-253)
//
// Three kernels live here: the streaming kernel (stream_step_kernel, further down: the default for
// grid sizes 256 ... 2048 and what the benchmark runs) and two tile kernels, the first fused
// versions, kept for other grid sizes and as independent implementations to test against.
//
// Data flow of the TILE kernels per CTA (one tile = TX consecutive rows x of one ensemble member, all N columns y):
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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "ivs_fused_common.cuh"

namespace {

constexpr int kPad = 16;   // zero columns on each side of a staged row (floats)
constexpr int kChunk = 8;  // outputs per lane and chunk in the row passes

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
    __shared__ int ctl[2];
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
        step_ctl_begin(g, ctl);
        mbar_init(&bar, 1);
        uint32_t valid = 0;
        for (uint32_t r = 0; r < R; r++) valid += ((uint32_t)(xs + (int)r) < n);
        if (ctl[0] == kStatusClean) mbar_expect_tx(&bar, valid * n * (uint32_t)sizeof(float));
    }
    __syncthreads();
    if (ctl[0] != kStatusClean) return;  // an earlier step overflowed: see step_ctl_begin
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
                    step_ctl_overflow(g, ctl);
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
    __shared__ int ctl[2];
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
        step_ctl_begin(g, ctl);
        mbar_init(&bar, 1);
        uint32_t valid = 0;
        for (uint32_t r = 0; r < R; r++) valid += ((uint32_t)(xs + (int)r) < n);
        if (ctl[0] == kStatusClean) mbar_expect_tx(&bar, valid * n * (uint32_t)sizeof(float));
    }
    __syncthreads();
    if (ctl[0] != kStatusClean) return;  // an earlier step overflowed: see step_ctl_begin
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
                    step_ctl_overflow(g, ctl);
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
// Streaming variant (grid sizes that are multiples of 256, up to 2048): a CTA owns a strip of
// consecutive rows of one member and streams them through a ring of row slots in shared memory.
// A thread owns ONE quad of columns (a float4) for the whole strip, a warp 128 adjacent columns;
// the n/128 warps of a CTA run independently of each other -- no block-wide barrier in the row loop:
//
//   staging : one TMA bulk copy per input row into the next free slot.  full[slot] completes when
//             the bytes have landed, empty[slot] counts the warps that are done with the row that
//             was there.  There is no producer warp (it would unbalance the four schedulers): the
//             duty rotates, row r is issued by lane 0 of warp r mod n_warps when its probe of
//             empty[slot] (started a row earlier) says the slot is free.
//   kick    : all lanes of a warp work on the same input row; a lane reads the 10..13 staged values
//             its quad depends on (3-4 LDS.128) and applies wake kick and RF kick back to back in
//             registers -- the 7 wake-kicked values the RF taps need are recomputed per lane instead
//             of being exchanged, so the row-local passes cost no shared-memory round trip;
//   drift   : column-local, a sliding window of kicked rows in registers.  The window holds
//             NT = 4 + spread rows, spread = how far the first drift tap varies over the warp's 128
//             columns; a column's 4 weights sit at its offset, the other slots hold 0;
//   Fokker-Planck: needs 2 neighbours left and right: warp shuffles inside the warp; across a
//             warp boundary the two edge lanes post their values in a small mailbox (sequence
//             number = row) that the neighbour warp's edge lane polls.  The stencil runs two rows
//             behind the drift, so a warp only waits if its neighbour is more than a row late.
//   then STG.128 and the Simpson-weighted partial row sum (next step's profile).
//
// The wake potential of the strip's rows (Toeplitz product in double) is formed by the whole CTA
// while the first copies are in flight.  Arithmetic per cell is identical to the tile kernels
// (same taps, same order, same summation tree of the profile).

constexpr int kPadS = 32;     // zero columns each side of a staged row: two kicks of up to 16 cells
constexpr uint32_t kMinStrip = 2;  // shortest strip the planner considers (rows), see ivs_fused_plan_create
constexpr int kMailSlots = 32; // mailbox depth per warp edge (> ring slots + kLag + 5, see the hand-back rule)
#ifndef IVS_WARP_CLOCKS
#define IVS_WARP_CLOCKS 0  // development aid: per-warp-index totals of loop time and poll counts (stderr at plan destroy)
#endif
#if IVS_WARP_CLOCKS
__device__ unsigned long long g_wstat[4][32];  // [loop-end clock, mailbox polls, row waits, strips][warp]
#endif
#ifndef IVS_LAG
#define IVS_LAG 4
#endif
#ifndef IVS_PROBE_NOKICK
#define IVS_PROBE_NOKICK 0  // development aids: timing probes that leave out part of the arithmetic
#endif
#ifndef IVS_PROBE_NOFP
#define IVS_PROBE_NOFP 0
#endif
#ifndef IVS_PROBE_FIXRR
#define IVS_PROBE_FIXRR 0  // timing probe (wrong results): one alignment variant of the kick for every row
#endif
#ifndef IVS_PROBE_NT6
#define IVS_PROBE_NT6 0    // timing probe (right results): every warp runs the 6-row drift window
#endif
constexpr int kLag = IVS_LAG;        // drifted rows wait this many iterations (in a per-thread FIFO in shared memory)
                               // for their Fokker-Planck step: the slack between neighbouring warps

// wake kick then RF kick of one quad of columns (y0 .. y0+3) of a staged row, in registers.
//   k1[m] = sum_t in[m + f1 + t] w1[t]         wake kick (src/SM/KickMap.cpp:243-266), 0 outside the grid
//   out[y] = sum_t k1[y + f2 + t] w2[t]        RF kick, same loop
// `p` is the shared address of in[y0 + a], a = (f1+f2) rounded down to a multiple of 4, RR the rest
// (A, B, C = the 12 staged values from p on, loaded by the caller);
// m0 = y0 + f2 is the first wake-kicked column the quad needs.
template <bool FMA, int RR>
__device__ __forceinline__ float4 kick_quad_rr(uint32_t p, const float4 A, const float4 B, const float4 C, int m0, uint32_t n,
                                               bool near_edge, const float4 w1, const float4 w2) {
    float in[16];
    {
        in[0] = A.x; in[1] = A.y; in[2] = A.z; in[3] = A.w;
        in[4] = B.x; in[5] = B.y; in[6] = B.z; in[7] = B.w;
        in[8] = C.x; in[9] = C.y; in[10] = C.z; in[11] = C.w;
        if (RR == 3) {
            const float4 D = lds128(p + 48);
            in[12] = D.x; in[13] = D.y; in[14] = D.z; in[15] = D.w;
        }
    }
    float k1[7];
#pragma unroll
    for (int i = 0; i < 7; i++) k1[i] = tap4<FMA>(in[RR + i], in[RR + i + 1], in[RR + i + 2], in[RR + i + 3], w1);
    // some of m0 .. m0+6 lie outside the grid: the wake-kicked grid is 0 there.  near_edge is a
    // per-warp constant (only the first and the last warp of a row can get there)
    if (near_edge && (uint32_t)m0 > n - 7) {
#pragma unroll
        for (int i = 0; i < 7; i++)
            if ((uint32_t)(m0 + i) >= n) k1[i] = 0.0f;
    }
    float4 o;
    o.x = tap4<FMA>(k1[0], k1[1], k1[2], k1[3], w2);
    o.y = tap4<FMA>(k1[1], k1[2], k1[3], k1[4], w2);
    o.z = tap4<FMA>(k1[2], k1[3], k1[4], k1[5], w2);
    o.w = tap4<FMA>(k1[3], k1[4], k1[5], k1[6], w2);
    return o;
}
template <bool FMA>
__device__ __forceinline__ float4 kick_quad(uint32_t row_q, int fsum, int m0, uint32_t n, bool near_edge, const float4 w1,
                                            const float4 w2) {
    const int a = fsum & ~3;  // floor to a multiple of 4 (two's complement: also for negatives)
    const uint32_t p = row_q + (uint32_t)(a * 4);
    // the three quads every alignment needs are requested before the switch (their latency then
    // overlaps the branch); only the shift by 3 needs a fourth.  (Requested before the tap tables
    // instead, or a whole stencil step earlier, measured 1 % and 6 % slower.)
    const float4 A = lds128(p), B = lds128(p + 16), C = lds128(p + 32);
    switch (fsum - a) {
    case 0: return kick_quad_rr<FMA, 0>(p, A, B, C, m0, n, near_edge, w1, w2);
    case 1: return kick_quad_rr<FMA, 1>(p, A, B, C, m0, n, near_edge, w1, w2);
    case 2: return kick_quad_rr<FMA, 2>(p, A, B, C, m0, n, near_edge, w1, w2);
    default: return kick_quad_rr<FMA, 3>(p, A, B, C, m0, n, near_edge, w1, w2);
    }
}

#if IVS_PROBE_NOFP
__device__ __forceinline__ float warp_sum_probe(float v) {
    for (int o = 8; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

// wake kick + RF kick of the QPT quads a thread owns in one staged row (quad j sits 32 quads =
// 512 bytes after quad j-1).  The alignment of the first tap is the same for all of them, so one
// warp-uniform switch serves QPT independent chains of arithmetic.
template <bool FMA, int QPT>
__device__ __forceinline__ void kick_quads(uint32_t row_q, int fsum, int m0, uint32_t n, const bool (&near_edge)[QPT],
                                           const float4 w1, const float4 w2, float4 (&kq)[QPT]) {
    const int a = fsum & ~3;  // floor to a multiple of 4 (two's complement: also for negatives)
    const uint32_t p = row_q + (uint32_t)(a * 4);
    // the three staged quads every alignment needs are requested before the switch (their latency
    // then overlaps the branch); only the shift by 3 needs a fourth
    float4 A[QPT], B[QPT], C[QPT];
#pragma unroll
    for (int j = 0; j < QPT; j++) {
        A[j] = lds128(p + j * 512);
        B[j] = lds128(p + j * 512 + 16);
        C[j] = lds128(p + j * 512 + 32);
    }
#define IVS_KICK_CASE(RR)                                                                                       \
    _Pragma("unroll") for (int j = 0; j < QPT; j++)                                                             \
        kq[j] = kick_quad_rr<FMA, RR>(p + j * 512, A[j], B[j], C[j], m0 + j * 128, n, near_edge[j], w1, w2);
#if IVS_PROBE_FIXRR
    IVS_KICK_CASE(1)
#else
    switch (fsum - a) {
    case 0: IVS_KICK_CASE(0) break;
    case 1: IVS_KICK_CASE(1) break;
    case 2: IVS_KICK_CASE(2) break;
    default: IVS_KICK_CASE(3) break;
    }
#endif
#undef IVS_KICK_CASE
}

// A thread owns QPT quads of columns (QPT float4 per row) for the whole strip: quad j of the thread
// is quad (warp * QPT + j) * 32 + lane of the row, i.e. a warp covers QPT adjacent segments of 128
// columns and a lane's quads are 128 columns apart (conflict-free LDS.128, coalesced STG.128).
// QPT = 2 halves the warps per row (8 warps of up to 255 registers at n = 2048 instead of 16 of
// 128): the per-row control (barrier probes, tap tables, cursors, staging duty, mailboxes) is paid
// once for twice the columns and every dependent chain has an independent twin to interleave with.
template <bool FMA, int QPT>
__global__ void __launch_bounds__(512 / QPT, 1) stream_step_kernel(const stream_args sa) {
    const fused_args& g = sa.f;
    extern __shared__ __align__(128) float smem[];
    __shared__ int ctl[2];
    const uint32_t n = g.n, nq = g.n >> 2, pitch = g.pitch, hl = g.hl;
    const uint32_t H = g.hl + g.hr;
    const uint32_t K = sa.kslots;
    const uint32_t T = blockDim.x;             // == nq / QPT
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
    const uint32_t NSEG = NW * QPT;            // segments of 128 columns in a row
    const uint32_t member = blockIdx.y;
    const uint32_t xa = blockIdx.x * sa.strip;
    const uint32_t S = min(sa.strip, n - xa);  // output rows of this strip
    const uint32_t RIN = S + H;                // input rows streamed through the ring
    const int xs = (int)xa - (int)hl;          // grid row of input row 0

    // shared memory carve-up (every block 16-byte aligned)
    unsigned char* sp = reinterpret_cast<unsigned char*>(smem);
    auto carve = [&sp](size_t bytes) {
        unsigned char* p = sp;
        sp += (bytes + 15) & ~(size_t)15;
        return p;
    };
    float* ring = reinterpret_cast<float*>(carve((size_t)K * pitch * 4));        // K row slots
    double* rho = reinterpret_cast<double*>(carve((size_t)(n + (n >> 5)) * 8));  // last step's profile (skewed)
    double* wwin = reinterpret_cast<double*>(carve((size_t)sa.wwin_len * 8));    // wake function window
    double* partial = reinterpret_cast<double*>(carve((size_t)NSEG * sa.max_rin * 8));  // wake partial sums
    float* red = reinterpret_cast<float*>(carve((size_t)sa.max_strip * NSEG * 4));       // row-sum partials
    // per input row: wake taps, RF taps, (f1 + f2, f2, -, -)
    float4* tab = reinterpret_cast<float4*>(carve((size_t)(sa.max_rin + 1) * 48));
    float4* mail = reinterpret_cast<float4*>(carve((size_t)NW * 2 * kMailSlots * 16)); // [warp][side][slot] = (a, b, seq, -)
    float4* fifo = reinterpret_cast<float4*>(carve((size_t)kLag * QPT * T * 16));      // [slot][quad j][thread] drifted quads
    uint64_t* full = reinterpret_cast<uint64_t*>(carve((size_t)K * 8));
    uint64_t* empty = reinterpret_cast<uint64_t*>(carve((size_t)K * 8));

    const float* gin = g.in + (size_t)member * n * n;
    const uint32_t pitchB = pitch * 4;
    const uint32_t full_b = smem_u32(full), empty_b = smem_u32(empty);

    // one staged row: copy it, or (outside the grid, where the RF taps are 0) complete the phase by hand
    auto issue_row = [&](uint32_t r, uint32_t slot) {
        const int x = xs + (int)r;
        if ((uint32_t)x < n) {
            fence_proxy_async();  // generic-proxy reads of the slot are ordered before the async write
            mbar_expect_tx(&full[slot], n * (uint32_t)sizeof(float));
            bulk_g2s(ring + (size_t)slot * pitch + kPadS, gin + (size_t)x * n, n * (uint32_t)sizeof(float), &full[slot]);
        } else {
            mbar_arrive_addr(full_b + slot * 8);
        }
    };

    if (tid == 0) step_ctl_begin(g, ctl);
    if (tid < K) {  // one thread per slot; a single fence covers both initialisations
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_b + tid * 8), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_b + tid * 8), "r"(NW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- per-thread constants, requested early: drift taps, Fokker-Planck and Simpson weights
    uint32_t q[QPT];  // this thread's quads of columns, y = 4q .. 4q+3
#pragma unroll
    for (int j = 0; j < QPT; j++) q[j] = (warp * QPT + j) * 32 + lane;
    int4 df[QPT];
    float4 dwl[QPT][4];
    float w5[QPT][4][5];
    float4 ws4[QPT];
#pragma unroll
    for (int j = 0; j < QPT; j++) {
        df[j] = __ldg(reinterpret_cast<const int4*>(g.dr_first) + q[j]);
#pragma unroll
        for (int c = 0; c < 4; c++) dwl[j][c] = __ldg(&g.dr_w[4 * q[j] + c]);
        const float4* wp = reinterpret_cast<const float4*>(g.fp_w5 + (size_t)q[j] * 20);
        float t[20];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float4 v = __ldg(wp + k);
            t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < 5; k++) w5[j][c][k] = t[c * 5 + k];
        ws4[j] = __ldg(reinterpret_cast<const float4*>(g.ws) + q[j]);
    }
    // last step's profile and the first batch of the wake-function window are requested in the same
    // round trip (T = n / (4 QPT) threads: 4 QPT profile entries and 5 QPT window entries each)
    // w[x - x'] is stored at table index x - x' + n - 1; window entry t holds the table entry
    // xs + t, so row r and x' meet at t = r - x' + n - 1
    constexpr int NRV = 4 * QPT, NWV = 5 * QPT;
    const int wtop = (int)(2 * n - 1);
    const uint32_t tcount = n + ((RIN + 63) & ~63u);  // <= NWV T for n >= 1024
    float rv[NRV];
    double wv0[NWV];
    // ... and so is the static part (RF taps) of this thread's first entry of the per-row table
    int pre_f2 = -1;
    float4 pre_w2 = make_float4(0, 0, 0, 0);  // rows outside the grid are zeros
    if (tid < RIN && (uint32_t)(xs + (int)tid) < n) {
        pre_f2 = __ldg(&g.rf_first[xs + (int)tid]);
        pre_w2 = __ldg(&g.rf_w[xs + (int)tid]);
    }
    if (g.has_wake) {
        const float* rin = g.rho_in + (size_t)member * n;
#pragma unroll
        for (int k = 0; k < NRV; k++) rv[k] = __ldg(rin + tid + k * T);
#pragma unroll
        for (int k = 0; k < NWV; k++)
            wv0[k] = __ldg(g.wakefn + min(max(xs + (int)(tid + k * T), 0), wtop));  // clamped entries belong to rows outside the grid
    }
    __syncthreads();
    // an earlier step of this run left the staged pad (the host redoes it on another path and
    // re-enqueues what follows): nothing may be written any more
    if (ctl[0] != kStatusClean) return;
    if (tid < min(K, RIN)) issue_row(tid, tid);  // the first ring-full flies during the set-up

    // ---- set-up: zero pads (and slots of rows outside the grid, whose stale contents would
    // otherwise meet the zero weights as NaN * 0), mailboxes, last step's profile and the
    // wake-function window in shared memory
    for (uint32_t r = warp; r < K; r += NW) {
        float* row = ring + (size_t)r * pitch;
        row[lane] = 0.0f;
        row[n + kPadS + lane] = 0.0f;
        if (r < RIN && (uint32_t)(xs + (int)r) >= n) {
            for (uint32_t i = lane; i < nq; i += 32) reinterpret_cast<float4*>(row + kPadS)[i] = make_float4(0, 0, 0, 0);
        }
    }
    for (uint32_t i = tid; i < NW * 2 * kMailSlots; i += T) mail[i] = make_float4(0, 0, 0, 0);  // seq 0 = nothing posted
    // static part of the per-row table (RF taps), in place before the wake potential is formed
    if (tid < RIN) {
        tab[3 * tid + 1] = pre_w2;
        tab[3 * tid + 2] = make_float4(0.0f, __int_as_float(pre_f2), 0.0f, 0.0f);
    }
    for (uint32_t r = tid + T; r < RIN; r += T) {  // small grids: strips longer than the block
        const int x = xs + (int)r;
        const bool valid = (uint32_t)x < n;
        const int f2 = valid ? __ldg(&g.rf_first[x]) : -1;
        tab[3 * r + 1] = valid ? __ldg(&g.rf_w[x]) : make_float4(0, 0, 0, 0);
        tab[3 * r + 2] = make_float4(0.0f, __int_as_float(f2), 0.0f, 0.0f);
    }
    if (g.has_wake) {
#pragma unroll
        for (int k = 0; k < NRV; k++) {
            const uint32_t i = tid + k * T;
            rho[i + (i >> 5)] = (double)rv[k];
        }
#pragma unroll
        for (int k = 0; k < NWV; k++) {
            const uint32_t t = tid + k * T;
            if (t < tcount) wwin[wwin_pos(t)] = wv0[k];
        }
        for (uint32_t t0 = tid + NWV * T; t0 < tcount; t0 += NWV * T) {  // small grids: the rest of the window
            double wv[NWV];
#pragma unroll
            for (int k = 0; k < NWV; k++)
                wv[k] = __ldg(g.wakefn + min(max(xs + (int)(t0 + k * T), 0), wtop));  // clamped entries belong to rows outside the grid
#pragma unroll
            for (int k = 0; k < NWV; k++) {
                const uint32_t t = t0 + k * T;
                if (t < tcount) wwin[wwin_pos(t)] = wv[k];
            }
        }
    }
    __syncthreads();
    if (g.has_wake) {
        // W[x] = scaling * sum_x' rho[x'] w[x-x']  (src/PS/ElectricField.cpp:186-253 as a Toeplitz
        // product).  Segment s sums x' in [128 s, 128 s + 128); a lane owns 8 consecutive rows and 32
        // consecutive x', so the 8 wake-function values it needs slide through registers: one
        // window load and one profile load per 8 multiply-adds.
        const uint32_t rbk = lane & 7, sub = lane >> 3;
#pragma unroll 1
        for (uint32_t seg = warp * QPT; seg < (warp + 1) * QPT; seg++) {
            const uint32_t xb = seg * 128 + sub * 32;
            const double* rp = rho + xb + (xb >> 5);
            for (uint32_t pass = 0; pass * 64 < RIN; pass++) {  // warp-uniform trip count
                const uint32_t r0 = pass * 64 + rbk * 8;
                const uint32_t t0 = r0 + n - 1 - xb;
                double wv[8], acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    wv[j] = wwin[wwin_pos(t0 + j)];
                    acc[j] = 0.0;
                }
#pragma unroll 8
                for (uint32_t k = 0; k < 32; k++) {
                    const double rh = rp[k];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = __fma_rn(rh, wv[j], acc[j]);
#pragma unroll
                    for (int j = 7; j > 0; j--) wv[j] = wv[j - 1];
                    wv[0] = wwin[wwin_pos(max((int)(t0 - k) - 1, 0))];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
                    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
                }
                if (sub == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (r0 + j < RIN) partial[(size_t)seg * sa.max_rin + r0 + j] = acc[j];
                }
            }
        }
    }
    __syncthreads();
    {
        const float scal = g.has_wake ? g.wakescaling[member] : 0.0f;
        for (uint32_t r = tid; r < RIN; r += T) {
            const int x = xs + (int)r;
            const bool valid = (uint32_t)x < n;
            // no wake map: the identity as a 4-tap row (0, 1, 0, 0) at first = -1
            ivs_taps t;
            t.first = -1;
            t.w[0] = 0.0f; t.w[1] = 1.0f; t.w[2] = 0.0f; t.w[3] = 0.0f;
            if (valid && g.has_wake) {
                double acc = 0.0;
                for (uint32_t k = 0; k < NSEG; k++) acc += partial[(size_t)k * sa.max_rin + r];
                const float wk = __fmul_rn(scal, (float)acc);  // src/PS/ElectricField.cpp:236-241
                t = ivs_kick_taps(wk, n, g.it);
                if (t.first < -16 || t.first > 8) {
                    // beyond the staged pad: record the step (the host redoes it through the kernel
                    // sequence, which has no such limit) and keep the addresses in range
                    step_ctl_overflow(g, ctl);
                    t.first = -1;
                    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
                }
                if (r >= hl && r < hl + S) g.wake_out[(size_t)member * n + x] = wk;
            }
            const int f2 = __float_as_int(tab[3 * r + 2].y);
            tab[3 * r] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
            tab[3 * r + 2].x = __int_as_float(t.first + f2);
        }
        if (tid == 0) tab[3 * RIN + 2] = make_float4(__int_as_float(-2), __int_as_float(-1), 0.0f, 0.0f);  // read ahead, never used
    }
    __syncthreads();  // tables, pads and mailboxes visible to everybody

    // =============================== the row loop ==============================================
    int wfmin = 0x7fffffff, wfmax = -0x7fffffff;
#pragma unroll
    for (int j = 0; j < QPT; j++) {
        wfmin = min(wfmin, min(min(df[j].x, df[j].y), min(df[j].z, df[j].w)));
        wfmax = max(wfmax, max(max(df[j].x, df[j].y), max(df[j].z, df[j].w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wfmin = min(wfmin, __shfl_xor_sync(0xffffffffu, wfmin, o));
        wfmax = max(wfmax, __shfl_xor_sync(0xffffffffu, wfmax, o));
    }
    const int spread = wfmax - wfmin;  // <= 2 (checked when the plan was made)

    const uint32_t ring_q = smem_u32(ring + kPadS) + q[0] * 16;  // this thread's first quad in slot 0
    const int y0 = (int)(q[0] << 2);
    float4* gout = reinterpret_cast<float4*>(g.out + ((size_t)member * n + xa) * n) + q[0];
    // mailboxes: this warp posts the drifted x,y of its first quad (for the left neighbour) and
    // the z,w of its last quad (for the right neighbour); lane 0 / lane 31 read the neighbours'
    const uint32_t mail_b = smem_u32(mail);
    const uint32_t post_l = mail_b + (warp * 2 + 0) * kMailSlots * 16;   // written by lane 0
    const uint32_t post_r = mail_b + (warp * 2 + 1) * kMailSlots * 16;   // written by lane 31
    const uint32_t post_me = lane == 0 ? post_l : post_r;
    const bool edge_l = lane == 0 && warp > 0, edge_r = lane == 31 && warp + 1 < NW;
    // lane 0 reads the left neighbour's right post, lane 31 the right neighbour's left post
    const uint32_t peek = edge_l ? mail_b + ((warp - 1) * 2 + 1) * kMailSlots * 16
                                 : mail_b + ((warp + 1) * 2 + 0) * kMailSlots * 16;

#if IVS_WARP_CLOCKS
    const long long wclk0 = clock64();
    unsigned long long n_mpoll = 0, n_rwait = 0;
#endif
    auto run = [&](auto nt_tag) {
        constexpr int NT = decltype(nt_tag)::value;
        float dw[QPT][4][NT];
#pragma unroll
        for (int j = 0; j < QPT; j++) {
            const int fs[4] = {df[j].x, df[j].y, df[j].z, df[j].w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float wv[4] = {dwl[j][c].x, dwl[j][c].y, dwl[j][c].z, dwl[j][c].w};
                const int off = fs[c] - wfmin;  // 0 .. NT-4
#pragma unroll
                for (int t = 0; t < NT; t++) {
                    float v = 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (t - u >= 0 && t - u <= NT - 4) v = (off == t - u) ? wv[u] : v;
                    dw[j][c][t] = v;
                }
            }
        }
        // rows r0 .. r0+NT-1 make output row 0; kick number k (k = 0, 1, ...) reads row r0 + k
        const uint32_t r0 = (uint32_t)((int)hl + wfmin);
        const uint32_t nkicks = S + NT - 1;

        // ---- staging duty: rows >= K are issued by lane 0 of warp (row mod NW).  The slot of row
        // `prow` cannot be free before this warp itself has handed back the row that is in it, so
        // the probing starts at iteration `pwin`.
        uint32_t prow = K + ((warp + NW - K % NW) % NW);  // first row >= K with row % NW == warp
        uint32_t pslot = prow % K, puse = prow / K;
        uint32_t pfree = 0;  // outcome of the last probe of empty[pslot]
        int pwin = prow < RIN ? (int)prow - (int)K - (int)r0 + 1 : 0x7fffffff;
        auto duty_probe = [&]() {  // warp-uniform: every lane reads the same barrier word
            pfree = mbar_test(empty_b + pslot * 8, (puse - 1) & 1);
        };
        auto duty_issue = [&]() {
            if (pfree) {
                if (lane == 0) issue_row(prow, pslot);
                prow += NW;
                pslot += NW;
                while (pslot >= K) { pslot -= K; puse++; }
                pfree = 0;
            }
            pwin = prow < RIN ? (int)prow - (int)K - (int)r0 + 1 : 0x7fffffff;
        };
        // blocking wait for a staged row; keeps serving the staging duty so that nobody starves
        auto wait_row = [&](uint32_t bar, uint32_t ph, int it) {
            uint32_t tries = 0;
            while (!mbar_try(bar, ph)) {
#if IVS_WARP_CLOCKS
                n_rwait++;
#endif
                if (it >= pwin) {
                    duty_probe();
                    duty_issue();
                }
                if (++tries > (1u << 22)) __trap();  // a copy that never lands is a bug: do not hang the GPU
            }
        };

        uint32_t aload = ring_q + (r0 % K) * pitchB;
        const uint32_t ring_back = pitchB - K * pitchB;
        uint32_t tp = smem_u32(tab) + r0 * 48;
        int2 f = lds64i(tp + 32);
        // drift as running sums: P[j] is the output row that has taken j + 1 of its NT taps so far.
        // A kicked row k feeds tap NT-1 of the oldest sum (which completes it), tap NT-2 of the next,
        // ..., and starts a new one with tap 0 -- the same products added in the same order as a
        // 4-tap sum over a window of rows, without the window (and its register moves).
        float4 P[QPT][NT - 1];
#pragma unroll
        for (int j = 0; j < QPT; j++)
#pragma unroll
            for (int t = 0; t < NT - 1; t++) P[j][t] = make_float4(0, 0, 0, 0);
        bool near_edge[QPT];
#pragma unroll
        for (int j = 0; j < QPT; j++) {
            const int yj = y0 + j * 128;
            near_edge[j] = __any_sync(0xffffffffu, yj < 16 || yj + 19 > (int)n);
        }
        // cursor: full[] entry of the staged row the next kick reads and its phase; empty[] sits eoff behind
        uint32_t kbar = full_b + (r0 % K) * 8, kpar = (r0 / K) & 1;
        const uint32_t full_end = full_b + K * 8, eoff = empty_b - full_b;

        // the rows below the window are never read by this warp
        if (lane == 0) {
            for (uint32_t r = 0; r < r0; r++) mbar_arrive_addr(empty_b + (r % K) * 8);
        }
        wait_row(kbar, kpar, 0);  // the first kick's row

        float4 d[QPT];  // the drifted row of this iteration
#pragma unroll
        for (int j = 0; j < QPT; j++) d[j] = make_float4(0, 0, 0, 0);
        const uint32_t fifo_t = smem_u32(fifo) + tid * 16;  // this thread's entry (quad 0) of FIFO slot 0
        const uint32_t fifo_j = T * 16;                     // from quad j to quad j+1 inside a FIFO slot
        float pprev[QPT];  // this lane's part of a finished row's sum (reduced an iteration later)
#pragma unroll
        for (int j = 0; j < QPT; j++) pprev[j] = 0.0f;
        // where the row sums reduced next go: lane 0 holds segment warp*QPT's, lane 16 the next one's (QPT = 2)
        uint32_t redp = smem_u32(red) + (warp * QPT + (QPT == 2 && lane >= 16 ? 1 : 0)) * 4 - NSEG * 4;
        const size_t gstride = (size_t)nq;
        const bool edge = edge_l || edge_r;
        const bool red_lane = QPT == 2 ? (lane & 15) == 0 : lane == 0;
        const uint32_t lane_up = (lane + 31) & 31, lane_dn = (lane + 1) & 31;

        // ---- one kick step: staged row r0 + it -> wake kick + RF kick -> window -> drifted row d
        // (output row it - NT + 1 once the window is full).  POST: d is a real output row whose edge
        // values the neighbour warps need.  MORE: another kick follows, its row has to have landed.
        auto kick_step = [&](const int it, const uint32_t fa, const bool POST, const bool MORE) {
            // probes whose answers are needed at the end of the step
            uint32_t nbar = kbar + 8, npar = kpar;
            const bool wrap = nbar == full_end;  // the next row sits in slot 0
            if (wrap) { nbar = full_b; npar ^= 1; }
            const uint32_t landed = mbar_test(MORE ? nbar : kbar, MORE ? npar : kpar);
            const float4 w1 = lds128(tp), w2 = lds128(tp + 16);
            float4 kq[QPT];
#if IVS_PROBE_NOKICK  // timing probe (wrong results): the staged quad passes through unkicked
#pragma unroll
            for (int j = 0; j < QPT; j++) {
                kq[j] = lds128(aload + j * 512);
                kq[j].x += 1e-30f * (w1.x + w2.x + (float)f.x);
            }
#else
            kick_quads<FMA, QPT>(aload, f.x, y0 + f.y, n, near_edge, w1, w2, kq);
#endif
            tp += 48;
            f = lds64i(tp + 32);  // next row's entry, a row ahead of its use
            aload += wrap ? ring_back : pitchB;  // ring_back = pitchB - K * pitchB: back to slot 0
#pragma unroll
            for (int j = 0; j < QPT; j++) {  // src/SM/KickMap.cpp:221-242, column-local 4-tap along x
#define IVS_DRIFT_COL(c, m)                                                                           \
                d[j].m = madd<FMA>(kq[j].m, dw[j][c][NT - 1], P[j][NT - 2].m);                        \
                _Pragma("unroll") for (int t = NT - 2; t > 0; t--) P[j][t].m = madd<FMA>(kq[j].m, dw[j][c][t], P[j][t - 1].m); \
                P[j][0].m = __fmul_rn(kq[j].m, dw[j][c][0]);
                IVS_DRIFT_COL(0, x)
                IVS_DRIFT_COL(1, y)
                IVS_DRIFT_COL(2, z)
                IVS_DRIFT_COL(3, w)
#undef IVS_DRIFT_COL
                // parked in this thread's FIFO entry until its stencil step kLag iterations later
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(fa + j * fifo_j), "f"(d[j].x), "f"(d[j].y), "f"(d[j].z), "f"(d[j].w) : "memory");
            }
            if (POST) {
                const uint32_t orow = (uint32_t)(it - NT + 1);
                const uint32_t ms = (orow & (kMailSlots - 1)) * 16;
                // lane 0 posts (x, y) of its first quad for the left neighbour, lane 31 (z, w) of its last for the right one
                const bool l0 = lane == 0;
                sts128_volatile_if(post_me + ms,
                                   make_float4(l0 ? d[0].x : d[QPT - 1].z, l0 ? d[0].y : d[QPT - 1].w, __uint_as_float(orow + 1), 0.0f),
                                   l0 || lane == 31);
            }
            // the staged row is in registers everywhere (its drift has been computed): hand it back
            mbar_arrive_if(kbar + eoff, lane == 0);
            kbar = nbar;
            kpar = npar;
            if (it >= pwin || !landed) {  // one rarely taken region for both slow paths
                if (it >= pwin) {  // this warp's turn to stage a row (one in NW iterations)
                    duty_probe();
                    duty_issue();
                }
                if (!landed) wait_row(nbar, npar, it);
            }
        };
        // ---- one stencil step on drifted output row `orow` (out of the FIFO): neighbours,
        // Fokker-Planck, store, partial row sum; the warp sum of the row before is reduced on the side
        auto fp_step = [&](const int it, const uint32_t fa, const uint32_t orow) {
            float4 dp[QPT];
#pragma unroll
            for (int j = 0; j < QPT; j++) dp[j] = lds128(fa + j * fifo_j);
            float4 m = make_float4(0, 0, 0, 0);
            const uint32_t pa = peek + (orow & (kMailSlots - 1)) * 16;
            if (edge) m = lds128_volatile(pa);
            // neighbours inside the warp: quad j of lane l sits between quad j of lanes l-1 and l+1;
            // lane 0's left neighbour is lane 31's quad j-1, lane 31's right one lane 0's quad j+1
            float Lz[QPT], Lw[QPT], Rx[QPT], Ry[QPT];
#pragma unroll
            for (int j = 0; j < QPT; j++) {
                const float sz = (j > 0 && lane == 31) ? dp[j - 1].z : dp[j].z;
                const float sw = (j > 0 && lane == 31) ? dp[j - 1].w : dp[j].w;
                const float sx = (j + 1 < QPT && lane == 0) ? dp[j + 1].x : dp[j].x;
                const float sy = (j + 1 < QPT && lane == 0) ? dp[j + 1].y : dp[j].y;
                Lz[j] = __shfl_sync(0xffffffffu, sz, lane_up);
                Lw[j] = __shfl_sync(0xffffffffu, sw, lane_up);
                Rx[j] = __shfl_sync(0xffffffffu, sx, lane_dn);
                Ry[j] = __shfl_sync(0xffffffffu, sy, lane_dn);
            }
            // the row sums of the row before: first level of the butterfly (QPT = 2: the two segments'
            // sums are transposed into the two half-warps, same tree per segment as with one quad)
            float ps;
            if (QPT == 2) {
                const bool upper = (lane & 16) != 0;
                const float send = upper ? pprev[0] : pprev[QPT - 1];
                const float keep = upper ? pprev[QPT - 1] : pprev[0];
                ps = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            } else {
                ps = pprev[0];
                ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 16));
            }
            if (!__all_sync(0xffffffffu, !edge || __float_as_uint(m.z) == orow + 1)) {
                // the neighbour is more than a row behind: poll (warp-uniformly, so that lane 0
                // keeps serving the staging duty)
                uint32_t tries = 0;
                do {
                    if (it >= pwin) {
                        duty_probe();
                        duty_issue();
                    }
                    if (edge) m = lds128_volatile(pa);
#if IVS_WARP_CLOCKS
                    n_mpoll++;
#endif
                    if (++tries > (1u << 22)) __trap();  // a neighbour that never posts is a bug
                } while (!__all_sync(0xffffffffu, !edge || __float_as_uint(m.z) == orow + 1));
            }
            if (edge_l) { Lz[0] = m.x; Lw[0] = m.y; }
            if (edge_r) { Rx[QPT - 1] = m.x; Ry[QPT - 1] = m.y; }
            // Fokker-Planck (src/SM/FokkerPlanckMap.cpp:197-211)
            float o[QPT][4];
#if IVS_PROBE_NOFP  // timing probe (wrong results): no stencil, no neighbours
#pragma unroll
            for (int j = 0; j < QPT; j++) {
                o[j][0] = dp[j].x + Lz[j]; o[j][1] = dp[j].y; o[j][2] = dp[j].z; o[j][3] = dp[j].w + Ry[j];
                gout[j * 32] = make_float4(o[j][0], o[j][1], o[j][2], o[j][3]);
            }
            ps = warp_sum_probe(ps);
#else
#define IVS_FP_COL(j, c)                                                 \
            {                                                            \
                float v = __fmul_rn(in[c], w5[j][c][0]);                 \
                v = madd<FMA>(in[1 + c], w5[j][c][1], v);                \
                v = madd<FMA>(in[2 + c], w5[j][c][2], v);                \
                v = madd<FMA>(in[3 + c], w5[j][c][3], v);                \
                v = madd<FMA>(in[4 + c], w5[j][c][4], v);                \
                o[j][c] = v;                                             \
            }
#pragma unroll
            for (int j = 0; j < QPT; j++) {
                const float in[8] = {Lz[j], Lw[j], dp[j].x, dp[j].y, dp[j].z, dp[j].w, Rx[j], Ry[j]};
                IVS_FP_COL(j, 0)
                if (j == 0) ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 8));
                IVS_FP_COL(j, 1)
                if (j == 0) ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 4));
                IVS_FP_COL(j, 2)
                if (j == 0) ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 2));
                IVS_FP_COL(j, 3)
                if (j == 0) ps = __fadd_rn(ps, __shfl_xor_sync(0xffffffffu, ps, 1));
                gout[j * 32] = make_float4(o[j][0], o[j][1], o[j][2], o[j][3]);
            }
#undef IVS_FP_COL
#endif
            gout += gstride;
            if (red_lane && orow > 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(redp), "f"(ps) : "memory");
            redp += NSEG * 4;
#pragma unroll
            for (int j = 0; j < QPT; j++) {
                float p = __fmul_rn(o[j][0], ws4[j].x);  // src/PS/PhaseSpace.cpp:316-319 (products rounded)
                p = __fadd_rn(p, __fmul_rn(o[j][1], ws4[j].y));
                p = __fadd_rn(p, __fmul_rn(o[j][2], ws4[j].z));
                p = __fadd_rn(p, __fmul_rn(o[j][3], ws4[j].w));
                pprev[j] = p;
            }
        };

        // Iteration `it` kicks row r0 + it; from it = NT-1 on a drifted output row (it - NT + 1) comes
        // out and is parked in FIFO slot it mod kLag; the stencil handles the row parked kLag
        // iterations before (output row it - NT + 1 - kLag, read from the same slot before this
        // iteration's kick overwrites it), so a warp only waits for a neighbour that is more than
        // kLag - 1 rows late.
        auto fifo_at = [&](int i) { return fifo_t + (uint32_t)(i & (kLag - 1)) * (QPT * T * 16); };
        int it = 0;
        for (; it < NT - 1 + kLag; it++) {  // fill the drift sums and the FIFO
            if (it < (int)nkicks) kick_step(it, fifo_at(it), it + 1 >= NT, it + 1 < (int)nkicks);
        }
        for (; it + 1 < (int)nkicks; it++) {  // steady state
            const uint32_t fa = fifo_at(it);
            fp_step(it, fa, (uint32_t)(it - (NT - 1 + kLag)));
            kick_step(it, fa, true, true);
        }
        for (; it < (int)nkicks + kLag; it++) {  // the last kick and the rows still parked
            const uint32_t fa = fifo_at(it);
            fp_step(it, fa, (uint32_t)(it - (NT - 1 + kLag)));
            if (it < (int)nkicks) kick_step(it, fa, true, false);
        }
#if IVS_WARP_CLOCKS
        if (lane == 0) {
            atomicAdd(&g_wstat[0][warp], (unsigned long long)(clock64() - wclk0));
            atomicAdd(&g_wstat[1][warp], n_mpoll);
            atomicAdd(&g_wstat[2][warp], n_rwait);
            atomicAdd(&g_wstat[3][warp], 1ull);
        }
#endif
        // the last row's sums
#pragma unroll
        for (int j = 0; j < QPT; j++) {
            float p = pprev[j];
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, o2));
            if (lane == 0) red[(S - 1) * NSEG + warp * QPT + j] = p;
        }
        // rows this warp still has to issue (its probe may have failed on the last iterations)
        {
            uint32_t tries = 0;
            while (prow < RIN) {
                duty_probe();
                duty_issue();
                if (++tries > (1u << 24)) __trap();
            }
        }
    };
#if IVS_PROBE_NT6
    run(std::integral_constant<int, 6>{});
#else
    if (spread == 0) run(std::integral_constant<int, 4>{});
    else if (spread == 1) run(std::integral_constant<int, 5>{});
    else run(std::integral_constant<int, 6>{});
#endif

    // the strip's row sums: one thread per row adds the segments' partials in fixed order
    __syncthreads();
    for (uint32_t i = tid; i < S; i += T) {
        float v = 0.0f;
        for (uint32_t k = 0; k < NSEG; k++) v = __fadd_rn(v, red[i * NSEG + k]);
        g.rho_out[(size_t)member * n + xa + i] = v;
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
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<TX, kRG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<TX, kRG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
    } else {
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_generic_kernel<TX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_generic_kernel<TX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
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
#if IVS_PROBE_NT6
    plan->halo_hi += 2;  // every warp reads a 6-row window
#endif
    plan->pad_y = kPad;
    // ---- streaming kernel (default for grid sizes that are multiples of 256, up to 2048)
    plan->stream = false;
    // quads of columns per thread: 2 (half the warps, twice the registers each) from 1024 columns on,
    // where a row still gives a CTA at least 4 warps; IVS_SIM_QPT1 / IVS_SIM_QPT2 override
    uint32_t qpt = n >= 1024 ? 2 : 1;
    if (s->desc.flags & IVS_SIM_QPT1) qpt = 1;
    if ((s->desc.flags & IVS_SIM_QPT2) && n >= 256) qpt = 2;
    bool quads_ok = n % 256 == 0 && n <= 2048;
    // over a warp's 128 * qpt columns the first drift tap may vary by at most 2 rows (window of up to 6 rows)
    auto spread_ok = [&](uint32_t cols) {
        for (uint32_t w = 0; w < n / cols; w++) {
            const auto mm = std::minmax_element(dr_first.begin() + cols * w, dr_first.begin() + cols * (w + 1));
            if (*mm.second - *mm.first > 2) return false;
        }
        return true;
    };
    if (quads_ok && qpt == 2 && !spread_ok(256)) qpt = 1;
    quads_ok = quads_ok && spread_ok(128 * qpt);
    if (quads_ok && !(s->desc.flags & IVS_SIM_TILED)) {
        const uint32_t H = plan->halo_lo + plan->halo_hi;
        const uint32_t threads_s = n / 4 / qpt;
        const uint32_t NW = threads_s / 32;  // warps of a CTA
        const uint32_t pitch = n + 2 * kPadS;
        auto a16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
        const bool fma = ctx->arith == IVS_ARITH_FMA;
        const void* kfn = qpt == 2 ? (fma ? (const void*)stream_step_kernel<true, 2> : (const void*)stream_step_kernel<false, 2>)
                                   : (fma ? (const void*)stream_step_kernel<true, 1> : (const void*)stream_step_kernel<false, 1>);
        cudaFuncAttributes fa;
        IVS_CUDA(ctx, cudaFuncGetAttributes(&fa, kfn));
        const uint32_t regs_thread = (std::max(fa.numRegs, 32) + 7) & ~7;
        const uint32_t by_regs = std::max<uint32_t>(1, 65536u / (regs_thread * threads_s));  // CTAs per SM
        // ring depth: the warps' windows start up to H rows apart, plus rows in flight; the
        // mailboxes (kMailSlots deep) bound it from above
        const uint32_t kmin = H + 4, kmax = kMailSlots - kLag - 6;
        // Search over the strip length S0 = ceil(n / P): the per-row tables in shared memory are sized
        // for the strip, so shorter strips leave room for more CTAs per SM (small grids) while longer
        // ones amortise the set-up and the halo rows.  Cost = rounds of CTAs over the machine x (rows
        // through the pipeline per CTA + a set-up worth ~16 rows), scaled by how well the resident
        // warps fill the four schedulers (16 quads per lane and SM saturate this kernel).
        double best = 1e300;
        uint32_t bestP = 0, bestK = 0, best_rin = 0;
        size_t best_sb = 0, best_wwin = 0;
        const uint32_t forceP = s->desc.strips;  // 0: the planner's choice
        for (uint32_t P = 1; P <= n / kMinStrip && kmin <= kmax; P++) {
            if (forceP && P != forceP) continue;
            const uint32_t S0 = ivs_div_up(n, P);
            if ((uint64_t)S0 * (P - 1) >= n || S0 > 128) continue;  // last strip empty / table cap
            // Strips shorter than 8 rows only pay when every strip gets an SM of its own (a single
            // small run: a step is then the latency of set-up + S0 + halo rows; measured 7.4 vs 9.9 us
            // per step at 256^2, 8.2 vs 9.8 at 512^2), not when CTAs share SMs (12.9 vs 11.5 us at 1024^2)
            if (S0 < 8 && (uint64_t)P * s->members > (uint64_t)ctx->sm_count) continue;
            const uint32_t max_rin = S0 + H;
            const uint32_t wcount = n + ((max_rin + 63) & ~63u);
            const uint32_t wwin_len = wcount + (wcount >> 3) + 8;
            const size_t fixed = a16((size_t)(n + n / 32) * 8) + a16((size_t)wwin_len * 8) + a16((size_t)(n / 128) * max_rin * 8) +
                                 a16((size_t)S0 * (n / 128) * 4) + a16((size_t)(max_rin + 1) * 48) +
                                 a16((size_t)NW * 2 * kMailSlots * 16) + a16((size_t)kLag * (n / 4) * 16);
            uint32_t want = std::min<uint32_t>(by_regs, 16), K = 0;
            size_t sb = 0;
            for (; want >= 1; want--) {
                const size_t budget = (size_t)227 * 1024 / want - 1024;  // 1 KB per CTA is reserved by the system
                if (budget < fixed + (size_t)kmin * (pitch * 4 + 16)) continue;
                K = (uint32_t)std::min<size_t>(std::min(kmin + 12, kmax), (budget - fixed) / ((size_t)pitch * 4 + 16));
                sb = a16((size_t)K * pitch * 4) + fixed + 2 * a16((size_t)K * 8);
                if (sb <= budget && K >= kmin) break;
                K = 0;
            }
            if (K < kmin) continue;
            const uint64_t units = (uint64_t)P * s->members;
            const uint64_t slots = (uint64_t)ctx->sm_count * want;
            const uint64_t rounds = (units + slots - 1) / slots;
            const uint32_t resident = (uint32_t)std::min<uint64_t>(want, ivs_div_up((uint32_t)units, ctx->sm_count));
            const double fill = std::min(1.0, resident * (n / 128) / 16.0);  // of 16 segments in flight per SM
            const double cost = (double)rounds * (S0 + H + 16) * resident / std::max(fill, 1.0 / 16);
            if (cost < best) { best = cost; bestP = P; bestK = K; best_sb = sb; best_rin = max_rin; best_wwin = wwin_len; }
        }
        if (bestP) {
            // Fokker-Planck weights with the taps that fall outside the grid zeroed: the shuffles
            // deliver another lane's value there instead of the tile kernels' zero pad
            std::vector<float> fps(fpw);
            fps[0] = fps[1] = 0.0f;
            fps[5] = 0.0f;
            fps[(size_t)(n - 1) * 5 + 3] = fps[(size_t)(n - 1) * 5 + 4] = 0.0f;
            fps[(size_t)(n - 2) * 5 + 4] = 0.0f;
            IVS_CUDA(ctx, cudaMalloc(&plan->fp5, (size_t)n * 20));
            IVS_CUDA(ctx, cudaMemcpy(plan->fp5, fps.data(), (size_t)n * 20, cudaMemcpyHostToDevice));
            plan->stream = true;
            plan->qpt = qpt;
            plan->strip = ivs_div_up(n, bestP);
            plan->kslots = bestK;
            plan->max_strip = plan->strip;
            plan->max_rin = best_rin;
            plan->wwin_len = (uint32_t)best_wwin;
            plan->threads = threads_s;
            plan->smem = best_sb;
            plan->tx = 0;
            IVS_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem));
        }
    }
    // ---- column-block kernel (ivs_stream8.cu): same domain of grid sizes, 8 adjacent columns per lane
    // (the default; IVS_SIM_QPT1 / IVS_SIM_QPT2 select the streaming kernel instead)
    const bool want8 = (s->desc.flags & IVS_SIM_STREAM8) || !(s->desc.flags & (IVS_SIM_QPT1 | IVS_SIM_QPT2));
    if (want8 && !(s->desc.flags & IVS_SIM_TILED) && n % 256 == 0 && n <= 2048) {
        size_t keep_smem = plan->smem;
        uint32_t keep_threads = plan->threads;
        if (ivs_stream8_plan(ctx, s, plan, dr_first.data())) {
            if (!plan->fp5) {
                std::vector<float> fps(fpw);  // out-of-grid taps zeroed, as for the streaming kernel
                fps[0] = fps[1] = 0.0f;
                fps[5] = 0.0f;
                fps[(size_t)(n - 1) * 5 + 3] = fps[(size_t)(n - 1) * 5 + 4] = 0.0f;
                fps[(size_t)(n - 2) * 5 + 4] = 0.0f;
                IVS_CUDA(ctx, cudaMalloc(&plan->fp5, (size_t)n * 20));
                IVS_CUDA(ctx, cudaMemcpy(plan->fp5, fps.data(), (size_t)n * 20, cudaMemcpyHostToDevice));
            }
            plan->stream = false;
        } else {
            plan->smem = keep_smem;
            plan->threads = keep_threads;
        }
    }
    ivs_resident_plan(ctx, s, plan);
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
    if (plan->stream || plan->stream8) { tx = 16; threads = plan->threads; smem = plan->smem; }
    if (tx == 0)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "drift displacement needs more halo rows than fit in shared memory: pass IVS_SIM_UNFUSED");
    if (!plan->stream && !plan->stream8) { plan->tx = tx; plan->threads = threads; plan->smem = smem; }
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
    if (plan->stream || plan->stream8) {
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

ivs_status ivs_fused_plan_set_rf(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s, const float* rf_offset) {
    if (!plan->ready) return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused plan missing");
    const uint32_t n = s->n;
    std::vector<int> rf_first(n);
    std::vector<float4> rf_w(n);
    for (uint32_t i = 0; i < n; i++) {
        const ivs_taps t = ivs_kick_taps(rf_offset[i], n, s->it);  // KickMap::updateSM arithmetic
        if (t.first < -kPad || t.first > kPad - 4)
            return ivs_fail(ctx, IVS_ERR_HALO, "RF kick displacement exceeds the fused kernel's staged pad: pass IVS_SIM_UNFUSED");
        rf_first[i] = t.first;
        rf_w[i] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
    }
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // steps in flight still read the old tables
    IVS_CUDA(ctx, cudaMemcpy(plan->rf_first, rf_first.data(), n * 4, cudaMemcpyHostToDevice));
    IVS_CUDA(ctx, cudaMemcpy(plan->rf_w, rf_w.data(), n * 16, cudaMemcpyHostToDevice));
    return IVS_OK;
}

void ivs_fused_plan_destroy(ivs_fused_plan* plan) {
#if IVS_WARP_CLOCKS
    if (plan->stream) {
        unsigned long long h[4][32];
        if (cudaMemcpyFromSymbol(h, g_wstat, sizeof(h)) == cudaSuccess) {
            for (int w = 0; w < 32; w++)
                if (h[3][w])
                    fprintf(stderr, "[ivs warp %2d] strips %llu: loop end at %.0f clocks, %.1f mailbox polls, %.1f row waits per strip\n", w,
                            h[3][w], (double)h[0][w] / h[3][w], (double)h[1][w] / h[3][w], (double)h[2][w] / h[3][w]);
            memset(h, 0, sizeof(h));
            cudaMemcpyToSymbol(g_wstat, h, sizeof(h));
        }
    }
#endif
    void** bufs[] = {&plan->rf_first, &plan->rf_w, &plan->dr_first, &plan->dr_w, &plan->fp_base, &plan->fp_w, &plan->fp5, &plan->res_tags};
    for (void** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
    }
    plan->ready = false;
}

// Many steps in ONE cooperative launch of the persistent kernel (ivs_resident.cu): for runs that fit
// the machine at once and that have nothing between two steps but the step itself
// (no tracked particles, no DynamicRFKickMap).  The two grids and profiles swap roles inside the kernel.
ivs_status ivs_fused_enqueue_run(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s, uint32_t nsteps) {
    if (!plan->ready || !plan->resident || s->np || s->dyn_rf)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "persistent step kernel does not apply");
    if (nsteps == 0) return IVS_OK;
    fused_args a;
    a.in = s->grid[s->cur];
    a.out = s->grid[s->cur ^ 1];
    a.rho_in = s->profile[s->pcur];
    a.rho_out = s->profile[s->pcur ^ 1];
    a.wake_out = s->wakebuf[0];
    a.wb = 0;
    a.wakefn = s->wakefunction;
    a.wakescaling = s->wakescaling;
    a.rf_first = (const int*)plan->rf_first; a.rf_w = (const float4*)plan->rf_w;
    a.dr_first = (const int*)plan->dr_first; a.dr_w = (const float4*)plan->dr_w;
    a.fp_base = (const int*)plan->fp_base; a.fp_w = (const float4*)plan->fp_w;
    a.fp_w5 = s->fpw;
    a.ws = s->ws;
    a.status = s->status;
    a.host_flag = s->host_flag;
    a.n = s->n; a.it = s->it; a.hl = plan->halo_lo; a.hr = plan->halo_hi; a.pitch = s->n + 2 * kPad;
    a.has_wake = s->has_wake ? 1 : 0;
    const ivs_status st = ivs_resident_run(ctx, plan, s, a, nsteps);
    if (st != IVS_OK) return st;
    s->cur ^= (int)(nsteps & 1u);
    s->pcur ^= (int)(nsteps & 1u);
    return IVS_OK;
}

ivs_status ivs_fused_enqueue_step(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s) {
    if (!plan->ready) return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "fused plan missing");
    fused_args a;
    a.in = s->grid[s->cur];
    a.out = s->grid[s->cur ^ 1];
    a.rho_in = s->profile[s->pcur];
    a.rho_out = s->profile[s->pcur ^ 1];
    // with tracked particles on the side stream the steps alternate between two wake buffers
    const int wb = s->track_overlap ? s->cur : 0;
    a.wake_out = s->wakebuf[wb];
    a.wb = wb;
    if (s->track_inflight[wb]) {  // the tracking launch that still reads this buffer (two steps ago)
        IVS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s->ev_track[wb], 0));
        s->track_inflight[wb] = false;
    }
    a.wakefn = s->wakefunction;
    a.wakescaling = s->wakescaling;
    a.rf_first = (const int*)plan->rf_first; a.rf_w = (const float4*)plan->rf_w;
    a.dr_first = (const int*)plan->dr_first; a.dr_w = (const float4*)plan->dr_w;
    a.fp_base = (const int*)plan->fp_base; a.fp_w = (const float4*)plan->fp_w;
    a.fp_w5 = s->fpw;
    a.ws = s->ws;
    a.status = s->status;
    a.host_flag = s->host_flag;
    a.n = s->n; a.it = s->it; a.hl = plan->halo_lo; a.hr = plan->halo_hi; a.pitch = s->n + 2 * kPad;
    a.has_wake = s->has_wake ? 1 : 0;
    a.train_nb = s->train_windows ? s->members : 0;
    a.train_windows = s->train_windows;
    ivs_status st;
    if (plan->stream8) {
        st = ivs_stream8_launch(ctx, plan, s, a);
        goto launched;
    }
    if (plan->stream) {
        stream_args sa;
        sa.f = a;
        sa.f.pitch = s->n + 2 * kPadS;
        sa.f.fp_w5 = (const float*)plan->fp5;
        sa.strip = plan->strip; sa.kslots = plan->kslots;
        sa.max_strip = plan->max_strip; sa.max_rin = plan->max_rin; sa.wwin_len = plan->wwin_len;
        const dim3 sgrid(ivs_div_up(s->n, plan->strip), s->members);
        const bool fma = ctx->arith == IVS_ARITH_FMA;
        if (plan->qpt == 2) {
            if (fma) stream_step_kernel<true, 2><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
            else stream_step_kernel<false, 2><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        } else {
            if (fma) stream_step_kernel<true, 1><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
            else stream_step_kernel<false, 1><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        }
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
        if (s->track_overlap) {
            // the particles' step runs beside the NEXT fused step (it only needs this step's wake
            // potential and the static tables), on a low-priority stream that fills the SMs the step
            // kernel's last round of CTAs leaves idle
            IVS_CUDA(ctx, cudaEventRecord(s->ev_step, ctx->stream));
            IVS_CUDA(ctx, cudaStreamWaitEvent(s->side, s->ev_step, 0));
            st = ivs_track_step_launch(ctx, s->particles, s->np, s->members, s->n, s->has_wake ? s->wakebuf[wb] : nullptr,
                                       s->rf, s->drift, s->has_fp ? s->fpw : nullptr, nullptr, s->desc.fptrack,
                                       s->desc.e1, s->desc.dp, s->desc.seed, s->track_counter, s->status, wb, s->side);
            if (st != IVS_OK) return st;
            IVS_CUDA(ctx, cudaEventRecord(s->ev_track[wb], s->side));
            s->track_inflight[wb] = true;
        } else {
            st = ivs_track_step_launch(ctx, s->particles, s->np, s->members, s->n, s->has_wake ? s->wakebuf[wb] : nullptr,
                                       s->rf, s->drift, s->has_fp ? s->fpw : nullptr, nullptr, s->desc.fptrack,
                                       s->desc.e1, s->desc.dp, s->desc.seed, s->track_counter, s->status);
        }
    }
    return st;
}

ivs_status ivs_fused_join(ivs_ctx* ctx, ivs_sim* s) {
    for (int b = 0; b < 2; b++) {
        if (s->track_inflight[b]) {
            IVS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s->ev_track[b], 0));
            s->track_inflight[b] = false;
        }
    }
    return IVS_OK;
}
