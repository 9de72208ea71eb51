// ivs_resident.cu -- the simulation loop as ONE persistent kernel for runs that fit the machine at once
// (single simulations and small ensembles up to 1024 x 1024: the state is L2-resident and a step is
// bound by latency, not by bandwidth) (sm_100a).
//
//   for (step ...) { wake potential -> wake kick (y) -> RF kick (y) -> drift (x) -> Fokker-Planck (y)
//                    -> X projection }                                   (src/main.cpp:948-1053)
//
// The step kernels of ivs_fused.cu / ivs_stream8.cu stream the rows of a strip through one warp per
// column block: the right shape when the grid has to come from HBM, the wrong one when a whole
// 256 x 256 run is two rows per SM and every serial row costs a pipeline turn (8.9 us per step,
// however short the strip).  Here a CTA (one per SM, cooperative launch) keeps a block of rows for the
// whole run and does every map as one block-wide pass over ALL its rows at once:
//
//   boundary | wait for every block's row sums of the last step (8-byte words carrying the step's tag:
//            | the profile IS the grid-wide barrier, see publish_row_sum) -> profile in shared memory
//   sync     | the wake potential (Toeplitz product in double: one chain of 32 x' per lane, 4 NSEG lanes
//            | per row, a row's taps by its first lane); beside it the warps without a row of it wait for
//            | the neighbouring blocks' "rows are out" flags and stage the drift halo rows (cp.async.cg, 16 B:
//            | L2 -> shared memory; the block's own rows are already there)
//   sync     | wake kick + RF kick, a thread per quad of cells: 4 aligned LDS.128, shifted by the row's
//            | displacement mod 4, 7 wake-kicked cells -> 4 RF-kicked
//   sync     | drift + Fokker-Planck, a thread per quad, a warp per 128 columns of a row: 4 LDS.128 of kicked
//            | rows, stencil neighbours by shuffle, STG.128 + the row in place for the next step,
//            | Simpson-weighted sums by warp butterfly
//   sync     | row sums of the segments added in order -> published (relaxed); one thread releases the
//            | "rows are out" flag
//
// The tables of the maps, the window of the wake function and the zero pads live in shared memory for the
// whole run; a warp keeps one segment of 128 columns, so a lane's drift and Fokker-Planck weights are
// fetched once per step.
//
// Per cell the arithmetic is that of the other step kernels operation by operation (tap4 / madd in the
// reference's order, the same summation trees for the wake potential and the row sums), so a run of k
// steps equals k launches of any of them bit for bit (tests/test_gpu_persistent.py).
//
// Measured (B200, tools/single_runs.py): 3.2 / 4.2 / 8.3 us per step at 256^2 / 512^2 / 1024^2 against
// 8.9 / 9.6 / 10.9 us with one launch of the column-block kernel per step (CUDA graph).  Where the time
// goes at 256^2 (clocks of one CTA, -DIVS_RES_CLOCKS=1): boundary 1400 - 1800, wake + taps + halo copies 2100,
// kicks 550, drift + Fokker-Planck 1250, row sums 700 (profiles/r2_experiments.txt has the road there).
#include <algorithm>
#include <cstdio>
#include <vector>

#include "ivs_fused_common.cuh"

namespace {

constexpr int kResThreads = 512;
constexpr int kTagStride = 8;  // published row sums sit 64 B apart: the polls of 148 CTAs spread over the L2 slices (1, 4, 16 measured slower)
constexpr uint32_t kMaxPolls = 1u << 24;  // about 10 s of polling: a step takes microseconds
constexpr int kResPad = 32;  // zero cells left and right of a staged row: wake first tap in [-16, 8], RF in [-16, 12]

struct resident_args {
    fused_args f;
    uint32_t rows;      // output rows per CTA
    uint32_t max_rin;   // rows + drift halo
    uint32_t wwin_len;  // doubles in the staged wake-function window
    uint32_t ord0;      // ordinal (since the host's last reset of the status words) of the run's first step
    uint32_t tickets_per_step;  // CTAs of a step-kernel launch of the same simulation (see step_ctl_begin)
    uint2* tags;        // [2][members][n] published row sums (value, step tag), zeroed before the launch
    uint2* cta_tags;    // [2][CTAs] one word per CTA and step (ensembles: members do not read each other's row sums)
    uint32_t* rows_out; // [CTAs] step count up to which a CTA's rows are in L2 (kTagStride words apart)
};

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// The step boundary.  Two things of step k + 1 wait for step k: the wake potential needs EVERY block's row
// sums, the drift the NEIGHBOURING blocks' rows.  Instead of a counter barrier followed by a read of the new
// profile, the profile IS the barrier: a row sum is published as one 8-byte word (value, step tag) with a
// relaxed store (value and tag travel together, nothing else hangs on it); every CTA polls the n words of
// the step it waits for until all carry the step's tag -- one trip through L2 instead of three -- and goes
// on to the wake potential.  The rows take the slower road beside it: publish_rows_out / await_rows_out,
// release / acquire, awaited only by the warps that copy the halo rows, just before they do.
// Two tag arrays alternate (a CTA can be one step ahead of another, never two: it needs every CTA's row
// sums of a step to leave it; by then every CTA's copies of that step's input rows have landed, so the
// next step may overwrite them).
__device__ __forceinline__ void publish_row_sum(uint2* slot, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// "the rows of step `tag - 1` of this CTA are in L2": a release store by one thread after the CTA barrier
// that follows the stores (its 1500 - 2000 cycles are nobody's critical path); the readers of those rows
// (the neighbouring blocks) poll it with acquire loads before they copy
__device__ __forceinline__ void publish_rows_out(uint32_t* flag, uint32_t tag) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(tag) : "memory");
}
__device__ __forceinline__ void await_rows_out(const uint32_t* flag, uint32_t tag) {
    uint32_t t;
    // acquire loads: a relaxed poll followed by an acquire fence costs 1300 cycles more per step (the fence
    // holds back the warp's next shared-memory accesses as well)
    uint32_t tries = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(t) : "l"(flag) : "memory");
        if (++tries > kMaxPolls) __trap();
    } while (t < tag);
}
// bit 31 of the tag: the publishing CTA's wake kick left the staged pad in that step (every CTA reads every
// word of a step, so all of them learn it at the same boundary and end the run together)
__device__ __forceinline__ float await_row_sum(const uint2* slot, uint32_t tag, uint32_t& flags) {
    // relaxed polls (acquire polls measured the same, sleeping between polls 10 % slower); the caller's
    // acquire fence after its last word orders what follows
    uint32_t v, t, tries = 0;
    do {
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(slot) : "memory");
        if (++tries > kMaxPolls) __trap();  // a block that never publishes is a bug: fail, do not hang
    } while ((t & 0x7fffffffu) != tag);
    flags |= t;
    return __uint_as_float(v);
}

// staged wake-function window: the lanes of the wake product read entries that are multiples of 32 apart
// (blocks of 32 x'); one pad double after every 32 puts 16 consecutive blocks on 16 different bank pairs
__device__ __forceinline__ uint32_t wres_pos(uint32_t t) { return t + (t >> 5); }

#ifndef IVS_RES_CLOCKS
#define IVS_RES_CLOCKS 0  // development aid: SM clocks per phase of CTA 0, printed at the end of the run
#endif
#if IVS_RES_CLOCKS
#define RES_MARK(k) do { if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0 && tid == 0) { const long long c_ = clock64(); clk[k] += c_ - last; last = c_; } } while (0)
#else
#define RES_MARK(k) do { } while (0)
#endif

// wake kick + RF kick of one quad of cells (src/SM/KickMap.cpp:243-266 twice): v[SH + k] = staged cell
// y + f1 + f2 + k, k = 0..9; the 4 RF outputs read the wake-kicked cells y + f2 + 0..6 (0 outside the grid)
template <bool FMA, int SH>
__device__ __forceinline__ float4 kick_quad(const float* v, const float4 w1, const float4 w2, int c0, uint32_t n, bool edge) {
    float k1[7];
#pragma unroll
    for (int k = 0; k < 7; k++) k1[k] = tap4<FMA>(v[SH + k], v[SH + k + 1], v[SH + k + 2], v[SH + k + 3], w1);
    if (edge) {
#pragma unroll
        for (int k = 0; k < 7; k++)
            if ((uint32_t)(c0 + k) >= n) k1[k] = 0.0f;
    }
    float4 o;
    o.x = tap4<FMA>(k1[0], k1[1], k1[2], k1[3], w2);
    o.y = tap4<FMA>(k1[1], k1[2], k1[3], k1[4], w2);
    o.z = tap4<FMA>(k1[2], k1[3], k1[4], k1[5], w2);
    o.w = tap4<FMA>(k1[3], k1[4], k1[5], k1[6], w2);
    return o;
}

template <bool FMA>
__global__ void __launch_bounds__(kResThreads, 1) resident_run_kernel(const resident_args ra, const uint32_t nsteps) {
    const fused_args& g = ra.f;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint32_t ovf;  // a wake kick of this CTA left the staged pad (bit 31)
    const uint32_t n = g.n, nq = n >> 2, hl = g.hl;
    const uint32_t H = g.hl + g.hr;
    constexpr uint32_t T = kResThreads;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t NW = T / 32;
    const uint32_t NSEG = n >> 7;  // segments of 128 columns in a row
    const uint32_t member = blockIdx.y;
    const uint32_t xa = blockIdx.x * ra.rows;
    const uint32_t S = min(ra.rows, n - xa);  // output rows of this block
    const uint32_t RIN = S + H;               // input rows it depends on
    const int xs = (int)xa - (int)hl;         // grid row of input row 0
    const uint32_t pitchA = n + 2 * kResPad;

    unsigned char* sp = reinterpret_cast<unsigned char*>(smem);
    auto carve = [&sp](size_t bytes) {
        unsigned char* p = sp;
        sp += (bytes + 15) & ~(size_t)15;
        return p;
    };
    float* A = reinterpret_cast<float*>(carve((size_t)ra.max_rin * pitchA * 4));  // staged input rows, zero pads
    float* K = reinterpret_cast<float*>(carve((size_t)ra.max_rin * n * 4));       // rows after the two kicks
    float4* drw = reinterpret_cast<float4*>(carve((size_t)n * 16));  // drift taps: column y at (y & 3) * nq + (y >> 2)
    int* drf = reinterpret_cast<int*>(carve((size_t)n * 4));
    float* fp5 = reinterpret_cast<float*>(carve((size_t)n * 20));                 // Fokker-Planck weights per column
    float* wsm = reinterpret_cast<float*>(carve((size_t)n * 4));                  // Simpson weights
    double* rho = reinterpret_cast<double*>(carve((size_t)(n + (n >> 5)) * 8));   // last step's profile (skewed)
    double* wwin = reinterpret_cast<double*>(carve((size_t)ra.wwin_len * 8));     // wake function window
    float* red = reinterpret_cast<float*>(carve((size_t)ra.rows * NSEG * 4));     // row-sum partials [row][segment]
    // per input row: wake taps, RF taps, (first wake + RF tap, first RF tap, row inside the grid, -)
    float4* tab = reinterpret_cast<float4*>(carve((size_t)ra.max_rin * 48));

    // ---- once per run: zero pads and rows outside the grid, the maps' tables, the wake-function window
    for (uint32_t i = tid; i < ra.max_rin * pitchA; i += T) A[i] = 0.0f;
    for (uint32_t i = tid; i < ra.max_rin * n; i += T) K[i] = 0.0f;
    for (uint32_t i = tid; i < n; i += T) {
        drw[(i & 3) * nq + (i >> 2)] = __ldg(&g.dr_w[i]);
        drf[i] = __ldg(&g.dr_first[i]);
        wsm[i] = __ldg(&g.ws[i]);
    }
    for (uint32_t i = tid; i < n * 5; i += T) fp5[i] = __ldg(&g.fp_w5[i]);
    for (uint32_t r = tid; r < RIN; r += T) {
        const int x = xs + (int)r;
        const bool valid = (uint32_t)x < n;
        const int f2 = valid ? __ldg(&g.rf_first[x]) : -1;
        // no wake map: the identity as a 4-tap row (0, 1, 0, 0) at first = -1
        tab[3 * r] = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
        tab[3 * r + 1] = valid ? __ldg(&g.rf_w[x]) : make_float4(0, 0, 0, 0);
        tab[3 * r + 2] = make_float4(__int_as_float(f2 - 1), __int_as_float(f2), valid ? 1.0f : 0.0f, 0.0f);
    }
    // w[x - x'] is stored at table index x - x' + n - 1; window entry t holds the table entry xs + t, so
    // row r and x' meet at t = r - x' + n - 1
    const int wtop = (int)(2 * n - 1);
    const uint32_t tcount = n + ((RIN + 63) & ~63u);
    if (g.has_wake)
        for (uint32_t t = tid; t < tcount; t += T)
            wwin[wres_pos(t)] = __ldg(g.wakefn + min(max(xs + (int)t, 0), wtop));  // clamped entries belong to rows outside the grid
    const float scal = g.has_wake ? __ldg(&g.wakescaling[member]) : 0.0f;
    const uint32_t A_b = smem_u32(A);
    // wake potential: a row's sum over x' is spread over 4 NSEG lanes (a chain of 32 x' each)
    const uint32_t LPR = NSEG * 4, RPW = 32 / LPR;  // lanes per row, rows per warp and pass
    const uint32_t wrow = lane / LPR, wli = lane - wrow * LPR;
    // kicks, drift, Fokker-Planck: a warp keeps one segment of 128 columns (a lane one quad) and takes every
    // (NW / NSEG)-th row
    const uint32_t segw = warp % NSEG, grp = warp / NSEG, ngrp = NW / NSEG;
    const uint32_t q = segw * 32 + lane;
    uint2* const tags = ra.tags + (size_t)member * n * kTagStride;  // [2][members][n] (value, step tag), one per line
    const size_t tag_stride = (size_t)gridDim.y * n * kTagStride;
    // ensembles: a CTA only reads its own member's row sums, but a step that leaves the staged pad in one
    // member has to end the run for all of them at the same step: every CTA also publishes one word per
    // step, and every CTA reads all of those
    const uint32_t n_ctas = gridDim.x * gridDim.y, cta_id = blockIdx.y * gridDim.x + blockIdx.x;
    // the blocks that own this block's halo rows
    const uint32_t nb_lo = (uint32_t)max(xs, 0) / ra.rows, nb_hi = min((uint32_t)(xs + (int)RIN) - 1u, n - 1u) / ra.rows;
    // the copies are dealt to the warps that form no row of the wake potential when there are at least four
    // of them (they reach the copies first), else to all warps
    const uint32_t wake_warps = g.has_wake ? min((RIN + RPW - 1) / RPW, NW) : 0u;
    const uint32_t copy_warp0 = NW - wake_warps >= 4 ? wake_warps : 0u, copy_warps = NW - copy_warp0;
    const uint32_t copy_first = warp >= copy_warp0 ? 0u : 0xffffff00u;  // warps before copy_warp0 take no copies
    const bool ensemble = gridDim.y > 1;
    // a step enqueued before this run left the supported range of displacements (the host redoes it on
    // another path and re-enqueues what follows): nothing may be written.  (status[0] = ordinal of the first
    // such step: one of THIS run is at least ord0, so every CTA decides the same whenever it looks.)
    uint32_t steps_run = 0;
    const bool dead = *reinterpret_cast<volatile int*>(g.status) < (int)ra.ord0;
    if (tid == 0) ovf = 0;
    __syncthreads();  // pads and tables in place before the first copies land
#if IVS_RES_CLOCKS
    long long clk[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, last = clock64();
#endif

#pragma unroll 1
    for (uint32_t step = 0; step < nsteps && !dead; step++) {
        steps_run = step + 1;
        // this step's buffers (they swap roles every step)
        const bool odd = (step & 1u) != 0;
        const float* gin = (odd ? g.out : g.in) + (size_t)member * n * n;
        float* const gout = (odd ? const_cast<float*>(g.in) : g.out) + (size_t)member * n * n;
        float* const rho_dst = (odd ? const_cast<float*>(g.rho_in) : g.rho_out) + (size_t)member * n;
        // ---- the step boundary: last step's profile, which is also the barrier (see publish_row_sum)
        uint32_t flags = 0;
        if (step == 0) {
            for (uint32_t i = tid; i < n; i += T) rho[i + (i >> 5)] = (double)__ldcg(g.rho_in + (size_t)member * n + i);
        } else {
            const uint2* tsrc = tags + (size_t)(odd ? 0 : 1) * tag_stride;  // step - 1 published with parity (step - 1) & 1
            for (uint32_t i = tid; i < n; i += T) rho[i + (i >> 5)] = (double)await_row_sum(tsrc + (size_t)i * kTagStride, step, flags);
            if (ensemble) {
                const uint2* csrc = ra.cta_tags + (size_t)(odd ? 0 : 1) * n_ctas * kTagStride;
                for (uint32_t i = tid; i < n_ctas; i += T) await_row_sum(csrc + (size_t)i * kTagStride, step, flags);
            }
        }
        // a step that left the staged pad ends the run for everybody; the host redoes it and what follows
        // (ivs_step.cu, settle)
        const bool over = __syncthreads_or((int)(flags >> 31)) != 0;
        RES_MARK(0);
        if (over) {
            steps_run = step;
            break;
        }
        // ---- wake potential and wake-kick taps of the block's rows while the copies fly
        // W[x] = scaling * sum_x' rho[x'] w[x-x']  (src/PS/ElectricField.cpp:186-253 as a Toeplitz product),
        // summed as in the step kernels: chains of 32 x', blocks of a segment of 128 as (b0 + b1) + (b2 + b3),
        // segments in order
        if (g.has_wake) {
            for (uint32_t r0 = warp * RPW; r0 < RIN; r0 += NW * RPW) {
                const uint32_t r = r0 + wrow;
                const uint32_t xb = (wli >> 2) * 128 + (wli & 3) * 32;
                double acc = 0.0;
                if (wrow < RPW) {
                    const double* rp = rho + xb + (xb >> 5);
                    const uint32_t t0 = r + n - 1 - xb;
#pragma unroll 8
                    for (uint32_t k = 0; k < 32; k++) acc = __fma_rn(rp[k], wwin[wres_pos(t0 - k)], acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                double tot = 0.0;
                for (uint32_t sg = 0; sg < NSEG; sg++) tot += __shfl_sync(0xffffffffu, acc, wrow * LPR + 4 * sg);
                const int x = xs + (int)r;
                if (wrow < RPW && wli == 0 && r < RIN && (uint32_t)x < n) {
                    const float wk = __fmul_rn(scal, (float)tot);  // src/PS/ElectricField.cpp:236-241
                    ivs_taps t = ivs_kick_taps(wk, n, g.it);
                    if (t.first < -16 || t.first > 8) {
                        // beyond the range the fused kernels agree on: record the step (the host redoes it
                        // through the kernel sequence, which has no such limit) and keep the addresses in range
                        atomicMin(g.status, (int)(ra.ord0 + step));
                        *reinterpret_cast<volatile int*>(g.host_flag) = 1;
                        __threadfence_system();
                        *reinterpret_cast<volatile uint32_t*>(&ovf) = 0x80000000u;
                        t.first = -1;
                        t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
                    }
                    if (r >= hl && r < hl + S) g.wake_out[(size_t)member * n + x] = wk;
                    const int f2 = __float_as_int(tab[3 * r + 2].y);
                    tab[3 * r] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
                    tab[3 * r + 2].x = __int_as_float(t.first + f2);
                }
            }
        }
        // ---- stage the rows (16-byte copies, L2 only: the rows were written by other SMs a step ago); a
        // warp copies one segment of 128 cells.  After the first step only the drift halo comes from L2 (the
        // block's own rows were put in place by the last step's Fokker-Planck pass), once the blocks that
        // own those rows have said that they are out.  The warps that had no row of the wake potential to
        // form get here first.
        if (step && warp >= copy_warp0) {
            for (uint32_t b = nb_lo + lane; b <= nb_hi; b += 32)
                if (b != blockIdx.x) await_rows_out(ra.rows_out + (size_t)(blockIdx.y * gridDim.x + b) * kTagStride, step);
            __syncwarp();
        }
        for (uint32_t w = copy_first + (warp + NW - copy_warp0) % NW; w < RIN * NSEG; w += copy_warps) {
            const uint32_t r = w / NSEG, qq = (w - r * NSEG) * 32 + lane;
            const int x = xs + (int)r;
            if ((uint32_t)x < n && (step == 0 || r < hl || r >= hl + S))
                cp_async16_cg(A_b + (r * pitchA + kResPad + 4 * qq) * 4, gin + (size_t)x * n + 4 * qq);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // rows staged, taps known
        RES_MARK(1);

        // ---- wake kick + RF kick, one thread per quad of cells: the staged cells from y + f1 + f2 on are
        // fetched as aligned quads and shifted by (f1 + f2) mod 4
        if (grp < ngrp) {
            for (uint32_t r = grp; r < RIN; r += ngrp) {
                const float4 m = tab[3 * r + 2];
                if (m.z == 0.0f) continue;  // rows outside the grid stay zeros
                const float4 w1 = tab[3 * r], w2 = tab[3 * r + 1];
                const int fs = __float_as_int(m.x), f2 = __float_as_int(m.y);
                const int sh = fs & 3;
                const float4* a = reinterpret_cast<const float4*>(A + (size_t)r * pitchA + kResPad + (fs - sh)) + q;
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 t = a[k];
                    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                }
                const int c0 = (int)(4 * q) + f2;  // column of the first wake-kicked cell
                const int s0 = (int)(segw * 128) + f2;
                const bool edge = s0 < 0 || s0 + 130 >= (int)n;  // some of the segment's wake-kicked cells lie outside the grid
                float4 o;
                if (sh == 0) o = kick_quad<FMA, 0>(v, w1, w2, c0, n, edge);
                else if (sh == 1) o = kick_quad<FMA, 1>(v, w1, w2, c0, n, edge);
                else if (sh == 2) o = kick_quad<FMA, 2>(v, w1, w2, c0, n, edge);
                else o = kick_quad<FMA, 3>(v, w1, w2, c0, n, edge);
                reinterpret_cast<float4*>(K + (size_t)r * n)[q] = o;
            }
        }
        __syncthreads();
        RES_MARK(2);
        // ---- drift (src/SM/KickMap.cpp:221-242) and Fokker-Planck (src/SM/FokkerPlanckMap.cpp:197-211), one
        // thread per quad of cells.  The stencil's neighbours come by shuffle; the two cells beyond each end
        // of the segment are drifted once more by the end lanes.
        if (grp < ngrp) {
            // the quad's tables, once per step
            const int4 f = reinterpret_cast<const int4*>(drf)[q];
            const float4 dw0 = drw[q], dw1 = drw[nq + q], dw2 = drw[2 * nq + q], dw3 = drw[3 * nq + q];
            const float4 sw = reinterpret_cast<const float4*>(wsm)[q];
            float wq[20];
            {
                const float4* wp = reinterpret_cast<const float4*>(fp5 + (size_t)q * 20);
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float4 t = wp[k];
                    wq[4 * k] = t.x; wq[4 * k + 1] = t.y; wq[4 * k + 2] = t.z; wq[4 * k + 3] = t.w;
                }
            }
            const bool uniform = __all_sync(0xffffffffu, f.x == f.y && f.y == f.z && f.z == f.w);
            for (uint32_t j = grp; j < S; j += ngrp) {
                const float* krow = K + (size_t)(hl + j) * n;  // the kicked row `first` = 0 points at
                auto drift_cell = [&](uint32_t y) {
                    const float* k0 = krow + (ptrdiff_t)drf[y] * (ptrdiff_t)n + y;
                    return tap4<FMA>(k0[0], k0[n], k0[2 * n], k0[3 * n], drw[(y & 3) * nq + (y >> 2)]);
                };
                float d[4];
                if (uniform) {
                    const float4* k0 = reinterpret_cast<const float4*>(krow + (ptrdiff_t)f.x * (ptrdiff_t)n) + q;
                    const float4 a0 = k0[0], a1 = k0[nq], a2 = k0[2 * nq], a3 = k0[3 * nq];
                    d[0] = tap4<FMA>(a0.x, a1.x, a2.x, a3.x, dw0);
                    d[1] = tap4<FMA>(a0.y, a1.y, a2.y, a3.y, dw1);
                    d[2] = tap4<FMA>(a0.z, a1.z, a2.z, a3.z, dw2);
                    d[3] = tap4<FMA>(a0.w, a1.w, a2.w, a3.w, dw3);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) d[c] = drift_cell(4 * q + c);
                }
                float lz = __shfl_up_sync(0xffffffffu, d[2], 1), lw = __shfl_up_sync(0xffffffffu, d[3], 1);
                float rx = __shfl_down_sync(0xffffffffu, d[0], 1), ry = __shfl_down_sync(0xffffffffu, d[1], 1);
                if (lane == 0 || lane == 31) {  // beyond the grid: any finite value (those taps are zero)
                    const uint32_t y0 = lane ? 4 * q + 4 : 4 * q - 2;
                    const bool inside = lane ? q + 1 < nq : q > 0;
                    const float e0 = inside ? drift_cell(y0) : 0.0f, e1 = inside ? drift_cell(y0 + 1) : 0.0f;
                    if (lane) { rx = e0; ry = e1; } else { lz = e0; lw = e1; }
                }
                const float fin[8] = {lz, lw, d[0], d[1], d[2], d[3], rx, ry};
                float o[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float t = __fmul_rn(fin[c], wq[5 * c]);
                    t = madd<FMA>(fin[1 + c], wq[5 * c + 1], t);
                    t = madd<FMA>(fin[2 + c], wq[5 * c + 2], t);
                    t = madd<FMA>(fin[3 + c], wq[5 * c + 3], t);
                    t = madd<FMA>(fin[4 + c], wq[5 * c + 4], t);
                    o[c] = t;
                }
                reinterpret_cast<float4*>(gout + (size_t)(xa + j) * n)[q] = make_float4(o[0], o[1], o[2], o[3]);
                // ... and in place as this block's input row of the next step (the kicks have read it)
                reinterpret_cast<float4*>(A + (size_t)(hl + j) * pitchA + kResPad)[q] = make_float4(o[0], o[1], o[2], o[3]);
                // Simpson-weighted sum over the segment (src/PS/PhaseSpace.cpp:316-319; same tree as everywhere
                // else: partner quad xor 16, 8, 4, 2, 1)
                float p = __fmul_rn(o[0], sw.x);
                p = __fadd_rn(p, __fmul_rn(o[1], sw.y));
                p = __fadd_rn(p, __fmul_rn(o[2], sw.z));
                p = __fadd_rn(p, __fmul_rn(o[3], sw.w));
#pragma unroll
                for (int x = 16; x > 0; x >>= 1) p = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, x));
                if (lane == 0) red[j * NSEG + segw] = p;
            }
        }
        __syncthreads();
        if (tid == 32 && step + 1 < nsteps) publish_rows_out(ra.rows_out + (size_t)cta_id * kTagStride, step + 1);
        RES_MARK(3);
        // ---- the block's row sums: segments added in order; published for the next step (value + step tag,
        // released after this CTA's rows: the barrier above orders them before it)
        for (uint32_t i = tid; i < S; i += T) {
            float v = 0.0f;
            for (uint32_t k = 0; k < NSEG; k++) v = __fadd_rn(v, red[i * NSEG + k]);
            rho_dst[xa + i] = v;
            if (step + 1 < nsteps)
                publish_row_sum(tags + (size_t)(odd ? 1 : 0) * tag_stride + (size_t)(xa + i) * kTagStride, v, (step + 1) | *reinterpret_cast<volatile uint32_t*>(&ovf));
        }
        if (ensemble && tid == 0 && step + 1 < nsteps)
            publish_row_sum(ra.cta_tags + ((size_t)(odd ? 1 : 0) * n_ctas + cta_id) * kTagStride, 0.0f,
                            (step + 1) | *reinterpret_cast<volatile uint32_t*>(&ovf));
        RES_MARK(4);
    }
    // the step kernels count the steps by CTA tickets (step_ctl_begin): leave the counter as they would have
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
        atomicAdd(reinterpret_cast<unsigned int*>(g.status) + 1, steps_run * ra.tickets_per_step);
#if IVS_RES_CLOCKS
    if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0 && tid == 0 && nsteps > 100)
        printf("[ivs resident n=%u rows=%u rin=%u cta %u] clocks/step: boundary+rho %lld load+wake+taps %lld kick %lld drift+fp %lld rowsum %lld\n",
               n, S, RIN, blockIdx.x, clk[0] / nsteps, clk[1] / nsteps, clk[2] / nsteps, clk[3] / nsteps, clk[4] / nsteps);
#endif
}

size_t resident_smem(uint32_t n, uint32_t rows, uint32_t H, uint32_t* wwin_len_out) {
    auto a16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const uint32_t NSEG = n / 128, max_rin = rows + H;
    const uint32_t wcount = n + ((max_rin + 63) & ~63u);
    const uint32_t wwin_len = wcount + (wcount >> 3) + 8;
    if (wwin_len_out) *wwin_len_out = wwin_len;
    return a16((size_t)max_rin * (n + 2 * kResPad) * 4) + a16((size_t)max_rin * n * 4) +
           a16((size_t)n * 16) + a16((size_t)n * 4) + a16((size_t)n * 20) + a16((size_t)n * 4) +
           a16((size_t)(n + n / 32) * 8) + a16((size_t)wwin_len * 8) +
           a16((size_t)rows * NSEG * 4) + a16((size_t)max_rin * 48) + 16;
}

// row sums [2][members][n] + one word per CTA [2][at most a CTA per SM], kTagStride words apart
size_t resident_tag_bytes(const ivs_ctx* ctx, uint32_t members, uint32_t n) {
    return ((size_t)2 * members * n + (size_t)3 * ctx->sm_count) * kTagStride * sizeof(uint2);
}

const void* resident_kernel(const ivs_ctx* ctx) {
    return ctx->arith == IVS_ARITH_FMA ? (const void*)resident_run_kernel<true> : (const void*)resident_run_kernel<false>;
}

}  // namespace

// Does the run fit the machine at once?  One CTA per SM, equal blocks of rows per member.
bool ivs_resident_plan(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan) {
    plan->resident = false;
    const uint32_t n = s->n;
    // a step kernel that was asked for by name (or direct launches) is what runs
    const uint32_t named = IVS_SIM_NO_PERSISTENT | IVS_SIM_TILED | IVS_SIM_QPT1 | IVS_SIM_QPT2 | IVS_SIM_STREAM8 | IVS_SIM_NO_GRAPH;
    if ((s->desc.flags & named) || n % 128 != 0 || n > 1024 || !plan->fp5) return false;
    if (s->members > (uint32_t)ctx->sm_count) return false;
    int coop = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device) != cudaSuccess || !coop) return false;
    const uint32_t H = plan->halo_lo + plan->halo_hi;
    const uint32_t blocks = std::min<uint32_t>((uint32_t)ctx->sm_count / s->members, n);  // per member
    const uint32_t rows = ivs_div_up(n, blocks);
    // (what bounds the block of rows is shared memory: rows + halo staged and kicked rows of n cells each)
    uint32_t wwin_len = 0;
    const size_t sb = resident_smem(n, rows, H, &wwin_len);
    if (sb > (size_t)kIvsMaxOptinSmem) return false;
    const void* kfn = resident_kernel(ctx);
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem) != cudaSuccess) return false;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kResThreads, sb) != cudaSuccess || per_sm < 1) return false;
    if (!plan->res_tags && cudaMalloc(&plan->res_tags, resident_tag_bytes(ctx, s->members, n)) != cudaSuccess) return false;
    plan->resident = true;
    plan->res_rows = rows;
    plan->res_smem = sb;
    plan->res_wwin = wwin_len;
    return true;
}

// nsteps steps in one cooperative launch; tags = device buffer of 2 * members * n (value, tag) words
ivs_status ivs_resident_run(ivs_ctx* ctx, const ivs_fused_plan* plan, const ivs_sim* s,
                            const ivs_fused_detail::fused_args& a, uint32_t nsteps) {
    resident_args ra;
    ra.f = a;
    ra.f.fp_w5 = (const float*)plan->fp5;
    ra.rows = plan->res_rows;
    ra.max_rin = plan->res_rows + plan->halo_lo + plan->halo_hi;
    ra.wwin_len = plan->res_wwin;
    ra.ord0 = (uint32_t)s->fused_enqueued;
    ra.tickets_per_step = plan->stream8 || plan->stream ? ivs_div_up(s->n, plan->strip) * s->members
                                                        : ivs_div_up(s->n, plan->tx) * s->members;
    ra.tags = (uint2*)plan->res_tags;
    ra.cta_tags = ra.tags + (size_t)2 * s->members * s->n * kTagStride;
    ra.rows_out = reinterpret_cast<uint32_t*>(ra.cta_tags + (size_t)2 * ctx->sm_count * kTagStride);
    const dim3 grid(ivs_div_up(s->n, plan->res_rows), s->members);
    IVS_CUDA(ctx, cudaMemsetAsync(plan->res_tags, 0, resident_tag_bytes(ctx, s->members, s->n), ctx->stream));
    void* args[] = {(void*)&ra, (void*)&nsteps};
    const cudaError_t e = cudaLaunchCooperativeKernel(resident_kernel(ctx), grid, dim3(kResThreads), args, plan->res_smem, ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
        // the device cannot hold the whole grid right now (shared with other work, MPS limits ...): not an
        // error, the caller goes on with one launch per step
        cudaGetLastError();
        return IVS_ERR_UNSUPPORTED;
    }
    IVS_CUDA(ctx, e);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
