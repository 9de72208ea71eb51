// ivs_stream8.cu -- the fused simulation step as a streaming kernel of fully independent warps
// with the kicks' displacement absorbed by the staging copy (sm_100a).
//
//   wake potential -> wake kick (y) -> RF kick (y) -> drift (x) -> Fokker-Planck (y) -> X projection
//   (src/main.cpp:948-1053; maps src/SM/KickMap.cpp:221-266, src/SM/FokkerPlanckMap.cpp:197-211,
//    projection src/PS/PhaseSpace.cpp:313-321, wake src/PS/ElectricField.cpp:186-253)
//
// Same arithmetic per cell, in the same order, as stream_step_kernel / the tile kernels of
// ivs_fused.cu and the kernel sequence of ivs_maps.cu (bit for bit, tested); what differs is how the
// data moves.  A CTA owns a strip of consecutive rows x of one member; a warp owns a block of 256
// adjacent columns y of that strip and streams its part of the rows through a ring of its own:
//
//   staging : the two kicks of a row read its cells at  y + f1(x) + f2(x) + 0..6  (first wake tap +
//             first RF tap).  That whole-cell displacement is applied by the copy into shared memory:
//             the warp copies the 264 cells from column  256 w + f1 + f2  on with 4-byte cp.async
//             (LDGSTS.32; taps outside the grid and rows outside the grid are zero-filled by the
//             copy).  The staged row then sits at a FIXED alignment relative to the lanes' columns,
//             for every row: one straight-line kick, no variants per tap alignment, no zero pads.
//             (TMA cannot do this: bulk copies and tensor-map coordinates move multiples of 16 B.)
//             Even and odd quads of the staged cells are kept in two halves of the slot so that a
//             lane's 14 cells (8 columns + 3 + 3 taps) are 3 LDS.128 + 1 LDS.64 without bank conflicts.
//             The ring is private to the warp: cp.async groups + __syncwarp, no mbarrier, no
//             producer role, no slot hand-back between warps.
//   kick    : a lane owns 8 ADJACENT columns: 11 wake-kicked values (44 multiply-adds) feed the 8 RF
//             outputs (32) -- 9.5 per cell against 11 with a quad per lane.
//   drift   : column-local running sums in registers, NT = 4 + spread taps (ivs_fused.cu).
//   hand-over: the drifted row is written to a small per-warp row FIFO in shared memory, again split
//             in even and odd quads.  The stencil stage reads it back kLag rows later in a different
//             thread mapping -- lane l owns quads l and 32 + l of the block, so the two STG.128 of a
//             row are fully coalesced -- and takes the neighbouring quads from the same FIFO row
//             (LDS.128) instead of by shuffles.  Only the two quads beyond the block's ends come from
//             the neighbour warps: their edge lanes post them in a mailbox (sequence stamp = row).
//             The lag of kLag rows between posting and reading is the only coupling between warps.
//   Fokker-Planck, STG.128, Simpson-weighted row sums (same summation tree as everywhere else).
//
// The wake potential of the strip's rows is formed by the whole CTA during the set-up exactly as in
// stream_step_kernel (Toeplitz product in double).
#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "ivs_fused_common.cuh"

namespace {

#ifndef IVS_S8_SLOTS
#define IVS_S8_SLOTS 8
#endif
#ifndef IVS_S8_LAG
#define IVS_S8_LAG 4
#endif
#ifndef IVS_S8_PROBE_NOSTAGE
#define IVS_S8_PROBE_NOSTAGE 0  // timing probes (wrong results): no copies into the ring / no global stores
#endif
#ifndef IVS_S8_PROBE_NOSTORE
#define IVS_S8_PROBE_NOSTORE 0
#endif
#ifndef IVS_S8_STCS
#define IVS_S8_STCS 1  // streaming (evict-first) stores of the output rows: 246.4 us against 247.5 with plain stores
#endif
#ifndef IVS_S8_UNROLL2
#define IVS_S8_UNROLL2 0
#endif
constexpr int kSlots8 = IVS_S8_SLOTS;  // staged rows per warp (ring depth, a power of two)
constexpr int kLag8 = IVS_S8_LAG;      // rows between drift and stencil = depth of the row FIFO (a power of two)
constexpr int kMail8 = 32;             // mailbox depth per warp edge (> 2 kLag8 + 2)
constexpr uint32_t kSlotB = 1152;      // bytes per slot: even quads at 0 (33 quads), odd quads at kOddB (33 quads)
constexpr uint32_t kOddB = 576;        // 144 floats: 16 banks away from the even half

template <int DOFF, int SOFF>
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 4;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
template <int DOFF, int SOFF>
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const float* src, uint32_t bytes) {
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 4, %4;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds32i(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// TRAIN: the members are the bunches of one train (ivs_sim_set_bunch_train): the Toeplitz product of the
// set-up runs once per source bunch against that pair's window of the wake function.  A separate
// instantiation, so that the code of the independent-members kernel (and its register allocation) is untouched.
template <bool FMA, bool TRAIN>
#ifdef IVS_S8_MAXNREG
__global__ void __maxnreg__(IVS_S8_MAXNREG) stream8_step_kernel(
#else
__global__ void __launch_bounds__(256, 1) stream8_step_kernel(
#endif
    const stream_args sa) {
    const fused_args& g = sa.f;
    extern __shared__ __align__(128) float smem[];
    __shared__ int ctl[2];
    const uint32_t n = g.n, nq = g.n >> 2, hl = g.hl;
    const uint32_t H = g.hl + g.hr;
    const uint32_t T = blockDim.x;             // == n / 8
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
    const uint32_t NSEG = NW * 2;              // segments of 128 columns in a row
    const uint32_t member = blockIdx.y;
    const uint32_t xa = blockIdx.x * sa.strip;
    const uint32_t S = min(sa.strip, n - xa);  // output rows of this strip
    const uint32_t RIN = S + H;                // input rows the strip depends on
    const int xs = (int)xa - (int)hl;          // grid row of input row 0

    // shared memory carve-up (every block 16-byte aligned)
    unsigned char* sp = reinterpret_cast<unsigned char*>(smem);
    auto carve = [&sp](size_t bytes) {
        unsigned char* p = sp;
        sp += (bytes + 15) & ~(size_t)15;
        return p;
    };
    unsigned char* ring = carve((size_t)NW * kSlots8 * kSlotB);                  // [warp][slot]
    unsigned char* fifo = carve((size_t)NW * kLag8 * kSlotB);                    // [warp][row mod kLag8]
    double* rho = reinterpret_cast<double*>(carve((size_t)(n + (n >> 5)) * 8));  // last step's profile (skewed)
    double* wwin = reinterpret_cast<double*>(carve((size_t)sa.wwin_len * 8));    // wake function window
    double* partial = reinterpret_cast<double*>(carve((size_t)NSEG * sa.max_rin * 8));  // wake partial sums
    float* red = reinterpret_cast<float*>(carve((size_t)sa.max_strip * NSEG * 4));       // row-sum partials
    // per input row: wake taps, RF taps, (f1 + f2, f2, -, -)
    float4* tab = reinterpret_cast<float4*>(carve((size_t)(sa.max_rin + 1) * 48));
    float4* mail = reinterpret_cast<float4*>(carve((size_t)NW * 2 * kMail8 * 16));  // [warp][side][slot]

    const float* gin = g.in + (size_t)member * n * n;

    if (tid == 0) step_ctl_begin(g, ctl);
    // ---- per-thread constants, requested early.  Drift stage: columns cb + 8 lane + 0..7
    const uint32_t cb = warp * 256;
    const uint32_t ycol = cb + 8 * lane;
    int dfs[8];
    float4 dwl[8];
    {
        const int4 a = __ldg(reinterpret_cast<const int4*>(g.dr_first + ycol));
        const int4 b = __ldg(reinterpret_cast<const int4*>(g.dr_first + ycol) + 1);
        dfs[0] = a.x; dfs[1] = a.y; dfs[2] = a.z; dfs[3] = a.w;
        dfs[4] = b.x; dfs[5] = b.y; dfs[6] = b.z; dfs[7] = b.w;
#pragma unroll
        for (int c = 0; c < 8; c++) dwl[c] = __ldg(&g.dr_w[ycol + c]);
    }
    // stencil stage: the same 8 columns (the lane's own drifted values never leave the lane)
    float w5[8][5];
    float ws8[8];
    {
        const float4* wp = reinterpret_cast<const float4*>(g.fp_w5 + (size_t)ycol * 5);
        float t[40];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const float4 v = __ldg(wp + k);
            t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int k = 0; k < 5; k++) w5[c][k] = t[c * 5 + k];
        const float4 wa = __ldg(reinterpret_cast<const float4*>(g.ws + ycol));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(g.ws + ycol) + 1);
        ws8[0] = wa.x; ws8[1] = wa.y; ws8[2] = wa.z; ws8[3] = wa.w;
        ws8[4] = wb.x; ws8[5] = wb.y; ws8[6] = wb.z; ws8[7] = wb.w;
    }
    // last step's profile and the first batch of the wake-function window are requested in the same
    // round trip (T = n / 8 threads: 8 profile entries and 10 window entries each).
    // w[x - x'] is stored at table index x - x' + n - 1; window entry t holds the table entry
    // xs + t, so row r and x' meet at t = r - x' + n - 1
    constexpr int NRV = 8, NWV = 10;
    const int wtop = (int)(2 * n - 1);
    const uint32_t tcount = n + ((RIN + 63) & ~63u);
    float rv[NRV];
    double wv0[NWV];
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
    // an earlier step of this run left the supported range of displacements (the host redoes it on
    // another path and re-enqueues what follows): nothing may be written any more
    if (ctl[0] != kStatusClean) return;

    // ---- set-up: mailboxes, per-row tables, last step's profile and the wake-function
    // window in shared memory
    for (uint32_t i = tid; i < NW * 2 * kMail8; i += T) mail[i] = make_float4(0, 0, 0, 0);  // seq 0 = nothing posted
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
                wv[k] = __ldg(g.wakefn + min(max(xs + (int)(t0 + k * T), 0), wtop));
#pragma unroll
            for (int k = 0; k < NWV; k++) {
                const uint32_t t = t0 + k * T;
                if (t < tcount) wwin[wwin_pos(t)] = wv[k];
            }
        }
    }
    __syncthreads();
    const uint32_t nsrc = TRAIN ? g.train_nb : 1u;  // source bunches of the wake potential
#pragma unroll 1
    for (uint32_t src = 0; src < nsrc; src++) {
    if (TRAIN && g.has_wake) {
        // the profile of source bunch `src` and the window of the pair (member, src) replace what the
        // set-up staged (src/PS/ElectricField.cpp:236-263: all bunches of the padded train act on each one)
        __syncthreads();  // the previous pair's product has read rho and wwin
        const float* rin = g.rho_in + (size_t)src * n;
        const double* wf = g.train_windows + ((size_t)member * g.train_nb + src) * 2 * n;
        // (T = n / 8 threads: 8 profile entries each, requested together; the window in batches of 8)
        float rr[8];
#pragma unroll
        for (int k = 0; k < 8; k++) rr[k] = __ldg(rin + tid + k * T);
        for (uint32_t t0 = tid; t0 < tcount; t0 += 8 * T) {
            double ww[8];
#pragma unroll
            for (int k = 0; k < 8; k++) ww[k] = __ldg(wf + min(max(xs + (int)(t0 + k * T), 0), wtop));
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (t0 + k * T < tcount) wwin[wwin_pos(t0 + k * T)] = ww[k];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t i = tid + k * T;
            rho[i + (i >> 5)] = (double)rr[k];
        }
        __syncthreads();
    }
    if (g.has_wake) {
        // W[x] = scaling * sum_x' rho[x'] w[x-x']  (src/PS/ElectricField.cpp:186-253 as a Toeplitz
        // product).  Segment s sums x' in [128 s, 128 s + 128); a lane owns 8 consecutive rows and 32
        // consecutive x', so the 8 wake-function values it needs slide through registers.
        const uint32_t rbk = lane & 7, sub = lane >> 3;
#pragma unroll 1
        for (uint32_t seg = warp * 2; seg < (warp + 1) * 2; seg++) {
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
#ifdef IVS_S8_WAKE_F32
                // TIMING PROBE ONLY (not the reference's numbers to the last bit): the 32-term inner sums in fp32,
                // i.e. with the accuracy of the reference's own fp32 transforms; the outer sums stay in double
                {
                    float wf[8], af[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { wf[j] = (float)wv[j]; af[j] = 0.0f; }
#pragma unroll 8
                    for (uint32_t k = 0; k < 32; k++) {
                        const float rh = (float)rp[k];
#pragma unroll
                        for (int j = 0; j < 8; j++) af[j] = __fmaf_rn(rh, wf[j], af[j]);
#pragma unroll
                        for (int j = 7; j > 0; j--) wf[j] = wf[j - 1];
                        wf[0] = (float)wwin[wwin_pos(max((int)(t0 - k) - 1, 0))];
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = (double)af[j];
                }
#else
#pragma unroll 8
                for (uint32_t k = 0; k < 32; k++) {
                    const double rh = rp[k];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = __fma_rn(rh, wv[j], acc[j]);
#pragma unroll
                    for (int j = 7; j > 0; j--) wv[j] = wv[j - 1];
                    wv[0] = wwin[wwin_pos(max((int)(t0 - k) - 1, 0))];
                }
#endif
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
                    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
                }
                if (sub == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (r0 + j < RIN) {
                            double* pp = partial + (size_t)seg * sa.max_rin + r0 + j;
                            *pp = (TRAIN && src) ? *pp + acc[j] : acc[j];  // (the lane that wrote it adds to it)
                        }
                }
            }
        }
    }
    }  // source bunches
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
                    // beyond the range the fused kernels agree on: record the step (the host redoes it
                    // through the kernel sequence, which has no such limit) and keep the addresses in range
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
    }
    __syncthreads();  // tables and mailboxes visible to everybody

    // =============================== the row loop ==============================================
    int wfmin = 0x7fffffff, wfmax = -0x7fffffff;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        wfmin = min(wfmin, dfs[c]);
        wfmax = max(wfmax, dfs[c]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wfmin = min(wfmin, __shfl_xor_sync(0xffffffffu, wfmin, o));
        wfmax = max(wfmax, __shfl_xor_sync(0xffffffffu, wfmax, o));
    }
    const int spread = wfmax - wfmin;  // <= 2 (checked when the plan was made)

    const uint32_t tab_b = smem_u32(tab);
    const uint32_t ring_w = smem_u32(ring) + warp * (kSlots8 * kSlotB);
    const uint32_t fifo_w = smem_u32(fifo) + warp * (kLag8 * kSlotB);
    const uint32_t mail_b = smem_u32(mail);
    // mailboxes: lane 0 posts (x, y) of the block's first quad for the left neighbour, lane 31 (z, w) of
    // its last quad for the right neighbour
    const uint32_t post_me = mail_b + (warp * 2 + (lane == 0 ? 0 : 1)) * kMail8 * 16;
    const bool edge_l = lane == 0 && warp > 0, edge_r = lane == 31 && warp + 1 < NW;
    const bool edge = edge_l || edge_r;
    // the two cells beyond each end of the block come from the neighbour warps' posts (zeros beyond the grid)
    const uint32_t peek = edge_l   ? mail_b + ((warp - 1) * 2 + 1) * kMail8 * 16
                          : edge_r ? mail_b + ((warp + 1) * 2 + 0) * kMail8 * 16
                                   : mail_b;
    const bool edge_warp = warp == 0 || warp + 1 == NW;
    const bool strip_edge = xs < 0 || xs + (int)RIN > (int)n;  // some of the strip's input rows lie outside the grid
    // staging: this lane's cell of every group of 32 staged cells (see the header comment)
    const uint32_t st_dst = ((lane >> 2) & 1) * kOddB + ((lane >> 3) * 4 + (lane & 3)) * 4;
    float4* gout = reinterpret_cast<float4*>(g.out + ((size_t)member * n + xa) * n) + (cb >> 2) + 2 * lane;

    auto run = [&](auto nt_tag) {
        constexpr int NT = decltype(nt_tag)::value;
        float dw[8][NT];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const float wv[4] = {dwl[c].x, dwl[c].y, dwl[c].z, dwl[c].w};
            const int off = dfs[c] - wfmin;  // 0 .. NT-4
#pragma unroll
            for (int t = 0; t < NT; t++) {
                float v = 0.0f;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (t - u >= 0 && t - u <= NT - 4) v = (off == t - u) ? wv[u] : v;
                dw[c][t] = v;
            }
        }
        // rows r0 .. r0+NT-1 make output row 0; kick number k (k = 0, 1, ...) reads input row r0 + k
        const uint32_t r0 = (uint32_t)((int)hl + wfmin);
        const int nkicks = (int)(S + NT - 1);

        // ---- staging of input row r0 + k into slot k mod kSlots8, one cp.async group per call.
        // Cursors: the next row's cells (this lane's first), its table entry, its grid row; the
        // whole-cell displacement f1 + f2 of the next row is fetched a call ahead.
        const float* srow = gin + ((long long)(xs + (int)r0) * (long long)n + (long long)(cb + lane));
        const uint32_t stab = tab_b + r0 * 48 + 32;
        int sx = xs + (int)r0;
        int fs_next = lds32i(stab);
        if (strip_edge) {  // stale shared memory must not meet the zero weights as NaN * 0
            for (uint32_t i = lane; i < kSlots8 * kSlotB / 16; i += 32)
                sts128(ring_w + i * 16, 0.0f, 0.0f, 0.0f, 0.0f);
            __syncwarp();
        }
        auto stage = [&](int k) {
            const int fs = fs_next;
            const uint32_t dst = ring_w + ((uint32_t)k & (kSlots8 - 1)) * kSlotB + st_dst;
            const float* src = srow + fs;
            // rows outside the grid (first and last strip of a member) are not copied: their RF taps are
            // zeros and the slots of such a strip were cleared during the set-up
            if (!IVS_S8_PROBE_NOSTAGE && (!strip_edge || (uint32_t)sx < n)) {
                // cells outside the grid are zeros (src-size 0: nothing is read).  Only the first 32 cells
                // can lie left of the grid and only the last 64 right of it (|f1 + f2| <= 32), and only in
                // the grid's first and last block; the three guards are the same code for every block so
                // that all warps of a CTA keep the same pace.
                const int col0 = (int)(cb + lane) + fs;
                cp_async4_zfill<0, 0>(dst, src, col0 >= 0 ? 4u : 0u);
                cp_async4<64, 128>(dst, src);
                cp_async4<128, 256>(dst, src);
                cp_async4<192, 384>(dst, src);
                cp_async4<256, 512>(dst, src);
                cp_async4<320, 640>(dst, src);
                cp_async4<384, 768>(dst, src);
                cp_async4_zfill<448, 896>(dst, src, col0 + 224 < (int)n ? 4u : 0u);
                cp_async4_zfill<512, 1024>(dst, src, col0 + 256 < (int)n ? 4u : 0u);  // 8 of these 32 cells are needed
            }
            cp_async_commit();
            srow += n;
            sx++;
            fs_next = lds32i(stab + (uint32_t)(k + 1) * 48);  // the table has an entry beyond the last row
        };

        // drift as running sums: P[c][j] is the output row that has taken j + 1 of its NT taps so far
        float P[8][NT - 1];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int t = 0; t < NT - 1; t++) P[c][t] = 0.0f;
        float pprev[2] = {0.0f, 0.0f};  // this lane's part of the sums of the row the stencil handled last (its two quads)
        // The row sums follow the same tree per 128-column segment as everywhere else: over the
        // segment's 32 quads, partner quad index xor 16, 8, 4, 2, 1.  Here quad Q of a segment sits in
        // lane Q / 2 of a half-warp (first or second quad of the lane), so the partners are lanes xor 8,
        // 4, 2, 1 and finally the lane's other quad.  The first level also transposes: lanes with bit 3
        // clear go on with the first quads' sums, the others with the second quads', which makes the
        // last level one more exchange with lane xor 8.  The five levels are spread over five
        // iterations, one level each for five different rows: every iteration issues five independent
        // shuffle + add pairs, so no add ever waits for its shuffle.
        // lvl[k] = the sum after k + 1 levels of the row handled k + 2 stencil steps ago.
        float lvl[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        // where the finished sums go: lane 0 holds segment 2 warp's, lane 16 the next one's
        uint32_t redp = smem_u32(red) + (warp * 2 + (lane >= 16 ? 1 : 0)) * 4 - 5 * NSEG * 4;
        const bool red_lane = (lane & 15) == 0;
        const size_t gstride = (size_t)nq;
        auto rowsum_step = [&]() {
            const bool second = (lane & 8) != 0;
            const float send = second ? pprev[0] : pprev[1];
            const float keep = second ? pprev[1] : pprev[0];
            const float t1 = __shfl_xor_sync(0xffffffffu, send, 8);
            const float t2 = __shfl_xor_sync(0xffffffffu, lvl[0], 4);
            const float t3 = __shfl_xor_sync(0xffffffffu, lvl[1], 2);
            const float t4 = __shfl_xor_sync(0xffffffffu, lvl[2], 1);
            const float t5 = __shfl_xor_sync(0xffffffffu, lvl[3], 8);
            const float done = __fadd_rn(lvl[3], t5);
            lvl[3] = __fadd_rn(lvl[2], t4);
            lvl[2] = __fadd_rn(lvl[1], t3);
            lvl[1] = __fadd_rn(lvl[0], t2);
            lvl[0] = __fadd_rn(keep, t1);
            return done;
        };
        // What crosses from one iteration to the next: the wake-kicked cells k1 of the row the next
        // kick works on (k1[i] = cell ycol + f2 + i of that row after the wake kick,
        // src/SM/KickMap.cpp:243-266, 0 outside the grid) and that row's RF taps.  Its staged cells and
        // taps are requested in the middle of an iteration and turned into k1 at its end.
        float k1[11];
        float4 w2 = make_float4(0, 0, 0, 0);
        float4 w1n, w2n;
        float in[14];
        int f2row = -1;  // first RF tap of that row (only the grid's first and last block look at it)
        auto load_row = [&](const int k) {
            const uint32_t tp = tab_b + (r0 + (uint32_t)k) * 48;
            w1n = lds128(tp);
            w2n = lds128(tp + 16);
            if (edge_warp) f2row = lds32i(tp + 36);
            const uint32_t sl = ring_w + ((uint32_t)k & (kSlots8 - 1)) * kSlotB + lane * 16;
            const float4 a0 = lds128(sl), b0 = lds128(sl + kOddB), a1 = lds128(sl + 16);
            const float2 b1 = lds64f(sl + kOddB + 16);
            in[0] = a0.x; in[1] = a0.y; in[2] = a0.z; in[3] = a0.w;
            in[4] = b0.x; in[5] = b0.y; in[6] = b0.z; in[7] = b0.w;
            in[8] = a1.x; in[9] = a1.y; in[10] = a1.z; in[11] = a1.w;
            in[12] = b1.x; in[13] = b1.y;
        };
        auto wake_kick = [&]() {
#pragma unroll
            for (int i = 0; i < 11; i++) k1[i] = tap4<FMA>(in[i], in[i + 1], in[i + 2], in[i + 3], w1n);
            w2 = w2n;
            if (edge_warp) {
                const int f2 = f2row;
                const int m0 = (int)ycol + f2;
                if (NW > 1 && f2 >= -3 && f2 <= 1) {
                    // only the block's first 3 (left edge of the grid) or last 4 (right edge) can lie outside
                    if (warp == 0) {
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            if (m0 + i < 0) k1[i] = 0.0f;
                    } else {
#pragma unroll
                        for (int i = 7; i < 11; i++)
                            if (m0 + i >= (int)n) k1[i] = 0.0f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 11; i++)
                        if ((uint32_t)(m0 + i) >= n) k1[i] = 0.0f;
                }
            }
        };

        // One iteration.  Iteration `it` kicks input row r0 + it (KICK); from it = NT-1 on a drifted
        // output row (it - NT + 1) comes out and is parked in FIFO row it mod kLag8; the stencil (FP)
        // handles the row parked kLag8 iterations before (output row it - NT + 1 - kLag8, read from the
        // same FIFO row before this iteration's kick overwrites it), so a warp only waits for a
        // neighbour that is more than kLag8 - 1 rows late.  Shared-memory loads run ahead of their
        // use: the stencil's are requested before the kick's arithmetic, the next kick's after this one's.
        auto iter = [&](const int it, auto fp_tag, auto kick_tag, auto stage_tag) {
            constexpr bool FP = decltype(fp_tag)::value, KICK = decltype(kick_tag)::value;
            constexpr bool STAGE = decltype(stage_tag)::value;
            const uint32_t frow = fifo_w + ((uint32_t)it & (kLag8 - 1)) * kSlotB;
            const uint32_t orow = (uint32_t)(it - (NT - 1 + kLag8));  // the stencil's output row
            if (KICK) cp_async_wait<kSlots8 - 3>();  // row r0 + it + 1 has landed (this lane's part of it)
            __syncwarp();  // ... and everybody's; the FIFO row written an iteration ago is visible; slot it-1 is free
            if (STAGE) stage(it + kSlots8 - 1);
            else if (KICK) cp_async_commit();  // keeps the group count in step
            // the stencil's row: the lane's own 8 drifted values out of its FIFO entry (written by this
            // lane kLag8 iterations ago), the edge lanes' neighbours out of the mailbox
            float4 fa, fb, mb;
            uint32_t pa = 0;
            if (FP) {
                fa = lds128(frow + lane * 16);
                fb = lds128(frow + kOddB + lane * 16);
                pa = peek + ((orow * 16) & ((kMail8 - 1) * 16));
                mb = make_float4(0, 0, 0, 0);
                if (edge) mb = lds128_volatile(pa);
            }
            // one level of the row sums' butterfly for each of the last five rows (see lvl above)
            float ps = 0.0f;
            if (FP) ps = rowsum_step();
            bool arrived = true;
            if (KICK) {
                const uint32_t prow = (uint32_t)(it - NT + 1);  // the output row this kick completes
                const float seq = __uint_as_float(prow + 1);
                const uint32_t post_a = post_me + (prow & (kMail8 - 1)) * 16;
                // has the neighbours' post for the stencil's row arrived?  Asked here, used after the drift
                if (FP) arrived = __all_sync(0xffffffffu, !edge || __float_as_uint(edge_l ? mb.x : mb.z) == orow + 1);
                float kq[8];
#pragma unroll
                for (int c = 0; c < 8; c++) kq[c] = tap4<FMA>(k1[c], k1[c + 1], k1[c + 2], k1[c + 3], w2);  // RF kick, same loop
                load_row(it + 1);  // the registers are free again: the next row's cells and taps
                float d[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    // src/SM/KickMap.cpp:221-242, column-local 4-tap along x
                    d[c] = madd<FMA>(kq[c], dw[c][NT - 1], P[c][NT - 2]);
#pragma unroll
                    for (int t = NT - 2; t > 0; t--) P[c][t] = madd<FMA>(kq[c], dw[c][t], P[c][t - 1]);
                    P[c][0] = __fmul_rn(kq[c], dw[c][0]);
                }
                // parked in the row FIFO until its stencil step kLag8 iterations later
                sts128(frow + lane * 16, d[0], d[1], d[2], d[3]);
                sts128(frow + kOddB + lane * 16, d[4], d[5], d[6], d[7]);
                if (it + 1 >= NT) {  // a real output row: its ends are posted for the neighbour warps
                    const bool l0 = lane == 0;
                    sts128_volatile_if(post_a,
                                       make_float4(l0 ? d[0] : seq, l0 ? d[1] : 0.0f, l0 ? seq : d[6], l0 ? 0.0f : d[7]),
                                       l0 || lane == 31);
                }
                wake_kick();  // of the row just loaded, for the next iteration
            }
            if (FP) {
                const uint32_t want = orow + 1;
                if (!KICK) arrived = __all_sync(0xffffffffu, !edge || __float_as_uint(edge_l ? mb.x : mb.z) == want);
                if (!arrived) {
                    // the neighbour is more than kLag8 - 1 rows behind: poll
                    uint32_t tries = 0;
                    do {
                        if (edge) mb = lds128_volatile(pa);
                        if (++tries > (1u << 22)) __trap();  // a neighbour that never posts is a bug
                    } while (!__all_sync(0xffffffffu, !edge || __float_as_uint(edge_l ? mb.x : mb.z) == want));
                }
                // the two cells left and right of the lane's 8: the neighbour lanes', the posts at the
                // block's ends (left post = (seq, -, z, w), right post = (x, y, seq, -)), zeros beyond the grid
                float lz = __shfl_up_sync(0xffffffffu, fb.z, 1), lw = __shfl_up_sync(0xffffffffu, fb.w, 1);
                float rx = __shfl_down_sync(0xffffffffu, fa.x, 1), ry = __shfl_down_sync(0xffffffffu, fa.y, 1);
                if (lane == 0) { lz = mb.z; lw = mb.w; }
                if (lane == 31) { rx = mb.x; ry = mb.y; }
                // Fokker-Planck (src/SM/FokkerPlanckMap.cpp:197-211)
                const float fin[12] = {lz, lw, fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w, rx, ry};
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float v = __fmul_rn(fin[c], w5[c][0]);
                    v = madd<FMA>(fin[1 + c], w5[c][1], v);
                    v = madd<FMA>(fin[2 + c], w5[c][2], v);
                    v = madd<FMA>(fin[3 + c], w5[c][3], v);
                    v = madd<FMA>(fin[4 + c], w5[c][4], v);
                    o[c] = v;
                }
#if IVS_S8_PROBE_NOSTORE
                if (o[0] == 123.456f) gout[0] = make_float4(o[0], o[1], o[2], o[3]);
                if (o[4] == 123.456f) gout[1] = make_float4(o[4], o[5], o[6], o[7]);
#else
#if IVS_S8_STCS
                __stcs(gout, make_float4(o[0], o[1], o[2], o[3]));
                __stcs(gout + 1, make_float4(o[4], o[5], o[6], o[7]));
#else
                gout[0] = make_float4(o[0], o[1], o[2], o[3]);
                gout[1] = make_float4(o[4], o[5], o[6], o[7]);
#endif
#endif
                gout += gstride;
                if (red_lane && orow >= 5) asm volatile("st.shared.f32 [%0], %1;" ::"r"(redp), "f"(ps) : "memory");
                redp += NSEG * 4;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    float p = __fmul_rn(o[4 * j], ws8[4 * j]);  // src/PS/PhaseSpace.cpp:316-319 (products rounded)
                    p = __fadd_rn(p, __fmul_rn(o[4 * j + 1], ws8[4 * j + 1]));
                    p = __fadd_rn(p, __fmul_rn(o[4 * j + 2], ws8[4 * j + 2]));
                    p = __fadd_rn(p, __fmul_rn(o[4 * j + 3], ws8[4 * j + 3]));
                    pprev[j] = p;
                }
            }
        };
        const std::true_type yes{};
        const std::false_type no{};
        // prologue: the first kSlots8 - 1 rows are requested, the first one is loaded
#pragma unroll 1
        for (int k = 0; k < kSlots8 - 1; k++) {
            if (k < nkicks) stage(k);
            else cp_async_commit();
        }
        cp_async_wait<kSlots8 - 2>();
        __syncwarp();
        load_row(0);
        wake_kick();
        const int n_stage = nkicks - (kSlots8 - 1);  // iterations that still have a row to request
        int it = 0;
#pragma unroll 1
        for (; it < NT - 1 + kLag8 && it < nkicks; it++) {  // fill the drift sums and the FIFO
            if (it < n_stage) iter(it, no, yes, yes);
            else iter(it, no, yes, no);
        }
#if IVS_S8_UNROLL2
#pragma unroll 1
        for (; it + 1 < n_stage; it += 2) {  // steady state, two rows per trip
            iter(it, yes, yes, yes);
            iter(it + 1, yes, yes, yes);
        }
#endif
#pragma unroll 1
        for (; it < n_stage; it++) iter(it, yes, yes, yes);  // steady state
#pragma unroll 1
        for (; it < nkicks; it++) iter(it, yes, yes, no);    // the last rows are in flight or staged
#pragma unroll 1
        for (; it < nkicks + kLag8; it++) {                  // the rows still parked
            if (it >= NT - 1 + kLag8) iter(it, yes, no, no);
        }
        // the five rows still in the butterfly
#pragma unroll 1
        for (int e = 0; e < 5; e++) {
            const float tot = rowsum_step();
            pprev[0] = pprev[1] = 0.0f;
            const int row = (int)S - 5 + e;
            if (red_lane && row >= 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(redp), "f"(tot) : "memory");
            redp += NSEG * 4;
        }
    };
    if (spread == 0) run(std::integral_constant<int, 4>{});
    else if (spread == 1) run(std::integral_constant<int, 5>{});
    else run(std::integral_constant<int, 6>{});

    // the strip's row sums: one thread per row adds the segments' partials in fixed order
    __syncthreads();
    for (uint32_t i = tid; i < S; i += T) {
        float v = 0.0f;
        for (uint32_t k = 0; k < NSEG; k++) v = __fadd_rn(v, red[i * NSEG + k]);
        g.rho_out[(size_t)member * n + xa + i] = v;
    }
}

size_t stream8_fixed_smem(uint32_t n, uint32_t S0, uint32_t H, uint32_t* wwin_len_out) {
    auto a16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const uint32_t NW = n / 256, NSEG = n / 128;
    const uint32_t max_rin = S0 + H;
    const uint32_t wcount = n + ((max_rin + 63) & ~63u);
    const uint32_t wwin_len = wcount + (wcount >> 3) + 8;
    if (wwin_len_out) *wwin_len_out = wwin_len;
    return a16((size_t)NW * kSlots8 * kSlotB) + a16((size_t)NW * kLag8 * kSlotB) + a16((size_t)(n + n / 32) * 8) +
           a16((size_t)wwin_len * 8) + a16((size_t)NSEG * max_rin * 8) + a16((size_t)S0 * NSEG * 4) +
           a16((size_t)(max_rin + 1) * 48) + a16((size_t)NW * 2 * kMail8 * 16) + 16;
}

}  // namespace

// Launch plan of the column-block kernel: strip length search with the cost model of the streaming
// kernel (ivs_fused.cu).  Returns false when the kernel does not apply (grid size, drift spread).
bool ivs_stream8_plan(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan, const int* dr_first) {
    const uint32_t n = s->n;
    if (n % 256 != 0 || n > 2048) return false;
    // over a warp's 256 columns the first drift tap may vary by at most 2 rows (window of up to 6 rows)
    for (uint32_t w = 0; w < n / 256; w++) {
        const auto mm = std::minmax_element(dr_first + 256 * w, dr_first + 256 * (w + 1));
        if (*mm.second - *mm.first > 2) return false;
    }
    const uint32_t H = plan->halo_lo + plan->halo_hi;
    const uint32_t threads = n / 8;
    const bool fma = ctx->arith == IVS_ARITH_FMA;
    const void* kfn = fma ? (const void*)stream8_step_kernel<true, false> : (const void*)stream8_step_kernel<false, false>;
    const void* kfn_train = fma ? (const void*)stream8_step_kernel<true, true> : (const void*)stream8_step_kernel<false, true>;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kfn) != cudaSuccess) return false;
    const uint32_t regs_thread = (std::max(fa.numRegs, 32) + 7) & ~7;
    const uint32_t by_regs = std::max<uint32_t>(1, 65536u / (regs_thread * threads));  // CTAs per SM
    double best = 1e300;
    uint32_t bestP = 0, best_rin = 0, best_wwin = 0;
    size_t best_sb = 0;
    const uint32_t forceP = s->desc.strips;  // 0: the planner's choice
    for (uint32_t P = 1; P <= n / 2; P++) {
        if (forceP && P != forceP) continue;
        const uint32_t S0 = ivs_div_up(n, P);
        if ((uint64_t)S0 * (P - 1) >= n || S0 > 128) continue;  // last strip empty / table cap
        if (S0 < 8 && (uint64_t)P * s->members > (uint64_t)ctx->sm_count) continue;
        uint32_t wwin_len = 0;
        const size_t sb = stream8_fixed_smem(n, S0, H, &wwin_len);
        uint32_t want = std::min<uint32_t>(by_regs, 16);
        for (; want >= 1; want--)
            if (sb <= (size_t)227 * 1024 / want - 1024) break;
        if (want < 1) continue;
        const uint64_t units = (uint64_t)P * s->members;
        const uint64_t slots = (uint64_t)ctx->sm_count * want;
        const uint64_t rounds = (units + slots - 1) / slots;
        const uint32_t resident = (uint32_t)std::min<uint64_t>(want, ivs_div_up((uint32_t)units, ctx->sm_count));
        const double fill = std::min(1.0, resident * (n / 128) / 16.0);
        const double cost = (double)rounds * (S0 + H + 16) * resident / std::max(fill, 1.0 / 16);
        if (cost < best) { best = cost; bestP = P; best_sb = sb; best_rin = S0 + H; best_wwin = wwin_len; }
    }
    if (!bestP) return false;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem) != cudaSuccess) return false;
    if (cudaFuncSetAttribute(kfn_train, cudaFuncAttributeMaxDynamicSharedMemorySize, kIvsMaxOptinSmem) != cudaSuccess) return false;
    plan->stream8 = true;
    plan->strip = ivs_div_up(n, bestP);
    plan->max_strip = plan->strip;
    plan->max_rin = best_rin;
    plan->wwin_len = best_wwin;
    plan->threads = threads;
    plan->smem = best_sb;
    plan->kslots = kSlots8;
    return true;
}

ivs_status ivs_stream8_launch(ivs_ctx* ctx, const ivs_fused_plan* plan, const ivs_sim* s, const fused_args& a) {
    stream_args sa;
    sa.f = a;
    sa.f.fp_w5 = (const float*)plan->fp5;
    sa.strip = plan->strip; sa.kslots = plan->kslots;
    sa.max_strip = plan->max_strip; sa.max_rin = plan->max_rin; sa.wwin_len = plan->wwin_len;
    const dim3 sgrid(ivs_div_up(s->n, plan->strip), s->members);
    if (sa.f.train_nb) {
        if (ctx->arith == IVS_ARITH_FMA) stream8_step_kernel<true, true><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
        else stream8_step_kernel<false, true><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
    } else if (ctx->arith == IVS_ARITH_FMA) {
        stream8_step_kernel<true, false><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
    } else {
        stream8_step_kernel<false, false><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
    }
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
