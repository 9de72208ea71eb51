// ivs_stream8ws.cu -- the column-block step kernel (ivs_stream8.cu) with the two halves of a row's
// work on different warps (sm_100a, grids of 1024 ... 2048 columns).
//
// Same data flow and the same arithmetic as stream8_step_kernel (bit for bit, tested); a block of
// 256 adjacent columns of a strip is served by TWO warps instead of one:
//
//   kick warp    stages the block's cells (4-byte cp.async with the kicks' whole-cell displacement
//                applied by the copy, warp-private ring), wake kick + RF kick on 8 adjacent columns
//                per lane, drift as running sums, drifted row -> the block's row FIFO in shared memory;
//   stencil warp reads the FIFO row in the quad-per-lane mapping (neighbour quads from the same row,
//                the two quads beyond the block's ends from the neighbour blocks' mailboxes),
//                Fokker-Planck, STG.128, Simpson-weighted row sums.
//
// A FIFO row changes hands through a pair of mbarriers (full: the kick warp's lane 0 arrives after the
// warp's stores; empty: the block's stencil warp arrives after its loads), which also carry the
// release / acquire ordering between the two warps of a block.  Blocks only meet in the mailboxes
// (kMailW rows deep, sequence stamp = row, as in ivs_stream8.cu): a stencil warp needs the two
// neighbour kick warps to have posted its row, nothing ever waits for another block's stencil warp.
// A kick warp runs at most kDepthW rows ahead of its own stencil warp and that one never ahead of the
// neighbours' kick warps, so a post is read long before its mailbox entry comes round again
// (kMailW > 2 kDepthW).
//
// Why: stream8_step_kernel needs 255 registers per thread, i.e. 8 warps per SM, two per scheduler,
// and is bound by the latency of each warp's own instruction stream (51 % of the issue slots used).
// The kick half needs ~150 registers, the stencil half ~100; with setmaxnreg the 16 warps of a CTA
// split the register file 152 : 104, every scheduler gets two warps of each kind (4 instead of 2),
// and the two halves of a row overlap in time.
#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "ivs_fused_common.cuh"

namespace {

#ifndef IVS_WS_SLOTS
#define IVS_WS_SLOTS 4
#endif
#ifndef IVS_WS_DEPTH
#define IVS_WS_DEPTH 8
#endif
#ifndef IVS_WS_KREGS
#define IVS_WS_KREGS 152
#endif
#ifndef IVS_WS_FREGS
#define IVS_WS_FREGS 104
#endif
constexpr int kSlotsW = IVS_WS_SLOTS;  // staged rows per kick warp (ring depth, a power of two)
constexpr int kDepthW = IVS_WS_DEPTH;  // rows of a block's FIFO (a power of two)
constexpr int kMailW = 32;             // mailbox depth per block edge (> 2 kDepthW)
constexpr uint32_t kSlotB = 1152;      // bytes per slot / FIFO row: even quads at 0 (33 quads), odd quads at kOddB
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
template <int REGS>
__device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
// blocking wait on an mbarrier phase; the warp is suspended by the hardware (up to the hint, in ns) instead
// of spinning in the issue slots of the warps it waits for (a row that never comes is a bug: trap
// instead of hanging the GPU)
#ifndef IVS_WS_HINT
#define IVS_WS_HINT 2000
#endif
#ifndef IVS_WS_SLEEP
#define IVS_WS_SLEEP 100
#endif
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t phase) {
    uint32_t tries = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(bar), "r"(phase), "r"((uint32_t)IVS_WS_HINT)
            : "memory");
        if (ok) break;
        if (++tries > (1u << 20)) __trap();
    }
}

template <bool FMA>
__global__ void __launch_bounds__(512, 1) stream8ws_step_kernel(const stream_args sa) {
    const fused_args& g = sa.f;
    extern __shared__ __align__(128) float smem[];
    __shared__ int ctl[2];
    const uint32_t n = g.n, nq = g.n >> 2, hl = g.hl;
    const uint32_t H = g.hl + g.hr;
    const uint32_t T = blockDim.x;             // == n / 4
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t NB = n >> 8;                // blocks of 256 columns = kick warps = stencil warps
    const uint32_t NSEG = NB * 2;              // segments of 128 columns in a row (== warps of the CTA)
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
    unsigned char* ring = carve((size_t)NB * kSlotsW * kSlotB);                  // [block][slot]
    unsigned char* fifo = carve((size_t)NB * kDepthW * kSlotB);                  // [block][row mod kDepthW]
    double* rho = reinterpret_cast<double*>(carve((size_t)(n + (n >> 5)) * 8));  // last step's profile (skewed)
    double* wwin = reinterpret_cast<double*>(carve((size_t)sa.wwin_len * 8));    // wake function window
    double* partial = reinterpret_cast<double*>(carve((size_t)NSEG * sa.max_rin * 8));  // wake partial sums
    float* red = reinterpret_cast<float*>(carve((size_t)sa.max_strip * NSEG * 4));       // row-sum partials
    // per input row: wake taps, RF taps, (f1 + f2, f2, -, -)
    float4* tab = reinterpret_cast<float4*>(carve((size_t)(sa.max_rin + 1) * 48));
    float4* zeroq = reinterpret_cast<float4*>(carve(16));                        // the quad beyond the grid's ends
    float4* mail = reinterpret_cast<float4*>(carve((size_t)NB * 2 * kMailW * 16));   // [block][side][row mod kMailW]
    uint64_t* full = reinterpret_cast<uint64_t*>(carve((size_t)NB * kDepthW * 8));   // [block][row mod kDepthW]
    uint64_t* empty = reinterpret_cast<uint64_t*>(carve((size_t)NB * kDepthW * 8));

    const float* gin = g.in + (size_t)member * n * n;

    if (tid == 0) step_ctl_begin(g, ctl);
    if (tid < NB * kDepthW) {
        const uint32_t blk = tid / kDepthW;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + tid)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + tid)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // last step's profile and the first batch of the wake-function window are requested in the same
    // round trip (T = n / 4 threads: 4 profile entries and 5 window entries each).
    // w[x - x'] is stored at table index x - x' + n - 1; window entry t holds the table entry
    // xs + t, so row r and x' meet at t = r - x' + n - 1
    constexpr int NRV = 4, NWV = 5;
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

    // ---- set-up: the zero quad, per-row tables, last step's profile and the wake-function
    // window in shared memory
    if (tid == 0) zeroq[0] = make_float4(0, 0, 0, 0);
    for (uint32_t i = tid; i < NB * 2 * kMailW; i += T) mail[i] = make_float4(0, 0, 0, 0);  // seq 0 = nothing posted
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
    if (g.has_wake) {
        // W[x] = scaling * sum_x' rho[x'] w[x-x']  (src/PS/ElectricField.cpp:186-253 as a Toeplitz
        // product).  Segment s sums x' in [128 s, 128 s + 128); a lane owns 8 consecutive rows and 32
        // consecutive x', so the 8 wake-function values it needs slide through registers.
        const uint32_t rbk = lane & 7, sub = lane >> 3;
#pragma unroll 1
        for (uint32_t seg = warp; seg < warp + 1; seg++) {  // one segment per warp (16 warps, 16 segments at n = 2048)
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
    __syncthreads();  // tables visible to everybody


    const uint32_t tab_b = smem_u32(tab);
    const uint32_t full_b = smem_u32(full), empty_b = smem_u32(empty);
    // The schedulers favour the warp with the highest index among the eligible ones: the kick warps,
    // whose row takes longer, are the upper half of the CTA, so a stencil warp that polls for its
    // neighbours never takes an issue slot a kick warp could use.
    if (warp >= NB) {
        // =============================== kick warps ============================================
        reg_alloc<IVS_WS_KREGS>();
        const uint32_t blk = warp - NB;
        const uint32_t cb = blk * 256;
        const uint32_t ycol = cb + 8 * lane;  // this lane's columns: ycol + 0..7
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
        const uint32_t ring_w = smem_u32(ring) + blk * (kSlotsW * kSlotB);
        const uint32_t fifo_w = smem_u32(fifo) + blk * (kDepthW * kSlotB);
        const uint32_t full_w = full_b + blk * (kDepthW * 8), empty_w = empty_b + blk * (kDepthW * 8);
        const bool edge_warp = blk == 0 || blk + 1 == NB;
        // mailboxes: lane 0 posts (x, y) of the block's first quad for the left neighbour, lane 31 (z, w) of
        // its last quad for the right neighbour
        const uint32_t post_me = smem_u32(mail) + (blk * 2 + (lane == 0 ? 0 : 1)) * kMailW * 16;
        const bool strip_edge = xs < 0 || xs + (int)RIN > (int)n;  // some of the strip's input rows lie outside the grid
        // staging: this lane's cell of every group of 32 staged cells (see ivs_stream8.cu)
        const uint32_t st_dst = ((lane >> 2) & 1) * kOddB + ((lane >> 3) * 4 + (lane & 3)) * 4;

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

            // ---- staging of input row r0 + k into slot k mod kSlotsW, one cp.async group per call
            const float* srow = gin + ((long long)(xs + (int)r0) * (long long)n + (long long)(cb + lane));
            const uint32_t stab = tab_b + r0 * 48 + 32;
            int sx = xs + (int)r0;
            int fs_next = lds32i(stab);
            if (strip_edge) {  // stale shared memory must not meet the zero weights as NaN * 0
                for (uint32_t i = lane; i < kSlotsW * kSlotB / 16; i += 32)
                    sts128(ring_w + i * 16, 0.0f, 0.0f, 0.0f, 0.0f);
                __syncwarp();
            }
            auto stage = [&](int k) {
                const int fs = fs_next;
                const uint32_t dst = ring_w + ((uint32_t)k & (kSlotsW - 1)) * kSlotB + st_dst;
                const float* src = srow + fs;
                // rows outside the grid (first and last strip of a member) are not copied: their RF taps
                // are zeros and the slots of such a strip were cleared above
                if (!strip_edge || (uint32_t)sx < n) {
                    // cells outside the grid are zeros (src-size 0: nothing is read); see ivs_stream8.cu
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
            // what crosses from one iteration to the next: the wake-kicked cells of the next row and its RF taps
            float k1[11];
            float4 w2 = make_float4(0, 0, 0, 0);
            float4 w1n, w2n;
            float in[14];
            int f2row = -1;
            auto load_row = [&](const int k) {
                const uint32_t tp = tab_b + (r0 + (uint32_t)k) * 48;
                w1n = lds128(tp);
                w2n = lds128(tp + 16);
                if (edge_warp) f2row = lds32i(tp + 36);
                const uint32_t sl = ring_w + ((uint32_t)k & (kSlotsW - 1)) * kSlotB + lane * 16;
                const float4 a0 = lds128(sl), b0 = lds128(sl + kOddB), a1 = lds128(sl + 16);
                const float2 b1 = lds64f(sl + kOddB + 16);
                in[0] = a0.x; in[1] = a0.y; in[2] = a0.z; in[3] = a0.w;
                in[4] = b0.x; in[5] = b0.y; in[6] = b0.z; in[7] = b0.w;
                in[8] = a1.x; in[9] = a1.y; in[10] = a1.z; in[11] = a1.w;
                in[12] = b1.x; in[13] = b1.y;
            };
            auto wake_kick = [&]() {  // k1[i] = cell ycol + f2 + i after the wake kick (src/SM/KickMap.cpp:243-266), 0 outside the grid
#pragma unroll
                for (int i = 0; i < 11; i++) k1[i] = tap4<FMA>(in[i], in[i + 1], in[i + 2], in[i + 3], w1n);
                w2 = w2n;
                if (edge_warp) {
                    const int f2 = f2row;
                    const int m0 = (int)ycol + f2;
                    if (f2 >= -3 && f2 <= 1) {
                        // only the block's first 3 (left edge of the grid) or last 4 (right edge) can lie outside
                        if (blk == 0) {
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

            // One iteration: kick of input row r0 + it; from it = NT-1 on the drifted output row
            // it - NT + 1 comes out and goes to FIFO row (it - NT + 1) mod kDepthW.
            auto iter = [&](const int it, auto stage_tag) {
                constexpr bool STAGE = decltype(stage_tag)::value;
                cp_async_wait<kSlotsW - 3>();  // row r0 + it + 1 has landed (this lane's part of it)
                __syncwarp();                  // ... and everybody's; slot it - 1 is free
                if (STAGE) stage(it + kSlotsW - 1);
                else cp_async_commit();        // keeps the group count in step
                float kq[8];
#pragma unroll
                for (int c = 0; c < 8; c++) kq[c] = tap4<FMA>(k1[c], k1[c + 1], k1[c + 2], k1[c + 3], w2);  // RF kick, same loop
                load_row(it + 1);  // the next row's cells and taps
                float d[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    // src/SM/KickMap.cpp:221-242, column-local 4-tap along x
                    d[c] = madd<FMA>(kq[c], dw[c][NT - 1], P[c][NT - 2]);
#pragma unroll
                    for (int t = NT - 2; t > 0; t--) P[c][t] = madd<FMA>(kq[c], dw[c][t], P[c][t - 1]);
                    P[c][0] = __fmul_rn(kq[c], dw[c][0]);
                }
                if (it + 1 >= NT) {  // a real output row: hand it to the stencil warps
                    const uint32_t orow = (uint32_t)(it - NT + 1);
                    const uint32_t slot = orow & (kDepthW - 1);
                    if (orow >= (uint32_t)kDepthW) mbar_wait_addr(empty_w + slot * 8, ((orow / kDepthW) - 1) & 1);
                    const uint32_t frow = fifo_w + slot * kSlotB + lane * 16;
                    sts128(frow, d[0], d[1], d[2], d[3]);
                    sts128(frow + kOddB, d[4], d[5], d[6], d[7]);
                    {
                        const float seq = __uint_as_float(orow + 1);
                        const bool l0 = lane == 0;
                        sts128_volatile_if(post_me + (orow & (kMailW - 1)) * 16,
                                           make_float4(l0 ? d[0] : seq, l0 ? d[1] : 0.0f, l0 ? seq : d[6], l0 ? 0.0f : d[7]),
                                           l0 || lane == 31);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_addr(full_w + slot * 8);
                }
                wake_kick();  // of the row just loaded, for the next iteration
            };
            const std::true_type yes{};
            const std::false_type no{};
#pragma unroll 1
            for (int k = 0; k < kSlotsW - 1; k++) {
                if (k < nkicks) stage(k);
                else cp_async_commit();
            }
            cp_async_wait<kSlotsW - 2>();
            __syncwarp();
            load_row(0);
            wake_kick();
            const int n_stage = nkicks - (kSlotsW - 1);  // iterations that still have a row to request
            int it = 0;
#pragma unroll 1
            for (; it < n_stage; it++) iter(it, yes);
#pragma unroll 1
            for (; it < nkicks; it++) iter(it, no);
        };
        if (spread == 0) run(std::integral_constant<int, 4>{});
        else if (spread == 1) run(std::integral_constant<int, 5>{});
        else run(std::integral_constant<int, 6>{});
    } else {
        // =============================== stencil warps =========================================
        reg_dealloc<IVS_WS_FREGS>();
        const uint32_t blk = warp;
        const uint32_t cb = blk * 256;
        // quads (cb / 4) + 32 j + lane, j = 0, 1
        float w5[2][4][5];
        float4 ws4[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t qf = (cb >> 2) + 32 * j + lane;
            const float4* wp = reinterpret_cast<const float4*>(g.fp_w5 + (size_t)qf * 20);
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
            ws4[j] = __ldg(reinterpret_cast<const float4*>(g.ws) + qf);
        }
        const uint32_t fifo_w = smem_u32(fifo) + blk * (kDepthW * kSlotB);
        const uint32_t full_w = full_b + blk * (kDepthW * 8), empty_w = empty_b + blk * (kDepthW * 8);
        const bool has_l = blk > 0, has_r = blk + 1 < NB;
        // offsets inside a FIFO row: own quad (j = 0), its left and right neighbours (see ivs_stream8.cu)
        const uint32_t own_off = (lane & 1) ? kOddB + (lane >> 1) * 16 : (lane >> 1) * 16;
        const int nb_l = (lane & 1) ? -(int)kOddB : (int)kOddB - 16;
        const int nb_r = (lane & 1) ? 16 - (int)kOddB : (int)kOddB;
        // the two quads beyond the block's ends: the neighbour kick warps' posts, or zeros beyond the grid
        const bool spec_l = lane == 0, spec_r = lane == 31;
        const bool edge_l = spec_l && has_l, edge_r = spec_r && has_r, edge = edge_l || edge_r;
        const uint32_t mail_b = smem_u32(mail);
        const uint32_t spec_base = edge_l   ? mail_b + ((blk - 1) * 2 + 1) * kMailW * 16
                                   : edge_r ? mail_b + ((blk + 1) * 2 + 0) * kMailW * 16
                                            : smem_u32(zeroq);
        const uint32_t spec_mask = edge ? (kMailW - 1) * 16 : 0;
        float4* gout = reinterpret_cast<float4*>(g.out + ((size_t)member * n + xa) * n) + (cb >> 2) + lane;
        const size_t gstride = (size_t)nq;
        float pprev[2] = {0.0f, 0.0f};
        float lvl[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // the row sums' butterfly, one level per row and iteration (ivs_stream8.cu)
        uint32_t redp = smem_u32(red) + (blk * 2 + (lane >= 16 ? 1 : 0)) * 4 - 5 * NSEG * 4;
        const bool red_lane = (lane & 15) == 0;
        auto rowsum_step = [&]() {
            const bool upper = (lane & 16) != 0;
            const float send = upper ? pprev[0] : pprev[1];
            const float keep = upper ? pprev[1] : pprev[0];
            const float s16 = __shfl_xor_sync(0xffffffffu, send, 16);
            const float s8 = __shfl_xor_sync(0xffffffffu, lvl[0], 8);
            const float s4 = __shfl_xor_sync(0xffffffffu, lvl[1], 4);
            const float s2 = __shfl_xor_sync(0xffffffffu, lvl[2], 2);
            const float s1 = __shfl_xor_sync(0xffffffffu, lvl[3], 1);
            const float done = __fadd_rn(lvl[3], s1);
            lvl[3] = __fadd_rn(lvl[2], s2);
            lvl[2] = __fadd_rn(lvl[1], s4);
            lvl[1] = __fadd_rn(lvl[0], s8);
            lvl[0] = __fadd_rn(keep, s16);
            return done;
        };
#pragma unroll 1
        for (uint32_t orow = 0; orow < S; orow++) {
            const uint32_t slot = orow & (kDepthW - 1), par = (orow / kDepthW) & 1;
            mbar_wait_addr(full_w + slot * 8, par);  // the block's kick warp has written the row
            const uint32_t fo = fifo_w + slot * kSlotB + own_off;
            const float4 o0 = lds128(fo), o1 = lds128(fo + 256);
            const float4 r0q = lds128(fo + nb_r), l1q = lds128(fo + 256 + nb_l);
            const uint32_t sp_a = spec_base + ((orow * 16) & spec_mask);
            const uint32_t la = spec_l ? sp_a : fo + nb_l;
            const uint32_t ra = spec_r ? sp_a : fo + 256 + nb_r;
            float4 l0q = lds128_volatile(la), r1q = lds128_volatile(ra);
            if (!__all_sync(0xffffffffu, !edge || __float_as_uint(edge_l ? l0q.x : r1q.z) == orow + 1)) {
                // a neighbour block's kick warp has not posted this row yet: poll
                uint32_t tries = 0;
                do {
                    __nanosleep(IVS_WS_SLEEP);
                    if (edge_l) l0q = lds128_volatile(la);
                    if (edge_r) r1q = lds128_volatile(ra);
                    if (++tries > (1u << 22)) __trap();  // a neighbour that never posts is a bug
                } while (!__all_sync(0xffffffffu, !edge || __float_as_uint(edge_l ? l0q.x : r1q.z) == orow + 1));
            }
            const float ps = rowsum_step();
            // Fokker-Planck (src/SM/FokkerPlanckMap.cpp:197-211)
            float o[2][4];
#define IVS_FP_COL(j, c)                                             \
            {                                                        \
                float v = __fmul_rn(fin[c], w5[j][c][0]);            \
                v = madd<FMA>(fin[1 + c], w5[j][c][1], v);           \
                v = madd<FMA>(fin[2 + c], w5[j][c][2], v);           \
                v = madd<FMA>(fin[3 + c], w5[j][c][3], v);           \
                v = madd<FMA>(fin[4 + c], w5[j][c][4], v);           \
                o[j][c] = v;                                         \
            }
            {
                const float fin[8] = {l0q.z, l0q.w, o0.x, o0.y, o0.z, o0.w, r0q.x, r0q.y};
                IVS_FP_COL(0, 0)
                IVS_FP_COL(0, 1)
                IVS_FP_COL(0, 2)
                IVS_FP_COL(0, 3)
                gout[0] = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
            }
            {
                const float fin[8] = {l1q.z, l1q.w, o1.x, o1.y, o1.z, o1.w, r1q.x, r1q.y};
                IVS_FP_COL(1, 0)
                IVS_FP_COL(1, 1)
                IVS_FP_COL(1, 2)
                IVS_FP_COL(1, 3)
                gout[32] = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
            }
#undef IVS_FP_COL
            // the loads have delivered (their values were used): the row may be overwritten
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_w + slot * 8);
            gout += gstride;
            if (red_lane && orow >= 5) asm volatile("st.shared.f32 [%0], %1;" ::"r"(redp), "f"(ps) : "memory");
            redp += NSEG * 4;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                float p = __fmul_rn(o[j][0], ws4[j].x);  // src/PS/PhaseSpace.cpp:316-319 (products rounded)
                p = __fadd_rn(p, __fmul_rn(o[j][1], ws4[j].y));
                p = __fadd_rn(p, __fmul_rn(o[j][2], ws4[j].z));
                p = __fadd_rn(p, __fmul_rn(o[j][3], ws4[j].w));
                pprev[j] = p;
            }
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
    }

    // the strip's row sums: one thread per row adds the segments' partials in fixed order
    __syncthreads();
    for (uint32_t i = tid; i < S; i += T) {
        float v = 0.0f;
        for (uint32_t k = 0; k < NSEG; k++) v = __fadd_rn(v, red[i * NSEG + k]);
        g.rho_out[(size_t)member * n + xa + i] = v;
    }
}

size_t stream8ws_smem(uint32_t n, uint32_t S0, uint32_t H, uint32_t* wwin_len_out) {
    auto a16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const uint32_t NB = n / 256, NSEG = n / 128;
    const uint32_t max_rin = S0 + H;
    const uint32_t wcount = n + ((max_rin + 63) & ~63u);
    const uint32_t wwin_len = wcount + (wcount >> 3) + 8;
    if (wwin_len_out) *wwin_len_out = wwin_len;
    return a16((size_t)NB * kSlotsW * kSlotB) + a16((size_t)NB * kDepthW * kSlotB) + a16((size_t)(n + n / 32) * 8) +
           a16((size_t)wwin_len * 8) + a16((size_t)NSEG * max_rin * 8) + a16((size_t)S0 * NSEG * 4) +
           a16((size_t)(max_rin + 1) * 48) + 16 + a16((size_t)NB * 2 * kMailW * 16) + 2 * a16((size_t)NB * kDepthW * 8);
}

}  // namespace

// Launch plan of the warp-specialised column-block kernel (1024 and 2048 columns: whole warpgroups
// of kick and of stencil warps, which setmaxnreg needs).  Returns false when it does not apply.
bool ivs_stream8ws_plan(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan, const int* dr_first) {
    const uint32_t n = s->n;
    if (n != 1024 && n != 2048) return false;
    for (uint32_t w = 0; w < n / 256; w++) {
        const auto mm = std::minmax_element(dr_first + 256 * w, dr_first + 256 * (w + 1));
        if (*mm.second - *mm.first > 2) return false;
    }
    const uint32_t H = plan->halo_lo + plan->halo_hi;
    const bool fma = ctx->arith == IVS_ARITH_FMA;
    const void* kfn = fma ? (const void*)stream8ws_step_kernel<true> : (const void*)stream8ws_step_kernel<false>;
    double best = 1e300;
    uint32_t bestP = 0, best_rin = 0, best_wwin = 0;
    size_t best_sb = 0;
    const uint32_t forceP = s->desc.strips;  // 0: the planner's choice
    const uint32_t per_sm = n == 2048 ? 1 : 2;  // CTAs per SM by registers (the whole file / half of it)
    for (uint32_t P = 1; P <= n / 2; P++) {
        if (forceP && P != forceP) continue;
        const uint32_t S0 = ivs_div_up(n, P);
        if ((uint64_t)S0 * (P - 1) >= n || S0 > 128) continue;  // last strip empty / table cap
        if (S0 < 8 && (uint64_t)P * s->members > (uint64_t)ctx->sm_count) continue;
        uint32_t wwin_len = 0;
        const size_t sb = stream8ws_smem(n, S0, H, &wwin_len);
        uint32_t want = per_sm;
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
    plan->stream8 = true;
    plan->stream8ws = true;
    plan->strip = ivs_div_up(n, bestP);
    plan->max_strip = plan->strip;
    plan->max_rin = best_rin;
    plan->wwin_len = best_wwin;
    plan->threads = n / 4;
    plan->smem = best_sb;
    plan->kslots = kSlotsW;
    return true;
}

ivs_status ivs_stream8ws_launch(ivs_ctx* ctx, const ivs_fused_plan* plan, const ivs_sim* s, const fused_args& a) {
    stream_args sa;
    sa.f = a;
    sa.f.fp_w5 = (const float*)plan->fp5;
    sa.strip = plan->strip; sa.kslots = plan->kslots;
    sa.max_strip = plan->max_strip; sa.max_rin = plan->max_rin; sa.wwin_len = plan->wwin_len;
    const dim3 sgrid(ivs_div_up(s->n, plan->strip), s->members);
    if (ctx->arith == IVS_ARITH_FMA) stream8ws_step_kernel<true><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
    else stream8ws_step_kernel<false><<<sgrid, plan->threads, plan->smem, ctx->stream>>>(sa);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
