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
    const float* wakefn;
    const float* wakescaling;
    const int* rf_first;
    const float4* rf_w;
    const int* dr_first;
    const float4* dr_w;
    const int* fp_base;
    const float4* fp_w;
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
__global__ void __launch_bounds__(1024, 1) fused_step_kernel(const fused_args g) {
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
            const float* wf = g.wakefn + (x + (int)n - 1);  // wf[-x'] = w[x - x']
            double acc = 0.0;
            for (uint32_t xp = lane; xp < n; xp += 32) acc += (double)rho[xp] * (double)__ldg(wf - (int)xp);
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

struct launch_cfg {
    int tx;
    uint32_t threads;
    size_t smem;
};

size_t smem_bytes(uint32_t n, uint32_t R, int tx, uint32_t threads) {
    const uint32_t pitch = n + 2 * kPad;
    size_t b = (size_t)R * pitch * 4 + (size_t)n * 4;
    b += ((R + 3) & ~3u) * 4 + (size_t)R * 16 + (size_t)tx * (threads / 32) * 4;
    return b;
}

template <int TX>
ivs_status launch_tx(ivs_ctx* ctx, const fused_args& a, dim3 grid, uint32_t threads, size_t smem) {
    if (ctx->arith == IVS_ARITH_FMA) fused_step_kernel<TX, true><<<grid, threads, smem, ctx->stream>>>(a);
    else fused_step_kernel<TX, false><<<grid, threads, smem, ctx->stream>>>(a);
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
    // tile height: minimise (rounds of CTAs over the machine) x (rows staged per CTA)
    const uint32_t halo = plan->halo_lo + plan->halo_hi;
    const int candidates[3] = {16, 8, 4};
    int tx = 0;
    uint32_t threads = 0;
    size_t smem = 0;
    uint64_t best = ~0ull;
    for (int c : candidates) {
        if ((uint32_t)c > n) continue;
        const uint32_t R = c + halo;
        uint32_t th = std::min<uint32_t>(1024, std::max<uint32_t>(128, ((R * n / 32) / 32) * 32));
        th = std::min<uint32_t>(th, std::max<uint32_t>(128, n / 2));
        const size_t sb = smem_bytes(n, R, c, th);
        if (sb > 227 * 1024) continue;
        // 64 registers per thread (launch bounds) -> at most 1024 resident threads per SM
        const uint32_t per_sm = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(227 * 1024 / (sb + 1024)), 1024 / th));
        const uint64_t ctas = (uint64_t)ivs_div_up(n, c) * s->members;
        const uint64_t rounds = (ctas + (uint64_t)ctx->sm_count * per_sm - 1) / ((uint64_t)ctx->sm_count * per_sm);
        const uint64_t cost = rounds * R * std::max<uint32_t>(1, n / th);
        if (cost < best) { best = cost; tx = c; threads = th; smem = sb; }
    }
    if (tx == 0)
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED, "drift displacement needs more halo rows than fit in shared memory: pass IVS_SIM_UNFUSED");
    plan->tx = tx; plan->threads = threads; plan->smem = smem; plan->launches_per_step = 1;
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
    switch (tx) {
    case 16:
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        break;
    case 8:
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        break;
    default:
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IVS_CUDA(ctx, cudaFuncSetAttribute(fused_step_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        break;
    }
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
    a.ws = s->ws;
    a.status = s->status;
    a.n = s->n; a.it = s->it; a.hl = plan->halo_lo; a.hr = plan->halo_hi; a.pitch = s->n + 2 * kPad;
    a.has_wake = s->has_wake ? 1 : 0;
    const dim3 grid(ivs_div_up(s->n, plan->tx), s->members);
    ivs_status st;
    switch (plan->tx) {
    case 16: st = launch_tx<16>(ctx, a, grid, plan->threads, plan->smem); break;
    case 8: st = launch_tx<8>(ctx, a, grid, plan->threads, plan->smem); break;
    default: st = launch_tx<4>(ctx, a, grid, plan->threads, plan->smem); break;
    }
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
