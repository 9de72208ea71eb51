// ivs_common.cuh -- shared declarations of the CUDA handler (sm_100a only).
//
// Arithmetic contract (DESIGN.md "Arithmetic contract"): the library is compiled with
// -fmad=false, so NO multiply-add is contracted implicitly; every rounding follows the
// reference's CPU source operation by operation.  The only fused operations are the explicit
// madd<true>() calls of the tap accumulations when the context runs in IVS_ARITH_FMA.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ivs.h"

#ifndef IVS_HD
#ifdef __CUDACC__
#define IVS_HD __host__ __device__ __forceinline__
#else
#define IVS_HD inline
#endif
#endif

struct ivs_ctx {
    int device = 0;
    int arith = IVS_ARITH_FMA;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint64_t launches = 0;
    std::string error;
    // cached twiddle tables for the shared-memory FFT, keyed by nmax (device, nmax/2 float2)
    void* twiddle[16] = {nullptr};
    uint32_t twiddle_n[16] = {0};
    float* scratch = nullptr;  // small device scratch (integrals, flags)
    uint32_t track_slots = 0, track_slots_n = 0;  // blocks of track_step_staged_kernel resident at once, for tables of track_slots_n
    uint32_t optin = 0;        // kernels already opted in to large dynamic shared memory (bit 0: rotate_tile_kernel, bit 1: field_kernel)
};

// error plumbing ---------------------------------------------------------------------------
ivs_status ivs_fail(ivs_ctx* ctx, ivs_status code, const std::string& msg);
ivs_status ivs_check_cuda(ivs_ctx* ctx, cudaError_t e, const char* what);
#define IVS_CUDA(ctx, call)                                                     \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ivs_check_cuda((ctx), e__, #call);       \
    } while (0)
#define IVS_LAUNCHED(ctx)                                                       \
    do {                                                                        \
        (ctx)->launches++;                                                      \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess) return ivs_check_cuda((ctx), e__, "kernel launch"); \
    } while (0)
#define IVS_REQUIRE(ctx, cond, msg)                                             \
    do {                                                                        \
        if (!(cond)) return ivs_fail((ctx), IVS_ERR_INVALID_ARGUMENT, (msg));   \
    } while (0)

// fp32 primitives that never contract -------------------------------------------------------
#ifdef __CUDA_ARCH__
#define IVS_MUL(a, b) __fmul_rn((a), (b))
#define IVS_ADD(a, b) __fadd_rn((a), (b))
#define IVS_SUB(a, b) __fsub_rn((a), (b))
#define IVS_DIV(a, b) __fdiv_rn((a), (b))
#else
// host translation units are compiled with -ffp-contract=off
#define IVS_MUL(a, b) ((a) * (b))
#define IVS_ADD(a, b) ((a) + (b))
#define IVS_SUB(a, b) ((a) - (b))
#define IVS_DIV(a, b) ((a) / (b))
#endif

// acc + a*w, fused or not: the tap accumulation `value += in*weight` of the reference's map
// loops (src/SM/KickMap.cpp:236,260; src/SM/FokkerPlanckMap.cpp:205; src/SM/SourceMap.cpp:103)
template <bool FMA>
__device__ __forceinline__ float madd(float a, float w, float acc) {
    if (FMA) return __fmaf_rn(a, w, acc);
    return __fadd_rn(__fmul_rn(a, w), acc);
}

// SourceMap::calcCoefficiants (src/SM/SourceMap.cpp:142-170), fp32, unfused, same order.
IVS_HD void ivs_calc_coefficients(float f, uint32_t it, float* ic) {
    switch (it) {
    case 1:
        ic[0] = 1.0f;
        break;
    case 2:
        ic[0] = IVS_SUB(1.0f, f);
        ic[1] = f;
        break;
    // (x / 2 as x * 0.5f: scaling by a power of two is exact, so the two round alike everywhere,
    // and the product is one instruction where the IEEE division is a subroutine)
    case 3:
        ic[0] = IVS_MUL(IVS_MUL(f, IVS_SUB(f, 1.0f)), 0.5f);
        ic[1] = IVS_SUB(1.0f, IVS_MUL(f, f));
        ic[2] = IVS_MUL(IVS_MUL(f, IVS_ADD(f, 1.0f)), 0.5f);
        break;
    default:
        ic[0] = IVS_MUL(IVS_MUL(IVS_MUL(IVS_SUB(f, 1.0f), IVS_SUB(f, 2.0f)), f), (float)(-1. / 6.));
        ic[1] = IVS_MUL(IVS_MUL(IVS_MUL(IVS_ADD(f, 1.0f), IVS_SUB(f, 1.0f)), IVS_SUB(f, 2.0f)), 0.5f);
        ic[2] = IVS_MUL(IVS_MUL(IVS_MUL(IVS_SUB(2.0f, f), f), IVS_ADD(f, 1.0f)), 0.5f);
        ic[3] = IVS_MUL(IVS_MUL(IVS_MUL(f, IVS_ADD(f, 1.0f)), IVS_SUB(f, 1.0f)), (float)(1. / 6.));
        break;
    }
}

// One row/column entry of KickMap::updateSM (src/SM/KickMap.cpp:339-389) in a form suited to
// a register window: the taps sit on 4 consecutive cells starting `first` cells from the
// output cell, weights w[0..3] in the reference's accumulation order.  Interpolation types
// with fewer than 4 points occupy the matching slots with zero weights elsewhere (adding
// x*0 = +0 does not change an fp32 sum), invalid taps (index outside the grid, :365-371) and
// invalid rows (:375-380) get weight 0 exactly as the reference stores {n/2, 0}.
struct ivs_taps {
    int first;
    float w[4];
};

IVS_HD ivs_taps ivs_kick_taps(float offset, uint32_t n_kd, uint32_t it) {
    ivs_taps t;
    const float poffs = IVS_ADD((float)(n_kd / 2), offset);  // fp32 add: quantises the fraction
    const float qp_int = truncf(poffs);
    const float xip = IVS_SUB(poffs, qp_int);                // == modf fraction (exact)
    // `meshindex_t jd = qp_int` on x86-64: cvttss2si (64 bit), low 32 bits kept
    const long long jl = (poffs == poffs && fabsf(qp_int) < 9.0e18f) ? (long long)qp_int : 0ll;
    const uint32_t jd = (uint32_t)jl;
    t.first = (int)jd - 1 - (int)(n_kd / 2);
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.0f;
    if (jd < n_kd) {
        float smc[4];
        ivs_calc_coefficients(xip, it, smc);
        const uint32_t lo = (it - 1) / 2;  // taps start at jd - lo
        for (uint32_t j1 = 0; j1 < it; j1++) {
            const uint32_t j0 = jd + j1 - lo;
            if (j0 < n_kd) t.w[j1 + 1 - lo] = smc[j1];
        }
    } else {
        t.first = -1;  // all-zero row; keep the window in range
    }
    return t;
}

static inline uint32_t ivs_div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Opt-in limit of dynamic shared memory per CTA on sm_100 (227 KB).  Kernels that need more than
// 48 KB are always opted in to this constant, never to the size one plan happens to need:
// the attribute belongs to the kernel, and two live simulations of different grid sizes share it.
constexpr int kIvsMaxOptinSmem = 227 * 1024 - 256;  // minus room for a kernel's few static bytes
