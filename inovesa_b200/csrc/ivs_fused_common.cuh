// ivs_fused_common.cuh -- pieces shared by the fused step kernels (ivs_fused.cu, ivs_stream8.cu):
// kernel arguments, step bookkeeping, shared-memory and mbarrier accessors.  Internal to libivs.so.
#pragma once
#include "ivs_internal.h"
#include "ivs_sim_struct.h"

namespace ivs_fused_detail {

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
    int* status;     // [0] ordinal of the first step that left the staged pad (kStatusClean: none), [1] CTA tickets
    int* host_flag;  // host-mapped copy of "status[0] is set" (read by the host without a sync)
    uint32_t n, it, hl, hr, pitch;
    int has_wake;
    int wb;          // which of the two wake buffers wake_out is: status[2 + wb] = ordinal of this step
    // train of bunches (ivs_sim_set_bunch_train; column-block kernel only): the members are the bunches, coupled
    // through train_windows[(b * train_nb + c) * 2 n + m + n - 1]; 0: independent members
    uint32_t train_nb = 0;
    const double* train_windows = nullptr;
};

struct stream_args {
    fused_args f;
    uint32_t strip;      // rows per strip
    uint32_t kslots;     // ring slots
    uint32_t max_strip;  // capacity of the partial-sum table (rows)
    uint32_t max_rin;    // capacity of the per-row tap tables (input rows)
    uint32_t wwin_len;   // doubles in the staged wake-function window
};

}  // namespace ivs_fused_detail

// the column-block kernel (ivs_stream8.cu); the plan's halo and table fields are filled by ivs_fused_plan_create
bool ivs_stream8_plan(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan, const int* dr_first);
ivs_status ivs_stream8_launch(ivs_ctx* ctx, const ivs_fused_plan* plan, const ivs_sim* s,
                              const ivs_fused_detail::fused_args& a);
// the persistent kernel for runs that fit the machine at once (ivs_resident.cu)
bool ivs_resident_plan(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan);
ivs_status ivs_resident_run(ivs_ctx* ctx, const ivs_fused_plan* plan, const ivs_sim* s,
                            const ivs_fused_detail::fused_args& a, uint32_t nsteps);

namespace {
using ivs_fused_detail::fused_args;
using ivs_fused_detail::stream_args;


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


// ---- step bookkeeping shared by the three fused kernels.  status[0] = ordinal (counted from the
// host's last reset of the two words) of the first step in which a wake kick left the staged pad,
// status[1] = CTA ticket counter (ticket / CTAs per step = this step's ordinal).  Once status[0] is
// set every later step exits without writing, so the input of the offending step stays intact
// in the ping-pong buffer and the host can redo it through the kernel sequence (ivs_step.cu).
constexpr int kStatusClean = 0x7fffffff;
__device__ __forceinline__ void step_ctl_begin(const fused_args& g, int* ctl) {
    ctl[0] = *reinterpret_cast<volatile int*>(g.status);
    ctl[1] = (int)atomicAdd(reinterpret_cast<unsigned int*>(g.status) + 1, 1u);
    // the step's first ticket notes which step fills wake buffer wb: the tracking kernel that reads
    // the buffer (possibly while the NEXT step runs and overflows) only stands down if this step
    // or an earlier one overflowed
    const unsigned int ctas = gridDim.x * gridDim.y;
    if ((unsigned int)ctl[1] % ctas == 0) g.status[2 + g.wb] = (int)((unsigned int)ctl[1] / ctas);
}
__device__ __forceinline__ void step_ctl_overflow(const fused_args& g, const int* ctl) {
    const unsigned int ord = (unsigned int)ctl[1] / (gridDim.x * gridDim.y);
    atomicMin(g.status, (int)ord);
    *reinterpret_cast<volatile int*>(g.host_flag) = 1;
    __threadfence_system();
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

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


// shared-memory accessors on 32-bit shared addresses (no 64-bit generic pointer arithmetic)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128_volatile(uint32_t addr) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ int2 lds64i(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(phase)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(phase)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the same as predicated instructions (no branch around a one-lane operation)
__device__ __forceinline__ void mbar_arrive_if(uint32_t bar, bool on) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p mbarrier.arrive.shared::cta.b64 _, [%0]; }" ::"r"(bar),
                 "r"((uint32_t)on)
                 : "memory");
}
__device__ __forceinline__ void sts128_volatile_if(uint32_t addr, const float4 v, bool on) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p st.volatile.shared.v4.f32 [%0], {%1,%2,%3,%4}; }" ::"r"(addr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((uint32_t)on)
                 : "memory");
}

// staged wake-function window: one pad double after every 8 keeps lanes that are 8 entries apart
// on different banks
__device__ __forceinline__ uint32_t wwin_pos(uint32_t t) { return t + (t >> 3); }

}  // namespace
