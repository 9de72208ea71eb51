// ivs_dynrf.cu -- DynamicRFKickMap on the device (src/SM/DynamicRFKickMap.cpp:80-109).
//
// The reference re-draws the RF phase and amplitude every step and recomputes KickMap::_offset on
// the host inside apply() (RFKickMap::_calcKick(phase, ampl), src/SM/RFKickMap.cpp:59-83), from a
// queue of (phase, amplitude) pairs it fills ahead of time for all steps of the run
// (__calcModulation, :80-94).  Here the queue is a device array and one small kernel per step turns
// its next entry into the offsets and the 4-tap tables the step kernels read -- inside the step
// graph, with the position in the queue kept on the device, so that a replayed graph walks through
// the queue.  The arithmetic is _calcKick's, fp32, operation by operation; the sine of the
// non-linear map is formed in double and rounded (glibc's sinf is correctly rounded for all but a
// handful of arguments, CUDA's sinf is not).
#include "ivs_common.cuh"
#include "ivs_internal.h"

namespace {

struct dynrf_args {
    ivs_rf_model m;
    const float* seq;       // [count][2] = (phase, amplitude factor)
    uint32_t count, n, it;
    unsigned int* pos;      // device: next entry of the queue
    const int* sim_status;  // fused-step bookkeeping (a poisoned run does not advance)
    float* offset;          // [n] KickMap::_offset (tracking and the unfused kick read it)
    int* first;             // [n] fused kernels' tap tables (may be null)
    float4* w;
    int* flag;              // set to 1 if a first tap leaves what the fused kernels stage
    int lo, hi;             // admissible range of the first tap
};

__global__ void __launch_bounds__(256) dynrf_kernel(const dynrf_args a) {
    if (a.sim_status && *a.sim_status != 0x7fffffff) return;
    const unsigned int pos = *reinterpret_cast<volatile unsigned int*>(a.pos);
    __syncthreads();  // everybody has read the position before thread 0 advances it
    if (pos >= a.count) return;  // queue exhausted: the last offsets stay (the reference would pop an empty queue)
    const float phase = a.seq[2 * pos], ampl = a.seq[2 * pos + 1];
    const ivs_rf_model& m = a.m;
    for (uint32_t x = threadIdx.x; x < a.n; x += blockDim.x) {
        float off;
        if (m.linear) {
            const float phaseoffs = IVS_SUB(m.syncphase, phase);
            off = IVS_MUL(m.tan_angle, IVS_SUB(m.xcenter, (float)x));
            off = IVS_ADD(off, IVS_DIV(IVS_DIV(IVS_MUL(m.tan_angle, phaseoffs), m.bl2phase), m.delta0));
            off = IVS_MUL(off, ampl);
        } else {
            const float q = IVS_ADD(m.qmin, IVS_MUL((float)x, m.delta0));  // Ruler::at(x)
            const float arg = IVS_ADD(IVS_MUL(q, m.bl2phase), phase);
            const float sn = (float)sin((double)arg);
            const float v = IVS_ADD(IVS_MUL(IVS_MUL(-ampl, m.V_RF), sn), m.V0);
            off = IVS_DIV(IVS_DIV(IVS_MUL(m.revolutionpart, v), m.delta1), m.dE_scale);
        }
        a.offset[x] = off;
        if (a.first) {
            const ivs_taps t = ivs_kick_taps(off, a.n, a.it);  // KickMap::updateSM
            if (t.first < a.lo || t.first > a.hi) *a.flag = 1;
            a.first[x] = t.first;
            a.w[x] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
        }
    }
    if (threadIdx.x == 0) *a.pos = pos + 1;
}

}  // namespace

ivs_status ivs_dynrf_launch(ivs_ctx* ctx, const ivs_rf_model* model, const float* d_seq, uint32_t count,
                            unsigned int* d_pos, const int* sim_status, uint32_t n, uint32_t it, float* d_offset,
                            int* d_first, void* d_w, int* d_flag, int lo, int hi) {
    dynrf_args a;
    a.m = *model; a.seq = d_seq; a.count = count; a.n = n; a.it = it; a.pos = d_pos; a.sim_status = sim_status;
    a.offset = d_offset; a.first = d_first; a.w = (float4*)d_w; a.flag = d_flag; a.lo = lo; a.hi = hi;
    dynrf_kernel<<<1, 256, 0, ctx->stream>>>(a);
    IVS_LAUNCHED(ctx);
    return IVS_OK;
}
