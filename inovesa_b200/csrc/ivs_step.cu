// ivs_step.cu -- the simulation object: device-resident state of `members` independent
// single-bunch simulations and the loop body of main() (src/main.cpp:948-1053) on the device.
//
//   wake potential -> wake kick (y) -> RF kick (y) -> drift (x) -> Fokker-Planck (y) -> X projection
//
// Two executions of the same step:
//   * unfused (IVS_SIM_UNFUSED): the reference's kernel sequence, one launch per map (ivs_maps.cu),
//     three grids t1/t2/t3 wired like src/main.cpp:593-763;
//   * fused (default): ONE kernel per step (ivs_fused.cuh) that reads the grid once and writes it
//     once -- 8 B per cell-update, the algorithmic minimum of SURVEY.md section 8(d).
// Either way the step sequence is captured in a CUDA graph and replayed.
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "ivs_common.cuh"
#include "ivs_internal.h"
#include "ivs_sim_struct.h"

namespace {

template <typename T>
ivs_status dalloc(ivs_ctx* ctx, T** p, size_t count) {
    IVS_CUDA(ctx, cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return IVS_OK;
}
#define IVS_TRY(expr)                       \
    do {                                    \
        ivs_status st__ = (expr);           \
        if (st__ != IVS_OK) return st__;    \
    } while (0)

// one step as the reference's kernel sequence (src/main.cpp:948-1053), grids wired
// wm: t1->t2, rfm: t2->t1, drm: t1->t3, fpm: t3->t1 (:593-763); t1 = the current buffer (it stays
// the current one), t2 = the other ping-pong buffer, t3 = the third grid
ivs_status enqueue_unfused_step(ivs_sim* s) {
    ivs_ctx* ctx = s->ctx;
    const uint32_t n = s->n, m = s->members;
    float *t1 = s->grid[s->cur], *t2 = s->grid[s->cur ^ 1], *t3 = s->grid[2];
    float* prof = s->profile[s->pcur];
    const size_t bytes = (size_t)m * n * n * sizeof(float);
    if (s->has_wake) {
        // wkm->update(): ElectricField::wakePotential -> offsets (src/SM/WakePotentialMap.cpp:34-50)
        if (s->train_windows) {
            IVS_TRY(ivs_train_wake_launch(ctx, prof, n, m, s->train_windows, s->wakescaling, 0.0f, s->wake));
        } else {
            IVS_TRY(ivs_field_launch(ctx, prof, n, m, s->impedance, s->nmax, s->wakescaling, 0, s->wake,
                                     nullptr, nullptr, nullptr, 0, 0, 0, 0));
        }
        IVS_TRY(ivs_kick_y_batched(ctx, t1, t2, n, m, s->wake, n, s->it));
    } else {
        IVS_CUDA(ctx, cudaMemcpyAsync(t2, t1, bytes, cudaMemcpyDeviceToDevice, ctx->stream));  // Identity
    }
    IVS_TRY(ivs_kick_y_batched(ctx, t2, t1, n, m, s->rf, 0, s->it));
    IVS_TRY(ivs_kick_x_batched(ctx, t1, t3, n, m, s->drift, 0, s->it));
    if (s->np) {
        IVS_TRY(ivs_track_step_launch(ctx, s->particles, s->np, m, n, s->has_wake ? s->wake : nullptr,
                                      s->rf, s->drift, s->has_fp ? s->fpw : nullptr, t3,
                                      s->desc.fptrack, s->desc.e1, s->desc.dp, s->desc.seed, s->track_counter));
    }
    if (s->has_fp) {
        IVS_TRY(ivs_fokker_planck_batched(ctx, t3, t1, n, m, s->fpw));
    } else {
        IVS_CUDA(ctx, cudaMemcpyAsync(t1, t3, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    IVS_TRY(ivs_project_x_batched(ctx, t1, n, m, s->ws, prof));
    return IVS_OK;
}

ivs_status enqueue_step(ivs_sim* s) {
    const bool unfused = (s->desc.flags & IVS_SIM_UNFUSED) != 0;
    if (s->dyn_rf) {
        // DynamicRFKickMap::apply: _calcKick(next modulation) before the map is applied
        // (src/SM/DynamicRFKickMap.cpp:102-109)
        IVS_TRY(ivs_dynrf_launch(s->ctx, &s->rf_model, s->rf_seq, s->rf_seq_len, s->rf_pos,
                                 unfused ? nullptr : s->status, s->n, s->it, s->rf,
                                 unfused ? nullptr : (int*)s->plan.rf_first, unfused ? nullptr : s->plan.rf_w,
                                 (int*)(s->rf_pos + 1), -16, 12));
    }
    if (unfused) return enqueue_unfused_step(s);
    return ivs_fused_enqueue_step(s->ctx, &s->plan, s);
}

void drop_graphs(ivs_sim* s) {
    for (int k = 0; k < 2; k++) {
        if (s->graph1[k]) cudaGraphExecDestroy(s->graph1[k]);
        if (s->graphK[k]) cudaGraphExecDestroy(s->graphK[k]);
        s->graph1[k] = s->graphK[k] = nullptr;
    }
}

ivs_status capture(ivs_sim* s, uint32_t steps, cudaGraphExec_t* out) {
    ivs_ctx* ctx = s->ctx;
    cudaGraph_t g = nullptr;
    const int cur0 = s->cur, pcur0 = s->pcur;
    const uint64_t launches0 = ctx->launches, done0 = s->steps_done, enq0 = s->fused_enqueued;
    IVS_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ivs_status st = IVS_OK;
    for (uint32_t i = 0; i < steps && st == IVS_OK; i++) {
        st = enqueue_step(s);
        s->steps_done++;
    }
    if (st == IVS_OK) st = ivs_fused_join(ctx, s);  // the side stream's launches rejoin the captured stream
    s->track_inflight[0] = s->track_inflight[1] = false;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    // capture only records; restore the bookkeeping it advanced
    s->cur = cur0; s->pcur = pcur0; ctx->launches = launches0; s->steps_done = done0; s->fused_enqueued = enq0;
    if (st != IVS_OK) { if (g) cudaGraphDestroy(g); return st; }
    if (e != cudaSuccess) return ivs_check_cuda(ctx, e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return ivs_check_cuda(ctx, e, "cudaGraphInstantiate");
    return IVS_OK;
}

// how many kernels one step launches (for ivs_ctx_launch_count bookkeeping under graphs)
uint32_t launches_per_step(const ivs_sim* s) {
    if (!(s->desc.flags & IVS_SIM_UNFUSED)) return s->plan.launches_per_step + (s->np ? 1 : 0) + (s->dyn_rf ? 1 : 0);
    uint32_t k = 3 + (s->dyn_rf ? 1 : 0);  // rf kick, drift, projection
    if (s->has_wake) k += 2;
    if (s->has_fp) k += 1;
    if (s->np) k += 1;
    return k;
}

}  // namespace

extern "C" {

static ivs_status reset_status(ivs_sim* s);

ivs_status ivs_wake_function(ivs_ctx* ctx, const void* d_impedance, uint32_t nmax,
                             float wakescaling, uint32_t n, void* d_wakefunction) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d_impedance && d_wakefunction, "ivs_wake_function: null pointer");
    IVS_REQUIRE(ctx, nmax >= 2 * n && (nmax & (nmax - 1)) == 0,
                "ivs_wake_function: nmax must be a power of two >= 2n");
    std::vector<float> z(2 * (size_t)nmax);
    IVS_TRY(ivs_download(ctx, z.data(), d_impedance, z.size() * sizeof(float)));
    std::vector<double> wd = ivs_wake_function_host(z.data(), nmax, n);
    std::vector<float> w(wd.size());
    for (size_t i = 0; i < wd.size(); i++) w[i] = wakescaling * (float)wd[i];
    IVS_TRY(ivs_upload(ctx, d_wakefunction, w.data(), w.size() * sizeof(float)));
    return ivs_sync(ctx);
}

ivs_status ivs_sim_create(ivs_ctx* ctx, const ivs_sim_desc* d, ivs_sim** out) {
    if (!ctx) return IVS_ERR_INVALID_ARGUMENT;
    IVS_REQUIRE(ctx, d && out, "ivs_sim_create: null argument");
    *out = nullptr;
    IVS_REQUIRE(ctx, d->n >= 8, "ivs_sim_create: grid size must be >= 8");
    IVS_REQUIRE(ctx, d->members >= 1, "ivs_sim_create: members must be >= 1");
    IVS_REQUIRE(ctx, d->it >= 1 && d->it <= 4, "ivs_sim_create: interpolation type must be 1..4");
    IVS_REQUIRE(ctx, d->rf_offset && d->drift_offset && d->grid, "ivs_sim_create: rf_offset, drift_offset and grid are required");
    IVS_REQUIRE(ctx, d->fptrack <= 3, "ivs_sim_create: FPTracking must be 0..3");
    const bool has_wake = d->impedance != nullptr && d->nmax != 0;
    if (has_wake) {
        IVS_REQUIRE(ctx, (d->nmax & (d->nmax - 1)) == 0 && d->nmax >= 2 * d->n && d->nmax <= 16384,
                    "ivs_sim_create: nmax must be a power of two with 2n <= nmax <= 16384");
        IVS_REQUIRE(ctx, d->wakescaling != nullptr, "ivs_sim_create: wakescaling is required with an impedance");
    }
    IVS_REQUIRE(ctx, d->n_particles == 0 || d->particles, "ivs_sim_create: particles missing");
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));

    ivs_sim* s = new ivs_sim;
    s->ctx = ctx; s->desc = *d;
    s->n = d->n; s->members = d->members; s->it = d->it; s->nmax = has_wake ? d->nmax : 0;
    s->np = d->n_particles; s->has_wake = has_wake; s->has_fp = d->fp_weights != nullptr;
    const uint32_t n = s->n, m = s->members;
    const size_t cells = (size_t)m * n * n;
    const bool unfused = (d->flags & IVS_SIM_UNFUSED) != 0;
    ivs_status st = IVS_OK;
    auto fail = [&](ivs_status code) { ivs_sim_destroy(s); return code; };
#define S_TRY(expr) do { st = (expr); if (st != IVS_OK) return fail(st); } while (0)
    for (int i = 0; i < (unfused ? 3 : 2); i++) S_TRY(dalloc(ctx, &s->grid[i], cells));
    S_TRY(dalloc(ctx, &s->profile[0], (size_t)m * n));
    S_TRY(dalloc(ctx, &s->profile[1], (size_t)m * n));
    S_TRY(dalloc(ctx, &s->wake, (size_t)m * n));
    s->wakebuf[0] = s->wakebuf[1] = s->wake;
    // tracked particles of a fused run step on a side stream, overlapping the next step (not with a
    // DynamicRFKickMap, whose per-step kernel rewrites the RF offsets the particles read)
    s->track_overlap = !unfused && d->n_particles > 0 && !(d->flags & IVS_SIM_TRACK_INLINE);
    if (s->track_overlap) {
        S_TRY(dalloc(ctx, &s->wakebuf[1], (size_t)m * n));
        S_TRY(ivs_memset_zero(ctx, s->wakebuf[1], (size_t)m * n * sizeof(float)));
        int lo = 0, hi = 0;
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least urgent
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_step, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_track[0], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_track[1], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(ivs_check_cuda(ctx, e, "side stream of the tracking kernel"));
    }
    S_TRY(dalloc(ctx, &s->rf, n));
    S_TRY(dalloc(ctx, &s->drift, n));
    S_TRY(dalloc(ctx, &s->fpw, (size_t)n * 5));
    S_TRY(dalloc(ctx, &s->ws, n));
    S_TRY(dalloc(ctx, &s->projy, (size_t)m * n));
    S_TRY(dalloc(ctx, &s->obs, (size_t)m * 8));
    S_TRY(dalloc(ctx, &s->integral, m));
    S_TRY(dalloc(ctx, &s->moments, (size_t)m * 6));
    S_TRY(dalloc(ctx, &s->power, m));
    S_TRY(dalloc(ctx, &s->status, 4));
    {
        void* hf = nullptr;
        cudaError_t e = cudaHostAlloc(&hf, 64, cudaHostAllocMapped);
        if (e != cudaSuccess) return fail(ivs_check_cuda(ctx, e, "cudaHostAlloc"));
        s->host_flag = (int*)hf;
    }
    S_TRY(reset_status(s));
    S_TRY(dalloc(ctx, &s->track_counter, 4));
    S_TRY(ivs_memset_zero(ctx, s->track_counter, 4 * sizeof(unsigned int)));
    S_TRY(ivs_memset_zero(ctx, s->wake, (size_t)m * n * sizeof(float)));
    S_TRY(ivs_memset_zero(ctx, s->power, m * sizeof(float)));
    S_TRY(ivs_upload(ctx, s->grid[0], d->grid, cells * sizeof(float)));
    S_TRY(ivs_upload(ctx, s->rf, d->rf_offset, n * sizeof(float)));
    S_TRY(ivs_upload(ctx, s->drift, d->drift_offset, n * sizeof(float)));
    std::vector<float> ws(n), fpw((size_t)n * 5, 0.0f);
    ivs_simpson_weights(n, d->dq, ws.data());
    S_TRY(ivs_upload(ctx, s->ws, ws.data(), n * sizeof(float)));
    if (s->has_fp) {
        S_TRY(ivs_upload(ctx, s->fpw, d->fp_weights, (size_t)n * 5 * sizeof(float)));
    } else {
        for (uint32_t y = 0; y < n; y++) fpw[y * 5 + 2] = 1.0f;  // Identity
        S_TRY(ivs_upload(ctx, s->fpw, fpw.data(), (size_t)n * 5 * sizeof(float)));
    }
    S_TRY(ivs_sync(ctx));  // host vectors above go out of scope
    std::vector<double> wf;
    if (has_wake) {
        S_TRY(dalloc(ctx, &s->impedance, (size_t)2 * s->nmax));
        S_TRY(dalloc(ctx, &s->wakescaling, m));
        S_TRY(dalloc(ctx, &s->wakefunction, (size_t)2 * n));
        S_TRY(ivs_upload(ctx, s->impedance, d->impedance, (size_t)2 * s->nmax * sizeof(float)));
        S_TRY(ivs_upload(ctx, s->wakescaling, d->wakescaling, m * sizeof(float)));
        S_TRY(ivs_field_prepare(ctx, s->nmax));
        wf = ivs_wake_function_host(d->impedance, s->nmax, n);
        wf.push_back(0.0);
        S_TRY(ivs_upload(ctx, s->wakefunction, wf.data(), (size_t)2 * n * sizeof(double)));
    }
    if (s->np) {
        S_TRY(dalloc(ctx, &s->particles, (size_t)m * s->np * 2));
        S_TRY(ivs_upload(ctx, s->particles, d->particles, (size_t)m * s->np * 2 * sizeof(float)));
    }
    // grid_t1->updateXProjection() before the loop (src/main.cpp:893-894)
    S_TRY(ivs_project_x_batched(ctx, s->grid[0], n, m, s->ws, s->profile[0]));
    if (!unfused) S_TRY(ivs_fused_plan_create(ctx, s, &s->plan));
    S_TRY(ivs_sync(ctx));
    // the step graphs are captured and instantiated here, one per buffer parity, so that no
    // ivs_sim_step ever pays for it
    if (!(d->flags & IVS_SIM_NO_GRAPH)) {
        S_TRY(capture(s, ivs_sim::kGraphSteps, &s->graphK[0]));
        if (unfused) {
            S_TRY(capture(s, 1, &s->graph1[0]));
        } else {
            s->cur = s->pcur = 1;
            st = capture(s, ivs_sim::kGraphSteps, &s->graphK[1]);
            s->cur = s->pcur = 0;
            if (st != IVS_OK) return fail(st);
        }
    }
#undef S_TRY
    // the descriptor's host pointers are not retained
    s->desc.rf_offset = s->desc.drift_offset = s->desc.fp_weights = s->desc.impedance = nullptr;
    s->desc.wakescaling = s->desc.grid = s->desc.particles = nullptr;
    *out = s;
    return IVS_OK;
}

ivs_status ivs_sim_destroy(ivs_sim* s) {
    if (!s) return IVS_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    drop_graphs(s);
    ivs_fused_plan_destroy(&s->plan);
    if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
    if (s->ev_step) cudaEventDestroy(s->ev_step);
    for (int b = 0; b < 2; b++) if (s->ev_track[b]) cudaEventDestroy(s->ev_track[b]);
    if (s->wakebuf[1] && s->wakebuf[1] != s->wake) cudaFree(s->wakebuf[1]);
    if (s->train_windows) cudaFree(s->train_windows);
    if (s->train_filling) cudaFree(s->train_filling);
    float* bufs[] = {s->grid[0], s->grid[1], s->grid[2], s->profile[0], s->profile[1], s->wake, s->rf,
                     s->drift, s->fpw, s->ws, s->impedance, s->rad_impedance, s->wakescaling,
                     (float*)s->wakefunction, s->particles, s->projy, s->obs, s->integral, s->moments, s->power,
                     (float*)s->track_counter, s->spectrum, s->wake_now, s->rf_seq, (float*)s->rf_pos};
    for (float* b : bufs) if (b) cudaFree(b);
    if (s->status) cudaFree(s->status);
    if (s->host_flag) cudaFreeHost(s->host_flag);
    delete s;
    return IVS_OK;
}

// enqueue nsteps steps: replays of the kGraphSteps-step graph of the current buffer parity plus
// direct launches for the remainder
static ivs_status enqueue_steps(ivs_sim* s, uint32_t nsteps) {
    ivs_ctx* ctx = s->ctx;
    const bool use_graph = !(s->desc.flags & IVS_SIM_NO_GRAPH);
    const bool unfused = (s->desc.flags & IVS_SIM_UNFUSED) != 0;
    const uint32_t lps = launches_per_step(s);
    uint32_t left = nsteps;
    // a run whose CTAs are all resident at once steps inside ONE persistent kernel (grid-wide barrier
    // instead of the kernel boundary); nothing else happens between two such steps
    if (!unfused && s->plan.resident && !s->np && !s->dyn_rf && left > 1) {
        const ivs_status st = ivs_fused_enqueue_run(ctx, &s->plan, s, left);
        if (st == IVS_OK) {
            s->steps_done += left;
            s->fused_enqueued += left;
            return IVS_OK;
        }
        if (st != IVS_ERR_UNSUPPORTED) return st;
        s->plan.resident = false;  // the cooperative launch was refused: one launch per step from now on
    }
    if (use_graph) {
        const uint32_t K = ivs_sim::kGraphSteps;  // even: a replay leaves the parity as it was
        while (left >= K) {
            cudaGraphExec_t& gk = s->graphK[s->cur];
            if (!gk) IVS_TRY(capture(s, K, &gk));
            IVS_CUDA(ctx, cudaGraphLaunch(gk, ctx->stream));
            ctx->launches += (uint64_t)lps * K;
            s->steps_done += K;
            if (!unfused) s->fused_enqueued += K;
            left -= K;
        }
        if (unfused) {
            while (left > 0) {
                cudaGraphExec_t& g1 = s->graph1[s->cur];
                if (!g1) IVS_TRY(capture(s, 1, &g1));
                IVS_CUDA(ctx, cudaGraphLaunch(g1, ctx->stream));
                ctx->launches += lps;
                s->steps_done++;
                left--;
            }
        }
    }
    while (left > 0) {
        IVS_TRY(enqueue_step(s));
        s->steps_done++;
        if (!unfused) s->fused_enqueued++;
        left--;
    }
    return ivs_fused_join(ctx, s);
}

static ivs_status reset_status(ivs_sim* s) {
    ivs_ctx* ctx = s->ctx;
    const int words[4] = {0x7fffffff, 0, 0, 0};
    IVS_CUDA(ctx, cudaMemcpyAsync(s->status, words, sizeof(words), cudaMemcpyHostToDevice, ctx->stream));
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *reinterpret_cast<volatile int*>(s->host_flag) = 0;
    s->fused_enqueued = 0;
    s->cur_at_reset = s->cur;
    return IVS_OK;
}

// Wait for the stream; if a fused step left the staged pad (see ivs_fused.cu, step_ctl_begin), rewind
// to the input of that step -- intact in the ping-pong buffer, since every later step exited without
// writing --, redo it through the kernel sequence and re-enqueue the steps after it.  On return
// the state is what `steps_done` steps of the reference's loop produce.
static ivs_status settle(ivs_sim* s) {
    ivs_ctx* ctx = s->ctx;
    for (;;) {
        IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if ((s->desc.flags & IVS_SIM_UNFUSED) || s->fused_enqueued == 0) return IVS_OK;
        if (s->dyn_rf) {
            unsigned int w[2];
            IVS_CUDA(ctx, cudaMemcpy(w, s->rf_pos, sizeof(w), cudaMemcpyDeviceToHost));
            if (w[1] != 0)
                return ivs_fail(ctx, IVS_ERR_HALO, "a dynamic RF kick displacement exceeds the fused kernel's staged pad: pass IVS_SIM_UNFUSED");
        }
        if (*reinterpret_cast<volatile int*>(s->host_flag) == 0) {
            // clean; keep the 32-bit ticket counter far from wrapping
            if (s->fused_enqueued > (1u << 20)) IVS_TRY(reset_status(s));
            return IVS_OK;
        }
        int words[4];
        IVS_CUDA(ctx, cudaMemcpy(words, s->status, sizeof(words), cudaMemcpyDeviceToHost));
        const uint64_t good = (uint64_t)words[0];  // fused steps since the reset that are valid
        if (words[0] < 0 || good >= s->fused_enqueued)
            return ivs_fail(ctx, IVS_ERR_HALO, "inconsistent overflow bookkeeping of the fused step kernel");
        const uint64_t redo = s->fused_enqueued - good;  // the offending step and those after it
        s->cur = s->pcur = s->cur_at_reset ^ (int)(good & 1);
        s->steps_done -= redo;
        IVS_TRY(reset_status(s));
        if (!s->grid[2]) IVS_TRY(dalloc(ctx, &s->grid[2], (size_t)s->members * s->n * s->n));
        if (s->dyn_rf) {
            // the queue position follows the steps that count: the offending step's entry is next again
            const unsigned int pos = (unsigned int)std::min<uint64_t>(s->steps_done - s->rf_step0, s->rf_seq_len);
            IVS_CUDA(ctx, cudaMemcpy(s->rf_pos, &pos, sizeof(pos), cudaMemcpyHostToDevice));
            IVS_TRY(ivs_dynrf_launch(ctx, &s->rf_model, s->rf_seq, s->rf_seq_len, s->rf_pos, nullptr, s->n, s->it,
                                     s->rf, (int*)s->plan.rf_first, s->plan.rf_w, (int*)(s->rf_pos + 1), -16, 12));
        }
        IVS_TRY(enqueue_unfused_step(s));
        // the kernel sequence leaves its wake in buffer 0: both buffers hold the last step's wake
        if (s->wakebuf[1] != s->wake)
            IVS_CUDA(ctx, cudaMemcpyAsync(s->wakebuf[1], s->wake, (size_t)s->members * s->n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        s->steps_done++;
        s->fallback_steps++;
        if (redo > 1) IVS_TRY(enqueue_steps(s, (uint32_t)(redo - 1)));
    }
}

ivs_status ivs_sim_step(ivs_sim* s, uint32_t nsteps) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));
    // an overflow seen so far (no sync: the flag is host-mapped) is repaired before more is enqueued
    if (s->host_flag && (*reinterpret_cast<volatile int*>(s->host_flag) != 0 ||
                         s->fused_enqueued + nsteps > (1u << 21)))
        IVS_TRY(settle(s));
    return enqueue_steps(s, nsteps);
}

// the wake buffer the last executed step filled (fused steps with tracked particles alternate)
static float* last_wake(const ivs_sim* s) { return s->track_overlap ? s->wakebuf[s->cur ^ 1] : s->wake; }

uint64_t ivs_sim_steps_done(const ivs_sim* s) { return s ? s->steps_done : 0; }
uint64_t ivs_sim_fallback_steps(const ivs_sim* s) { return s ? s->fallback_steps : 0; }

ivs_status ivs_sim_state(ivs_sim* s, void** d_grid, void** d_profile, void** d_wake,
                         void** d_particles) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    IVS_TRY(settle(s));
    if (d_grid) *d_grid = s->grid[s->cur];
    if (d_profile) *d_profile = s->profile[s->pcur];
    if (d_wake) *d_wake = last_wake(s);
    if (d_particles) *d_particles = s->particles;
    return IVS_OK;
}

ivs_status ivs_sim_download(ivs_sim* s, float* grid, float* profile, float* wake, float* particles) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    const size_t n = s->n, m = s->members;
    IVS_TRY(settle(s));
    if (grid) IVS_CUDA(ctx, cudaMemcpyAsync(grid, s->grid[s->cur], m * n * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (profile) IVS_CUDA(ctx, cudaMemcpyAsync(profile, s->profile[s->pcur], m * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (wake) IVS_CUDA(ctx, cudaMemcpyAsync(wake, last_wake(s), m * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (particles && s->np) IVS_CUDA(ctx, cudaMemcpyAsync(particles, s->particles, m * s->np * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    IVS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IVS_OK;
}

ivs_status ivs_sim_upload_grid(ivs_sim* s, const float* grid) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, grid != nullptr, "ivs_sim_upload_grid: null pointer");
    const size_t n = s->n, m = s->members;
    if (s->host_flag && *reinterpret_cast<volatile int*>(s->host_flag) != 0) IVS_TRY(settle(s));
    IVS_CUDA(ctx, cudaMemcpyAsync(s->grid[s->cur], grid, m * n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return ivs_project_x_batched(ctx, s->grid[s->cur], s->n, s->members, s->ws, s->profile[s->pcur]);
}

ivs_status ivs_sim_set_radiation(ivs_sim* s, const float* rad_impedance, float f_scale, float cutoff) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, rad_impedance != nullptr, "ivs_sim_set_radiation: null pointer");
    uint32_t nmax = s->nmax ? s->nmax : s->desc.nmax;
    IVS_REQUIRE(ctx, nmax != 0 && (nmax & (nmax - 1)) == 0 && nmax <= 16384 && nmax >= s->n,
                "ivs_sim_set_radiation: the simulation needs a power-of-two nmax");
    IVS_TRY(ivs_field_prepare(ctx, nmax));
    if (!s->rad_impedance) IVS_TRY(dalloc(ctx, &s->rad_impedance, (size_t)2 * nmax));
    IVS_TRY(ivs_upload(ctx, s->rad_impedance, rad_impedance, (size_t)2 * nmax * sizeof(float)));
    IVS_TRY(ivs_sync(ctx));
    s->has_rad = true; s->f_scale = f_scale; s->cutoff = cutoff;
    if (!s->nmax) s->nmax = nmax;
    return IVS_OK;
}

// out-step block of src/main.cpp:961-1033: integrate, variance(0), updateYProjection,
// variance(1), updateCSR -- all on the device, 8 floats per member come back
ivs_status ivs_sim_observe(ivs_sim* s, float* obs) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, obs != nullptr, "ivs_sim_observe: null pointer");
    const uint32_t n = s->n, m = s->members;
    IVS_TRY(settle(s));
    const float* g = s->grid[s->cur];
    const float* prof = s->profile[s->pcur];
    IVS_TRY(ivs_project_y_batched(ctx, g, n, m, s->ws, s->projy));
    if (s->has_rad) {
        // one launch: CSR spectrum + power of the radiation impedance, integral and the moments of
        // both projections, 8 floats per member
        ivs_field_obs o{s->obs, s->projy, s->ws, s->desc.qmin, s->desc.pmin, s->desc.dp};
        IVS_TRY(ivs_field_launch(ctx, prof, n, m, s->rad_impedance, s->nmax, nullptr, 0, nullptr, nullptr,
                                 nullptr, s->power, s->desc.dq, s->f_scale, s->cutoff, 1, &o));
        return ivs_download(ctx, obs, s->obs, (size_t)8 * m * sizeof(float));
    }
    IVS_TRY(ivs_integrate_batched(ctx, prof, n, m, s->ws, s->integral));
    IVS_TRY(ivs_moments_batched(ctx, prof, n, m, s->desc.qmin, s->desc.dq, s->integral, s->moments));
    IVS_TRY(ivs_moments_batched(ctx, s->projy, n, m, s->desc.pmin, s->desc.dp, s->integral, s->moments + (size_t)3 * m));
    std::vector<float> integral(m), mom((size_t)6 * m);
    IVS_TRY(ivs_download(ctx, integral.data(), s->integral, m * sizeof(float)));
    IVS_TRY(ivs_download(ctx, mom.data(), s->moments, (size_t)6 * m * sizeof(float)));
    for (uint32_t i = 0; i < m; i++) {
        obs[i * 8 + 0] = integral[i];
        for (int k = 0; k < 3; k++) {
            obs[i * 8 + 1 + k] = mom[(size_t)i * 3 + k];
            obs[i * 8 + 4 + k] = mom[(size_t)3 * m + (size_t)i * 3 + k];
        }
        obs[i * 8 + 7] = 0.0f;
    }
    return IVS_OK;
}

// ElectricField::updateCSR's spectrum (src/PS/ElectricField.cpp:133-184) and
// ElectricField::wakePotential (:186-253) of the CURRENT bunch profile -- what main() appends to
// /CSR/Spectrum/data and /WakePotential/data at an out-step (src/main.cpp:949-952,1000-1001)
ivs_status ivs_sim_fields(ivs_sim* s, float* csr_spectrum, float* wake) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    const uint32_t n = s->n, m = s->members;
    IVS_TRY(settle(s));
    const float* prof = s->profile[s->pcur];
    if (csr_spectrum) {
        IVS_REQUIRE(ctx, s->has_rad, "ivs_sim_fields: the CSR spectrum needs ivs_sim_set_radiation");
        if (!s->spectrum) IVS_TRY(dalloc(ctx, &s->spectrum, (size_t)m * s->nmax));
        IVS_TRY(ivs_field_launch(ctx, prof, n, m, s->rad_impedance, s->nmax, nullptr, 0, nullptr, nullptr,
                                 s->spectrum, s->power, s->desc.dq, s->f_scale, s->cutoff, 1));
        IVS_TRY(ivs_download(ctx, csr_spectrum, s->spectrum, (size_t)m * s->nmax * sizeof(float)));
    }
    if (wake) {
        if (s->has_wake) {
            if (!s->wake_now) IVS_TRY(dalloc(ctx, &s->wake_now, (size_t)m * n));
            if (s->train_windows) {
                IVS_TRY(ivs_train_wake_launch(ctx, prof, n, m, s->train_windows, s->wakescaling, 0.0f, s->wake_now));
            } else {
                IVS_TRY(ivs_field_launch(ctx, prof, n, m, s->impedance, s->nmax, s->wakescaling, 0, s->wake_now,
                                         nullptr, nullptr, nullptr, 0, 0, 0, 0));
            }
            IVS_TRY(ivs_download(ctx, wake, s->wake_now, (size_t)m * n * sizeof(float)));
        } else {
            for (size_t i = 0; i < (size_t)m * n; i++) wake[i] = 0.0f;
        }
    }
    return IVS_OK;
}

ivs_status ivs_sim_renormalize(ivs_sim* s) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_TRY(settle(s));
    IVS_TRY(ivs_integrate_batched(ctx, s->profile[s->pcur], s->n, s->members, s->ws, s->integral));
    // (a train: every bunch back to the charge fraction it carried when the train was set, PhaseSpace::normalize
    // with `filling`, src/PS/PhaseSpace.cpp:342-362)
    return ivs_normalize_batched(ctx, s->grid[s->cur], s->n, s->members, 1.0f, s->integral, s->train_filling);
}

ivs_status ivs_sim_set_rf_offset(ivs_sim* s, const float* rf_offset) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, rf_offset != nullptr, "ivs_sim_set_rf_offset: null pointer");
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));
    IVS_TRY(settle(s));
    if (!(s->desc.flags & IVS_SIM_UNFUSED)) IVS_TRY(ivs_fused_plan_set_rf(ctx, &s->plan, s, rf_offset));
    // the stream orders the copy after the steps already enqueued; the host array is consumed here
    IVS_TRY(ivs_upload(ctx, s->rf, rf_offset, s->n * sizeof(float)));
    return ivs_sync(ctx);
}

ivs_status ivs_sim_set_rf_modulation(ivs_sim* s, const ivs_rf_model* model, const float* phase_ampl,
                                     uint32_t count) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, count == 0 || (model && phase_ampl), "ivs_sim_set_rf_modulation: null pointer");
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));
    IVS_TRY(settle(s));
    const bool was = s->dyn_rf;
    if (s->rf_seq) { cudaFree(s->rf_seq); s->rf_seq = nullptr; }
    s->dyn_rf = count != 0;
    // the per-step RF kernel rewrites the offsets the particles read: their step runs in line then
    if (s->side) {
        if (s->track_overlap && s->dyn_rf && s->wakebuf[s->cur ^ 1] != s->wake)  // the last step's wake moves to buffer 0
            IVS_CUDA(ctx, cudaMemcpyAsync(s->wake, s->wakebuf[1], (size_t)s->members * s->n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        s->track_overlap = !s->dyn_rf;
    }
    s->rf_seq_len = count;
    s->rf_step0 = s->steps_done;
    if (s->dyn_rf) {
        s->rf_model = *model;
        IVS_TRY(dalloc(ctx, &s->rf_seq, (size_t)2 * count));
        IVS_TRY(ivs_upload(ctx, s->rf_seq, phase_ampl, (size_t)2 * count * sizeof(float)));
        if (!s->rf_pos) IVS_TRY(dalloc(ctx, &s->rf_pos, 4));
        IVS_TRY(ivs_memset_zero(ctx, s->rf_pos, 4 * sizeof(unsigned int)));
        IVS_TRY(ivs_sync(ctx));
    }
    if (was != s->dyn_rf) {
        // the step graphs gain or lose a node per step
        drop_graphs(s);
        if (!(s->desc.flags & IVS_SIM_NO_GRAPH)) {
            const bool unfused = (s->desc.flags & IVS_SIM_UNFUSED) != 0;
            const int cur = s->cur;
            IVS_TRY(capture(s, ivs_sim::kGraphSteps, &s->graphK[cur]));
            if (unfused) {
                IVS_TRY(capture(s, 1, &s->graph1[cur]));
            } else {
                s->cur = s->pcur = cur ^ 1;
                ivs_status st = capture(s, ivs_sim::kGraphSteps, &s->graphK[cur ^ 1]);
                s->cur = s->pcur = cur;
                if (st != IVS_OK) return st;
            }
        }
    }
    return IVS_OK;
}

ivs_status ivs_sim_set_bunch_train(ivs_sim* s, const uint32_t* bucket, uint32_t spacing_bins, const float* z,
                                   uint32_t nmax, float wakescaling) {
    if (!s) return IVS_ERR_INVALID_ARGUMENT;
    ivs_ctx* ctx = s->ctx;
    IVS_REQUIRE(ctx, bucket && z, "ivs_sim_set_bunch_train: null pointer");
    const bool unfused = (s->desc.flags & IVS_SIM_UNFUSED) != 0;
    if (!s->has_wake || !(unfused || s->plan.stream8))
        return ivs_fail(ctx, IVS_ERR_UNSUPPORTED,
                        "ivs_sim_set_bunch_train: a train of bunches needs a wake impedance and either the column-block step "
                        "kernel (N a multiple of 256 up to 2048) or the kernel sequence (IVS_SIM_UNFUSED)");
    IVS_REQUIRE(ctx, nmax >= 2 && !(nmax & (nmax - 1)), "ivs_sim_set_bunch_train: nmax must be a power of two");
    const uint32_t n = s->n, m = s->members;
    std::vector<double> w((size_t)m * m * 2 * n);
    const int rc = ivs_train_wake_windows(z, nmax, n, m, bucket, spacing_bins, w.data());
    IVS_REQUIRE(ctx, rc == 0, "ivs_sim_set_bunch_train: a bunch does not fit the padded profile "
                              "(bucket * spacing_bins + n > nmax)");
    IVS_CUDA(ctx, cudaSetDevice(ctx->device));
    IVS_TRY(settle(s));
    const bool was = s->train_windows != nullptr;
    if (!s->train_windows) IVS_TRY(dalloc(ctx, &s->train_windows, w.size()));
    IVS_TRY(ivs_upload(ctx, s->train_windows, w.data(), w.size() * sizeof(double)));
    // the charge fractions the grids carry now are what ivs_sim_renormalize restores
    if (!s->train_filling) IVS_TRY(dalloc(ctx, &s->train_filling, m));
    IVS_TRY(ivs_integrate_batched(ctx, s->profile[s->pcur], n, m, s->ws, s->integral));
    IVS_CUDA(ctx, cudaMemcpyAsync(s->train_filling, s->integral, m * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<float> sc(m, wakescaling);  // one ElectricField::_wakescaling for the whole beam
    IVS_TRY(ivs_upload(ctx, s->wakescaling, sc.data(), m * sizeof(float)));
    IVS_TRY(ivs_sync(ctx));
    s->plan.resident = false;  // (the persistent kernel keeps its members independent)
    if (!was) {  // the step graphs call another wake kernel / another instantiation of the step kernel
        drop_graphs(s);
        if (!(s->desc.flags & IVS_SIM_NO_GRAPH)) {
            const int cur = s->cur;
            IVS_TRY(capture(s, ivs_sim::kGraphSteps, &s->graphK[cur]));
            if (unfused) {
                IVS_TRY(capture(s, 1, &s->graph1[cur]));
            } else {
                s->cur = s->pcur = cur ^ 1;
                ivs_status st = capture(s, ivs_sim::kGraphSteps, &s->graphK[cur ^ 1]);
                s->cur = s->pcur = cur;
                if (st != IVS_OK) return st;
            }
        }
    }
    return IVS_OK;
}

}  // extern "C"
