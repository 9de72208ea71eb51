// ivs_sim_struct.h -- the simulation object (internal; the C ABI only sees an opaque ivs_sim*)
#pragma once
#include "ivs_internal.h"

struct ivs_sim {
    ivs_ctx* ctx = nullptr;
    ivs_sim_desc desc{};
    uint32_t n = 0, members = 0, it = 4, nmax = 0, np = 0;
    bool has_wake = false, has_fp = false, has_rad = false;
    float* grid[3] = {nullptr, nullptr, nullptr};  // t1, t2, t3 (unfused) / ping, pong (fused)
    int cur = 0;                                   // index of the buffer holding grid_t1
    float* profile[2] = {nullptr, nullptr};        // X projection (double buffered for the fused path)
    int pcur = 0;
    float* wake = nullptr;        // [members][n] offsets applied by the last wake kick (== wakebuf[0])
    // Fused steps with tracked particles alternate between two wake buffers (index = cur of the step)
    // so that the tracking kernel of step t can run on a low-priority side stream while step t + 1
    // already fills the other buffer; ev_track[b] = "the tracking launch that reads buffer b is done"
    float* wakebuf[2] = {nullptr, nullptr};
    bool track_overlap = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_step = nullptr, ev_track[2] = {nullptr, nullptr};
    bool track_inflight[2] = {false, false};
    float* rf = nullptr;          // [n]
    float* drift = nullptr;       // [n]
    float* fpw = nullptr;         // [n][5]
    float* ws = nullptr;          // [n]
    float* impedance = nullptr;   // [nmax] complex
    float* rad_impedance = nullptr;
    float* wakescaling = nullptr; // [members]
    double* wakefunction = nullptr;// [2n] unit-scaled wake function in double (fused path)
    float* particles = nullptr;   // [members][np][2]
    float* projy = nullptr;       // [members][n]
    float* obs = nullptr;         // [members][8] + scratch
    float* integral = nullptr;    // [members]
    float* moments = nullptr;     // [members][2][3]
    float* power = nullptr;       // [members]
    float* spectrum = nullptr;    // [members][nmax] CSR spectrum (ivs_sim_fields, allocated on first use)
    float* wake_now = nullptr;    // [members][n] wake potential of the current profile (ivs_sim_fields)
    // bookkeeping of the fused kernels (ivs_fused.cu, step_ctl_begin): status[0] = ordinal of the first
    // step whose wake kick left the staged pad (0x7fffffff: none), status[1] = CTA tickets
    int* status = nullptr;
    int* host_flag = nullptr;     // host-mapped word, set together with status[0]: peeked without a sync
    uint64_t fused_enqueued = 0;  // fused steps enqueued since the status words were last reset ...
    int cur_at_reset = 0;         // ... and the buffer parity at that moment
    uint64_t fallback_steps = 0;  // steps redone through the kernel sequence after an overflow
    unsigned int* track_counter = nullptr;  // device step counter of the stochastic tracking mode (+ block ticket)
    // DynamicRFKickMap: model constants, device queue of (phase, amplitude) pairs, position in it
    // (device word 0; word 1 = "an RF tap left the staged pad") and the step count at which it was set
    bool dyn_rf = false;
    ivs_rf_model rf_model{};
    float* rf_seq = nullptr;
    uint32_t rf_seq_len = 0;
    unsigned int* rf_pos = nullptr;
    uint64_t rf_step0 = 0;
    // train of bunches (ivs_sim_set_bunch_train): the members are coupled through these wake-function windows
    double* train_windows = nullptr;  // [members][members][2 n]
    float* train_filling = nullptr;   // [members] the bunches' charge fractions (their integrals when the train was set)
    float f_scale = 0, cutoff = 0;
    uint64_t steps_done = 0;
    ivs_fused_plan plan{};
    // fused steps ping-pong two buffers: one graph per buffer parity (captured at creation)
    cudaGraphExec_t graph1[2] = {nullptr, nullptr};   // one step (unfused sequence only)
    cudaGraphExec_t graphK[2] = {nullptr, nullptr};   // kGraphSteps steps
    static constexpr uint32_t kGraphSteps = 16;
};

