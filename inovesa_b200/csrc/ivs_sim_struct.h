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
    float* wake = nullptr;        // [members][n] offsets applied by the last wake kick
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
    int* status = nullptr;        // device error flag of the fused kernel
    unsigned int* track_counter = nullptr;  // device step counter of the stochastic tracking mode (+ block ticket)
    float f_scale = 0, cutoff = 0;
    uint64_t steps_done = 0;
    ivs_fused_plan plan{};
    cudaGraphExec_t graph1 = nullptr;   // one step
    cudaGraphExec_t graphK = nullptr;   // kGraphSteps steps
    int graph_cur = -1;                 // buffer parity the graphs were captured for
    static constexpr uint32_t kGraphSteps = 16;
};

