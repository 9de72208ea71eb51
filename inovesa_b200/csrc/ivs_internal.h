// ivs_internal.h -- declarations shared between the translation units of libivs.so
// (not part of the C ABI).
#pragma once
#include "ivs_common.cuh"

#include <vector>

struct ivs_sim;

// unit wake function of an impedance table in double (ivs_host.cpp): w[m + n - 1], m = -(n-1)..(n-1)
std::vector<double> ivs_wake_function_host(const float* z, uint32_t nmax, uint32_t n);
// include/ivs_host.h
extern "C" int ivs_train_wake_windows(const float* z, uint32_t nmax, uint32_t n, uint32_t nb, const uint32_t* bucket,
                                      uint32_t spacing_bins, double* windows);

// member-batched launchers (ivs_maps.cu, ivs_reduce.cu, ivs_field.cu, ivs_track.cu)
ivs_status ivs_kick_y_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n, uint32_t members,
                              const float* offset, uint32_t offset_stride, uint32_t it);
ivs_status ivs_kick_x_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n, uint32_t members,
                              const float* offset, uint32_t offset_stride, uint32_t it);
ivs_status ivs_fokker_planck_batched(ivs_ctx* ctx, const float* in, float* out, uint32_t n,
                                     uint32_t members, const float* w5);
ivs_status ivs_project_x_batched(ivs_ctx* ctx, const float* grid, uint32_t n, uint32_t members,
                                 const float* ws, float* proj);
ivs_status ivs_project_y_batched(ivs_ctx* ctx, const float* grid, uint32_t n, uint32_t members,
                                 const float* ws, float* proj);
ivs_status ivs_integrate_batched(ivs_ctx* ctx, const float* proj, uint32_t n, uint32_t members,
                                 const float* ws, float* integral);
ivs_status ivs_normalize_batched(ivs_ctx* ctx, float* grid, uint32_t n, uint32_t members,
                                 float filling_set, const float* integral, const float* filling_sets = nullptr);
ivs_status ivs_moments_batched(ivs_ctx* ctx, const float* proj, uint32_t n, uint32_t members,
                               float amin, float delta, const float* integral, float* out3);
// out-step observables formed by the field kernel in CSR mode: obs[member*8 + {integral, mean_q,
// var_q, sigma_q, mean_p, var_p, sigma_p, csr_power}] from the X profile and projy
struct ivs_field_obs {
    float* obs;
    const float* projy;
    const float* ws;
    float qmin, pmin, dp;
};
ivs_status ivs_field_launch(ivs_ctx* ctx, const float* profile, uint32_t n, uint32_t members,
                            const float* impedance, uint32_t nmax, const float* d_scaling,
                            float scaling, float* wake, float* formfactor, float* spectrum,
                            float* power, float dq, float f_scale, float cutoff, int mode,
                            const ivs_field_obs* obs = nullptr);
// allocates the twiddle table for an FFT length (must not happen inside a graph capture)
ivs_status ivs_field_prepare(ivs_ctx* ctx, uint32_t nmax);
// all four maps' applyTo for every tracked particle of every member, one launch per step;
// step_counter: 4 device words {step lo, step hi, blocks done, -} kept by the kernel itself
ivs_status ivs_track_step_launch(ivs_ctx* ctx, float* particles, uint32_t np, uint32_t members,
                                 uint32_t n, const float* wake, const float* rf, const float* drift,
                                 const float* fpw, const float* grid_fp_in, uint32_t fptrack, float e1,
                                 float dp, uint64_t seed, unsigned int* step_counter,
                                 const int* sim_status = nullptr, int wb = -1, cudaStream_t stream = nullptr);

// DynamicRFKickMap: one block that turns the next (phase, amplitude) entry of the device queue into
// KickMap::_offset and the fused kernels' tap tables (ivs_dynrf.cu)
ivs_status ivs_dynrf_launch(ivs_ctx* ctx, const ivs_rf_model* model, const float* d_seq, uint32_t count,
                            unsigned int* d_pos, const int* sim_status, uint32_t n, uint32_t it, float* d_offset,
                            int* d_first, void* d_w, int* d_flag, int lo, int hi);

// the fused step kernel (ivs_fused.cu)
struct ivs_fused_plan {
    uint32_t launches_per_step = 1;
    uint32_t tx = 0;          // output rows per CTA
    uint32_t halo_lo = 0, halo_hi = 0;  // extra rows the drift reads below / above a tile
    uint32_t pad_y = 0;       // zero columns kept left and right of every staged row
    uint32_t threads = 0;
    size_t smem = 0;
    // device tables: first tap position and 4 weights per row / column
    void* rf_first = nullptr; void* rf_w = nullptr;
    void* dr_first = nullptr; void* dr_w = nullptr;
    void* fp_base = nullptr; void* fp_w = nullptr;
    bool vec = false;   // float4-per-lane kernel (grid size multiple of 256)
    bool stream = false;  // streaming kernel (independent warps, qpt quads of columns per thread)
    uint32_t qpt = 1;
    bool stream8 = false; // column-block kernel (ivs_stream8.cu): 8 adjacent columns per lane, private rings
    // the whole run fits the machine at once: many steps per (cooperative) launch of the persistent
    // kernel (ivs_resident.cu), res_rows output rows per CTA
    bool resident = false;
    uint32_t res_rows = 0, res_wwin = 0;
    size_t res_smem = 0;
    void* res_tags = nullptr;  // [2][members][n] row sums published step by step (value, step tag)
    uint32_t strip = 0, kslots = 0, max_strip = 0, max_rin = 0, wwin_len = 0;
    void* fp5 = nullptr;  // Fokker-Planck weights, out-of-grid taps zeroed (streaming kernel)
    bool ready = false;
};
ivs_status ivs_fused_plan_create(ivs_ctx* ctx, ivs_sim* s, ivs_fused_plan* plan);
void ivs_fused_plan_destroy(ivs_fused_plan* plan);
ivs_status ivs_fused_enqueue_step(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s);
// nsteps steps in one launch of the persistent kernel (plan->resident, no tracked particles, static RF)
ivs_status ivs_fused_enqueue_run(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s, uint32_t nsteps);
// tracking launches that overlap the following step (side stream) are joined into the context's stream
ivs_status ivs_fused_join(ivs_ctx* ctx, ivs_sim* s);
ivs_status ivs_train_wake_launch(ivs_ctx* ctx, const float* profiles, uint32_t n, uint32_t nb, const double* windows,
                                 const float* scaling, float scal, float* wake);
// new RF kick offsets: rebuild the fused kernels' tap tables in place (same device buffers)
ivs_status ivs_fused_plan_set_rf(ivs_ctx* ctx, ivs_fused_plan* plan, ivs_sim* s, const float* rf_offset);
