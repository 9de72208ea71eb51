/* ivs.h -- C ABI of the B200-native Vlasov-Fokker-Planck step ("Inovesa hot path").
 *
 * This is the drop-in boundary: a thin CUDA handler that takes the place of the reference's
 * OpenCL handler (OCLH, /root/reference/inc/CL/OpenCLHandler.hpp:63-314) and of its clFFT/FFTW
 * wrappers (/root/reference/inc/FFTWWrapper.hpp:24-107).  Every entry point names the reference
 * interface it replaces.  Plain pointers and sizes only; no C++ types, no exceptions, no torch.
 *
 * Conventions
 *  - every function returns an ivs_status (0 = ok); ivs_last_error() gives the message
 *  - `void* d_*` arguments are DEVICE pointers obtained from ivs_malloc()
 *  - grids are fp32, laid out f[x][y] with y (energy) contiguous, exactly like
 *    PhaseSpace::_data (/root/reference/inc/PS/PhaseSpace.hpp:445-447); one bunch per grid
 *  - all work is asynchronous on the context's stream (the reference's single in-order
 *    queue, /root/reference/src/CL/OpenCLHandler.cpp:56-60); ivs_sync() == OCLH::finish()
 *  - arithmetic is fp32 and follows the reference's CPU path operation by operation
 *    (see DESIGN.md "Arithmetic contract"); IVS_ARITH_FMA contracts a*b+c exactly where
 *    GCC contracts the reference's loops when built with its own flags (-O3 -march=native),
 *    IVS_ARITH_STRICT rounds every product (a reference build with -ffp-contract=off).
 */
#ifndef IVS_H
#define IVS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int ivs_status;
enum {
    IVS_OK = 0,
    IVS_ERR_INVALID_ARGUMENT = 1, /* mirrors std::invalid_argument of the reference */
    IVS_ERR_CUDA = 2,             /* mirrors cl::Error */
    IVS_ERR_UNSUPPORTED = 3,
    IVS_ERR_HALO = 4,             /* a displacement exceeded what the fused kernel staged */
    IVS_ERR_NO_DEVICE = 5
};

enum { IVS_ARITH_FMA = 0, IVS_ARITH_STRICT = 1 };

typedef struct ivs_ctx ivs_ctx; /* replaces oclhptr_t (inc/CL/OpenCLHandler.hpp:36-41) */

/* ------------------------------------------------------------------ context / memory */

/* OCLH::OCLH(device,...) src/CL/OpenCLHandler.cpp:23-92.  `device` is the CUDA ordinal. */
ivs_status ivs_ctx_create(int device, int arith, ivs_ctx** out);
/* OCLH::~OCLH src/CL/OpenCLHandler.cpp:163-169 (finish, then release) */
ivs_status ivs_ctx_destroy(ivs_ctx* ctx);
/* message of the last failing call on this context ("" if none); ctx may be NULL for
 * errors of ivs_ctx_create */
const char* ivs_last_error(const ivs_ctx* ctx);
/* OCLH::listCLDevices src/CL/OpenCLHandler.cpp:171-200: number of CUDA devices (0 if none) */
int ivs_device_count(void);
/* library version string and the CUDA arch it was built for */
const char* ivs_version(void);
/* the raw cudaStream_t of the context (for callers that interleave their own work) */
void* ivs_ctx_stream(ivs_ctx* ctx);
/* number of kernels this library launched on the context since creation */
uint64_t ivs_ctx_launch_count(const ivs_ctx* ctx);

/* cl::Buffer(context, CL_MEM_READ_WRITE, bytes) */
ivs_status ivs_malloc(ivs_ctx* ctx, size_t bytes, void** d_ptr);
ivs_status ivs_free(ivs_ctx* ctx, void* d_ptr);
/* OCLH::enqueueWriteBuffer inc/CL/OpenCLHandler.hpp:236-262 (blocking=CL_TRUE semantics
 * for pageable memory; asynchronous if `host` is pinned via ivs_host_alloc) */
ivs_status ivs_upload(ivs_ctx* ctx, void* d_dst, const void* host, size_t bytes);
/* OCLH::enqueueReadBuffer inc/CL/OpenCLHandler.hpp:209-234 (always blocking, as CL_TRUE) */
ivs_status ivs_download(ivs_ctx* ctx, void* host, const void* d_src, size_t bytes);
/* OCLH::enqueueCopyBuffer inc/CL/OpenCLHandler.hpp:190-207 */
ivs_status ivs_copy(ivs_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
ivs_status ivs_memset_zero(ivs_ctx* ctx, void* d_dst, size_t bytes);
/* pinned host staging memory for the asynchronous transfers of the e2e path */
ivs_status ivs_host_alloc(ivs_ctx* ctx, size_t bytes, void** host_ptr);
ivs_status ivs_host_free(ivs_ctx* ctx, void* host_ptr);
/* OCLH::flush / OCLH::finish inc/CL/OpenCLHandler.hpp:283-288 */
ivs_status ivs_flush(ivs_ctx* ctx);
ivs_status ivs_sync(ivs_ctx* ctx);

/* ------------------------------------------------------------------ source maps (unfused ops)
 * 1:1 replacements of the reference's per-map kernels; results follow the CPU loops.   */

/* KickMap::apply(), Axis::y  (CPU src/SM/KickMap.cpp:243-266; OpenCL kernel apply_yKick
 * :113-163).  out[x][y] = sum_j w_j(x) in[x][y+first(x)+j], taps outside the grid dropped.
 * d_offset holds n floats: KickMap::_offset (what swapOffset()/getForce() exchange,
 * inc/SM/KickMap.hpp:47-55); the tap table of KickMap::updateSM() (src/SM/KickMap.cpp:339-389)
 * is rebuilt on the fly with the reference's fp32 arithmetic.  it = InterpolationType 1..4. */
ivs_status ivs_kick_y(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n,
                      const void* d_offset, uint32_t it);
/* KickMap::apply(), Axis::x  (CPU src/SM/KickMap.cpp:221-242; kernel apply_xKick :63-110) */
ivs_status ivs_kick_x(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n,
                      const void* d_offset, uint32_t it);
/* FokkerPlanckMap::apply() (CPU src/SM/FokkerPlanckMap.cpp:170-213; kernel applySM_Y :122-141).
 * d_weights: n rows of 5 floats = stencil weights on y-2..y+2 (built by
 * ivs_fokker_planck_weights from the constructor arithmetic, src/SM/FokkerPlanckMap.cpp:31-118). */
ivs_status ivs_fokker_planck(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n,
                             const void* d_weights);
/* host helper: fills weights[n*5] like the FokkerPlanckMap constructor
 * (fptype 0..3 = FPType, derivation 3|4 = DerivationType; e1 = damping decrement;
 * pmin/pmax describe the energy axis Ruler).  Pure host arithmetic, no GPU needed. */
ivs_status ivs_fokker_planck_weights(uint32_t n, float pmin, float pmax, float e1,
                                     uint32_t fptype, uint32_t derivation, float* weights);
/* RotationMap::apply() (CPU src/SM/RotationMap.cpp:97-176 with genHInfo :190-260; kernels
 * applyRotation :321-438 / applySM4sat :264-319).  Axis rulers are given by (min,max) per
 * axis; clamp != 0 requires it == 4 (std::invalid_argument otherwise, :83-87) */
ivs_status ivs_rotate(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n, float angle,
                      float qmin, float qmax, float pmin, float pmax, uint32_t it, int clamp);
/* SourceMap::apply() generic gather (CPU src/SM/SourceMap.cpp:75-109; kernel applySM1D
 * :121-139): out[i] = sum_{j<ip} in[index[i*ip+j]]*weight[i*ip+j]; table = array of
 * {uint32 index; float weight} (the reference's `hi`, inc/SM/SourceMap.hpp:37-40) */
ivs_status ivs_apply_source_map(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t ncells,
                                const void* d_hinfo, uint32_t ip);
/* Identity::apply() inc/SM/Identity.hpp:36-55 */
ivs_status ivs_identity(ivs_ctx* ctx, const void* d_in, void* d_out, uint32_t n);

/* ------------------------------------------------------------------ phase-space reductions */

/* PhaseSpace::simpsonWeights src/PS/PhaseSpace.cpp:486-501 (host arithmetic) */
ivs_status ivs_simpson_weights(uint32_t n, float delta, float* ws);
/* PhaseSpace::updateXProjection (CPU src/PS/PhaseSpace.cpp:313-321; kernel projectionX :538-550) */
ivs_status ivs_project_x(ivs_ctx* ctx, const void* d_grid, uint32_t n, const void* d_ws,
                         void* d_proj);
/* PhaseSpace::updateYProjection src/PS/PhaseSpace.cpp:325-340 (host loop in the reference) */
ivs_status ivs_project_y(ivs_ctx* ctx, const void* d_grid, uint32_t n, const void* d_ws,
                         void* d_proj);
/* PhaseSpace::integrate (CPU src/PS/PhaseSpace.cpp:229-238; kernels integral/integralSum
 * :509-532): *d_integral = sum_x proj[x]*ws[x] */
ivs_status ivs_integrate(ivs_ctx* ctx, const void* d_proj, uint32_t n, const void* d_ws,
                         void* d_integral);
/* PhaseSpace::normalize src/PS/PhaseSpace.cpp:342-372: grid *= filling_set / *d_integral */
ivs_status ivs_normalize(ivs_ctx* ctx, void* d_grid, uint32_t n, float filling_set,
                         const void* d_integral);
/* PhaseSpace::average + variance src/PS/PhaseSpace.cpp:241-289 on one projection:
 * d_out[0] = mean, d_out[1] = variance, d_out[2] = rms; axis values = amin + i*delta */
ivs_status ivs_moments(ivs_ctx* ctx, const void* d_proj, uint32_t n, float amin, float delta,
                       const void* d_integral, void* d_out3);

/* ------------------------------------------------------------------ electric field */

/* ElectricField::wakePotential (CPU src/PS/ElectricField.cpp:186-253; OpenCL chain
 * copy -> clFFT r2c -> wakeloss -> clFFT c2r -> scalewp :188-205): zero-pad the bunch
 * profile (n floats) to nmax, r2c FFT, multiply bins < nmax/2 by the impedance
 * (nmax complex<float>, upper half unused, Nyquist = 0), c2r FFT, scale by `wakescaling`
 * (already divided by nmax like ElectricField::_wakescaling :58) and keep the first n
 * samples.  One CTA, hand-written shared-memory FFT; nmax a power of two, n <= nmax <= 16384.
 * d_formfactor (nmax/2+1 complex, may be NULL) receives the spectrum of the profile. */
ivs_status ivs_wake_potential(ivs_ctx* ctx, const void* d_profile, uint32_t n,
                              const void* d_impedance, uint32_t nmax, float wakescaling,
                              void* d_wake, void* d_formfactor);
/* ElectricField::wakePotential for a TRAIN of nb bunches (PhaseSpace::nb > 1, src/PS/ElectricField.cpp:
 * 186-263 with padBunchProfiles :255-263; the filling pattern of src/main.cpp:248-321): d_profiles [nb][n]
 * floats (each bunch's X projection, carrying its charge fraction), d_windows = the nb * nb * 2 n doubles of
 * ivs_train_wake_windows (ivs_host.h) on the device, d_wake [nb][n] floats.  The reference transforms one
 * padded array of nmax = spaced_bins points per step; here the same circular convolution is a Toeplitz
 * product per pair of bunches against windows of the static wake function, accumulated in double
 * (agrees with the reference's fp32 transforms to their rounding, ~1e-6 of the largest value). */
ivs_status ivs_train_wake_potential(ivs_ctx* ctx, const void* d_profiles, uint32_t n, uint32_t nb,
                                    const void* d_windows, float wakescaling, void* d_wake);
/* The wake FUNCTION of the same operator, w[m], m = -(n-1)..(n-1) (2n-1 floats, w[0] at
 * index n-1), such that wakePotential()[x] == sum_x' profile[x'] * w[x - x'] exactly in
 * real arithmetic (the zero padding nmax >= 2n removes every circular wrap).  Used by the
 * fused step, which applies the kick from a direct Toeplitz product per row block. */
ivs_status ivs_wake_function(ivs_ctx* ctx, const void* d_impedance, uint32_t nmax,
                             float wakescaling, uint32_t n, void* d_wakefunction);
/* ElectricField::updateCSR src/PS/ElectricField.cpp:133-184: spectrum[i] (nmax floats) and
 * *d_power = sum_i delta_f * spectrum[i]; dq = grid spacing, f_scale = c/sigma_z ("Hertz"
 * scale of the frequency ruler), cutoff <= 0 disables the shielding factor */
ivs_status ivs_csr_power(ivs_ctx* ctx, const void* d_profile, uint32_t n,
                         const void* d_impedance, uint32_t nmax, float dq, float f_scale,
                         float cutoff, void* d_spectrum, void* d_power);

/* ------------------------------------------------------------------ passive particles */

/* SourceMap::applyToAll with KickMap::applyTo (src/SM/SourceMap.cpp:111-117,
 * src/SM/KickMap.cpp:280-303): particles = np pairs {x,y} in grid units
 * (PhaseSpace::Position); axis 0 = x-kick (drift), 1 = y-kick (RF, wake) */
ivs_status ivs_track_kick(ivs_ctx* ctx, void* d_particles, uint32_t np, const void* d_offset,
                          uint32_t n, int axis);
/* FokkerPlanckMap::applyTo src/SM/FokkerPlanckMap.cpp:214-268.  fptrack 0..3; mode 3
 * (stochastic) draws from a counter-based generator keyed by (seed, step, particle) --
 * the reference's unseeded mt19937 (:22-24) cannot be reproduced */
ivs_status ivs_track_fokker_planck(ivs_ctx* ctx, void* d_particles, uint32_t np,
                                   const void* d_grid, uint32_t n, const void* d_weights,
                                   uint32_t fptrack, float e1, float dp, uint64_t seed,
                                   uint64_t step);
/* RotationMap::applyTo src/SM/RotationMap.cpp:178-188 */
ivs_status ivs_track_rotate(ivs_ctx* ctx, void* d_particles, uint32_t np, float angle,
                            float zerobin_x, float zerobin_y);

/* ------------------------------------------------------------------ whole simulation step
 * The loop body of main() (src/main.cpp:948-1053) for one bunch:
 *   wake potential -> wake kick -> RF kick -> drift -> Fokker-Planck -> X projection
 * as ONE fused kernel per step (plus nothing else), captured in a CUDA graph.            */

typedef struct ivs_sim ivs_sim;

typedef struct ivs_sim_desc {
    uint32_t n;            /* GridSize (nx = ny) */
    uint32_t members;      /* independent simulations advanced in lock-step (ensemble); >= 1 */
    uint32_t it;           /* InterpolationPoints 1..4 */
    uint32_t nmax;         /* padded FFT length (power of two, >= 2n), 0 = no wake map */
    uint32_t n_particles;  /* tracked particles per member (0 = none) */
    uint32_t fptrack;      /* FPTracking for the particles */
    uint32_t flags;        /* IVS_SIM_* */
    uint32_t strips;       /* tuning: strips per member of the streaming step kernel (0 = planner's choice) */
    float dq;              /* axis 0 delta (Simpson weights are built from it, as the reference) */
    float dp;              /* axis 1 delta */
    float e1;              /* damping decrement (tracking only; the stencil is in fp_weights) */
    float qmin;            /* axis 0 minimum (moments) */
    float pmin;            /* axis 1 minimum (moments) */
    float reserved_f;
    const float* rf_offset;     /* host, n floats: RFKickMap::_offset (getForce()) */
    const float* drift_offset;  /* host, n floats: DriftMap::_offset */
    const float* fp_weights;    /* host, n*5 floats (ivs_fokker_planck_weights) or NULL = Identity */
    const float* impedance;     /* host, nmax complex<float> (wake impedance) or NULL */
    const float* wakescaling;   /* host, `members` floats: ElectricField::_wakescaling per member */
    const float* grid;          /* host, members*n*n floats: initial density (grid_t1) */
    const float* particles;     /* host, members*n_particles*2 floats or NULL */
    uint64_t seed;
} ivs_sim_desc;

enum {
    /* 1u: reserved (was IVS_SIM_WAKE_FFT, never implemented: the kernel sequence of IVS_SIM_UNFUSED is the
       path that forms the wake potential with the FFT kernel) */
    IVS_SIM_UNFUSED = 2u,       /* run the step as the reference's kernel sequence (debug / parity) */
    IVS_SIM_NO_GRAPH = 4u,      /* launch kernels directly instead of replaying a CUDA graph */
    IVS_SIM_TILED = 8u,         /* fused step as the tile kernel instead of the streaming pipeline */
    IVS_SIM_QPT1 = 16u,         /* tuning: the streaming kernel (shared TMA ring) with one quad of columns per thread ... */
    IVS_SIM_QPT2 = 32u,         /* ... or two, instead of the column-block kernel */
    IVS_SIM_NO_PERSISTENT = 256u, /* one launch per step (CUDA graph) even where the whole run fits the machine at once and
                                    ivs_sim_step would run its steps inside one persistent kernel (same results);
                                    implied by IVS_SIM_TILED, _QPT1, _QPT2, _STREAM8 and _NO_GRAPH */
    IVS_SIM_TRACK_INLINE = 128u, /* tracked particles step in the context's stream after every fused step instead of
                                   on a side stream beside the next step (same results) */
    IVS_SIM_STREAM8 = 64u       /* the column-block kernel (8 adjacent columns per lane, warp-private rings): the
                                   default for grid sizes that are multiples of 256 up to 2048 */
};

ivs_status ivs_sim_create(ivs_ctx* ctx, const ivs_sim_desc* desc, ivs_sim** out);
ivs_status ivs_sim_destroy(ivs_sim* sim);
/* Advance every member by nsteps (asynchronous): the loop body of main() (src/main.cpp:948-1053).
 * A run that fits the machine at once -- grid up to 1024 x 1024, at most one block of rows per SM and
 * member, no tracked particles, no DynamicRFKickMap -- does all nsteps inside ONE cooperative launch
 * of the persistent kernel (csrc/ivs_resident.cu); everything else replays 16-step CUDA graphs of the
 * fused step kernel (+ tracking / DynamicRFKickMap kernels), captured when the simulation was
 * created.  The results do not depend on the route or on how the steps are split over calls. */
ivs_status ivs_sim_step(ivs_sim* sim, uint32_t nsteps);
/* steps taken so far */
uint64_t ivs_sim_steps_done(const ivs_sim* sim);
/* The fused step kernel stages 32 zero columns beside every row, which bounds the wake kick's
 * first tap to [-16, 8] cells; KickMap::updateSM (src/SM/KickMap.cpp:339-389) has no such limit.
 * A step that exceeds it is detected on the device, redone through the unfused kernel sequence
 * (any displacement) and the steps after it are re-enqueued -- at the next call that touches the
 * simulation, transparently.  This counts the steps that were redone that way. */
uint64_t ivs_sim_fallback_steps(const ivs_sim* sim);
/* device pointers to the current state (valid until the next ivs_sim_step):
 * grid [members][n][n], x-projection [members][n], wake potential [members][n]
 * (the offsets the last step applied), particles [members][np][2] */
ivs_status ivs_sim_state(ivs_sim* sim, void** d_grid, void** d_profile, void** d_wake,
                         void** d_particles);
/* PhaseSpace::syncCLMem(dev2cpu) src/PS/PhaseSpace.cpp:395-412: blocking copies to host
 * (any pointer may be NULL) */
ivs_status ivs_sim_download(ivs_sim* sim, float* grid, float* profile, float* wake,
                            float* particles);
/* replace the density of all members (PhaseSpace::syncCLMem(cpu2dev)); recomputes the
 * X projection on the device */
ivs_status ivs_sim_upload_grid(ivs_sim* sim, const float* grid);
/* out-step observables of main.cpp:961-1033 for every member, computed on the device:
 * obs[m*8 + {0: integral, 1: mean_q, 2: var_q, 3: sigma_q, 4: mean_p, 5: var_p, 6: sigma_p,
 * 7: csr_power}]; csr needs `rad_impedance` given to ivs_sim_set_radiation */
ivs_status ivs_sim_observe(ivs_sim* sim, float* obs);
ivs_status ivs_sim_set_radiation(ivs_sim* sim, const float* rad_impedance, float f_scale,
                                 float cutoff);
/* the field rows main() appends at an out-step (src/main.cpp:949-952,1000-1001; HDF5File::append(ef)
 * src/IO/HDF5File.cpp:360-423), for the CURRENT bunch profile of every member, to HOST arrays:
 * csr_spectrum [members][nmax] = ElectricField::updateCSR's spectrum (src/PS/ElectricField.cpp:133-184),
 * wake [members][n] = ElectricField::wakePotential (:186-253).  Either may be NULL. */
ivs_status ivs_sim_fields(ivs_sim* sim, float* csr_spectrum, float* wake);
/* PhaseSpace::integrateAndNormalize (main.cpp:953-956) on every member */
ivs_status ivs_sim_renormalize(ivs_sim* sim);
/* Replace the RF kick offsets (n floats) before the next step: the hook DynamicRFKickMap needs
 * (src/SM/DynamicRFKickMap.cpp:80-109 re-draws phase and amplitude and recomputes KickMap::_offset
 * on the host every apply()).  Takes effect for all following steps, also inside replayed graphs
 * (the device tables are updated in place).  IVS_ERR_HALO if a displacement exceeds what the
 * fused kernel stages. */
ivs_status ivs_sim_set_rf_offset(ivs_sim* sim, const float* rf_offset);

/* DynamicRFKickMap (src/SM/DynamicRFKickMap.cpp:14-109): the RF map whose phase and amplitude
 * change every step.  The constants of RFKickMap::_calcKick(phase, ampl) (src/SM/RFKickMap.cpp:59-83)
 * as the constructors derive them (:15-57); ivs_rf_model_init (ivs_host.h) fills it from ring
 * parameters.  linear != 0: offset[x] = (tan(angle) (xcenter - x) + tan(angle) (syncphase - phase) /
 * bl2phase / delta0) * ampl; else offset[x] = revolutionpart (-ampl V_RF sin(q_x bl2phase + phase) + V0)
 * / delta1 / dE_scale. */
typedef struct ivs_rf_model {
    uint32_t linear;
    float tan_angle;      /* std::tan(_angle), fp32 */
    float xcenter;        /* axis 0 zerobin */
    float syncphase;      /* 0 (linear) or asin(V0 / V_RF) */
    float bl2phase;       /* scale(0,"Meter") / c * f_RF * 2 pi */
    float qmin, delta0;   /* axis 0 */
    float delta1;         /* axis 1 delta */
    float dE_scale;       /* scale(1,"ElectronVolt") */
    float revolutionpart, V_RF, V0;
} ivs_rf_model;
/* Makes the RF kick dynamic: step i (counted from this call) applies _calcKick(phase_ampl[2 i],
 * phase_ampl[2 i + 1]) -- the queue DynamicRFKickMap::__calcModulation fills for all steps of the run
 * (:80-94; ivs_rf_modulation in ivs_host.h generates one from a seed).  The queue is copied to the
 * device and walked by a one-block kernel in front of every step, inside the step graph; past its end
 * the last offsets stay in force.  count = 0 returns to the static map set by ivs_sim_set_rf_offset /
 * ivs_sim_create.  The reference applies a DynamicRFKickMap on the CPU only (no OpenCL path, hpp:75). */
ivs_status ivs_sim_set_rf_modulation(ivs_sim* sim, const ivs_rf_model* model, const float* phase_ampl,
                                     uint32_t count);

/* Makes the members of a simulation the bunches of ONE train (PhaseSpace::nb = members): from the next
 * step on, the wake potential of member b is the train's (ivs_train_wake_potential: every bunch acts on
 * every other through the wake function at their distance bucket * spacing_bins), the other maps stay per
 * bunch as in SourceMap::apply's loop over PhaseSpace::nb.  bucket: `members` bucket numbers as
 * src/main.cpp:258-268 enumerates them; z: the wake impedance at the SPACED length nmax (:318-321, :718),
 * a power of two with bucket * spacing_bins + n <= nmax; wakescaling = ElectricField::_wakescaling of the
 * whole beam (Ib = sum of the bunch currents, already divided by nmax), applied to every bunch -- the
 * grids carry the charge fractions (src/PS/PhaseSpace.cpp:17-66 with `filling`).  A train steps through the
 * column-block step kernel (N a multiple of 256 up to 2048: its set-up forms the Toeplitz product once per
 * source bunch) or, with IVS_SIM_UNFUSED, through the kernel sequence; tracked particles of member b move in
 * bunch b's fields; it needs a wake impedance.  ivs_sim_renormalize restores the charge
 * fraction every bunch carries at this call (PhaseSpace::normalize with `filling`, src/PS/PhaseSpace.cpp:342-362);
 * else IVS_ERR_UNSUPPORTED.  The kick of bunch b uses bunch b's wake potential -- the entry
 * KickMap::updateSM stores for it (src/SM/KickMap.cpp:348-372: bunch n, row x at (n N + x) _ip + j).  The
 * reference's CPU loop reads the entries of bunches > 0 back from the wrong place (:244-259 lacks the factor
 * _ip in the bunch offset) and its OpenCL kernels hold a single bunch; that slip is not reproduced
 * (DESIGN.md section 10). */
ivs_status ivs_sim_set_bunch_train(ivs_sim* sim, const uint32_t* bucket, uint32_t spacing_bins, const float* z,
                                   uint32_t nmax, float wakescaling);

#ifdef __cplusplus
}
#endif
#endif /* IVS_H */
