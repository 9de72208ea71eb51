/* vfps_oracle.h -- CPU restatement of the reference's per-step Vlasov-Fokker-Planck update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under inovesa_b200/ may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it, as the checker.
 *
 * Plain C restatement of the algorithm of Inovesa's CPU (non-OpenCL) path, written from the
 * behaviour of the sources cited at each function (paths relative to /root/reference/).
 * Parity of this restatement is PINNED two ways (tests/test_oracle_*.py):
 *   1. against every golden vector of the reference's own unit tests for the path
 *      (test/TestKickMap.cpp, TestRotationMap.cpp, TestElectricField.cpp, TestPhaseSpace.cpp,
 *      TestConstImpedance.cpp, TestRuler.cpp), committed as tests/golden/reference_tests.json;
 *   2. bit-for-bit against the reference's own sources compiled in place (oracle/_ref/, see
 *      oracle/Makefile) on seeded inputs, and against fixtures generated from that build
 *      (tests/golden/*.npz with tests/golden/make_golden.py).
 * Arithmetic flavour: compiled twice, always with -ffp-contract=off; with -DORACLE_FMA the
 * multiply-adds that GCC contracts in the reference's loops at -O3 on FMA hardware are
 * written as explicit fmaf() (liboracle_fma.so), otherwise every product is rounded
 * (liboracle.so).
 */
#ifndef VFPS_ORACLE_H
#define VFPS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#include "../include/ivs_host.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 1 if built with -DORACLE_FMA */
int orc_is_fma(void);

/* Ruler (inc/PS/Ruler.hpp:31-51) */
float orc_ruler_delta(uint32_t steps, float min, float max);
float orc_ruler_zerobin(uint32_t steps, float min, float max);
void orc_ruler_axis(uint32_t steps, float min, float max, float* axis);

/* SourceMap::calcCoefficiants (src/SM/SourceMap.cpp:142-170) */
void orc_calc_coefficients(float f, uint32_t it, float* ic);

/* KickMap::updateSM (src/SM/KickMap.cpp:339-389): per offset entry `it` taps {index,weight} */
void orc_kick_update_sm(const float* offset, uint32_t n_entries, uint32_t n_kd, uint32_t it,
                        uint32_t* index, float* weight);
/* KickMap::apply (src/SM/KickMap.cpp:194-271); axis 0 = x, 1 = y */
void orc_kick_apply(uint32_t n, uint32_t it, int axis, const float* offset, const float* in,
                    float* out);
/* KickMap::applyTo (src/SM/KickMap.cpp:280-303) */
void orc_kick_track(uint32_t n, int axis, const float* offset, float* particles, uint32_t np);

/* RFKickMap::_calcKick (src/SM/RFKickMap.cpp:59-83) with the constructors' members (:15-57) */
void orc_rf_offsets_linear(uint32_t n, float qmin, float qmax, float bl_scale, float angle,
                           float f_RF, float* offset);
void orc_rf_offsets_nonlinear(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                              float bl_scale, float dE_scale, float revolutionpart, float V_RF,
                              float f_RF, float V0, float* offset);
/* DriftMap::DriftMap (src/SM/DriftMap.cpp:21-30) */
void orc_drift_offsets(uint32_t n, float qmin, float qmax, float pmin, float pmax,
                       float dE_scale, const float* slip, uint32_t nslip, float E0,
                       float* offset);

/* FokkerPlanckMap constructor (src/SM/FokkerPlanckMap.cpp:31-118): table of {index,weight},
 * `derivation` taps per row */
void orc_fp_table(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, uint32_t* index, float* weight);
/* the same as 5 dense weights on y-2..y+2 */
void orc_fp_weights5(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                     uint32_t derivation, float* w5);
/* FokkerPlanckMap::apply (src/SM/FokkerPlanckMap.cpp:170-213) */
void orc_fp_apply(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, const float* in, float* out);
/* FokkerPlanckMap::applyTo modes 0,1,2 (src/SM/FokkerPlanckMap.cpp:214-268) */
void orc_fp_track(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                  uint32_t derivation, uint32_t fptrack, const float* grid, float* particles,
                  uint32_t np);

/* RotationMap (src/SM/RotationMap.cpp:12-88 ctor checks, :97-176 apply, :190-260 genHInfo).
 * returns 1 for the invalid_argument cases (:83-87) */
int orc_rotation_apply(uint32_t n, float qmin, float qmax, float pmin, float pmax, float angle,
                       uint32_t it, int clamp, uint32_t rotmapsize, const float* in, float* out);
/* RotationMap::applyTo (src/SM/RotationMap.cpp:178-188) */
void orc_rotation_track(uint32_t n, float qmin, float qmax, float pmin, float pmax, float angle,
                        float* particles, uint32_t np);

/* PhaseSpace (src/PS/PhaseSpace.cpp): simpsonWeights :486-501, updateXProjection :313-321,
 * updateYProjection :325-340, integrate :229-238, normalize :342-372, average/variance :241-289,
 * gaus + createFromProjections :457-484 */
void orc_simpson_weights(uint32_t n, float delta, float* ws);
void orc_project_x(uint32_t n, const float* grid, const float* ws, float* proj);
void orc_project_y(uint32_t n, const float* grid, const float* ws, float* proj);
float orc_integrate(uint32_t n, const float* proj, const float* ws);
void orc_normalize(uint32_t n, float* grid, float filling_set, float filling);
void orc_moments(uint32_t n, const float* proj, const float* axis, float delta, float filling,
                 float* mean, float* var, float* rms);
void orc_initial_grid(uint32_t n, float qmin, float qmax, float pmin, float pmax, double zoom,
                      float* grid);

/* ElectricField (src/PS/ElectricField.cpp): wakePotential :186-253 (+padBunchProfiles :255-263),
 * updateCSR :133-184.  FFT = FFTW semantics (r2c forward, c2r backward, unnormalised) evaluated
 * in double and rounded to float, like oracle/fftw_shim.cpp.  z = nmax complex (re,im). */
void orc_wake_potential(uint32_t n, const float* profile, const float* z, uint32_t nmax,
                        float wakescaling, float* wake, float* formfactor, float* padded_wake);
void orc_csr_power(uint32_t n, const float* profile, const float* z, uint32_t nmax, float dq,
                   float f_scale, float cutoff, float* spectrum, float* power);

/* Impedance models (src/Z/*.cpp) and factory (src/Z/ImpedanceFactory.cpp:21-84) */
void orc_z_free_space(size_t n, float f_rev, float f_max, float* z);
void orc_z_parallel_plates(size_t n, float f0, float f_max, double g, float* z);
void orc_z_resistive_wall(size_t n, float f0, float f_max, double L, double s, double xi,
                          double b, float* z);
void orc_z_const(size_t n, float re, float im, float* z);
void orc_z_collimator(size_t n, double outer, double inner, float* z);
int orc_make_impedance(size_t nfreqs, float fmax, double R_bend, double f_rev, double gap,
                       int use_csr, double s, double xi, double collimator, float* z);
/* Airy functions for x >= 0 by quadrature of the modified-Bessel integrals; returns 1 on
 * Bi overflow */
int orc_airy(double x, double* ai, double* aip, double* bi, double* bip);

/* scalar derivation of main() (src/main.cpp:163-355,663-664,682) */
void orc_derive(const ivs_ring_params* p, ivs_derived* d);

/* whole simulation (object wiring src/main.cpp:478-763, loop body :948-1053) */
typedef struct orc_sim orc_sim;
orc_sim* orc_sim_create(const ivs_ring_params* p, const float* grid, const float* particles,
                        uint32_t np);
/* the same with impedance tables made earlier by orc_make_impedance (NULL = compute here) */
orc_sim* orc_sim_create_z(const ivs_ring_params* p, const float* grid, const float* particles,
                          uint32_t np, const float* zwake, const float* zrad);
void orc_sim_destroy(orc_sim* s);
/* DynamicRFKickMap (src/SM/DynamicRFKickMap.cpp:80-109) with a given modulation queue */
void orc_sim_set_rf_modulation(orc_sim* s, const float* phase_ampl, uint32_t count);
void orc_rf_calc_kick(int linear, uint32_t n, float qmin, float qmax, float pmin, float pmax,
                      float bl_scale, float dE_scale, float angle, float revolutionpart, float V_RF,
                      float f_RF, float V0, float phase, float ampl, float* offset);
void orc_rf_modulation(uint32_t steps, float phasespread, float amplspread, float modampl,
                       double modtimeincrement, double revolutionpart, float syncphase,
                       const float* gauss, float* out);
void orc_sim_step(orc_sim* s, uint32_t nsteps);
void orc_sim_derived(const orc_sim* s, ivs_derived* d);
void orc_sim_get(const orc_sim* s, float* grid, float* profile, float* wake, float* particles);
void orc_sim_tables(const orc_sim* s, float* rf, float* drift, float* fp5, float* zwake,
                    float* zrad);
void orc_sim_observe(orc_sim* s, float* obs8, float* csr_spectrum, float* projy);
/* replace tables (to run the restatement on tables taken from another build) */
void orc_sim_set_tables(orc_sim* s, const float* rf, const float* drift, const float* fp5,
                        const float* zwake);

#ifdef __cplusplus
}
#endif
#endif
