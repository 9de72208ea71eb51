/* ivs_host.h -- host-side (CPU, one-time) helpers of the B200 VFP step library.
 *
 * These restate the set-up arithmetic the reference performs in main() and in the map /
 * impedance constructors, so that a caller can go from ring parameters to the tables the
 * device path consumes (ivs_sim_desc in ivs.h).  They run on the CPU, need no GPU, and are
 * deterministic.  Citations are into /root/reference/.
 */
#ifndef IVS_HOST_H
#define IVS_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The subset of the reference's program options (src/IO/ProgramOptions.cpp:22-305) that
 * reaches the hot path, with the reference's member types (inc/IO/ProgramOptions.hpp:239-290:
 * float for meshaxis_t/frequency_t options, double for the others). */
typedef struct ivs_ring_params {
    uint32_t grid_size;        /* GridSize            (256)   */
    uint32_t steps_per_Ts;     /* StepsPerTsync       (1000)  */
    uint32_t interpolation;    /* InterpolationPoints (4)     */
    uint32_t derivation;       /* DerivationType      (4)     */
    uint32_t fp_type;          /* FPType              (3)     */
    uint32_t fp_track;         /* FPTrack             (3)     */
    uint32_t linear_rf;        /* LinearRF            (1)     */
    uint32_t round_padding;    /* RoundPadding        (1)     */
    uint32_t use_csr;          /* UseCSR              (1)     */
    uint32_t reserved;
    float pq_size;             /* PhaseSpaceSize      (12)    */
    float shift_x, shift_y;    /* PhaseSpaceShiftX/Y  (0)     */
    float alpha0, alpha1, alpha2; /* AlphaZero/One/Two        */
    float f_s;                 /* SynchrotronFrequency (0 = derive from alpha0) */
    float f_rev;               /* RevolutionFrequency (9e6)   */
    float harmonic;            /* HarmonicNumber      (50)    */
    float f_cutoff;            /* CutoffFrequency     (23e9)  */
    double padding;            /* padding             (8)     */
    double steps_per_Trev;     /* StepsPerTrev        (0)     */
    double E0;                 /* BeamEnergy          (1.3e9) */
    double sE;                 /* BeamEnergySpread    (4.7e-4)*/
    double V_RF;               /* AcceleratingVoltage (1e6)   */
    double R_bend;             /* BendingRadius       (-1)    */
    double t_damp;             /* DampingTime         (-1)    */
    double gap;                /* VacuumGap           (0.03)  */
    double collimator;         /* CollimatorRadius    (0)     */
    double wall_conductivity;  /* WallConductivity    (0)     */
    double wall_susceptibility;/* WallSusceptibility  (0)     */
    double zoom;               /* InitialDistZoom     (1)     */
    float bunch_current;       /* BunchCurrent (single bunch) */
    float reserved_f;
} ivs_ring_params;

/* Scalars main() derives before building its objects (src/main.cpp:163-355,663-664,682). */
typedef struct ivs_derived {
    float qmin, qmax, pmin, pmax;   /* :179-187 */
    float dq, dp;                   /* Ruler delta (inc/PS/Ruler.hpp:38) */
    float zerobin_x, zerobin_y;     /* Ruler zerobin (inc/PS/Ruler.hpp:40) */
    float angle;                    /* :353 */
    float fmax;                     /* :309 */
    float e1;                       /* :682 */
    float slip[3];                  /* :663-664 */
    float wakescaling;              /* ElectricField::_wakescaling (hpp:99-100, cpp:58) */
    float bl_f, dE_f;               /* axis scales as stored (float map values) */
    double dE, R_bend, f_rev, f_RF, V0, V_eff, fs, alpha0, bl, Ib, Qb, steps, t_damp, dt,
        revolutionpart;
    uint32_t padded_bins;           /* :311-316 */
    uint32_t spacing_bins;          /* :304-307 */
    uint32_t laststep_per_Ts;       /* ceil(steps) */
    uint32_t reserved;
} ivs_derived;

/* fills `p` with the reference's defaults (the values in the comments above) */
void ivs_default_params(ivs_ring_params* p);
/* src/main.cpp:163-355,663-664,682 */
int ivs_derive(const ivs_ring_params* p, ivs_derived* d);
/* RFKickMap::_calcKick (src/SM/RFKickMap.cpp:59-83) at the synchronous phase: n floats */
int ivs_rf_kick_offsets(const ivs_ring_params* p, const ivs_derived* d, float* offset);
/* DriftMap::DriftMap (src/SM/DriftMap.cpp:21-30): n floats */
int ivs_drift_offsets(const ivs_ring_params* p, const ivs_derived* d, float* offset);
/* the same three with the map constructors' own arguments (axis limits, axis scales "Meter" =
 * bl and "ElectronVolt" = dE, and what src/main.cpp:640-679 passes), for host code that mirrors
 * the reference's classes (include/vfps_b200.hpp) */
int ivs_rf_kick_linear(uint32_t n, float qmin, float qmax, float bl, float angle, float f_RF,
                       float* offset);
int ivs_rf_kick_sine(uint32_t n, float qmin, float qmax, float pmin, float pmax, float bl, float dE,
                     float revolutionpart, float V_RF, float f_RF, float V0, float* offset);
int ivs_drift_slip(uint32_t n, float qmin, float qmax, float pmin, float pmax, float dE,
                   const float* slip, uint32_t nslip, float E0, float* offset);
/* makeImpedance (src/Z/ImpedanceFactory.cpp:21-84) without the file contribution:
 * z = nfreqs complex<float> (interleaved re,im); *nonzero = 0 if "no impedance is used" */
int ivs_make_impedance(size_t nfreqs, float fmax, double R_bend, double f_rev, double gap,
                       int use_csr, double s, double xi, double collimator, float* z,
                       int* nonzero);
/* Impedance::readData (src/Z/Impedance.cpp:94-111): an impedance table from a text file of
 * `line-number real imag` records (whitespace separated; a record whose line number repeats the
 * previous one is skipped).  Up to `cap` complex<float> values are stored in z (interleaved re,im),
 * *count = records in the file.  IVS_ERR_INVALID_ARGUMENT if the file cannot be opened. */
int ivs_read_impedance_file(const char* path, float* z, size_t cap, size_t* count);
/* `*rv += Impedance(impedance_file, fmax)` (src/Z/ImpedanceFactory.cpp:70-75): adds the first nfreqs
 * records of the file to z.  A file with fewer records than nfreqs is refused (the reference reads past
 * the end of its vector there). */
int ivs_add_impedance_file(const char* path, float* z, size_t nfreqs);
/* the wake impedance (src/main.cpp:717-721) and the radiation impedance (:724-727) */
int ivs_wake_impedance(const ivs_ring_params* p, const ivs_derived* d, float* z, int* nonzero);
int ivs_radiation_impedance(const ivs_ring_params* p, const ivs_derived* d, float* z);
/* PhaseSpace::gaus (src/PS/PhaseSpace.cpp:471-484): exp(-a_i^2 / (2 zoom^2)) / sqrt(2 pi) on the
 * axis values of a Ruler(n, amin, amax): n floats */
int ivs_gaussian_projection(uint32_t n, float amin, float amax, double zoom, float* proj);
/* PhaseSpace::PhaseSpace with data == nullptr followed by main()'s initial
 * renormalisation (src/PS/PhaseSpace.cpp:52-65,457-484; src/main.cpp:528-532): n*n floats */
int ivs_initial_grid(const ivs_derived* d, uint32_t n, double zoom, float* grid);
/* the constants of RFKickMap::_calcKick for these ring parameters (src/SM/RFKickMap.cpp:15-57 as
 * src/main.cpp:600-650 calls the constructors): `model` points at an ivs_rf_model (ivs.h) */
int ivs_rf_model_init(const ivs_ring_params* p, const ivs_derived* d, void* model);
/* DynamicRFKickMap::__calcModulation (src/SM/DynamicRFKickMap.cpp:80-94): out[2 i] = syncphase +
 * N(0,1) phasespread / sqrt(revolutionpart) + modampl sin(float(2 pi modtimeincrement) i),
 * out[2 i + 1] = 1 + N(0,1) amplspread / sqrt(revolutionpart).  The normal deviates come from `gauss`
 * (2 per step) if given, else from a counter-based generator keyed on `seed` (the reference uses an
 * unseeded mt19937, :36, which cannot be reproduced). */
int ivs_rf_modulation(uint64_t seed, uint32_t steps, float phasespread, float amplspread, float modampl,
                      double modtimeincrement, double revolutionpart, float syncphase, const float* gauss,
                      float* out);
/* Haissinski start distribution f = rho(q) exp(-p^2/2) (BASELINE configs[2]; the reference parses
 * HaissinskiIterations and ignores it, src/IO/ProgramOptions.cpp:270-272): the stationary bunch
 * profile rho ~ exp((dp/angle) int (rf + W[rho]) dq) of the step map, W[rho] = ElectricField::
 * wakePotential of rho with `zwake` (nmax complex<float>, may be NULL: no wake) and `wakescaling`,
 * rf_offset = RFKickMap offsets (n floats).  Damped fixed-point iteration (`relax` in (0,1]) in
 * double.  grid: n*n floats for PhaseSpace's `data` constructor argument (inc/PS/PhaseSpace.hpp:60).
 * Returns 0 if the last iteration changed rho by less than 1e-9 (relative L1, also in *residual), else 1. */
int ivs_haissinski_grid(const ivs_derived* d, uint32_t n, const float* rf_offset, const float* zwake,
                        uint32_t nmax, float wakescaling, uint32_t iterations, double relax, float* grid,
                        double* residual);
/* Wake-function windows of a train of bunches (PhaseSpace::nb > 1, src/main.cpp:248-321): the reference
 * places the profile of bunch b at bucket[b] * spacing_bins of ONE padded array of nmax = nFreqs bins and
 * reads its wake potential back from there (src/PS/ElectricField.cpp:236-263), so bunch b' acts on bunch b
 * through the (circular) wake function at lags (bucket[b] - bucket[b']) * spacing_bins + x - x'.
 * windows: nb * nb * 2 n doubles, window (b, b') entry m + n - 1 = lag (..) + m, m = -(n-1)..(n-1); input of
 * ivs_train_wake_potential / ivs_sim_set_bunch_train (ivs.h).  z: nmax complex<float> (ivs_make_impedance
 * with nfreqs = the spaced length), nmax a power of two.  Returns 0, -1 for bad arguments, -2 if a bunch
 * does not fit the padded array. */
int ivs_train_wake_windows(const float* z, uint32_t nmax, uint32_t n, uint32_t nb, const uint32_t* bucket,
                           uint32_t spacing_bins, double* windows);
/* Airy functions used by the parallel-plates impedance (src/Z/ParallelPlatesCSR.cpp:48-53);
 * returns 0, or 1 if Bi/Bi' overflow a double (Boost would throw there) */
int ivs_airy(double x, double* ai, double* aip, double* bi, double* bip);

#ifdef __cplusplus
}
#endif
#endif /* IVS_HOST_H */
