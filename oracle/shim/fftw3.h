/* Minimal stand-in for <fftw3.h> -- TEST INFRASTRUCTURE ONLY.
 *
 * FFTW3 is not installed in this image.  The reference calls it at
 * /root/reference/src/FFTWWrapper.cpp:11-107 (alloc, r2c/c2r plans, wisdom) and
 * /root/reference/src/PS/ElectricField.cpp:163,227,234 (execute).  This header
 * declares the subset those call sites need; oracle/fftw_shim.cpp backs it with a
 * plain radix-2 FFT evaluated in DOUBLE precision and rounded to float on output,
 * i.e. the correctly rounded transform that any single-precision FFT (FFTW's, or the
 * CUDA one under test) approximates to within its own round-off.  Normalisation and
 * sign conventions are FFTW's: r2c = forward (exp(-i...)), unnormalised, bins 0..n/2
 * written; c2r = backward (exp(+i...)), unnormalised, reads bins 0..n/2 and ignores the
 * imaginary parts of bins 0 and n/2.
 */
#ifndef IVS_ORACLE_FFTW3_SHIM_H
#define IVS_ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef double fftw_complex[2];
struct ivs_fftw_shim_plan;
typedef struct ivs_fftw_shim_plan* fftwf_plan;
typedef struct ivs_fftw_shim_plan_d* fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_WISDOM_ONLY (1U << 21)

void* fftwf_malloc(size_t n);
float* fftwf_alloc_real(size_t n);
fftwf_complex* fftwf_alloc_complex(size_t n);
void fftwf_free(void* p);
double* fftw_alloc_real(size_t n);
fftw_complex* fftw_alloc_complex(size_t n);
void fftw_free(void* p);

fftwf_plan fftwf_plan_dft_r2c_1d(int n, float* in, fftwf_complex* out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex* in, float* out, unsigned flags);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_import_wisdom_from_filename(const char* fn);
int fftwf_export_wisdom_to_filename(const char* fn);

fftw_plan fftw_plan_dft_r2c_1d(int n, double* in, fftw_complex* out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex* in, double* out, unsigned flags);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
int fftw_import_wisdom_from_filename(const char* fn);
int fftw_export_wisdom_to_filename(const char* fn);
#ifdef __cplusplus
}
#endif
#endif
