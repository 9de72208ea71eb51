// Backing implementation of oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
//
// Replaces the FFTW3 calls of the reference (/root/reference/src/FFTWWrapper.cpp:11-107,
// /root/reference/src/PS/ElectricField.cpp:163,227,234) when its sources are compiled in
// place into oracle/_ref/.  Transforms are evaluated in double precision (radix-2 for
// powers of two, direct O(n^2) summation otherwise) and rounded once to the output type.
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
typedef std::complex<double> cd;
const double kTwoPi = 6.283185307179586476925286766559005768;

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

// in-place complex DFT, sign = -1 forward / +1 backward, unnormalised
void dft(std::vector<cd>& a, int sign) {
    const size_t n = a.size();
    if (n <= 1) return;
    if (!is_pow2(n)) {
        std::vector<cd> out(n);
        for (size_t k = 0; k < n; k++) {
            cd s = 0;
            for (size_t j = 0; j < n; j++) {
                const double ph = sign * kTwoPi * double((k * j) % n) / double(n);
                s += a[j] * cd(std::cos(ph), std::sin(ph));
            }
            out[k] = s;
        }
        a.swap(out);
        return;
    }
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<cd> w(len / 2);
        for (size_t k = 0; k < len / 2; k++) {
            const double ph = sign * kTwoPi * double(k) / double(len);
            w[k] = cd(std::cos(ph), std::sin(ph));
        }
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; k++) {
                const cd u = a[i + k], v = a[i + k + len / 2] * w[k];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}

enum Kind { R2C, C2R, C2C };
template <typename real>
struct plan_t {
    Kind kind;
    int n;
    int sign;
    void* in;
    void* out;
};

template <typename real>
void execute(const plan_t<real>* p) {
    const size_t n = p->n;
    std::vector<cd> a(n);
    typedef real cplx[2];
    if (p->kind == R2C) {
        const real* in = static_cast<const real*>(p->in);
        cplx* out = static_cast<cplx*>(p->out);
        for (size_t i = 0; i < n; i++) a[i] = cd(in[i], 0);
        dft(a, -1);
        for (size_t k = 0; k <= n / 2; k++) {
            out[k][0] = static_cast<real>(a[k].real());
            out[k][1] = static_cast<real>(a[k].imag());
        }
    } else if (p->kind == C2R) {
        const cplx* in = static_cast<const cplx*>(p->in);
        real* out = static_cast<real*>(p->out);
        a[0] = cd(in[0][0], 0);
        for (size_t k = 1; k <= n / 2; k++) {
            cd v(in[k][0], in[k][1]);
            if (2 * k == n) v = cd(in[k][0], 0);
            a[k] = v;
            a[n - k] = std::conj(v);
        }
        dft(a, +1);
        for (size_t i = 0; i < n; i++) out[i] = static_cast<real>(a[i].real());
    } else {
        const cplx* in = static_cast<const cplx*>(p->in);
        cplx* out = static_cast<cplx*>(p->out);
        for (size_t i = 0; i < n; i++) a[i] = cd(in[i][0], in[i][1]);
        dft(a, p->sign);
        for (size_t i = 0; i < n; i++) {
            out[i][0] = static_cast<real>(a[i].real());
            out[i][1] = static_cast<real>(a[i].imag());
        }
    }
}

void* aligned(size_t bytes) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) return nullptr;
    return p;
}
}  // namespace

struct ivs_fftw_shim_plan : plan_t<float> {};
struct ivs_fftw_shim_plan_d : plan_t<double> {};

extern "C" {
void* fftwf_malloc(size_t n) { return aligned(n); }
float* fftwf_alloc_real(size_t n) { return static_cast<float*>(aligned(n * sizeof(float))); }
fftwf_complex* fftwf_alloc_complex(size_t n) { return static_cast<fftwf_complex*>(aligned(n * sizeof(fftwf_complex))); }
void fftwf_free(void* p) { free(p); }
double* fftw_alloc_real(size_t n) { return static_cast<double*>(aligned(n * sizeof(double))); }
fftw_complex* fftw_alloc_complex(size_t n) { return static_cast<fftw_complex*>(aligned(n * sizeof(fftw_complex))); }
void fftw_free(void* p) { free(p); }

static fftwf_plan mkf(Kind k, int n, int sign, void* in, void* out, unsigned flags) {
    if (flags & FFTW_WISDOM_ONLY) return nullptr;  // there never is wisdom
    ivs_fftw_shim_plan* p = new ivs_fftw_shim_plan;
    p->kind = k; p->n = n; p->sign = sign; p->in = in; p->out = out;
    return p;
}
static fftw_plan mkd(Kind k, int n, int sign, void* in, void* out, unsigned flags) {
    if (flags & FFTW_WISDOM_ONLY) return nullptr;
    ivs_fftw_shim_plan_d* p = new ivs_fftw_shim_plan_d;
    p->kind = k; p->n = n; p->sign = sign; p->in = in; p->out = out;
    return p;
}
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float* in, fftwf_complex* out, unsigned f) { return mkf(R2C, n, -1, in, out, f); }
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex* in, float* out, unsigned f) { return mkf(C2R, n, +1, in, out, f); }
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int s, unsigned f) { return mkf(C2C, n, s, in, out, f); }
void fftwf_execute(const fftwf_plan p) { execute<float>(p); }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
int fftwf_import_wisdom_from_filename(const char*) { return 0; }
int fftwf_export_wisdom_to_filename(const char*) { return 0; }

fftw_plan fftw_plan_dft_r2c_1d(int n, double* in, fftw_complex* out, unsigned f) { return mkd(R2C, n, -1, in, out, f); }
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex* in, double* out, unsigned f) { return mkd(C2R, n, +1, in, out, f); }
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int s, unsigned f) { return mkd(C2C, n, s, in, out, f); }
void fftw_execute(const fftw_plan p) { execute<double>(p); }
void fftw_destroy_plan(fftw_plan p) { delete p; }
int fftw_import_wisdom_from_filename(const char*) { return 0; }
int fftw_export_wisdom_to_filename(const char*) { return 0; }
}
