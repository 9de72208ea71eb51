// ivs_host.cpp -- one-time host-side set-up arithmetic (include/ivs_host.h and the host
// helpers of include/ivs.h): ring parameters -> derived scalars -> kick offsets, Fokker-Planck
// stencil, Simpson weights, impedance tables, start distribution.
//
// These follow the reference's set-up code expression by expression (citations into
// /root/reference/), because the per-step device arithmetic consumes the tables bit for bit:
// e.g. a 1-ulp difference in an RF offset is a systematic shift that accumulates over the
// steps.  Compiled with -ffp-contract=off: every fp32 operation is rounded once, like a
// reference build without contraction.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <fstream>
#include <limits>
#include <vector>

#include "../../include/ivs.h"
#include "../../include/ivs_host.h"

namespace {

namespace phys {  // inc/defines.hpp:47-66
constexpr double c = 2.99792458e8;
constexpr double epsilon0 = 8.854187817e-12;
constexpr double Z0 = 1 / (epsilon0 * c);
constexpr double e = 1.602e-19;
constexpr double me = 510998.9;
}  // namespace phys
constexpr double kZ0Impedance = 376.730313461;  // inc/Z/Impedance.hpp:94
constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kTwoPi = 2 * kPi;
constexpr double kPiSqr = 9.869604401089358618834490999876151135;

// inc/PS/Ruler.hpp:31-51
struct Axis {
    uint32_t steps;
    float min, max, delta, zerobin;
    Axis(uint32_t n, float lo, float hi)
        : steps(n), min(lo), max(hi), delta((hi - lo) / static_cast<float>(n - 1)),
          zerobin(((lo + hi) / (lo - hi) + 1) * (n - 1) / 2) {}
    float at(uint32_t i) const { return min + (static_cast<float>(i) * delta); }
};

uint64_t next_pow2(uint64_t v) {  // src/HelperFunctions.cpp:178-189
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- Airy functions for x >= 0: double results from long double arithmetic.
//   x <= 12.5 : Bi, Bi' from the Maclaurin series (all terms positive, no cancellation);
//   x <= 4.5  : Ai, Ai' from the same series (cancellation e^(2 zeta) <= 3e5, harmless at 64 bit
//               mantissa);
//   4.5 < x <= 12.5 : Ai, Ai' from the Taylor series of y'' = x y about the nearest integer,
//               started from tabulated 28-digit values (step <= 0.5, error growth <= e^(3.5));
//   x > 12.5  : asymptotic series (smallest term < e^(-59)).
// Accuracy ~1e-15 relative everywhere, which the parallel-plates sum needs: at low frequency
// Ai'Bi' + u Ai Bi cancels to O(u^-3) of its terms.
int airy_positive(double xd, double* ai, double* aip, double* bi, double* bip) {
    const long double c1 = 0.355028053887817239260063186004183176L;  // Ai(0)
    const long double c2 = 0.258819403792806798405183560189203963L;  // -Ai'(0)
    const long double sqrt3 = 1.732050807568877293527446341505872367L;
    const long double x = xd;
    if (x <= 12.5L) {
        const long double x3 = x * x * x;
        long double f = 1, g = x, fp = 0, gp = 1;
        long double tf = 1, tg = x;
        for (int k = 1; k < 400; k++) {
            tf *= x3 / ((3.0L * k - 1) * (3.0L * k));
            tg *= x3 / ((3.0L * k) * (3.0L * k + 1));
            f += tf;
            g += tg;
            if (x > 0) {
                fp += tf * (3.0L * k) / x;
                gp += tg * (3.0L * k + 1) / x;
            }
            if (tf < 1e-22L * f && tg < 1e-22L * g) break;
        }
        *bi = static_cast<double>(sqrt3 * (c1 * f + c2 * g));
        *bip = static_cast<double>(sqrt3 * (c1 * fp + c2 * gp));
        if (x <= 4.5L) {
            *ai = static_cast<double>(c1 * f - c2 * g);
            *aip = static_cast<double>(c1 * fp - c2 * gp);
            return 0;
        }
        // Ai(x0), Ai'(x0) for x0 = 4..12
        static const long double A0[9] = {
            0.0009515638512048018736214999689L, 0.000108344428136074417349865025L,
            0.000009947694360252889570238847669L, 0.0000007492128863997167080771040272L,
            0.00000004692207616099231625649081703L, 2.471168430872489843289241134e-9L,
            1.104753255289868593355020566e-10L, 4.226275864960359591298835451e-12L,
            1.393184688875360839049034503e-13L};
        static const long double A1[9] = {
            -0.001958640950204178900138140918L, -0.0002474138908684624760002361721L,
            -0.00002476520039703495475418182539L, -0.000002008150894738791991169305312L,
            -0.0000001341439297906786574291153708L, -7.480641389658946412759545273e-9L,
            -3.520633676738923636620644825e-10L, -1.411144124662851733545119127e-11L,
            -4.854736554985308462993653998e-13L};
        int c = static_cast<int>(x + 0.5L);
        if (c < 4) c = 4;
        if (c > 12) c = 12;
        const long double x0 = c, h = x - x0;
        long double am1 = 0, a0 = A0[c - 4], a1 = A1[c - 4];  // a_{k-1}, a_k, a_{k+1}
        long double y = a0 + a1 * h, yp = a1, hk = h;         // hk = h^(k+1)
        for (int k = 0; k < 200; k++) {
            const long double a2 = (x0 * a0 + am1) / ((k + 1.0L) * (k + 2.0L));  // a_{k+2}
            yp += (k + 2.0L) * a2 * hk;
            hk *= h;
            y += a2 * hk;
            am1 = a0; a0 = a1; a1 = a2;
            if (k > 4 && fabsl(a2 * hk) < 1e-24L * fabsl(y)) break;
        }
        *ai = static_cast<double>(y);
        *aip = static_cast<double>(yp);
        return 0;
    }
    const long double zeta = 2.0L / 3.0L * x * sqrtl(x);
    const long double x14 = sqrtl(sqrtl(x));
    const long double sqrtpi = 1.772453850905516027298167483341145183L;
    long double su = 1, sv = 1, au = 1, av = 1;  // sums with + signs (Bi) and alternating (Ai)
    long double u = 1, zk = 1, last = 1;
    for (int k = 1; k < 60; k++) {
        u *= (6.0L * k - 5) * (6.0L * k - 3) * (6.0L * k - 1) / (216.0L * k * (2.0L * k - 1));
        const long double v = -(6.0L * k + 1) / (6.0L * k - 1) * u;
        zk *= zeta;
        const long double tu = u / zk, tv = v / zk;
        if (fabsl(tu) > last) break;  // asymptotic series: stop at the smallest term
        last = fabsl(tu);
        su += tu;
        sv += tv;
        au += (k & 1) ? -tu : tu;
        av += (k & 1) ? -tv : tv;
        if (last < 1e-21L) break;
    }
    const long double em = expl(-zeta);
    *ai = static_cast<double>(em / (2 * sqrtpi * x14) * au);
    *aip = static_cast<double>(-x14 * em / (2 * sqrtpi) * av);
    // Boost raises overflow_error where the result leaves the double range; the reference
    // uses that to end its series (src/Z/ParallelPlatesCSR.cpp:54-57)
    if (zeta + 0.25L * logl(x) - logl(sqrtpi) > 709.78L) {
        *bi = *bip = INFINITY;
        return 1;
    }
    const long double ep = expl(zeta);
    *bi = static_cast<double>(ep / (sqrtpi * x14) * su);
    *bip = static_cast<double>(x14 * ep / sqrtpi * sv);
    return 0;
}

typedef std::vector<std::complex<float>> ztable;

// src/Z/FreeSpaceCSR.cpp:18-43
ztable z_free_space(size_t n, float f_rev, float f_max) {
    ztable rv(n, 0);
    const std::complex<float> Z0(306.3f, 176.9f);
    const float delta = f_max / f_rev / (n - 1);
    for (size_t i = 0; i <= n / 2; i++) rv[i] = Z0 * std::pow(i * delta, float(1.0 / 3.0));
    return rv;
}

// src/Z/ParallelPlatesCSR.cpp:25-70
ztable z_parallel_plates(size_t nfreqs, float f0, float f_max, double g) {
    ztable rv(nfreqs, 0);
    const double delta = f_max / f0 / (nfreqs - 1.0);
    const double r_bend = phys::c / (2 * kPi * f0);
    for (size_t i = 1; i <= nfreqs / 2; i++) {
        std::complex<double> Z = 0;
        const double n = i * delta;
        const double m = n * std::pow(g / r_bend, 3. / 2.);
        const uint32_t maxp = 2 * m * std::pow(r_bend / g, 3. / 2.) * f0 * g / phys::c;
        const double b = std::pow(m, -4. / 3.);
        for (uint32_t p = 1; p <= maxp; p += 2) {
            const double u = kPiSqr * p * p / std::pow(2, 2. / 3.) * b;
            double ai, aip, bi, bip;
            if (airy_positive(u, &ai, &aip, &bi, &bip)) break;  // overflow ends the sum
            const std::complex<double> zinc =
                aip * std::complex<double>(aip, -bip) + (u * ai) * std::complex<double>(ai, -bi);
            Z += zinc;
        }
        Z *= 4.0 * b * n * g / r_bend * kPiSqr * std::pow(2, 1. / 3.) * phys::Z0;
        rv[i] = static_cast<std::complex<float>>(Z);
    }
    return rv;
}

// src/Z/ResistiveWall.cpp:24-57
ztable z_resistive_wall(size_t n, float f0, float f_max, double L, double s, double xi, double b) {
    ztable rv(n, 0);
    const double mu_r = (1 + xi);
    const std::complex<float> Z1 =
        static_cast<float>(std::sqrt(kZ0Impedance * mu_r * f0 / s / kPi / phys::c) * L / 2 / b) *
        std::complex<float>(1, -1);
    const float delta = f_max / f0 / (n - 1.0);
    for (size_t i = 0; i <= n / 2; i++) rv[i] = Z1 * std::sqrt(i * delta);
    return rv;
}

// src/Z/CollimatorImpedance.cpp:12-25 over src/Z/ConstImpedance.cpp:18-28
ztable z_collimator(size_t n, double outer, double inner) {
    ztable rv(n, 0);
    const std::complex<float> Z(static_cast<float>(kZ0Impedance / kPi * std::log(outer / inner)), 0.0f);
    for (size_t i = 0; i < n / 2; i++) rv[i] = Z;
    return rv;
}

}  // namespace

// Unit wake function w[m], m = -(n-1)..(n-1) (index m + n - 1): c2r of the Hermitian-extended impedance
// with the reference's conventions (imaginary part of bin 0 ignored, bins >= nmax/2 zero,
// src/PS/ElectricField.cpp:229-234), in double.  wakePotential()[x] = scaling sum_x' rho[x'] w[x - x']
// exactly in real arithmetic, because nmax >= 2n leaves no circular wrap.  Shared with ivs_step.cu.
// all nmax lags of the (circular) unit wake function: c2r of the Hermitian-extended impedance, in double
static std::vector<double> wake_function_lags(const float* z, uint32_t nmax) {
    std::vector<std::complex<double>> a(nmax, 0.0);
    a[0] = std::complex<double>(z[0], 0.0);
    for (uint32_t k = 1; k < nmax / 2; k++) {
        a[k] = std::complex<double>(z[2 * k], z[2 * k + 1]);
        a[nmax - k] = std::conj(a[k]);
    }
    // radix-2 FFT, sign +
    for (uint32_t i = 1, j = 0; i < nmax; i++) {
        uint32_t bit = nmax >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (uint32_t len = 2; len <= nmax; len <<= 1) {
        for (uint32_t k = 0; k < len / 2; k++) {
            const double ph = kTwoPi * (double)k / (double)len;
            const std::complex<double> w(std::cos(ph), std::sin(ph));
            for (uint32_t i = k; i < nmax; i += len) {
                const std::complex<double> u = a[i], v = a[i + len / 2] * w;
                a[i] = u + v;
                a[i + len / 2] = u - v;
            }
        }
    }
    std::vector<double> r(nmax);
    for (uint32_t i = 0; i < nmax; i++) r[i] = a[i].real();
    return r;
}

std::vector<double> ivs_wake_function_host(const float* z, uint32_t nmax, uint32_t n) {
    const std::vector<double> a = wake_function_lags(z, nmax);
    std::vector<double> w(2 * n - 1);
    for (int m = -(int)(n - 1); m <= (int)(n - 1); m++)
        w[m + (int)(n - 1)] = a[(uint32_t)((m + (int)nmax) % (int)nmax)];
    return w;
}

extern "C" {

// Train of bunches (PhaseSpace::nb > 1): with the profiles at bucket[b] * spacing_bins in ONE padded array
// (src/PS/ElectricField.cpp:255-263) the circular convolution couples bunch b' to bunch b through the
// wake function at lags (bucket[b] - bucket[b']) * spacing_bins + (x - x').  windows[(b * nb + b') * 2 n + m
// + n - 1] = that lag's value for m = x - x' in -(n-1)..(n-1) (entry 2 n - 1 of a window is padding).
int ivs_train_wake_windows(const float* z, uint32_t nmax, uint32_t n, uint32_t nb, const uint32_t* bucket,
                           uint32_t spacing_bins, double* windows) {
    if (!z || !bucket || !windows || n < 2 || nb < 1 || nmax < 2 || (nmax & (nmax - 1))) return -1;
    for (uint32_t b = 0; b < nb; b++)
        if ((uint64_t)bucket[b] * spacing_bins + n > nmax) return -2;  // the reference would write past its array
    const std::vector<double> a = wake_function_lags(z, nmax);
    for (uint32_t b = 0; b < nb; b++)
        for (uint32_t c = 0; c < nb; c++) {
            double* w = windows + ((size_t)b * nb + c) * 2 * n;
            const int64_t lag0 = ((int64_t)bucket[b] - (int64_t)bucket[c]) * (int64_t)spacing_bins;
            for (int64_t m = -(int64_t)(n - 1); m <= (int64_t)(n - 1); m++) {
                int64_t l = (lag0 + m) % (int64_t)nmax;
                if (l < 0) l += nmax;
                w[m + (int64_t)(n - 1)] = a[(size_t)l];
            }
            w[2 * n - 1] = 0.0;
        }
    return 0;
}

int ivs_airy(double x, double* ai, double* aip, double* bi, double* bip) {
    if (!(x >= 0) || !ai || !aip || !bi || !bip) return -1;
    return airy_positive(x, ai, aip, bi, bip);
}

void ivs_default_params(ivs_ring_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->grid_size = 256; p->steps_per_Ts = 1000; p->interpolation = 4; p->derivation = 4;
    p->fp_type = 3; p->fp_track = 3; p->linear_rf = 1; p->round_padding = 1; p->use_csr = 1;
    p->pq_size = 12; p->alpha0 = 4e-3f; p->f_rev = 9e6f; p->harmonic = 50; p->f_cutoff = 23e9f;
    p->padding = 8; p->E0 = 1.3e9; p->sE = 4.7e-4; p->V_RF = 1e6; p->R_bend = -1; p->t_damp = -1;
    p->gap = 0.03; p->zoom = 1; p->bunch_current = 1e-3f;
}

// src/main.cpp:163-355, 663-664, 682; member scalars of PhaseSpace/ElectricField constructors
int ivs_derive(const ivs_ring_params* o, ivs_derived* d) {
    if (!o || !d || o->grid_size < 4) return IVS_ERR_INVALID_ARGUMENT;
    std::memset(d, 0, sizeof(*d));
    const uint32_t ps_bins = o->grid_size;
    const float pqsize = o->pq_size;
    const float qcenter = -o->shift_x * pqsize / (ps_bins - 1);
    const float pcenter = -o->shift_y * pqsize / (ps_bins - 1);
    const float pqhalf = pqsize / 2;
    d->qmax = qcenter + pqhalf;
    d->qmin = qcenter - pqhalf;
    d->pmax = pcenter + pqhalf;
    d->pmin = pcenter - pqhalf;
    const double sE = o->sE, E0 = o->E0;
    d->dE = sE * E0;
    d->f_rev = static_cast<double>(o->f_rev);
    d->R_bend = (o->R_bend > 0) ? o->R_bend : phys::c / (kTwoPi * d->f_rev);
    const double harmonic = static_cast<double>(o->harmonic);
    d->f_RF = d->f_rev * harmonic;
    const double bunchspacing = 1.0 / d->f_RF;
    const double lorentzgamma = E0 / phys::me;
    d->V0 = phys::e * std::pow(lorentzgamma, 4) / (3 * phys::epsilon0 * d->R_bend);
    const double W0 = d->V0 * phys::e;
    d->V_eff = std::sqrt(o->V_RF * o->V_RF - d->V0 * d->V0);
    double fs = static_cast<double>(o->f_s);
    double alpha0 = static_cast<double>(o->alpha0);
    if (std::fpclassify(fs) == FP_ZERO) {
        fs = d->f_rev * std::sqrt(alpha0 * harmonic * d->V_eff / (kTwoPi * E0));
    } else {
        const int sgn = (fs < 0) ? -1 : 1;
        alpha0 = sgn * kTwoPi * E0 / (harmonic * d->V_eff) * std::pow(fs / d->f_rev, 2);
    }
    d->fs = fs;
    d->alpha0 = alpha0;
    const float alpha[3] = {static_cast<float>(alpha0), o->alpha1, o->alpha2};
    d->bl = phys::c * d->dE / harmonic / std::pow(d->f_rev, 2.0) / d->V_eff * fs;
    d->Ib = 0.0 + static_cast<double>(o->bunch_current);
    d->Qb = d->Ib / d->f_rev;
    d->steps = (o->steps_per_Trev > 0) ? o->steps_per_Trev * d->f_rev / fs
                                       : static_cast<double>(std::max(o->steps_per_Ts, 1u));
    const double calc_damp = E0 * phys::e / W0 / d->f_rev;
    d->t_damp = (o->t_damp < 0) ? calc_damp : o->t_damp;
    d->dt = 1.0 / (fs * d->steps);
    d->revolutionpart = d->f_rev * d->dt;
    const double spacing_ps = bunchspacing * phys::c / d->bl / pqsize;
    d->spacing_bins = static_cast<uint32_t>(std::round(ps_bins * spacing_ps));
    d->fmax = ps_bins * phys::c / (pqsize * d->bl);
    const double padding = std::max(o->padding, 1.0);
    size_t padded = std::ceil(ps_bins * padding);
    if (o->round_padding) padded = next_pow2(padded);
    d->padded_bins = static_cast<uint32_t>(padded);
    d->angle = kTwoPi / d->steps;
    d->laststep_per_Ts = static_cast<uint32_t>(std::ceil(d->steps));
    d->e1 = (d->t_damp > 0) ? 2.0 / (fs * d->t_damp * d->steps) : 0;
    d->slip[0] = d->angle;
    d->slip[1] = alpha[1] / alpha[0] * d->angle;
    d->slip[2] = alpha[2] / alpha[0] * d->angle;
    const Axis ax0(ps_bins, d->qmin, d->qmax), ax1(ps_bins, d->pmin, d->pmax);
    d->dq = ax0.delta;
    d->dp = ax1.delta;
    d->zerobin_x = ax0.zerobin;
    d->zerobin_y = ax1.zerobin;
    d->bl_f = static_cast<float>(d->bl);  // Ruler scales are stored as meshaxis_t
    d->dE_f = static_cast<float>(d->dE);
    // inc/PS/ElectricField.hpp:99-100 then src/PS/ElectricField.cpp:58
    const float raw = d->Ib * d->dt * phys::c / d->bl_f / (d->dp * sE * E0);
    d->wakescaling = raw / padded;
    return IVS_OK;
}

// RFKickMap's linear constructor + _calcKick (src/SM/RFKickMap.cpp:15-33,59-68) with the
// constructor's own arguments; bl = "Meter" scale of the x axis
int ivs_rf_kick_linear(uint32_t n, float qmin, float qmax, float bl, float angle, float f_RF, float* offset) {
    if (!offset || n < 2) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax0(n, qmin, qmax);
    const float bl2phase = bl / phys::c * f_RF * kTwoPi;
    const float syncphase = 0, phase = 0, ampl = 1;
    const float phaseoffs = syncphase - phase;
    const float xcenter = ax0.zerobin;
    for (uint32_t x = 0; x < n; x++) {
        offset[x] = std::tan(angle) * (xcenter - x);
        offset[x] += std::tan(angle) * phaseoffs / bl2phase / ax0.delta;
        offset[x] *= ampl;
    }
    return IVS_OK;
}

// RFKickMap's sine constructor + _calcKick (src/SM/RFKickMap.cpp:35-57,69-80); dE = "ElectronVolt"
// scale of the y axis
int ivs_rf_kick_sine(uint32_t n, float qmin, float qmax, float pmin, float pmax, float bl, float dE,
                     float revolutionpart, float V_RF, float f_RF, float V0, float* offset) {
    if (!offset || n < 2) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax0(n, qmin, qmax), ax1(n, pmin, pmax);
    const float bl2phase = bl / phys::c * f_RF * kTwoPi;
    const float syncphase = std::asin(V0 / V_RF);
    const float phase = syncphase, ampl = 1;
    for (uint32_t x = 0; x < n; x++) {
        offset[x] = revolutionpart * (-ampl * V_RF * std::sin(ax0.at(x) * bl2phase + phase) + V0) /
                    ax1.delta / dE;
    }
    return IVS_OK;
}

// src/SM/RFKickMap.cpp:15-83 from the ring parameters
int ivs_rf_kick_offsets(const ivs_ring_params* o, const ivs_derived* d, float* offset) {
    if (!o || !d || !offset) return IVS_ERR_INVALID_ARGUMENT;
    const uint32_t n = o->grid_size;
    const float f_RF = d->f_RF;
    if (o->linear_rf) return ivs_rf_kick_linear(n, d->qmin, d->qmax, d->bl_f, d->angle, f_RF, offset);
    // main() passes V_eff as V_RF (src/main.cpp:646)
    return ivs_rf_kick_sine(n, d->qmin, d->qmax, d->pmin, d->pmax, d->bl_f, d->dE_f, (float)d->revolutionpart,
                            (float)d->V_eff, f_RF, (float)d->V0, offset);
}

// DriftMap::DriftMap (src/SM/DriftMap.cpp:21-30) with the constructor's own arguments
int ivs_drift_slip(uint32_t n, float qmin, float qmax, float pmin, float pmax, float dE, const float* slip,
                   uint32_t nslip, float E0, float* offset) {
    if (!offset || !slip || n < 2) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax0(n, qmin, qmax), ax1(n, pmin, pmax);
    for (uint32_t y = 0; y < n; y++) {
        float acc = 0;
        for (uint32_t i = 0; i < nslip; i++) {
            acc += slip[i] * ax1.at(y) * std::pow(ax1.at(y) * dE / E0, static_cast<float>(i));
        }
        offset[y] = acc / ax0.delta;
    }
    return IVS_OK;
}

// src/SM/DriftMap.cpp:21-30 with slip = {angle, a1/a0*angle, a2/a0*angle} (src/main.cpp:663-664)
int ivs_drift_offsets(const ivs_ring_params* o, const ivs_derived* d, float* offset) {
    if (!o || !d || !offset) return IVS_ERR_INVALID_ARGUMENT;
    const float E0 = o->E0;
    return ivs_drift_slip(o->grid_size, d->qmin, d->qmax, d->pmin, d->pmax, d->dE_f, d->slip, 3, E0, offset);
}

// src/SM/FokkerPlanckMap.cpp:31-118, written out as 5 dense weights per row (y-2..y+2)
ivs_status ivs_fokker_planck_weights(uint32_t n, float pmin, float pmax, float e1, uint32_t fptype,
                                     uint32_t derivation, float* w) {
    if (!w || n < 4 || !(pmin < pmax) || fptype > 3 || (derivation != 3 && derivation != 4))
        return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax(n, pmin, pmax);
    const float e1_2d = e1 / (2.0f * ax.delta);
    const float e1_6d = e1 / (6.0f * ax.delta);
    const float e1_d2 = e1 / (ax.delta * ax.delta);
    const bool damping = (fptype == 1 || fptype == 3);
    const bool diffusion = (fptype == 2 || fptype == 3);
    std::fill(w, w + static_cast<size_t>(n) * 5, 0.0f);
    auto at = [&](uint32_t row, int rel) -> float& { return w[row * 5 + (rel + 2)]; };
    if (derivation == 3) {  // two_sided: taps y-1, y, y+1
        for (uint32_t j = 1; j + 1 < n; j++) {
            at(j, 0) = 1;
            if (damping) {
                const float pos = ax.at(j);
                at(j, -1) += -e1_2d * pos;
                at(j, 0) += e1;
                at(j, +1) += e1_2d * pos;
            }
            if (diffusion) {
                at(j, -1) += e1_d2;
                at(j, 0) += -2.0f * e1_d2;
                at(j, +1) += e1_d2;
            }
        }
        return IVS_OK;
    }
    // cubic: rows below the centre use taps y-2..y+1, rows from floor(centre) on y-1..y+2;
    // the second loop overwrites row floor(centre) (:63-116)
    const float ycenter = ax.zerobin;
    for (uint32_t j = 2; j < ycenter; j++) {
        const float pos = ax.at(j);
        at(j, 0) = 1;
        if (damping) {
            at(j, -2) += e1_6d * 1.0f * pos;
            at(j, -1) += e1_6d * -6.0f * pos;
            at(j, 0) += e1 + e1_6d * 3.0f * pos;
            at(j, +1) += e1_6d * 2.0f * pos;
        }
        if (diffusion) {
            at(j, -1) += e1_d2;
            at(j, 0) += -2.0f * e1_d2;
            at(j, +1) += e1_d2;
        }
    }
    for (uint32_t j = static_cast<uint32_t>(ycenter); j + 2 < n; j++) {
        const float pos = ax.at(j);
        for (int r = -2; r <= 2; r++) at(j, r) = 0;
        at(j, 0) = 1;
        if (damping) {
            at(j, -1) += e1_6d * -2.0f * pos;
            at(j, 0) += e1 + e1_6d * -3.0f * pos;
            at(j, +1) += e1_6d * 6.0f * pos;
            at(j, +2) += e1_6d * -1.0f * pos;
        }
        if (diffusion) {
            at(j, -1) += e1_d2;
            at(j, 0) += -2.0f * e1_d2;
            at(j, +1) += e1_d2;
        }
    }
    for (int r = -2; r <= 2; r++) {  // rows 0,1,n-2,n-1 are zero (:63-70,109-116)
        at(0, r) = 0; at(1, r) = 0; at(n - 2, r) = 0; at(n - 1, r) = 0;
    }
    return IVS_OK;
}

// src/PS/PhaseSpace.cpp:486-501
ivs_status ivs_simpson_weights(uint32_t n, float delta, float* ws) {
    if (!ws || n < 2) return IVS_ERR_INVALID_ARGUMENT;
    const float ca = 3.f;
    float dc = 1;
    const float h03 = delta / 3.0f;
    ws[0] = h03;
    for (uint32_t x = 1; x + 1 < n; x++) {
        ws[x] = h03 * (ca + dc);
        dc = -dc;
    }
    ws[n - 1] = h03;
    return IVS_OK;
}

// src/Z/ImpedanceFactory.cpp:21-84
int ivs_make_impedance(size_t nfreqs, float fmax, double R_bend, double f_rev, double gap,
                       int use_csr, double s, double xi, double collimator, float* z,
                       int* nonzero) {
    if (!z || nfreqs < 2) return IVS_ERR_INVALID_ARGUMENT;
    const double f0 = phys::c / (kTwoPi * R_bend);
    ztable rv(nfreqs, 0);
    bool changed = false;
    auto add = [&](const ztable& t) {
        for (size_t i = 0; i < nfreqs; i++) rv[i] += t[i];
        changed = true;
    };
    if (gap != 0) {
        if (use_csr) {
            if (gap > 0) add(z_parallel_plates(nfreqs, f0, fmax, gap));
            else add(z_free_space(nfreqs, f0, fmax));
        }
        const double radius = std::abs(gap / 2);
        if (s > 0 && xi >= -1) add(z_resistive_wall(nfreqs, f_rev, fmax, phys::c / f_rev, s, xi, radius));
        if (0 < collimator && collimator < radius) add(z_collimator(nfreqs, radius, collimator));
    }
    for (size_t i = 0; i < nfreqs; i++) {
        z[2 * i] = rv[i].real();
        z[2 * i + 1] = rv[i].imag();
    }
    if (nonzero) *nonzero = changed ? 1 : 0;
    return IVS_OK;
}

// src/main.cpp:717-721
int ivs_wake_impedance(const ivs_ring_params* o, const ivs_derived* d, float* z, int* nonzero) {
    if (!o || !d) return IVS_ERR_INVALID_ARGUMENT;
    return ivs_make_impedance(d->padded_bins, d->fmax, d->R_bend, d->f_rev, o->gap, o->use_csr != 0,
                              o->wall_conductivity, o->wall_susceptibility, o->collimator, z, nonzero);
}

// Impedance::readData (src/Z/Impedance.cpp:94-111): `lineno real imag` records; the loop runs while the
// stream is good, so the last record is extracted once more at the end of the file -- its repeated line
// number is what keeps it from being stored twice, and what makes any repeated number skip its record
int ivs_read_impedance_file(const char* path, float* z, size_t cap, size_t* count) {
    if (!path || (!z && cap)) return IVS_ERR_INVALID_ARGUMENT;
    std::ifstream is(path);
    if (!is) return IVS_ERR_INVALID_ARGUMENT;
    size_t lineno = 0, old_lineno = std::numeric_limits<size_t>::max(), n = 0;
    float real = 0, imag = 0;
    while (is.good()) {
        is >> lineno >> real >> imag;
        if (lineno != old_lineno) {
            if (n < cap) { z[2 * n] = real; z[2 * n + 1] = imag; }
            n++;
        }
        old_lineno = lineno;
    }
    if (count) *count = n;
    return IVS_OK;
}

// src/Z/ImpedanceFactory.cpp:70-75 with Impedance::operator+= (src/Z/Impedance.cpp:67-76)
int ivs_add_impedance_file(const char* path, float* z, size_t nfreqs) {
    if (!z) return IVS_ERR_INVALID_ARGUMENT;
    std::vector<float> f(2 * nfreqs);
    size_t count = 0;
    const int rc = ivs_read_impedance_file(path, f.data(), nfreqs, &count);
    if (rc != IVS_OK) return rc;
    if (count < nfreqs) return IVS_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < 2 * nfreqs; i++) z[i] += f[i];  // complex<float> += complex<float>
    return IVS_OK;
}

// src/main.cpp:724-727
int ivs_radiation_impedance(const ivs_ring_params* o, const ivs_derived* d, float* z) {
    if (!o || !d) return IVS_ERR_INVALID_ARGUMENT;
    return ivs_make_impedance(d->padded_bins, d->fmax, d->R_bend, d->f_rev, (o->gap > 0) ? o->gap : -1,
                              1, 0, 0, 0, z, nullptr);
}

// PhaseSpace::gaus (src/PS/PhaseSpace.cpp:471-484): one projection of the Gaussian start distribution
int ivs_gaussian_projection(uint32_t n, float amin, float amax, double zoom, float* proj) {
    if (!proj || n < 2 || !(amin < amax)) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax(n, amin, amax);
    const double zoom2 = zoom * zoom;
    const float norm = 0.398942280401432677939946059934381868L;
    for (uint32_t i = 0; i < n; i++) proj[i] = norm * std::exp((-0.5) * ax.at(i) * ax.at(i) / zoom2);
    return IVS_OK;
}

// src/PS/PhaseSpace.cpp:52-65,457-484 (Gaussian start distribution, normalised), then the
// initial renormalisation of src/main.cpp:528-532.  Sequential fp32 sums like the reference.
int ivs_initial_grid(const ivs_derived* d, uint32_t n, double zoom, float* grid) {
    if (!d || !grid || n < 4) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax0(n, d->qmin, d->qmax), ax1(n, d->pmin, d->pmax);
    std::vector<float> ws(n), p0(n), p1(n), proj(n);
    ivs_simpson_weights(n, ax0.delta, ws.data());
    const double zoom2 = zoom * zoom;
    const float norm = 0.398942280401432677939946059934381868L;
    for (uint32_t i = 0; i < n; i++) {
        p0[i] = norm * std::exp((-0.5) * ax0.at(i) * ax0.at(i) / zoom2);
        p1[i] = norm * std::exp((-0.5) * ax1.at(i) * ax1.at(i) / zoom2);
    }
    for (uint32_t x = 0; x < n; x++)
        for (uint32_t y = 0; y < n; y++) grid[static_cast<size_t>(x) * n + y] = p0[x] * p1[y];
    auto integral = [&]() {
        for (uint32_t x = 0; x < n; x++) {
            float acc = 0;
            const float* row = grid + static_cast<size_t>(x) * n;
            for (uint32_t y = 0; y < n; y++) acc += row[y] * ws[y];
            proj[x] = acc;
        }
        float acc = 0;
        for (uint32_t x = 0; x < n; x++) acc += proj[x] * ws[x];
        return acc;
    };
    auto scale = [&](float filling) {
        for (size_t i = 0; i < static_cast<size_t>(n) * n; i++) grid[i] *= 1.0f / filling;
    };
    scale(integral());  // createFromProjections: normalize()
    scale(integral());  // constructor integrate(), then main(): normalize()
    return IVS_OK;
}

// RFKickMap constructors (src/SM/RFKickMap.cpp:15-57) with the arguments of src/main.cpp:600-650
int ivs_rf_model_init(const ivs_ring_params* o, const ivs_derived* d, void* model) {
    if (!o || !d || !model) return IVS_ERR_INVALID_ARGUMENT;
    ivs_rf_model* m = static_cast<ivs_rf_model*>(model);
    const Axis ax0(o->grid_size, d->qmin, d->qmax), ax1(o->grid_size, d->pmin, d->pmax);
    const float f_RF = static_cast<float>(d->f_RF);
    std::memset(m, 0, sizeof(*m));
    m->linear = o->linear_rf ? 1u : 0u;
    m->bl2phase = static_cast<float>(static_cast<double>(d->bl_f) / phys::c * static_cast<double>(f_RF) * kTwoPi);
    m->xcenter = ax0.zerobin;
    m->qmin = ax0.min;
    m->delta0 = ax0.delta;
    m->delta1 = ax1.delta;
    m->dE_scale = d->dE_f;
    if (o->linear_rf) {
        m->tan_angle = std::tan(d->angle);
        m->syncphase = 0;
    } else {
        m->revolutionpart = static_cast<float>(d->revolutionpart);
        m->V_RF = static_cast<float>(d->V_eff);
        m->V0 = static_cast<float>(d->V0);
        m->syncphase = std::asin(m->V0 / m->V_RF);
    }
    return IVS_OK;
}

namespace {
inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
}  // namespace

// DynamicRFKickMap::__calcModulation (src/SM/DynamicRFKickMap.cpp:80-94)
int ivs_rf_modulation(uint64_t seed, uint32_t steps, float phasespread, float amplspread, float modampl,
                      double modtimeincrement, double revolutionpart, float syncphase, const float* gauss,
                      float* out) {
    if (!out || !(revolutionpart > 0)) return IVS_ERR_INVALID_ARGUMENT;
    const float phasenoise_w = static_cast<float>(phasespread / std::sqrt(revolutionpart));
    const float amplnoise_w = static_cast<float>(amplspread / std::sqrt(revolutionpart));
    const float modtimedelta = static_cast<float>(kTwoPi * modtimeincrement);
    for (uint32_t i = 0; i < steps; i++) {
        float g0, g1;
        if (gauss) {
            g0 = gauss[2 * i];
            g1 = gauss[2 * i + 1];
        } else {  // Box-Muller on two counter-based uniforms
            const uint64_t h = mix64(seed ^ mix64(i));
            const double u1 = (static_cast<double>(h >> 11) + 1.0) * (1.0 / 9007199254740993.0);
            const double u2 = static_cast<double>(mix64(h) >> 11) * (1.0 / 9007199254740992.0);
            const double r = std::sqrt(-2.0 * std::log(u1));
            g0 = static_cast<float>(r * std::cos(kTwoPi * u2));
            g1 = static_cast<float>(r * std::sin(kTwoPi * u2));
        }
        const float phasenoise = g0 * phasenoise_w;
        const float amplnoise = g1 * amplnoise_w;
        const float phasemod = modampl * std::sin(modtimedelta * static_cast<float>(i));
        out[2 * i] = syncphase + phasenoise + phasemod;
        out[2 * i + 1] = 1 + amplnoise;
    }
    return IVS_OK;
}

// Haissinski start distribution (BASELINE.json configs[2] "Haissinski-initialised"; the reference
// parses HaissinskiIterations and ignores it, src/IO/ProgramOptions.cpp:270-272, so this is an
// addition: the grid goes to PhaseSpace's `data` constructor argument, inc/PS/PhaseSpace.hpp:60).
// Per step a particle moves by  q' = -angle p  (DriftMap, slip[0] = angle) and
// p' = -(rf[x] + W[x]) dp  (RF + wake kick in cells, KickMap::applyTo subtracts the offset,
// src/SM/KickMap.cpp:280-303), i.e. H = -angle p^2/2 + dp int (rf + W) dq, and with the
// Fokker-Planck term the stationary density is  f = exp(H / angle) = exp(-p^2/2) rho(q),
//     rho(q) ~ exp( (dp / angle) int^q (rf + W[rho]) dq' ),
// W[rho] = wakePotential() of rho.  Solved by damped fixed-point iteration in double
// (trapezoidal cumulative integral from the grid centre, Simpson-normalised like PhaseSpace).
// Returns 0 when the relative L1 change of the last iteration is below 1e-9, 1 otherwise.
int ivs_haissinski_grid(const ivs_derived* d, uint32_t n, const float* rf_offset, const float* zwake,
                        uint32_t nmax, float wakescaling, uint32_t iterations, double relax, float* grid,
                        double* residual) {
    if (!d || !grid || !rf_offset || n < 4 || !(relax > 0 && relax <= 1)) return IVS_ERR_INVALID_ARGUMENT;
    if (zwake && (nmax < 2 * n || (nmax & (nmax - 1)))) return IVS_ERR_INVALID_ARGUMENT;
    const Axis ax0(n, d->qmin, d->qmax), ax1(n, d->pmin, d->pmax);
    std::vector<float> wsf(n);
    ivs_simpson_weights(n, ax0.delta, wsf.data());
    std::vector<double> ws(wsf.begin(), wsf.end()), rho(n), next(n), force(n), pot(n), w;
    if (zwake) w = ivs_wake_function_host(zwake, nmax, n);
    const double dq = ax0.delta, k = (double)ax1.delta / (double)d->angle;
    auto normalise = [&](std::vector<double>& r) {
        double q = 0;
        for (uint32_t x = 0; x < n; x++) q += r[x] * ws[x];
        for (uint32_t x = 0; x < n; x++) r[x] /= q;
    };
    for (uint32_t x = 0; x < n; x++) rho[x] = std::exp(-0.5 * (double)ax0.at(x) * (double)ax0.at(x));
    normalise(rho);
    double res = 1.0;
    for (uint32_t it = 0; it < iterations; it++) {
        for (uint32_t x = 0; x < n; x++) {
            double acc = 0;
            if (zwake)
                for (uint32_t xp = 0; xp < n; xp++) acc += rho[xp] * w[x + (n - 1) - xp];
            force[x] = (double)rf_offset[x] + (double)wakescaling * acc;
        }
        pot[0] = 0;
        for (uint32_t x = 1; x < n; x++) pot[x] = pot[x - 1] + 0.5 * (force[x] + force[x - 1]) * dq;
        const double p0 = pot[n / 2];
        for (uint32_t x = 0; x < n; x++) next[x] = std::exp(k * (pot[x] - p0));
        normalise(next);
        double num = 0, den = 0;
        for (uint32_t x = 0; x < n; x++) {
            num += std::fabs(next[x] - rho[x]);
            den += std::fabs(rho[x]);
            rho[x] = (1 - relax) * rho[x] + relax * next[x];
        }
        res = num / den;
        if (res < 1e-12) break;
    }
    if (residual) *residual = res;
    // f = rho(q) gauss(p), normalised to integral 1 with the Simpson weights on both axes
    std::vector<double> py(n);
    for (uint32_t y = 0; y < n; y++) py[y] = std::exp(-0.5 * (double)ax1.at(y) * (double)ax1.at(y));
    normalise(py);
    for (uint32_t x = 0; x < n; x++)
        for (uint32_t y = 0; y < n; y++) grid[(size_t)x * n + y] = (float)(rho[x] * py[y]);
    return res < 1e-9 ? 0 : 1;
}

}  // extern "C"
