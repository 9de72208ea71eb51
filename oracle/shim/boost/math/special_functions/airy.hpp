// Minimal stand-in for <boost/math/special_functions/airy.hpp> -- TEST INFRASTRUCTURE ONLY.
//
// Airy functions Ai, Bi, Ai', Bi' in double for real arguments, expressed through the
// modified Bessel functions of libstdc++ (C++17 <cmath> special functions):
//   x > 0, z = 2/3 x^(3/2):
//     Ai (x) =  1/pi sqrt(x/3) K_{1/3}(z)          Ai'(x) = -x/(pi sqrt 3) K_{2/3}(z)
//     Bi (x) =  sqrt(x/3) (I_{-1/3}(z)+I_{1/3}(z))  Bi'(x) =  x/sqrt 3 (I_{-2/3}(z)+I_{2/3}(z))
//     with I_{-v} = I_v + 2/pi sin(v pi) K_v
//   x < 0 through J/Y of order 1/3, 2/3 (not reached by the reference's call site,
//   /root/reference/src/Z/ParallelPlatesCSR.cpp:48-53, where the argument is > 0).
// Like Boost's default policy, Bi and Bi' raise std::overflow_error when the result
// does not fit a double (the reference relies on this to stop its series), while the
// exponentially small Ai, Ai' silently underflow to 0.
// Parity note: this is not Boost's implementation; values agree with scipy.special.airy
// to ~1e-13 relative (tests/test_impedance.py), the overflow threshold is approximate.
#pragma once
#include <cmath>
#include <limits>
#include <stdexcept>

namespace boost { namespace math {
namespace airy_shim {
constexpr double pi = 3.141592653589793238462643383279502884;
constexpr double sqrt3 = 1.732050807568877293527446341505872367;
constexpr double ai0 = 0.355028053887817239260063186004183176;
constexpr double aip0 = -0.258819403792806798405183560189203963;
inline double zeta(double x) { return 2.0 / 3.0 * x * std::sqrt(x); }
inline void check_overflow(double x, double z) {
    // Bi ~ exp(z)/(sqrt(pi) x^(1/4)), Bi' ~ x^(1/4) exp(z)/sqrt(pi)
    if (z + 0.25 * std::log(x) - 0.5723649429247001 > 709.78) {
        throw std::overflow_error("airy_bi: overflow");
    }
}
}  // namespace airy_shim

inline double airy_ai(double x) {
    using namespace airy_shim;
    if (x == 0) return ai0;
    if (x > 0) {
        const double z = zeta(x);
        if (z > 740) return 0;
        return std::sqrt(x / 3.0) / pi * std::cyl_bessel_k(1.0 / 3.0, z);
    }
    const double ax = -x, z = zeta(ax);
    const double j = std::cyl_bessel_j(1.0 / 3.0, z), y = std::cyl_neumann(1.0 / 3.0, z);
    // J_{-1/3} = cos(pi/3) J_{1/3} - sin(pi/3) Y_{1/3}
    const double jm = 0.5 * j - sqrt3 / 2 * y;
    return std::sqrt(ax) / 3.0 * (j + jm);
}
inline double airy_ai_prime(double x) {
    using namespace airy_shim;
    if (x == 0) return aip0;
    if (x > 0) {
        const double z = zeta(x);
        if (z > 740) return 0;
        return -x / (pi * sqrt3) * std::cyl_bessel_k(2.0 / 3.0, z);
    }
    const double ax = -x, z = zeta(ax);
    const double j = std::cyl_bessel_j(2.0 / 3.0, z), y = std::cyl_neumann(2.0 / 3.0, z);
    const double jm = -0.5 * j - sqrt3 / 2 * y;  // J_{-2/3}
    return ax / 3.0 * (j - jm);
}
inline double airy_bi(double x) {
    using namespace airy_shim;
    if (x == 0) return ai0 * sqrt3;
    if (x > 0) {
        const double z = zeta(x);
        check_overflow(x, z);
        const double i = std::cyl_bessel_i(1.0 / 3.0, z);
        const double k = (z > 740) ? 0.0 : std::cyl_bessel_k(1.0 / 3.0, z);
        const double im = i + 2.0 / pi * (sqrt3 / 2) * k;
        return std::sqrt(x / 3.0) * (im + i);
    }
    const double ax = -x, z = zeta(ax);
    const double j = std::cyl_bessel_j(1.0 / 3.0, z), y = std::cyl_neumann(1.0 / 3.0, z);
    const double jm = 0.5 * j - sqrt3 / 2 * y;
    return std::sqrt(ax / 3.0) * (jm - j);
}
inline double airy_bi_prime(double x) {
    using namespace airy_shim;
    if (x == 0) return -aip0 * sqrt3;
    if (x > 0) {
        const double z = zeta(x);
        check_overflow(x, z);
        const double i = std::cyl_bessel_i(2.0 / 3.0, z);
        const double k = (z > 740) ? 0.0 : std::cyl_bessel_k(2.0 / 3.0, z);
        const double im = i + 2.0 / pi * (sqrt3 / 2) * k;
        return x / sqrt3 * (im + i);
    }
    const double ax = -x, z = zeta(ax);
    const double j = std::cyl_bessel_j(2.0 / 3.0, z), y = std::cyl_neumann(2.0 / 3.0, z);
    const double jm = -0.5 * j - sqrt3 / 2 * y;
    return ax / sqrt3 * (jm + j);
}
}}  // namespace boost::math
