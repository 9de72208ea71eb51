// Minimal stand-in for <boost/math/constants/constants.hpp> -- TEST INFRASTRUCTURE ONLY
// (see oracle/shim/boost/multi_array.hpp for why this exists).
#pragma once
namespace boost { namespace math { namespace constants {
template <typename T> constexpr T pi() { return static_cast<T>(3.141592653589793238462643383279502884L); }
template <typename T> constexpr T two_pi() { return static_cast<T>(6.283185307179586476925286766559005768L); }
template <typename T> constexpr T half_pi() { return static_cast<T>(1.570796326794896619231321691639751442L); }
template <typename T> constexpr T pi_sqr() { return static_cast<T>(9.869604401089358618834490999876151135L); }
template <typename T> constexpr T one_div_root_two_pi() { return static_cast<T>(0.398942280401432677939946059934381868L); }
}}}
