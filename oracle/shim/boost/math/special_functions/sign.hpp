// Minimal stand-in for <boost/math/special_functions/sign.hpp> -- TEST INFRASTRUCTURE ONLY.
#pragma once
namespace boost { namespace math {
template <typename T> inline int sign(const T& z) { return (z == 0) ? 0 : ((z < 0) ? -1 : 1); }
}}
