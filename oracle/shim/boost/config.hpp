// Minimal stand-in for <boost/config.hpp> -- TEST INFRASTRUCTURE ONLY.
#pragma once
#define BOOST_COMPILER "GNU C++ (shimmed boost)"
