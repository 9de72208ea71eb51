// Stand-in for the CMake-generated InovesaConfig.hpp (InovesaConfig.hpp.in) -- TEST INFRASTRUCTURE ONLY.
#pragma once
#define INOVESA_VERSION_MAJOR 1
#define INOVESA_VERSION_MINOR 2
#define INOVESA_VERSION_FIX -1
