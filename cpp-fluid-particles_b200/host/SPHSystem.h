// SPHSystem.h -- forwarding header: the reference's include name (/root/reference/src/SPHSystem.h) kept so that
// its call sites (main.cpp:26-33) compile unchanged; the B200-native classes live in sph_api.hpp.
#pragma once
#include "sph_api.hpp"
