// Forwarding header: the multi-GPU system of the class layer (no counterpart in the reference, which is single-GPU).
#pragma once
#include "sph_slab.hpp"
