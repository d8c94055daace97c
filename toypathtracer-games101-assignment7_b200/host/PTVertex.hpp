// PTVertex.hpp — forwarding header: the host API lives in tpt_api.hpp (the functions this header
// declares in the reference run on the GPU here, see include/tpt.h).
#pragma once
#include "tpt_api.hpp"
