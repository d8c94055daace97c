// Bounds3.hpp — forwarding header: the host API lives in tpt_api.hpp.
#pragma once
#include "tpt_api.hpp"
