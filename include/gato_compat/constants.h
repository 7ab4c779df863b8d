// Shim for gato/constants.h:10-26 (sizes derived from the plant and KNOT_POINTS).
#pragma once
#include <cstdint>
#include "settings.h"
using namespace sqp;
namespace gato { namespace constants {
constexpr uint32_t REFERENCE_TRAJ_SIZE = grid::EE_POS_SIZE * KNOT_POINTS;
constexpr uint32_t STATE_SIZE = grid::NUM_JOINTS * 2;
constexpr uint32_t CONTROL_SIZE = grid::NUM_JOINTS;
constexpr uint32_t TRAJ_SIZE = (STATE_SIZE + CONTROL_SIZE) * KNOT_POINTS - CONTROL_SIZE;
constexpr uint32_t VEC_SIZE_PADDED = (KNOT_POINTS + 2) * STATE_SIZE;
}}  // namespace gato::constants
