#include "ufo_tc_inst.cuh"
namespace ufo {
UFO_TC_DEFINE_PASS(tc_pass_f16_lo, false, UFO_TC_CASE(2, false) UFO_TC_CASE(3, false) UFO_TC_CASE(4, false) UFO_TC_CASE(5, false))
}  // namespace ufo
