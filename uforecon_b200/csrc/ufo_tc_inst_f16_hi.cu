#include "ufo_tc_inst.cuh"
namespace ufo {
UFO_TC_DEFINE_PASS(tc_pass_f16_hi, false, UFO_TC_CASE(6, false) UFO_TC_CASE(7, false) UFO_TC_CASE(8, false) UFO_TC_CASE(9, false) UFO_TC_CASE(10, false))
}  // namespace ufo
