#include "ufo_tc_inst.cuh"
namespace ufo {
UFO_TC_DEFINE_PASS(tc_pass_bf16_hi, true, UFO_TC_CASE(6, true) UFO_TC_CASE(7, true) UFO_TC_CASE(8, true) UFO_TC_CASE(9, true) UFO_TC_CASE(10, true))
}  // namespace ufo
