#include "ufo_tc_inst.cuh"
namespace ufo {
UFO_TC_DEFINE_PASS(tc_pass_bf16_lo, true, UFO_TC_CASE(2, true) UFO_TC_CASE(3, true) UFO_TC_CASE(4, true) UFO_TC_CASE(5, true))
}  // namespace ufo
