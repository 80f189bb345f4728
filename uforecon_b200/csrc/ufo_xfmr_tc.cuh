// Tensor-core (tcgen05) path of kernel 3 - placeholder until the fused kernel lands.
#pragma once
#include "ufo_common.cuh"

namespace ufo {
struct TcWeights { void* blob = nullptr; };
inline int tc_weights_create(const UfoWeightsDesc*, TcWeights*, cudaStream_t) { return UFO_OK; }
inline void tc_weights_destroy(TcWeights*) {}
inline void tc_scene_release(int) {}
inline int tc_render_rays(const SceneDev&, const TcWeights&, float, const int64_t*, int64_t, int32_t, const float*, const float*,
                          int64_t, const UfoRenderOut*, const UfoDebugTaps*, int, int, cudaStream_t) {
  return fail(UFO_EINVAL, "UFO_MODE_TC is not built in this revision");
}
}  // namespace ufo
