// Launchers of the tensor-core pipeline, instantiated per (operand format, view-count group) in separate
// translation units so that the library builds in parallel.
#pragma once
#include "ufo_handles.cuh"
#include "ufo_xfmr_tc.cuh"

namespace ufo {
template <int NV, bool BF16>
static int launch_tc_pass(const UfoScene* sc, const UfoWeights* w, int R, int SN, const float* z, float* sim8_tap, float* pts,
                          float* ray_out_tap, int sms, cudaStream_t st) {
  const TcWorkspace& ws = sc->tws;
  const long long P = (long long)R * SN;
  const int f = BF16 ? 0 : 1;
  UFO_KERNEL("k_gather_tc", st, k_gather_tc<NV, BF16><<<cdiv(P, 256), 256, 0, st>>>(sc->d, ws.rayinfo, z, R, SN, w->freqs, w->phases, w->pre_sim,
                                                                                  ws.tok, ws.rgbm, ws.dirs, sim8_tap, pts));
  {
    static bool attr = false;
    if (!attr) {
      UFO_CUDA(cudaFuncSetAttribute(k_view_tc<NV, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::V_SMEM));
      attr = true;
    }
    constexpr int PPT = 128 / (NV + 1);
    const long long tiles = (P + PPT - 1) / PPT;
    const int grid = (int)(tiles < sms ? tiles : sms);
    UFO_KERNEL("k_view_tc", st, k_view_tc<NV, BF16><<<grid, tc::kThreads, tc::V_SMEM, st>>>(w->tc.view_img[f], w->tc.vp, ws.tok, ws.rgbm, ws.dirs, P,
                                                                                           ws.vout0, ws.radiance));
  }
  {
    const long long tiles = (P + 127) / 128;
    const int grid = (int)(tiles < sms ? tiles : sms);
    if (SN == kNC) {
      static bool attr = false;
      if (!attr) { UFO_CUDA(cudaFuncSetAttribute(k_ray_tc<kNC, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::R_SMEM)); attr = true; }
      UFO_KERNEL("k_ray_tc", st, k_ray_tc<kNC, BF16><<<grid, tc::kThreads, tc::R_SMEM, st>>>(w->tc.ray_img[f], w->tc.rp, ws.vout0, w->pe_table, P, ws.srdf, ray_out_tap));
    } else {
      static bool attr = false;
      if (!attr) { UFO_CUDA(cudaFuncSetAttribute(k_ray_tc<kNS, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::R_SMEM)); attr = true; }
      UFO_KERNEL("k_ray_tc", st, k_ray_tc<kNS, BF16><<<grid, tc::kThreads, tc::R_SMEM, st>>>(w->tc.ray_img[f], w->tc.rp, ws.vout0, w->pe_table, P, ws.srdf, ray_out_tap));
    }
  }
  return UFO_OK;
}


#define UFO_TC_DEFINE_PASS(fn, BF16, ...)                                                                                   \
  int fn(const UfoScene* sc, const UfoWeights* w, int R, int SN, const float* z, float* sim8_tap, float* pts,              \
         float* ray_out_tap, int sms, cudaStream_t st) {                                                                     \
    switch (sc->d.nv) { __VA_ARGS__ }                                                                                        \
    return fail(UFO_EINVAL, "unsupported n_views");                                                                          \
  }
#define UFO_TC_CASE(NV, BF16) \
  case NV: return launch_tc_pass<NV, BF16>(sc, w, R, SN, z, sim8_tap, pts, ray_out_tap, sms, st);
}  // namespace ufo
