// Launchers of the tensor-core pipeline, instantiated per (operand format, view-count group) in separate
// translation units so that the library builds in parallel.
#pragma once
#include "ufo_handles.cuh"
#include "ufo_xfmr_tc.cuh"
#include "ufo_view_tc2.cuh"
#include "ufo_ray_tc2.cuh"
#include <cstdlib>

namespace ufo {
// One sample2rgb pass of the tensor-core pipeline over the R*64 points of half `half` (0: coarse samples,
// 1: importance samples): gather -> view stage, then the ray stage over the 64 coarse samples (half 0) or over
// all 128 samples in sorted order through ws.perm (half 1).
template <int NV, bool BF16>
static int launch_tc_pass(const UfoScene* sc, const UfoWeights* w, int R, int half, const float* z, bool want_sim8,
                          float* ray_out_tap, int sms, cudaStream_t st) {
  const TcWorkspace& ws = sc->tws;
  const long long P = (long long)R * kNC;
  const int f = BF16 ? 0 : 1;
  {
    UFO_SMEM_ATTR((k_gather_tc<NV, BF16>), gather_tc_smem<NV>());
  }
  UFO_KERNEL("k_gather_tc", st, k_gather_tc<NV, BF16><<<cdiv(P, 256), 256, gather_tc_smem<NV>(), st>>>(sc->d, ws.rayinfo, z, R, half, w->freqs, w->phases, w->pre_sim,
                                                                                  ws.tok, ws.rgbm, ws.dirs, want_sim8 ? ws.sim8 : nullptr));
  // Generation 2 (two tiles in flight; the token rows of a point aligned to a warp, or to a warp pair where a single warp would leave
  // > 5 % more rows idle: NV = 6, 8, 10) is the default for every view count.  Measured on a B200 at 1600x1216, ms per depth map in
  // the view stage, generation 2 vs generation 1 (lock step): NV=3 243 vs 360, NV=5 395 vs 611, NV=10 1372 vs 1467 (with warp-aligned
  // points only 22 of 32 rows were in use at NV=10 and generation 2 lost: 1640).  UFO_VIEW_KERNEL=1 selects generation 1.
  static const int view_env = getenv("UFO_VIEW_KERNEL") ? atoi(getenv("UFO_VIEW_KERNEL")) : 0;
  const int view_gen = view_env ? view_env : 2;
  static_assert(tc::V2_WEND == 141312, "k_view_tc2 weight image size (mirrored in ufo_api.cu)");
  if (view_gen == 2) {     // two tiles in flight per CTA (ufo_view_tc2.cuh)
    UFO_SMEM_ATTR((k_view_tc2<NV, BF16>), (int)tc::V2_SMEM);
    constexpr int PPT = tc::view2_points_per_tile(NV);
    const long long tiles = (P + PPT - 1) / PPT;
    const long long pairs = (tiles + 1) / 2;
    const int grid = (int)(pairs < sms ? pairs : sms);
    UFO_KERNEL("k_view_tc", st, k_view_tc2<NV, BF16><<<grid, 512, tc::V2_SMEM, st>>>(w->tc.view_img2[f], w->tc.vp, ws.tok, ws.rgbm, ws.dirs, (int)P, half,
                                                                                    ws.vout0, ws.radiance));
  } else {
    UFO_SMEM_ATTR((k_view_tc<NV, BF16>), (int)tc::V_SMEM);
    constexpr int PPT = 128 / (NV + 1);
    const long long tiles = (P + PPT - 1) / PPT;
    const int grid = (int)(tiles < sms ? tiles : sms);
    UFO_KERNEL("k_view_tc", st, k_view_tc<NV, BF16><<<grid, tc::kThreads, tc::V_SMEM, st>>>(w->tc.view_img[f], w->tc.vp, ws.tok, ws.rgbm, ws.dirs, (int)P, half,
                                                                                           ws.vout0, ws.radiance));
  }
  static const int ray_gen = getenv("UFO_RAY_KERNEL") ? atoi(getenv("UFO_RAY_KERNEL")) : 2;
  static_assert(tc::R2W_END == 178688, "k_ray_tc2 weight image size (mirrored in ufo_api.cu)");
  const long long PR = half == 0 ? P : (long long)R * kNS;           // tokens of the ray stage: 64 or all 128 per ray
  const long long rtiles = (PR + 127) / 128;
  if (ray_gen == 2) {      // two CTAs per SM (ufo_ray_tc2.cuh)
    const int grid = (int)(rtiles < 2 * sms ? rtiles : 2 * sms);
    if (half == 0) {
      UFO_SMEM_ATTR((k_ray_tc2<kNC, BF16>), (int)tc::R2_SMEM);
      UFO_KERNEL("k_ray_tc", st, k_ray_tc2<kNC, BF16><<<grid, 256, tc::R2_SMEM, st>>>(w->tc.ray_img2[f], w->tc.rp, ws.vout0, w->pe_table, nullptr, PR, ws.srdf, ray_out_tap));
    } else {
      UFO_SMEM_ATTR((k_ray_tc2<kNS, BF16>), (int)tc::R2_SMEM);
      UFO_KERNEL("k_ray_tc", st, k_ray_tc2<kNS, BF16><<<grid, 256, tc::R2_SMEM, st>>>(w->tc.ray_img2[f], w->tc.rp, ws.vout0, w->pe_table, ws.perm, PR, ws.srdf, ray_out_tap));
    }
  } else if (half == 0) {
    const int grid = (int)(rtiles < sms ? rtiles : sms);
    UFO_SMEM_ATTR((k_ray_tc<kNC, BF16>), (int)tc::R_SMEM);
    UFO_KERNEL("k_ray_tc", st, k_ray_tc<kNC, BF16><<<grid, tc::kThreads, tc::R_SMEM, st>>>(w->tc.ray_img[f], w->tc.rp, ws.vout0, w->pe_table, nullptr, PR, ws.srdf, ray_out_tap));
  } else {
    const int grid = (int)(rtiles < sms ? rtiles : sms);
    UFO_SMEM_ATTR((k_ray_tc<kNS, BF16>), (int)tc::R_SMEM);
    UFO_KERNEL("k_ray_tc", st, k_ray_tc<kNS, BF16><<<grid, tc::kThreads, tc::R_SMEM, st>>>(w->tc.ray_img[f], w->tc.rp, ws.vout0, w->pe_table, ws.perm, PR, ws.srdf, ray_out_tap));
  }
  return UFO_OK;
}

#define UFO_TC_DEFINE_PASS(fn, BF16, ...)                                                                                   \
  int fn(const UfoScene* sc, const UfoWeights* w, int R, int half, const float* z, bool want_sim8,                          \
         float* ray_out_tap, int sms, cudaStream_t st) {                                                                     \
    switch (sc->d.nv) { __VA_ARGS__ }                                                                                        \
    return fail(UFO_EINVAL, "unsupported n_views");                                                                          \
  }
#define UFO_TC_CASE(NV, BF16) \
  case NV: return launch_tc_pass<NV, BF16>(sc, w, R, half, z, want_sim8, ray_out_tap, sms, st);
}  // namespace ufo
