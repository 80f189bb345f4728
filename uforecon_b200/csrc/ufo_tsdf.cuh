// "Next" row N3 of SURVEY.md section 8f: TSDF integration of rendered depth maps.
//
// Replaces the reference's only native kernel, the PyCUDA `integrate` (tsdf_fusion.py:77-152, launch :241-266):
// one thread per voxel - voxel -> world -> camera (cam_pose is camera-to-world, the kernel applies R^T (p - t)) ->
// pixel (roundf) -> frustum / depth tests -> truncated distance -> running average.  The colour branch of the
// reference is dead code (`return` at :137).  Differences by design: (1) up to kTsdfMaxViews depth maps are
// integrated per launch, in order, with the voxel's (tsdf, weight) held in registers - the reference reads and writes
// the two volumes once per view and ships its scalar arguments through host<->device copies on every call
// (`cuda.InOut`); (2) voxel coordinates come from integer division (the reference divides in float, which misplaces
// voxels once the volume has more than 2^24 of them); (3) the range guard is `>=` (the reference's `>` lets one
// out-of-range thread through).  Arithmetic is fp32 with one rounding per operation in the reference's order, so
// the result is bit-identical to oracle/tsdf_oracle.py.
#pragma once
#include "ufo_common.cuh"

namespace ufo {

constexpr int kTsdfMaxViews = 16;

struct TsdfViewDev {
  const float* depth;   // [H][W]
  int im_h, im_w;
  float fx, cx, fy, cy;
  float r[9];           // cam_pose[:3,:3] row-major
  float t[3];           // cam_pose[:3,3]
};
struct TsdfLaunch {
  int n_views;
  TsdfViewDev v[kTsdfMaxViews];
};

__global__ void __launch_bounds__(256) k_tsdf_integrate(float* __restrict__ tsdf, float* __restrict__ weight, int X, int Y, int Z,
                                                        float ox, float oy, float oz, float voxel_size, float trunc,
                                                        float obs_weight, const __grid_constant__ TsdfLaunch L) {
  const long long total = (long long)X * Y * Z;
  const int yz = Y * Z;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int vx = (int)(idx / yz);
    const int rem = (int)(idx - (long long)vx * yz);
    const int vy = rem / Z, vz = rem - vy * Z;
    const float px = __fadd_rn(ox, __fmul_rn((float)vx, voxel_size));
    const float py = __fadd_rn(oy, __fmul_rn((float)vy, voxel_size));
    const float pz = __fadd_rn(oz, __fmul_rn((float)vz, voxel_size));
    float t_cur = 0.f, w_cur = 0.f;
    bool loaded = false;
    for (int i = 0; i < L.n_views; ++i) {
      const TsdfViewDev& v = L.v[i];
      const float tx = __fsub_rn(px, v.t[0]), ty = __fsub_rn(py, v.t[1]), tz = __fsub_rn(pz, v.t[2]);
      const float cx = __fadd_rn(__fadd_rn(__fmul_rn(v.r[0], tx), __fmul_rn(v.r[3], ty)), __fmul_rn(v.r[6], tz));
      const float cy = __fadd_rn(__fadd_rn(__fmul_rn(v.r[1], tx), __fmul_rn(v.r[4], ty)), __fmul_rn(v.r[7], tz));
      const float cz = __fadd_rn(__fadd_rn(__fmul_rn(v.r[2], tx), __fmul_rn(v.r[5], ty)), __fmul_rn(v.r[8], tz));
      const float fxp = roundf(__fadd_rn(__fmul_rn(v.fx, __fdiv_rn(cx, cz)), v.cx));
      const float fyp = roundf(__fadd_rn(__fmul_rn(v.fy, __fdiv_rn(cy, cz)), v.cy));
      if (!(fxp >= 0.f && fxp < (float)v.im_w && fyp >= 0.f && fyp < (float)v.im_h) || cz < 0.f) continue;   // also NaN/inf
      const float d = __ldg(v.depth + (size_t)((int)fyp) * v.im_w + (int)fxp);
      if (d == 0.f) continue;
      const float diff = __fsub_rn(d, cz);
      if (diff < -trunc) continue;
      const float dist = fminf(1.0f, __fdiv_rn(diff, trunc));
      if (!loaded) {
        t_cur = tsdf[idx];
        w_cur = weight[idx];
        loaded = true;
      }
      const float w_new = __fadd_rn(w_cur, obs_weight);
      t_cur = __fdiv_rn(__fadd_rn(__fmul_rn(t_cur, w_cur), __fmul_rn(obs_weight, dist)), w_new);
      w_cur = w_new;
    }
    if (loaded) {
      tsdf[idx] = t_cur;
      weight[idx] = w_cur;
    }
  }
}

}  // namespace ufo
