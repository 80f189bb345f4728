// Row a19 of SURVEY.md section 8a (alternative --volume_type featuregrid): the variance flavour of the feature-volume
// build, FeatureVolume.forward up to its 3-D regulariser (code1/feature_volume.py:40-92): project the reso^3 grid of
// the unit cube into every source view, bilinear gather (align_corners=False, zeros) of the 32-channel features,
// Linear 32->32->16->8 per (voxel, view), in-image / in-front masked mean and variance over views -> [16][Z][Y][X].
// One thread per voxel; the feature maps are channel-last (one texel = one 128-byte line), the MLP weights sit in
// shared memory.  The reference materialises [NV, 32, reso^3] gathered features and three MLP activations in HBM.
#pragma once
#include "ufo_common.cuh"
#include "ufo_gather.cuh"

namespace ufo {

struct FGridViews {
  int nv;
  float P[kMaxV][12];     // rows 0..2 of source_poses
};

constexpr int kFGridWFloats = 32 * 32 + 32 + 16 * 32 + 16 + 8 * 16 + 8;

__global__ void __launch_bounds__(128) k_feature_grid(const float* __restrict__ feat_cl /*[NV][h][w][32]*/, int h, int w, int reso,
                                                      const __grid_constant__ FGridViews V, const float* __restrict__ wts /*packed*/,
                                                      float* __restrict__ out /*[16][Z][Y][X]*/) {
  __shared__ __align__(16) float s_w[kFGridWFloats];
  for (int i = threadIdx.x; i < kFGridWFloats; i += blockDim.x) s_w[i] = __ldg(wts + i);
  __syncthreads();
  const float* w0 = s_w;                 // [32][32]
  const float* b0 = w0 + 32 * 32;
  const float* w2 = b0 + 32;             // [16][32]
  const float* b2 = w2 + 16 * 32;
  const float* w4 = b2 + 16;             // [8][16]
  const float* b4 = w4 + 8 * 16;
  const long long total = (long long)reso * reso * reso;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (ix*reso + iy)*reso + iz, feature_volume.py:28-29
  if (idx >= total) return;
  const int iz = (int)(idx % reso), iy = (int)((idx / reso) % reso), ix = (int)(idx / ((long long)reso * reso));
  // grid line: linspace(0, reso-1, reso) * 2 / (reso-1) - 1 in float64, then cast (feature_volume.py:23-25, :48)
  const float x = (float)((double)ix * 2.0 / (double)(reso - 1) - 1.0);
  const float y = (float)((double)iy * 2.0 / (double)(reso - 1) - 1.0);
  const float z = (float)((double)iz * 2.0 / (double)(reso - 1) - 1.0);
  float c[kMaxV][8];
  float m[kMaxV];
  float msum = 0.f;
  const size_t fstride = (size_t)h * w * kFeatC;
  for (int n = 0; n < V.nv; ++n) {
    float u, v, qz;
    project_pt(V.P[n], x, y, z, u, v, qz);
    const bool inb = (u <= 1.f) && (u >= -1.f) && (v <= 1.f) && (v >= -1.f);
    m[n] = (inb && qz > 0.f) ? 1.f : 0.f;                                   // :58, :74
    msum += m[n];
    const BilTaps t = bil_setup<false, false>(u, v, h, w);
    float f[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = bil_fetch32(feat_cl + n * fstride, t, j);
      f[4 * j] = a.x; f[4 * j + 1] = a.y; f[4 * j + 2] = a.z; f[4 * j + 3] = a.w;
    }
    float h1[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      float acc = b0[o];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(f[i], w0[o * 32 + i], acc);
      h1[o] = fmaxf(acc, 0.f);
    }
    float h2[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float acc = b2[o];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(h1[i], w2[o * 32 + i], acc);
      h2[o] = fmaxf(acc, 0.f);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float acc = b4[o];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc = fmaf(h2[i], w4[o * 16 + i], acc);
      c[n][o] = acc;
    }
  }
  const float den = msum + 1e-8f;                                            // :79
  float mean[8], var[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float a = 0.f;
    for (int n = 0; n < V.nv; ++n) a += c[n][o] * (m[n] / den);              // :86
    mean[o] = a;
    float q = 0.f;
    for (int n = 0; n < V.nv; ++n) {
      const float d = c[n][o] - a;
      q += (m[n] / den) * (d * d);                                           // :87
    }
    var[o] = q;
  }
  const size_t vox = ((size_t)iz * reso + iy) * reso + ix;                   // permute(0,4,3,2,1): [C][Z][Y][X]  :92
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    out[(size_t)o * total + vox] = mean[o];
    out[(size_t)(8 + o) * total + vox] = var[o];
  }
}

}  // namespace ufo
