// Kernel 1: fused correlation-frustum (cost-volume) build of one cascade stage.
//
// Replaces, per reference rotation n and source view i (code1/encoder_utils/fmt/TransMVSNet.py:76-100):
//   homo_warping_trans (fmt/module.py:329-367)  - homography warp of the source features to every
//                                                 depth hypothesis (bilinear, align_corners=True, zeros)
//   (warped * ref).mean(C)                      - channel-mean correlation
//   PixelwiseNet (TransMVSNet.py:23-41)         - stage 1: 1x1x1 conv MLP + sigmoid + max over depth
//   weighted accumulation over source views     - S = sum_i sim_i*vw_i / (1e-5 + sum_i vw_i)
// without materialising the [N,C,D,h,w] warped volume (2.2 GB per source view at stage 1, 1600x1216).
//
// Thread mapping: C/4 lanes own one reference pixel; each lane holds 4 channels (one float4) of the
// reference feature and of every bilinear tap, so a tap is one contiguous C*4-byte read per pixel.
// The channel dot product is a log2(C/4)-step xor-shuffle; lane l then owns depth planes
// k = l, l+C/4, ... for the pixel-wise net and the output writes.
#pragma once
#include "ufo_common.cuh"

namespace ufo {

struct PixelwiseDev {   // BatchNorm (eval) folded into the 1x1x1 convs
  float s0[16], t0[16];     // h0 = relu(s0*x + t0)
  float w1[8][16], t1[8];   // h1 = relu(w1.h0 + t1)
  float w2[8], b2;          // y  = sigmoid(w2.h1 + b2)
};

struct WarpMats {           // per (rotation, source view): rot (3x3) and trans (3) of src_proj * ref_proj^-1
  float r[9], t[3];
};

#ifndef UFO_COSTVOL_MINB
#define UFO_COSTVOL_MINB 3      // <= 80 registers: 3 blocks per SM (measured 1.59 / 2.85 / 1.63 ms vs 1.68 / 3.34 / 1.67 ms unbounded, NV=3 at 1600x1216)
#endif
template <int C, int D, bool kComputeVW>
__global__ void __launch_bounds__(256, UFO_COSTVOL_MINB) k_costvol(const float* __restrict__ ref_cl /*[N][h][w][C]*/,
                                                const float* const* __restrict__ src_cl /*[V-1] x [N][h][w][C]*/,
                                                const WarpMats* __restrict__ mats /*[N][V-1]*/,
                                                const float* __restrict__ hyp /*[N][D][h][w]*/,
                                                const float* __restrict__ vw_in /*[N][V-1][h][w]*/, PixelwiseDev pw,
                                                int N, int V, int h, int w, float* __restrict__ sim_out /*[N][D][h][w]*/,
                                                float* __restrict__ vw_out /*[N][V-1][h][w]*/) {
  constexpr int LPP = C / 4;                  // lanes per pixel
  constexpr int PPL = (D + LPP - 1) / LPP;    // planes per lane
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pix_lin = gid / LPP;
  const int l = (int)(gid % LPP);
  const long long hw = (long long)h * w;
  if (pix_lin >= (long long)N * hw) return;   // whole lane groups exit together
  const int n = (int)(pix_lin / hw);
  const int pix = (int)(pix_lin % hw);
  const int py = pix / w, px = pix % w;
  const unsigned gmask = (LPP == 32) ? 0xffffffffu : (((1u << LPP) - 1u) << ((threadIdx.x & 31) / LPP * LPP));

  const float4 rf = __ldg(reinterpret_cast<const float4*>(ref_cl + ((size_t)n * hw + pix) * C) + l);
  float dk[PPL], acc[PPL];
#pragma unroll
  for (int t = 0; t < PPL; ++t) {
    const int k = l + t * LPP;
    dk[t] = (k < D) ? __ldg(hyp + ((size_t)n * D + k) * hw + pix) : 0.f;
    acc[t] = 0.f;
  }
  float wsum = 1e-5f;                                           // TransMVSNet.py:74
  const float xf = (float)px, yf = (float)py;
  for (int i = 0; i < V - 1; ++i) {
    const WarpMats m = mats[n * (V - 1) + i];
    // rot . [x, y, 1] accumulated in k order with fused multiply-adds like the reference's sgemm (module.py:350); the
    // depth scaling and the translation below are separate roundings like torch's elementwise ops (module.py:351-353)
    const float rx = fmaf(m.r[2], 1.f, fmaf(m.r[1], yf, __fmul_rn(m.r[0], xf)));
    const float ry = fmaf(m.r[5], 1.f, fmaf(m.r[4], yf, __fmul_rn(m.r[3], xf)));
    const float rz = fmaf(m.r[8], 1.f, fmaf(m.r[7], yf, __fmul_rn(m.r[6], xf)));
    const float* src = src_cl[i] + (size_t)n * hw * C + 4 * l;
    float sim[PPL];
#pragma unroll
    for (int t = 0; t < PPL; ++t) sim[t] = 0.f;
    if constexpr (LPP >= 4) {
      const int lane0 = (threadIdx.x & 31) / LPP * LPP;           // first lane of this pixel's group
  #pragma unroll
      for (int t = 0; t < PPL; ++t) {
        // ---- lane l sets up the bilinear taps of ITS plane k = t*LPP + l once (the first version evaluated the homography of every plane
        //      in every lane: 186 warp instructions per lane and plane, issue-bound; stage 1 is 1.55 x faster this way); same arithmetic, same
        //      roundings, so the result is bit-identical.  Shipped to the group: base texel offset, validity bits, four weights.
        int s_off = 0;
        unsigned s_ok = 0u;
        float s_g00 = 0.f, s_g01 = 0.f, s_g10 = 0.f, s_g11 = 0.f;
        if (l + t * LPP < D) {
          const float d = dk[t];
          const float X = __fadd_rn(__fmul_rn(rx, d), m.t[0]), Y = __fadd_rn(__fmul_rn(ry, d), m.t[1]), Z = __fadd_rn(__fmul_rn(rz, d), m.t[2]);
          if (!(Z < 1e-6f)) {                                     // invalid -> grid -99 -> zero sample
            const float gx = __fsub_rn(__fdiv_rn(__fdiv_rn(X, Z), (float)(w - 1) / 2.f), 1.f);
            const float gy = __fsub_rn(__fdiv_rn(__fdiv_rn(Y, Z), (float)(h - 1) / 2.f), 1.f);
            const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1)), iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
            if ((ix > -1.f) && (ix < (float)w) && (iy > -1.f) && (iy < (float)h)) {
              const float fx = floorf(ix), fy = floorf(iy);
              const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
              const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
              s_off = y0 * w + x0;                                // may be negative / past the row: only valid taps are read
              s_ok = (x0 >= 0 && y0 >= 0 ? 1u : 0u) | (x1 < w && y0 >= 0 ? 2u : 0u) | (x0 >= 0 && y1 < h ? 4u : 0u) | (x1 < w && y1 < h ? 8u : 0u);
              s_g00 = wx0 * wy0; s_g01 = wx1 * wy0; s_g10 = wx0 * wy1; s_g11 = wx1 * wy1;
            }
          }
        }
        for (int kk = 0; kk < LPP; ++kk) {
          const int k = t * LPP + kk;
          if (k >= D) break;                                        // uniform across the lane group
          const int src_lane = lane0 + kk;
          const int off = __shfl_sync(gmask, s_off, src_lane);
          const unsigned ok = __shfl_sync(gmask, s_ok, src_lane);
          float part = 0.f;
          if (ok) {                                                 // uniform across the lane group
            const float g00 = __shfl_sync(gmask, s_g00, src_lane), g01 = __shfl_sync(gmask, s_g01, src_lane);
            const float g10 = __shfl_sync(gmask, s_g10, src_lane), g11 = __shfl_sync(gmask, s_g11, src_lane);
            const float* tp = src + (long long)off * C;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok & 1u) { const float4 v = __ldg(reinterpret_cast<const float4*>(tp)); a.x = v.x * g00; a.y = v.y * g00; a.z = v.z * g00; a.w = v.w * g00; }
            if (ok & 2u) { const float4 v = __ldg(reinterpret_cast<const float4*>(tp + C)); a.x = fmaf(v.x, g01, a.x); a.y = fmaf(v.y, g01, a.y); a.z = fmaf(v.z, g01, a.z); a.w = fmaf(v.w, g01, a.w); }
            if (ok & 4u) { const float4 v = __ldg(reinterpret_cast<const float4*>(tp + (long long)w * C)); a.x = fmaf(v.x, g10, a.x); a.y = fmaf(v.y, g10, a.y); a.z = fmaf(v.z, g10, a.z); a.w = fmaf(v.w, g10, a.w); }
            if (ok & 8u) { const float4 v = __ldg(reinterpret_cast<const float4*>(tp + (long long)(w + 1) * C)); a.x = fmaf(v.x, g11, a.x); a.y = fmaf(v.y, g11, a.y); a.z = fmaf(v.z, g11, a.z); a.w = fmaf(v.w, g11, a.w); }
            part = a.x * rf.x + a.y * rf.y + a.z * rf.z + a.w * rf.w;
          }
  #pragma unroll
          for (int o = LPP / 2; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o);
          const float s = part / (float)C;                          // .mean(1)
          if (kk == l) sim[t] = s;
        }
      }
    } else {
      // two lanes per pixel (stage 3, C = 8): sharing the set-up between two lanes costs more shuffles than it saves (measured +19 %)
  #pragma unroll
      for (int t = 0; t < PPL; ++t)
      for (int kk = 0; kk < LPP; ++kk) {
        const int k = t * LPP + kk;
        if (k >= D) break;                                        // uniform across the lane group
        // every lane of the group needs plane k's depth: broadcast from its owner lane kk
        const float d = __shfl_sync(gmask, dk[t], (threadIdx.x & 31) / LPP * LPP + kk);
        const float X = __fadd_rn(__fmul_rn(rx, d), m.t[0]), Y = __fadd_rn(__fmul_rn(ry, d), m.t[1]), Z = __fadd_rn(__fmul_rn(rz, d), m.t[2]);
        float part = 0.f;
        if (!(Z < 1e-6f)) {                                       // invalid -> grid -99 -> zero sample
          const float gx = __fsub_rn(__fdiv_rn(__fdiv_rn(X, Z), (float)(w - 1) / 2.f), 1.f);
          const float gy = __fsub_rn(__fdiv_rn(__fdiv_rn(Y, Z), (float)(h - 1) / 2.f), 1.f);
          const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1)), iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
          if ((ix > -1.f) && (ix < (float)w) && (iy > -1.f) && (iy < (float)h)) {
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
            const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x0 >= 0 && y0 >= 0) { const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)y0 * w + x0) * C)); const float g = wx0 * wy0; a.x = v.x * g; a.y = v.y * g; a.z = v.z * g; a.w = v.w * g; }
            if (x1 < w && y0 >= 0) { const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)y0 * w + x1) * C)); const float g = wx1 * wy0; a.x = fmaf(v.x, g, a.x); a.y = fmaf(v.y, g, a.y); a.z = fmaf(v.z, g, a.z); a.w = fmaf(v.w, g, a.w); }
            if (x0 >= 0 && y1 < h) { const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)y1 * w + x0) * C)); const float g = wx0 * wy1; a.x = fmaf(v.x, g, a.x); a.y = fmaf(v.y, g, a.y); a.z = fmaf(v.z, g, a.z); a.w = fmaf(v.w, g, a.w); }
            if (x1 < w && y1 < h) { const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)y1 * w + x1) * C)); const float g = wx1 * wy1; a.x = fmaf(v.x, g, a.x); a.y = fmaf(v.y, g, a.y); a.z = fmaf(v.z, g, a.z); a.w = fmaf(v.w, g, a.w); }
            part = a.x * rf.x + a.y * rf.y + a.z * rf.z + a.w * rf.w;
          }
        }
  #pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o);
        const float s = part / (float)C;                          // .mean(1)
        if (kk == l) sim[t] = s;
      }
    }
    float vw;
    if (kComputeVW) {
      float best = -INFINITY;
#pragma unroll
      for (int t = 0; t < PPL; ++t) {
        if (l + t * LPP < D) {
          float h0[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) h0[c] = fmaxf(fmaf(pw.s0[c], sim[t], pw.t0[c]), 0.f);
          float y = pw.b2;
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            float a1 = pw.t1[o];
#pragma unroll
            for (int c = 0; c < 16; ++c) a1 = fmaf(pw.w1[o][c], h0[c], a1);
            y = fmaf(pw.w2[o], fmaxf(a1, 0.f), y);
          }
          best = fmaxf(best, 1.f / (1.f + expf(-y)));
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(gmask, best, o));
      vw = best;
      if (l == 0) vw_out[((size_t)n * (V - 1) + i) * hw + pix] = vw;
    } else {
      vw = __ldg(vw_in + ((size_t)n * (V - 1) + i) * hw + pix);
    }
#pragma unroll
    for (int t = 0; t < PPL; ++t) acc[t] = fmaf(sim[t], vw, acc[t]);   // TransMVSNet.py:96
    wsum += vw;
  }
#pragma unroll
  for (int t = 0; t < PPL; ++t) {
    const int k = l + t * LPP;
    if (k < D) sim_out[((size_t)n * D + k) * hw + pix] = acc[t] / wsum;
  }
}

}  // namespace ufo
