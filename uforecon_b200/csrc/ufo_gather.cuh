// Kernel 2 building blocks: projection of ray samples into the source views, the three bilinear /
// trilinear gather flavours of the reference and the pairwise-similarity prior.
//
// Reference semantics:
//   projection            code1/misc/camera.py:378-407
//   match-map gather      align_corners=True,  border   code1/model.py:251 (utils/gmflow_utils.py:83)
//   feature/rgb/depth     align_corners=False, zeros    code1/encoder_utils/grid_sample.py:18
//   frustum volumes       align_corners=True,  zeros    code1/model.py:370-371 (3-D)
//   cosine similarity     8 groups x 4 channels, mean over pairs   code1/model.py:273-285
//
// Thread mapping: 8 lanes own one sample point.  Lane j holds channels 4j..4j+3 of every 32-channel
// texel (one float4; the 8 lanes read one 128-byte line), similarity group j, frustum-volume
// channel j and depth-PE component j - so none of the per-point reductions needs a shuffle.
#pragma once
#include "ufo_common.cuh"

namespace ufo {

// torch grid_sample's unnormalisation (ATen GridSampler.h grid_sampler_unnormalize)
template <bool kAlign>
__device__ __forceinline__ float gs_unnorm(float c, int size) {
  return kAlign ? ((c + 1.f) / 2.f) * (float)(size - 1) : ((c + 1.f) * (float)size - 1.f) / 2.f;
}

struct BilTaps {
  int i00, i01, i10, i11;      // texel indices y*W+x (clamped to a valid address when out of range)
  float w00, w01, w10, w11;    // nw, ne, sw, se weights (0 for out-of-range taps)
};

template <bool kAlign, bool kBorder>
__device__ __forceinline__ BilTaps bil_setup(float u, float v, int H, int W) {
  float ix = gs_unnorm<kAlign>(u, W), iy = gs_unnorm<kAlign>(v, H);
  if (kBorder) {
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  }
  BilTaps t;
  const bool near_img = (ix > -1.f) && (ix < (float)W) && (iy > -1.f) && (iy < (float)H);  // false for NaN/inf
  if (!near_img) {
    t.i00 = t.i01 = t.i10 = t.i11 = 0;
    t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
    return t;
  }
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const bool vx0 = x0 >= 0, vx1 = x1 < W, vy0 = y0 >= 0, vy1 = y1 < H;
  const int cx0 = max(x0, 0), cx1 = min(x1, W - 1), cy0 = max(y0, 0), cy1 = min(y1, H - 1);
  t.i00 = cy0 * W + cx0; t.i01 = cy0 * W + cx1; t.i10 = cy1 * W + cx0; t.i11 = cy1 * W + cx1;
  t.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.f;
  t.w01 = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  t.w10 = (vx0 && vy1) ? wx0 * wy1 : 0.f;
  t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// texel * weight (+ accumulator) on the packed fp32x2 pipes of sm_100 (same IEEE results as four scalar mul / fma)
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
  const float2 s2 = make_float2(s, s);
  const float2 lo = __fmul2_rn(make_float2(a.x, a.y), s2), hi = __fmul2_rn(make_float2(a.z, a.w), s2);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {
  const float2 s2 = make_float2(s, s);
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), s2, make_float2(c.x, c.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), s2, make_float2(c.z, c.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// 4 channels (lane j) of a bilinear sample of a channel-last [H][W][32] map.
__device__ __forceinline__ float4 bil_fetch32(const float* __restrict__ map, const BilTaps& t, int j) {
  const float* b = map + 4 * j;
  float4 a = f4_scale(ldg4(b + (size_t)t.i00 * kFeatC), t.w00);
  a = f4_fma(ldg4(b + (size_t)t.i01 * kFeatC), t.w01, a);
  a = f4_fma(ldg4(b + (size_t)t.i10 * kFeatC), t.w10, a);
  a = f4_fma(ldg4(b + (size_t)t.i11 * kFeatC), t.w11, a);
  return a;
}

__device__ __forceinline__ float4 bil_fetch_rgbd(const float4* __restrict__ map, const BilTaps& t) {
  float4 a = f4_scale(__ldg(map + t.i00), t.w00);
  a = f4_fma(__ldg(map + t.i01), t.w01, a);
  a = f4_fma(__ldg(map + t.i10), t.w10, a);
  a = f4_fma(__ldg(map + t.i11), t.w11, a);
  return a;
}

// Trilinear sample (align_corners=True, zeros) of channel j of a [D][H][W][8] volume and of the
// matching single-channel weight volume [D][H][W].  (u,v,zn) in [-1,1] NDC.
__device__ __forceinline__ void tri_fetch(const float* __restrict__ vf, const float* __restrict__ vw, int D, int H,
                                          int W, float u, float v, float zn, int j, float& f_out, float& w_out) {
  // ATen's scalar 3-D kernel rounds the source index before it takes floor() and the tap weights; written with explicit
  // roundings so that the compiler cannot fuse the last multiply into "ix - floor(ix)" (it did: vol24 was off by 2e-5 at
  // 1600 pixels, where one ulp of ix is 1e-4 of a voxel)
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.f), 0.5f), (float)(W - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), (float)(H - 1));
  const float iz = __fmul_rn(__fmul_rn(__fadd_rn(zn, 1.f), 0.5f), (float)(D - 1));
  f_out = 0.f;
  w_out = 0.f;
  const bool near_vol = (ix > -1.f) && (ix < (float)W) && (iy > -1.f) && (iy < (float)H) && (iz > -1.f) && (iz < (float)D);
  if (!near_vol) return;
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const float wx[2] = {(fx + 1.f) - ix, ix - fx}, wy[2] = {(fy + 1.f) - iy, iy - fy}, wz[2] = {(fz + 1.f) - iz, iz - fz};
  float fa = 0.f, wa = 0.f;
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {  // ATen order: tnw, tne, tsw, tse, bnw, bne, bsw, bse
        const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
        const bool ok = (x >= 0) && (x < W) && (y >= 0) && (y < H) && (z >= 0) && (z < D);
        const float wgt = ok ? wx[dx] * wy[dy] * wz[dz] : 0.f;
        const size_t idx = ok ? ((size_t)z * H + y) * W + x : 0;
        // ATen's 3-D grid_sample (GridSampler.cpp, the non-vectorised CPU kernel) adds rounded products: no fused multiply-add
        fa = __fadd_rn(fa, __fmul_rn(__ldg(vf + idx * kVolC + j), wgt));
        wa = __fadd_rn(wa, __fmul_rn(__ldg(vw + idx), wgt));
      }
  f_out = fa;
  w_out = wa;
}

__device__ __forceinline__ void project_pt(const float* __restrict__ P, float x, float y, float z, float& u, float& v,
                                           float& qz) {
  // accumulated in k order with fused multiply-adds, like the sgemm behind the reference's torch.matmul(P, [x;1])
  // (camera.py:387-388): at 1600 pixels one ulp of q moves u by 1e-4 pixel, so the order of the roundings decides
  // whether the gathers agree with the reference to 1e-5 or to 1e-4
  const float q0 = fmaf(P[3], 1.f, fmaf(P[2], z, fmaf(P[1], y, __fmul_rn(P[0], x))));
  const float q1 = fmaf(P[7], 1.f, fmaf(P[6], z, fmaf(P[5], y, __fmul_rn(P[4], x))));
  const float q2 = fmaf(P[11], 1.f, fmaf(P[10], z, fmaf(P[9], y, __fmul_rn(P[8], x))));
  u = q0 / q2;
  v = q1 / q2;
  qz = q2;
}

// Cosine similarity of one 4-channel group (torch>=2.0 clamps each norm separately at eps=1e-8).
__device__ __forceinline__ float cos4(float4 a, float4 b) {
  const float dot = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  const float na = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
  const float nb = sqrtf(b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w);
  return dot / (fmaxf(na, 1e-8f) * fmaxf(nb, 1e-8f));
}

// Everything kernel 2 produces for one sample point, as held by lane j of its 8-lane group.
template <int NV>
struct PointGather {
  float4 feat[NV];   // channels 4j..4j+3 of the FPN feature sample, per view
  float vol[3];      // channel j of the blended frustum feature, per stage
  float sim;         // similarity group j
  float pe[NV];      // depth-difference PE component j, per view
  float4 rgbm[NV];   // (r,g,b,mask) per view  (same in all 8 lanes)
  float4 dir[NV];    // relative direction per view (same in all 8 lanes)
};

template <int NV>
__device__ __forceinline__ void gather_point(const SceneDev& sc, float x, float y, float z, int j, float freq_j,
                                             float phase_j, PointGather<NV>& g) {
  float u[NV], v[NV], qz[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) project_pt(sc.P[n], x, y, z, u[n], v[n], qz[n]);

  // ---- a7/a8/a9: per-view feature, colour, MVS-depth gathers, depth PE, mask, relative direction
  const float rx = x - sc.ref_o[0], ry = y - sc.ref_o[1], rz = z - sc.ref_o[2];
  const float rn = sqrtf(rx * rx + ry * ry + rz * rz);
  const size_t fstride = (size_t)sc.h * sc.w * kFeatC, istride = (size_t)sc.H * sc.W;
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    const BilTaps tf = bil_setup<false, false>(u[n], v[n], sc.h, sc.w);
    g.feat[n] = bil_fetch32(sc.feat_cl + n * fstride, tf, j);
    const BilTaps ti = bil_setup<false, false>(u[n], v[n], sc.H, sc.W);
    const float4 c = bil_fetch_rgbd(sc.rgbd_cl + n * istride, ti);
    const float zc = fmaf(sc.w2c_z[n][0], x, fmaf(sc.w2c_z[n][1], y, fmaf(sc.w2c_z[n][2], z, sc.w2c_z[n][3])));
    const float delta = c.w - zc;                                  // ray_transformer.py:245
    g.pe[n] = sinf(fmaf(delta, freq_j, phase_j));                  // ray_transformer.py:66
    const bool inb = (u[n] <= 1.f) && (u[n] >= -1.f) && (v[n] <= 1.f) && (v[n] >= -1.f);
    g.rgbm[n] = make_float4(c.x, c.y, c.z, (inb && qz[n] > 0.f) ? 1.f : 0.f);
    const float sx = x - sc.cam_o[n][0], sy = y - sc.cam_o[n][1], sz = z - sc.cam_o[n][2];
    const float sn = sqrtf(sx * sx + sy * sy + sz * sz);
    g.dir[n] = make_float4(rx / rn - sx / sn, ry / rn - sy / sn, rz / rn - sz / sn, 0.f);
  }

  // ---- a5: pairwise similarity prior.  pair (a,b): view a's slot (b-1) sampled at uv_a against view
  // b's slot a sampled at uv_b (the reference stores the same map in both slots).
  {
    float acc = 0.f;
    int npairs = 0;
#pragma unroll
    for (int a = 0; a < NV - 1; ++a)
#pragma unroll
      for (int b = a + 1; b < NV; ++b) {
        const BilTaps ta = bil_setup<true, true>(u[a], v[a], sc.h, sc.w);
        const BilTaps tb = bil_setup<true, true>(u[b], v[b], sc.h, sc.w);
        const float4 fa = bil_fetch32(sc.match_cl + (size_t)sc.match_slot[a][b] * fstride, ta, j);
        const float4 fb = bil_fetch32(sc.match_cl + (size_t)sc.match_slot[b][a] * fstride, tb, j);
        acc += cos4(fa, fb);
        ++npairs;
      }
    g.sim = acc / (float)npairs;
  }

  // ---- a6: frustum volumes, blended over views with the summed per-stage weights
  {
    float G[3] = {0.f, 0.f, 0.f}, Wsum = 0.f;
    const float inv_range = sc.far0 - sc.near0;
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      const float zn = ((qz[n] - sc.near0) / inv_range) * 2.f - 1.f;   // camera.py:399-400
      float f[3], wl = 0.f;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const size_t vox = (size_t)sc.vd[s] * sc.vh[s] * sc.vw[s];
        float ws;
        tri_fetch(sc.vol_feat_cl[s] + n * vox * kVolC, sc.vol_w[s] + n * vox, sc.vd[s], sc.vh[s], sc.vw[s], u[n], v[n],
                  zn, j, f[s], ws);
        wl = (s == 0) ? ws : wl + ws;                                   // model.py:375-378
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) G[s] = (n == 0) ? __fmul_rn(f[s], wl) : __fadd_rn(G[s], __fmul_rn(f[s], wl));  // model.py:381-386
      Wsum = (n == 0) ? wl : Wsum + wl;
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) g.vol[s] = G[s] / (Wsum + 1e-8f);       // model.py:388
  }
}

// Exact-path kernel 2: writes the gathered quantities to the fp32 workspaces of the layer-by-layer
// transformer.  XV is the [P][NV+1][160] concat buffer (row 0 = view token, columns 0..79 = token).
template <int NV>
__global__ void __launch_bounds__(256) k_gather(SceneDev sc, const float* __restrict__ rayinfo,
                                                const float* __restrict__ zbuf, int R, int SN,
                                                const float* __restrict__ freqs, const float* __restrict__ phases,
                                                float* __restrict__ XV, float* __restrict__ sim8,
                                                float4* __restrict__ rgbm, float4* __restrict__ dirs,
                                                float* __restrict__ pts_out) {
  constexpr int L = NV + 1;
  const int sub = threadIdx.x >> 3, j = threadIdx.x & 7;
  const long long p = (long long)blockIdx.x * 32 + sub;
  if (p >= (long long)R * SN) return;
  const int r = (int)(p / SN);
  const float* ri = rayinfo + (size_t)r * 8;
  const float zz = zbuf[p];
  const float x = __fadd_rn(sc.ray_o[0], __fmul_rn(zz, ri[0]));   // sampler.py:47
  const float y = __fadd_rn(sc.ray_o[1], __fmul_rn(zz, ri[1]));
  const float z = __fadd_rn(sc.ray_o[2], __fmul_rn(zz, ri[2]));
  PointGather<NV> g;
  gather_point<NV>(sc, x, y, z, j, __ldg(freqs + j), __ldg(phases + j), g);

  float* xv = XV + (size_t)p * L * 160;
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    float* row = xv + (size_t)(n + 1) * 160;
    *reinterpret_cast<float4*>(row + 4 * j) = g.feat[n];
    row[32 + j] = g.vol[0];
    row[40 + j] = g.vol[1];
    row[48 + j] = g.vol[2];
    row[72 + j] = g.pe[n];
    if (j == 0) {
      rgbm[(size_t)p * NV + n] = g.rgbm[n];
      dirs[(size_t)p * NV + n] = g.dir[n];
    }
  }
  sim8[(size_t)p * 8 + j] = g.sim;
  if (pts_out != nullptr && j < 3) pts_out[(size_t)p * 3 + j] = (j == 0) ? x : (j == 1 ? y : z);
}

}  // namespace ufo
