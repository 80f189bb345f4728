// Kernel 4 and the two samplers: warp-per-ray scans.
//   FixedSampler.sample_ray        code1/encoder_utils/sampler.py:15-50
//   ImportanceSampler.sample_ray   code1/encoder_utils/sampler.py:74-108  (+ merge, code1/model.py:466-470)
//   VolumeRenderer.render          code1/encoder_utils/renderer.py:7-48   (NeuS alpha, iter_cos = -1.5)
#pragma once
#include "ufo_common.cuh"

namespace ufo {

// rayinfo[r] = {d.x, d.y, d.z, cam_d.z, near, far, 0, 0}   (code1/model.py:409-427)
static __global__ void __launch_bounds__(256) k_ray_setup(SceneDev sc, const long long* __restrict__ ray_idx,
                                                  long long ray_begin, int R, float* __restrict__ rayinfo) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const long long HW = (long long)sc.H * sc.W;
  long long pix = ray_idx ? ray_idx[r] : ray_begin + r;
  pix = pix < 0 ? 0 : (pix >= HW ? HW - 1 : pix);      // memory safety for a caller-supplied index list (the host wrapper rejects it)
  const float cz = __ldg(sc.cam_ray_d + 2 * HW + pix);
  float* o = rayinfo + (size_t)r * 8;
  o[0] = __ldg(sc.ray_d + pix);
  o[1] = __ldg(sc.ray_d + HW + pix);
  o[2] = __ldg(sc.ray_d + 2 * HW + pix);
  o[3] = cz;
  o[4] = sc.near0 / cz;
  o[5] = sc.far0 / cz;
  o[6] = 0.f;
  o[7] = 0.f;
}

// z[r][i] = lin_i*(far-near)+near + (u-0.5)*(1/63)*(far-near);  u is [64][u_stride], column r.
static __global__ void __launch_bounds__(256) k_coarse_z(const float* __restrict__ rayinfo, const float* __restrict__ u,
                                                 long long u_stride, int R, float* __restrict__ z) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * kNC) return;
  const int r = (int)(t / kNC), i = (int)(t % kNC);
  const float near = rayinfo[(size_t)r * 8 + 4], far = rayinfo[(size_t)r * 8 + 5];
  const float lin = (i == kNC - 1) ? 1.f : (float)((double)i * (1.0 / (double)(kNC - 1)));  // np.linspace(0,1,64)
  const float span = far - near;
  const float interval = (float)(1.0 / (double)(kNC - 1));
  const float base = __fadd_rn(__fmul_rn(lin, span), near);
  const float jit = __fmul_rn(__fmul_rn(u[(size_t)i * u_stride + r] - 0.5f, interval), span);
  z[t] = __fadd_rn(base, jit);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// One warp per ray; lane l owns the SN/32 consecutive samples l*K .. l*K+K-1.
template <int SN>
__global__ void __launch_bounds__(256) k_render(const float* __restrict__ z, const float* __restrict__ srdf,
                                               const float4* __restrict__ radiance, float inv_s, int R,
                                               float* __restrict__ weight_out, float* __restrict__ depth_out,
                                               float* __restrict__ rgb_out, float* __restrict__ depthz_out,
                                               const float* __restrict__ rayinfo, int rad_stride = SN,
                                               const uint8_t* __restrict__ perm = nullptr) {
  // radiance of sample i of ray r is radiance[r*rad_stride + (perm ? perm[r*SN+i] : i)]  (the tensor-core path keeps
  // per-point results in evaluation order: 64 coarse then 64 importance samples, see k_importance)
  constexpr int K = SN / 32;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float* zr = z + (size_t)r * SN;
  const float* sr = srdf + (size_t)r * SN;
  float alpha[K], zi[K];
  float prod = 1.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i = lane * K + k;
    const float zc = zr[i];
    const float zp = zr[max(i - 1, 0)], zn = zr[min(i + 1, SN - 1)];
    const float d_prev = (i == 0) ? (zr[1] - zr[0]) : (zc - zp);
    const float d_next = (i == SN - 1) ? (zr[SN - 1] - zr[SN - 2]) : (zn - zc);
    const float interval = (d_prev + d_next) / 2.f;                    // renderer.py:19-21
    const float s = sr[i];
    const float half = (-1.5f * interval) * 0.5f;                     // iter_cos * interval * 0.5
    const float prev_cdf = sigmoidf_((s - half) * inv_s);
    const float next_cdf = sigmoidf_((s + half) * inv_s);
    const float a = fminf(fmaxf(((prev_cdf - next_cdf) + 1e-5f) / (prev_cdf + 1e-5f), 0.f), 1.f);
    alpha[k] = a;
    zi[k] = zc;
    prod *= (1.f - a) + 1e-7f;
  }
  // exclusive prefix product across lanes
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  float acc_d = 0.f, acc_o = 0.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int i = lane * K + k;
    const float w = alpha[k] * T;
    T *= (1.f - alpha[k]) + 1e-7f;
    if (weight_out) weight_out[(size_t)r * SN + i] = w;
    const float4 c = radiance[(size_t)r * rad_stride + (perm ? (int)perm[(size_t)r * SN + i] : i)];
    acc_r = fmaf(c.x, w, acc_r);
    acc_g = fmaf(c.y, w, acc_g);
    acc_b = fmaf(c.z, w, acc_b);
    acc_d = fmaf(w, zi[k], acc_d);
    acc_o += w;
  }
  acc_r = warp_sum(acc_r);
  acc_g = warp_sum(acc_g);
  acc_b = warp_sum(acc_b);
  acc_d = warp_sum(acc_d);
  if (lane == 0) {
    if (depth_out) depth_out[r] = acc_d;
    if (depthz_out) depthz_out[r] = acc_d * rayinfo[(size_t)r * 8 + 3];   // model.py:821
    if (rgb_out) {
      rgb_out[(size_t)r * 3 + 0] = acc_r;
      rgb_out[(size_t)r * 3 + 1] = acc_g;
      rgb_out[(size_t)r * 3 + 2] = acc_b;
    }
  }
}

// Importance sampling + merge.  One warp per ray, 8 warps per block.
static __global__ void __launch_bounds__(256) k_importance(const float* __restrict__ weight, const float* __restrict__ zc,
                                                   const float* __restrict__ u, long long u_stride, int R,
                                                   float* __restrict__ z_fine_out, float* __restrict__ z_all,
                                                   uint8_t* __restrict__ perm = nullptr) {
  // perm [R][128] (optional): evaluation-order index of the sample at each sorted position
  // (0..63 = coarse sample i, 64..127 = importance sample i of the sorted fine list)
  __shared__ float s_cdf[8][kNC], s_zc[8][kNC], s_zf[8][kNC];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + wid;
  if (r >= R) return;
  float* cdf = s_cdf[wid];
  float* zcs = s_zc[wid];
  float* zf = s_zf[wid];
  const float w0 = weight[(size_t)r * kNC + lane], w1 = weight[(size_t)r * kNC + lane + 32];
  zcs[lane] = zc[(size_t)r * kNC + lane];
  zcs[lane + 32] = zc[(size_t)r * kNC + lane + 32];
  cdf[lane] = w0;
  cdf[lane + 32] = w1;
  const float total = warp_sum(w0 + w1);
  __syncwarp();
  if (lane == 0) {  // sequential cumsum, like torch.cumsum on a row (sampler.py:84)
    float run = 0.f;
    const float den = total + 1e-6f;
    for (int i = 0; i < kNC; ++i) {
      run += cdf[i];
      cdf[i] = run / den;
    }
  }
  __syncwarp();
  const float c_first = cdf[0], c_last = cdf[kNC - 1];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int s = lane + 32 * h;
    float su = u[(size_t)s * u_stride + r];                       // reference draws [64,RN] then transposes
    su = fminf(fmaxf(su, c_first), c_last);                       // sampler.py:88
    int lo = 0, hi = kNC;                                         // searchsorted (left): first cdf[i] >= su
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] < su) lo = mid + 1; else hi = mid;
    }
    int ri = lo;
    if (ri == 0) ri = 1;
    if (ri > kNC - 1) ri = kNC - 1;
    const float lc = cdf[ri - 1], rc = cdf[ri], zl = zcs[ri - 1], zr = zcs[ri];
    zf[s] = (su - lc) / (rc - lc + 1e-6f) * (zr - zl) + zl;       // sampler.py:101
  }
  __syncwarp();
  // bitonic sort of the 64 fine samples
  for (int k = 2; k <= kNC; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h, p = i ^ jj;
        if (p > i) {
          const float a = zf[i], b = zf[p];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { zf[i] = b; zf[p] = a; }
        }
      }
      __syncwarp();
    }
  // rank-merge of the two sorted lists (coarse first on ties)
  float* out = z_all + (size_t)r * kNS;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    {  // coarse element i: count fine < zc
      const float val = zcs[i];
      int lo = 0, hi = kNC;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (zf[mid] < val) lo = mid + 1; else hi = mid; }
      out[i + lo] = val;
      if (perm) perm[(size_t)r * kNS + i + lo] = (uint8_t)i;
    }
    {  // fine element i: count coarse <= zf
      const float val = zf[i];
      int lo = 0, hi = kNC;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (zcs[mid] <= val) lo = mid + 1; else hi = mid; }
      out[i + lo] = val;
      if (perm) perm[(size_t)r * kNS + i + lo] = (uint8_t)(kNC + i);
    }
    if (z_fine_out) z_fine_out[(size_t)r * kNC + i] = zf[i];
  }
}

// points_x_all of infer (model.py:466-470): x = o + z d for the merged, sorted samples
static __global__ void __launch_bounds__(256) k_points(SceneDev sc, const float* __restrict__ rayinfo, const float* __restrict__ z,
                                                       long long n, int SN, float* __restrict__ pts) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float* ri = rayinfo + (size_t)(p / SN) * 8;
  const float zz = z[p];
  pts[p * 3 + 0] = __fadd_rn(sc.ray_o[0], __fmul_rn(zz, ri[0]));
  pts[p * 3 + 1] = __fadd_rn(sc.ray_o[1], __fmul_rn(zz, ri[1]));
  pts[p * 3 + 2] = __fadd_rn(sc.ray_o[2], __fmul_rn(zz, ri[2]));
}

}  // namespace ufo
