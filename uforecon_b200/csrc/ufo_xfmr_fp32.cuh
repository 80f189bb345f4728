// Exact (fp32, CUDA-core) path of kernel 3: the view transformer, ray transformer and the SRDF /
// radiance heads, layer by layer through HBM workspaces.  This is the 1e-5 parity path and the
// on-device comparator for the tensor-core path; it is not the throughput path.
//
//   LoFTREncoderLayer   code1/attention/transformer.py:35-58
//   LinearAttention     code1/attention/linear_attention.py:20-47
//   token assembly      code1/ray_transformer.py:258-305
//   heads               code1/ray_transformer.py:307-320
#pragma once
#include "ufo_common.cuh"
#include "ufo_sampler_render.cuh"

namespace ufo {

// Y[M][N] = act(X[M][K] . W[N][K]^T).  W is staged transposed in shared memory once per CTA
// (persistent grid); each thread owns a 4-row x ceil(N/16)-column register tile.
template <int K, int N, bool kRelu>
__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                               float* __restrict__ Y, int ldy, long long M) {
  constexpr int BM = 64, TN = (N + 15) / 16;
  extern __shared__ float smem[];
  float* Wt = smem;            // [K][N]
  float* Xs = smem + K * N;    // [BM][K]
  for (int i = threadIdx.x; i < N * K; i += 256) {
    const int n = i / K, k = i % K;
    Wt[k * N + n] = __ldg(W + i);
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long n_tiles = (M + BM - 1) / BM;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = tile * BM;
    __syncthreads();
    for (int i = threadIdx.x; i < BM * K; i += 256) {
      const int rr = i / K, k = i % K;
      const long long m = m0 + rr;
      Xs[i] = (m < M) ? X[m * ldx + k] : 0.f;
    }
    __syncthreads();
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jn = 0; jn < TN; ++jn) acc[i][jn] = 0.f;
    const float* x0 = Xs + (ty * 4) * K;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float xv0 = x0[k], xv1 = x0[K + k], xv2 = x0[2 * K + k], xv3 = x0[3 * K + k];
      const float* wr = Wt + k * N + tx;
#pragma unroll
      for (int jn = 0; jn < TN; ++jn) {
        const float wv = (jn * 16 + tx < N) ? wr[jn * 16] : 0.f;
        acc[0][jn] = fmaf(xv0, wv, acc[0][jn]);
        acc[1][jn] = fmaf(xv1, wv, acc[1][jn]);
        acc[2][jn] = fmaf(xv2, wv, acc[2][jn]);
        acc[3][jn] = fmaf(xv3, wv, acc[3][jn]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ty * 4 + i;
      if (m < M) {
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) {
          const int n = jn * 16 + tx;
          if (n < N) Y[m * ldy + n] = kRelu ? fmaxf(acc[i][jn], 0.f) : acc[i][jn];
        }
      }
    }
  }
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.f : expf(x); }  // elu(x)+1

// Linear attention for one (sequence, head) per thread.  QKV is [rows][3*8*D] = [q | k | v];
// sequence s covers rows s*L .. s*L+L-1.  msg [rows][8*D].
template <int D>
__global__ void __launch_bounds__(128) k_linattn(const float* __restrict__ QKV, float* __restrict__ MSG,
                                                long long n_seq, int L) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_seq * kHeads) return;
  const long long s = t / kHeads;
  const int hd = (int)(t % kHeads);
  constexpr int d = D * kHeads;
  const float* base = QKV + s * L * (3 * d) + hd * D;
  float KV[D][D], Ksum[D];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    Ksum[a] = 0.f;
#pragma unroll
    for (int b = 0; b < D; ++b) KV[a][b] = 0.f;
  }
  const float invL = 1.f / (float)L;
  for (int l = 0; l < L; ++l) {
    const float* row = base + (size_t)l * 3 * d;
    float kf[D], vf[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      kf[a] = elu1(row[d + a]);
      vf[a] = row[2 * d + a] / (float)L;          // linear_attention.py:41
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      Ksum[a] += kf[a];
#pragma unroll
      for (int b = 0; b < D; ++b) KV[a][b] = fmaf(kf[a], vf[b], KV[a][b]);
    }
  }
  (void)invL;
  for (int l = 0; l < L; ++l) {
    const float* row = base + (size_t)l * 3 * d;
    float qf[D];
    float den = 0.f;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      qf[a] = elu1(row[a]);
      den = fmaf(qf[a], Ksum[a], den);
    }
    const float Z = 1.f / (den + 1e-6f);          // linear_attention.py:44
    float* o = MSG + (s * L + l) * d + hd * D;
#pragma unroll
    for (int b = 0; b < D; ++b) {
      float m = 0.f;
#pragma unroll
      for (int a = 0; a < D; ++a) m = fmaf(qf[a], KV[a][b], m);
      o[b] = m * Z * (float)L;                    // linear_attention.py:45
    }
  }
}

// Warp-per-row LayerNorm (eps 1e-5, affine).  out[m*ldo + c] = (res ? res[m*ldr+c] : 0) + LN(in[m*ldi + :])[c]
template <int DIM>
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ in, int ldi, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, const float* __restrict__ res, int ldr,
                                                  float* __restrict__ out, int ldo, long long M) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  constexpr int PER = (DIM + 31) / 32;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < DIM) ? in[m * ldi + c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)DIM;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    const float dlt = (c < DIM) ? v[i] - mean : 0.f;
    q = fmaf(dlt, dlt, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)DIM + 1e-5f);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    if (c < DIM) {
      const float y = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      out[m * ldo + c] = res ? res[m * ldr + c] + y : y;
    }
  }
}

struct Mlp3Dev {
  const float *w0, *b0, *w2, *b2, *w4, *b4;
};

// Generic 3-layer ReLU MLP evaluated by one thread; weights read through the read-only path
// (all threads of a warp read the same address).  The input is IN_A values read in place from `xa`
// (global memory, L1-resident across the H1 passes) followed by IN_B values held in registers.
template <int IN_A, int IN_B, int H1, int H2, int OUT>
__device__ __forceinline__ void mlp3_eval(const Mlp3Dev& w, const float* __restrict__ xa, const float* xb, float* out) {
  constexpr int IN = IN_A + IN_B;
  float h1[H1];
#pragma unroll
  for (int o = 0; o < H1; ++o) h1[o] = __ldg(w.b0 + o);
  for (int i = 0; i < IN_A; ++i) {
    const float xv = xa[i];
#pragma unroll
    for (int o = 0; o < H1; ++o) h1[o] = fmaf(xv, __ldg(w.w0 + o * IN + i), h1[o]);
  }
#pragma unroll
  for (int i = 0; i < IN_B; ++i)
#pragma unroll
    for (int o = 0; o < H1; ++o) h1[o] = fmaf(xb[i], __ldg(w.w0 + o * IN + IN_A + i), h1[o]);
  float h2[H2];
#pragma unroll
  for (int o = 0; o < H2; ++o) {
    float a = __ldg(w.b2 + o);
#pragma unroll
    for (int i = 0; i < H1; ++i) a = fmaf(fmaxf(h1[i], 0.f), __ldg(w.w2 + o * H1 + i), a);
    h2[o] = fmaxf(a, 0.f);
  }
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    float a = __ldg(w.b4 + o);
#pragma unroll
    for (int i = 0; i < H2; ++i) a = fmaf(h2[i], __ldg(w.w4 + o * H2 + i), a);
    out[o] = a;
  }
}

// pre_sim_mlp (8->32->32->16, ray_transformer.py:128-132,268) per point; writes the 16 values into
// columns 56..71 of every view row of XV and the learnable view token into row 0 (:286-288).
static __global__ void __launch_bounds__(128) k_presim(const float* __restrict__ sim8, Mlp3Dev w,
                                               const float* __restrict__ view_token, int L, long long P,
                                               float* __restrict__ XV) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float o[16];
  mlp3_eval<8, 0, 32, 32, 16>(w, sim8 + p * 8, nullptr, o);
  float* xv = XV + p * L * 160;
  for (int c = 0; c < kDView; ++c) xv[c] = __ldg(view_token + c);
  for (int n = 1; n < L; ++n)
#pragma unroll
    for (int i = 0; i < 16; ++i) xv[(size_t)n * 160 + 56 + i] = o[i];
}

// Ray-stage input: [token-0 output of the view stage | sinusoid(sample index)] (ray_transformer.py:301-303)
static __global__ void __launch_bounds__(256) k_ray_tokens(const float* __restrict__ VOUT, int L, int SN, long long P,
                                                   const float* __restrict__ pe_table /*[SN][8]*/,
                                                   float* __restrict__ XR /*[P][176]*/) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * kDRay) return;
  const long long p = t / kDRay;
  const int c = (int)(t % kDRay);
  XR[p * 176 + c] = (c < kDView) ? VOUT[p * L * kDView + c] : pe_table[(p % SN) * 8 + (c - kDView)];
}

// DensityMLP 88->32->16->1 (ray_transformer.py:147-150,307)
static __global__ void __launch_bounds__(128) k_density(const float* __restrict__ ROUT, Mlp3Dev w, long long P,
                                                float* __restrict__ srdf) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float o[1];
  mlp3_eval<kDRay, 0, 32, 16, 1>(w, ROUT + p * kDRay, nullptr, o);
  srdf[p] = o[0];
}

// Radiance blend: per view MLP 83->16->8->1 on [view feature | relative direction], masked softmax over
// views, weighted colour (ray_transformer.py:310-320).
static __global__ void __launch_bounds__(128) k_radiance(const float* __restrict__ VOUT, const float4* __restrict__ dirs,
                                                 const float4* __restrict__ rgbm, Mlp3Dev w, int NV, long long P,
                                                 float4* __restrict__ radiance) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int L = NV + 1;
  float om[kMaxV];
  float mx = -INFINITY;
  for (int n = 0; n < NV; ++n) {
    float o[1];
    const float4 d = dirs[p * NV + n];
    const float xb[3] = {d.x, d.y, d.z};
    mlp3_eval<kDView, 3, 16, 8, 1>(w, VOUT + (p * L + n + 1) * kDView, xb, o);
    om[n] = (rgbm[p * NV + n].w == 0.f) ? -1e9f : o[0];     // ray_transformer.py:316
    mx = fmaxf(mx, om[n]);
  }
  float den = 0.f;
  for (int n = 0; n < NV; ++n) { om[n] = expf(om[n] - mx); den += om[n]; }
  float r = 0.f, g = 0.f, b = 0.f;
  for (int n = 0; n < NV; ++n) {
    const float pw = om[n] / den;
    const float4 c = rgbm[p * NV + n];
    r = fmaf(c.x, pw, r); g = fmaf(c.y, pw, g); b = fmaf(c.z, pw, b);
  }
  radiance[p] = make_float4(r, g, b, 0.f);
}

}  // namespace ufo
