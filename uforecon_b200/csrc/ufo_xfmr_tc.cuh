// Tensor-core (tcgen05 / TMEM) path of kernel 3: view transformer, ray transformer, SRDF and radiance
// heads, plus the 16-bit-token flavour of kernel 2 that feeds it.
//
//   LoFTREncoderLayer   code1/attention/transformer.py:35-58
//   LinearAttention     code1/attention/linear_attention.py:20-47
//   token assembly      code1/ray_transformer.py:258-305
//   heads               code1/ray_transformer.py:307-320
//
// Design: one persistent CTA per SM, 512 threads.  A tile is 128 token rows = one
// tcgen05 M=128 accumulator: floor(128/(NV+1)) sample points x (NV+1) view tokens in the view stage, one
// ray x 128 samples (or two rays x 64) in the ray stage.  Every Linear layer is a tcgen05.mma.kind::f16
// (bf16 or fp16 operands, fp32 accumulate in TMEM) with the weight matrix as the K-major B operand in
// shared memory and the activations staged as the K-major A operand by the epilogue of the previous
// layer; TMEM lane r == token row r, so an epilogue thread owns one row (warp w -> lanes 32*(w%4)...,
// column group w/4).  LayerNorm, elu+1, the linear-attention normaliser, softmax over views and the
// SRDF tail run in fp32 on the CUDA cores between the MMAs.
#pragma once
#include "ufo_common.cuh"
#include "ufo_gather.cuh"
#include "ufo_umma.cuh"
#include "ufo_xfmr_fp32.cuh"
#include "ufo_tc_params.cuh"

namespace ufo {
namespace tc {


// ---- small device helpers ----------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if (BF16) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  const __half2 h = *reinterpret_cast<const __half2*>(&u);
  return __half22float2(h);
}

template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = umma::pack2<BF16>(v[0], v[1]);
  u.y = umma::pack2<BF16>(v[2], v[3]);
  u.z = umma::pack2<BF16>(v[4], v[5]);
  u.w = umma::pack2<BF16>(v[6], v[7]);
  return u;
}

// 16-byte piece (row, chunk) of a 128-row K-major operand tile
__device__ __forceinline__ uint8_t* tile_ptr(uint8_t* tile, int row, int chunk) { return tile + chunk * kChunk + row * 16; }

template <bool BF16>
__device__ __forceinline__ void st_chunk(uint8_t* tile, int row, int chunk, const float* v) {
  *reinterpret_cast<uint4*>(tile_ptr(tile, row, chunk)) = pack8<BF16>(v);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float elu1_fast(float x) { return x > 0.f ? x + 1.f : __expf(x); }

// split a float into a 16-bit head and the 16-bit remainder (hi + lo ~ x to ~2^-17 / 2^-22 relative)
template <bool BF16>
__device__ __forceinline__ void split_hi_lo(float x, float& hi, float& lo) {
  if (BF16) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
  } else {
    hi = __half2float(__float2half_rn(x));
  }
  lo = x - hi;
}

// one-thread bulk copy global -> shared, completing on an mbarrier (weights are stored in global memory as
// the exact shared-memory operand image, so no tensor map is needed)
__device__ __forceinline__ void bulk_load(uint8_t* dst_smem, const uint8_t* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
  for (uint32_t off = 0; off < bytes; off += 16384u) {
    const uint32_t n = (bytes - off) < 16384u ? (bytes - off) : 16384u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     umma::smem_u32(dst_smem + off)),
                 "l"(src + off), "r"(n), "r"(umma::smem_u32(bar))
                 : "memory");
  }
}

// K-major GEMM issue with an explicit B row offset / row count (sub-block of a weight tile).
// a_base: A tile (128 rows); b_base: B tile with b_rows_total rows; uses rows [b_row0, b_row0+N).
__device__ __forceinline__ void issue_gemm_sub(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t b_rows_total,
                                               uint32_t b_row0, uint32_t k_chunks, uint32_t idesc, uint32_t acc_first) {
  const uint32_t b_lbo = b_rows_total * 16u;
  for (uint32_t c = 0; c < k_chunks; c += 2) {
    const uint64_t ad = umma::make_smem_desc(a_base + c * kChunk, kChunk, 128u);
    const uint64_t bd = umma::make_smem_desc(b_base + c * b_lbo + b_row0 * 16u, b_lbo, 128u);
    umma::mma_f16(tmem_d, ad, bd, idesc, (c > 0) ? 1u : acc_first);
  }
}

}  // namespace tc

// =================================================================================================
// kernel 2, 16-bit token flavour: same gathers as k_gather, tokens written as the 16-bit operand rows the
// view-stage kernel loads, pre_sim_mlp (ray_transformer.py:128-132,268) fused in.
// tok [P][NV][80] 16-bit: [feat 32 | vol 24 | sim 16 | depth-PE 8]
// =================================================================================================
// Per-point buffers of the tensor-core path are laid out [ray][128 slots]: slots 0..63 hold the coarse samples,
// 64..127 the importance samples, both in evaluation order, so that the fine pass of infer only evaluates the 64
// NEW points of a ray (the view stage is point-wise: re-evaluating the coarse points, as the reference does, gives
// identical values).  A pass over R*64 points with half = 0 | 1 addresses slot(p).
__device__ __forceinline__ long long tc_slot(long long p, int half) { return (p >> 6) * kNS + half * kNC + (p & 63); }

template <int NV, bool BF16>
__global__ void __launch_bounds__(256) k_gather_tc(SceneDev sc, const float* __restrict__ rayinfo,
                                                   const float* __restrict__ zbuf, int R, int half,
                                                   const float* __restrict__ freqs, const float* __restrict__ phases,
                                                   Mlp3Dev presim, uint16_t* __restrict__ tok, float4* __restrict__ rgbm,
                                                   float4* __restrict__ dirs, float* __restrict__ sim8_out) {
  constexpr int SN = kNC;
  __shared__ float s_sim[256][9];
  __shared__ float s_w[8 * 32 + 32 + 32 * 32 + 32 + 32 * 16 + 16];
  {  // pre_sim_mlp weights -> shared memory
    const int n0 = 8 * 32, n1 = 32, n2 = 32 * 32, n3 = 32, n4 = 32 * 16, n5 = 16;
    for (int i = threadIdx.x; i < n0; i += 256) s_w[i] = __ldg(presim.w0 + i);
    for (int i = threadIdx.x; i < n1; i += 256) s_w[n0 + i] = __ldg(presim.b0 + i);
    for (int i = threadIdx.x; i < n2; i += 256) s_w[n0 + n1 + i] = __ldg(presim.w2 + i);
    for (int i = threadIdx.x; i < n3; i += 256) s_w[n0 + n1 + n2 + i] = __ldg(presim.b2 + i);
    for (int i = threadIdx.x; i < n4; i += 256) s_w[n0 + n1 + n2 + n3 + i] = __ldg(presim.w4 + i);
    for (int i = threadIdx.x; i < n5; i += 256) s_w[n0 + n1 + n2 + n3 + n4 + i] = __ldg(presim.b4 + i);
  }
  const int sub = threadIdx.x >> 3, j = threadIdx.x & 7;
  const long long P = (long long)R * SN;
  const long long p0 = (long long)blockIdx.x * 256;
  const float fj = __ldg(freqs + j), pj = __ldg(phases + j);
  for (int round = 0; round < 8; ++round) {
    const long long p = p0 + round * 32 + sub;
    if (p >= P) break;
    const int r = (int)(p / SN);
    const float* ri = rayinfo + (size_t)r * 8;
    const float zz = zbuf[p];
    const float x = __fadd_rn(sc.ray_o[0], __fmul_rn(zz, ri[0]));
    const float y = __fadd_rn(sc.ray_o[1], __fmul_rn(zz, ri[1]));
    const float z = __fadd_rn(sc.ray_o[2], __fmul_rn(zz, ri[2]));
    PointGather<NV> g;
    gather_point<NV>(sc, x, y, z, j, fj, pj, g);
    const size_t sl = (size_t)tc_slot(p, half);
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      uint16_t* row = tok + (sl * NV + n) * kDView;
      uint2 f;
      f.x = umma::pack2<BF16>(g.feat[n].x, g.feat[n].y);
      f.y = umma::pack2<BF16>(g.feat[n].z, g.feat[n].w);
      *reinterpret_cast<uint2*>(row + 4 * j) = f;
      row[32 + j] = (uint16_t)(umma::pack2<BF16>(g.vol[0], 0.f) & 0xffffu);
      row[40 + j] = (uint16_t)(umma::pack2<BF16>(g.vol[1], 0.f) & 0xffffu);
      row[48 + j] = (uint16_t)(umma::pack2<BF16>(g.vol[2], 0.f) & 0xffffu);
      row[72 + j] = (uint16_t)(umma::pack2<BF16>(g.pe[n], 0.f) & 0xffffu);
      if (j == 0) {
        rgbm[sl * NV + n] = g.rgbm[n];
        dirs[sl * NV + n] = g.dir[n];
      }
    }
    s_sim[round * 32 + sub][j] = g.sim;
    if (sim8_out != nullptr) sim8_out[sl * 8 + j] = g.sim;
  }
  __syncthreads();
  const long long p = p0 + threadIdx.x;
  if (p >= P) return;
  // pre_sim_mlp 8 -> 32 -> 32 -> 16, one thread per point, weights broadcast from shared memory
  const float* w0 = s_w;
  const float* b0 = w0 + 8 * 32;
  const float* w2 = b0 + 32;
  const float* b2 = w2 + 32 * 32;
  const float* w4 = b2 + 32;
  const float* b4 = w4 + 32 * 16;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = s_sim[threadIdx.x][i];
  float h1[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    float a = b0[o];
#pragma unroll
    for (int i = 0; i < 8; ++i) a = fmaf(s[i], w0[o * 8 + i], a);
    h1[o] = fmaxf(a, 0.f);
  }
  float h2[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    float a = b2[o];
#pragma unroll
    for (int i = 0; i < 32; ++i) a = fmaf(h1[i], w2[o * 32 + i], a);
    h2[o] = fmaxf(a, 0.f);
  }
  float o16[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    float a = b4[o];
#pragma unroll
    for (int i = 0; i < 32; ++i) a = fmaf(h2[i], w4[o * 32 + i], a);
    o16[o] = a;
  }
  const uint4 lo = tc::pack8<BF16>(o16), hi = tc::pack8<BF16>(o16 + 8);
  const size_t sl = (size_t)tc_slot(p, half);
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    uint16_t* row = tok + (sl * NV + n) * kDView;
    *reinterpret_cast<uint4*>(row + 56) = lo;
    *reinterpret_cast<uint4*>(row + 64) = hi;
  }
}

// =================================================================================================
// view stage: density_view_transformer (d = 80) + radiance-weight head, tokens = views of one point
// =================================================================================================
template <int NV, bool BF16>
__global__ void __launch_bounds__(tc::kThreads, 1)
k_view_tc(const uint8_t* __restrict__ wimg, const __grid_constant__ ViewParams prm, const uint16_t* __restrict__ tok,
          const float4* __restrict__ rgbm, const float4* __restrict__ dirs, long long P, int half,
          float* __restrict__ vout0, float4* __restrict__ radiance) {
  using namespace tc;
  constexpr int L = NV + 1, PPT = 128 / L, ROWS = PPT * L;
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + V_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + V_BAR + 16);
  float2* red = reinterpret_cast<float2*>(smem + V_RED);
  float* omg = reinterpret_cast<float*>(smem + V_OMG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2, r = q * 32 + lane;
  const int pl = r / L, l = r - pl * L;
  const bool row_ok = r < ROWS;

  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_barrier_init();
  }
  for (uint32_t i = tid; i < V_WEND / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
  // token rows: row l == 0 of every point is the learnable view token (constant), pad rows are zero
  for (int i = tid; i < 128 * 20; i += kThreads) {
    const int rr = i & 127, c = i >> 7;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (c < 10 && rr < ROWS && (rr % L) == 0) ? prm.vtok[c * 8 + k] : 0.f;
    st_chunk<BF16>(smem + V_X, rr, c, v);
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t sm_base = umma::smem_u32(smem);
  uint32_t ph = 0;
  const long long n_tiles = (P + PPT - 1) / PPT;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long pbase = tile * PPT;
    // ---- P0: token rows of this tile -> X (A operand, K = 80)
    for (int i = tid; i < 128 * 10; i += kThreads) {
      const int rr = i & 127, c = i >> 7;
      const int pr = rr / L, ll = rr - pr * L;
      if (rr < ROWS && ll > 0) {
        const long long p = pbase + pr;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (p < P) v = __ldg(reinterpret_cast<const uint4*>(tok + ((size_t)tc_slot(p, half) * NV + (ll - 1)) * kDView) + c);
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_X, rr, c)) = v;
      }
    }
    const long long my_p = pbase + pl;
    const size_t my_slot = (size_t)tc_slot(my_p, half);
    const bool view_row = row_ok && l > 0 && my_p < P;
    float4 my_dir = make_float4(0.f, 0.f, 0.f, 0.f);
    float my_mask = 0.f;
    if (g == 0 && view_row) {
      my_dir = __ldg(dirs + my_slot * NV + (l - 1));
      my_mask = __ldg(rgbm + my_slot * NV + (l - 1)).w;
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P1: q|k|v = X . Wqkv^T                                  (transformer.py:47)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + 0, sm_base + V_X, sm_base + V_WQKV, 240, 0, 10, umma::make_idesc(128, 240, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P2: elu+1 on q, k; stage K', V' (16-bit) for the per-point attention   (linear_attention.py:36-41)
    float qv[20];
    {
      float kv[40];
      umma::tmem_ld16(tlane + 20 * g, qv);
      tmem_ld4(tlane + 20 * g + 16, qv + 16);
      umma::tmem_ld16(tlane + 80 + 20 * g, kv);
      tmem_ld4(tlane + 80 + 20 * g + 16, kv + 16);
      umma::tmem_ld16(tlane + 160 + 20 * g, kv + 20);
      tmem_ld4(tlane + 160 + 20 * g + 16, kv + 36);
      umma::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 20; ++i) {
        qv[i] = elu1_fast(qv[i]);
        kv[i] = elu1_fast(kv[i]);
      }
      uint4* dst = reinterpret_cast<uint4*>(smem + V_KV + (size_t)(g * 128 + r) * 80);
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = pack8<BF16>(kv + 8 * i);
    }
    __syncthreads();
    // ---- P3: msg_l = sum_s (Q_l.K_s) V_s / (sum_s Q_l.K_s + 1e-6)  per head   (== Q (K^T V) Z, linear_attention.py:43-45)
    {
      float msg[20];
      if (row_ok) {
        float den0 = 0.f, den1 = 0.f;
#pragma unroll
        for (int i = 0; i < 20; ++i) msg[i] = 0.f;
#pragma unroll
        for (int s = 0; s < L; ++s) {
          const uint4* src = reinterpret_cast<const uint4*>(smem + V_KV + (size_t)(g * 128 + pl * L + s) * 80);
          float kk[40];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const uint4 u = src[i];
            const float2 a = unpack2<BF16>(u.x), b = unpack2<BF16>(u.y), c = unpack2<BF16>(u.z), d = unpack2<BF16>(u.w);
            kk[8 * i + 0] = a.x; kk[8 * i + 1] = a.y; kk[8 * i + 2] = b.x; kk[8 * i + 3] = b.y;
            kk[8 * i + 4] = c.x; kk[8 * i + 5] = c.y; kk[8 * i + 6] = d.x; kk[8 * i + 7] = d.y;
          }
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int a = 0; a < 10; ++a) {
            s0 = fmaf(qv[a], kk[a], s0);
            s1 = fmaf(qv[10 + a], kk[10 + a], s1);
          }
          den0 += s0;
          den1 += s1;
#pragma unroll
          for (int b = 0; b < 10; ++b) {
            msg[b] = fmaf(s0, kk[20 + b], msg[b]);
            msg[10 + b] = fmaf(s1, kk[30 + b], msg[10 + b]);
          }
        }
        const float z0 = 1.f / (den0 + 1e-6f), z1 = 1.f / (den1 + 1e-6f);
#pragma unroll
        for (int b = 0; b < 10; ++b) {
          msg[b] *= z0;
          msg[10 + b] *= z1;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 20; ++i) msg[i] = 0.f;
      }
      // columns 20g .. 20g+19 of the message tile: two full chunks and one half chunk
      const int c0 = (20 * g) >> 3;
      if ((g & 1) == 0) {  // 20g % 8 == 0: chunks c0, c0+1 full, first half of c0+2
        st_chunk<BF16>(smem + V_M, r, c0, msg);
        st_chunk<BF16>(smem + V_M, r, c0 + 1, msg + 8);
        uint2 h;
        h.x = umma::pack2<BF16>(msg[16], msg[17]);
        h.y = umma::pack2<BF16>(msg[18], msg[19]);
        *reinterpret_cast<uint2*>(tile_ptr(smem + V_M, r, c0 + 2)) = h;
      } else {             // 20g % 8 == 4: second half of c0, chunks c0+1, c0+2 full
        uint2 h;
        h.x = umma::pack2<BF16>(msg[0], msg[1]);
        h.y = umma::pack2<BF16>(msg[2], msg[3]);
        *reinterpret_cast<uint2*>(tile_ptr(smem + V_M, r, c0) + 8) = h;
        st_chunk<BF16>(smem + V_M, r, c0 + 1, msg + 4);
        st_chunk<BF16>(smem + V_M, r, c0 + 2, msg + 12);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P4: merge                                                (transformer.py:55)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + 256, sm_base + V_M, sm_base + V_WMRG, 80, 0, 10, umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P5: LayerNorm 1 -> second half of the concat operand     (transformer.py:56)
    {
      float v[24];
      int nc = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 10) {
          umma::tmem_ld8(tlane + 256 + 8 * c, v + 8 * i);
          nc = i + 1;
        }
      }
      umma::tmem_ld_wait();
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 24; ++i)
        if (i < 8 * nc) {
          s += v[i];
          ss = fmaf(v[i], v[i], ss);
        }
      red[g * 128 + r] = make_float2(s, ss);
      __syncthreads();
      const float2 a0 = red[r], a1 = red[128 + r], a2 = red[256 + r], a3 = red[384 + r];
      const float mean = (a0.x + a1.x + a2.x + a3.x) * (1.f / 80.f);
      const float var = fmaxf((a0.y + a1.y + a2.y + a3.y) * (1.f / 80.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 10) {
          float o[8];
  #pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * prm.n1w[8 * c + k] + prm.n1b[8 * c + k];
          st_chunk<BF16>(smem + V_M, r, c, o);
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P6: mlp.0 on [x | msg]                                   (transformer.py:57)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + 0, sm_base + V_X, sm_base + V_WML0, 160, 0, 20, umma::make_idesc(128, 160, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P7: ReLU -> H1 operand (aliases the K'/V' staging)
    {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = g + 4 * i;
        if (c < 20) {
          float v[8];
          umma::tmem_ld8(tlane + 8 * c, v);
          umma::tmem_ld_wait();
  #pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
          st_chunk<BF16>(smem + V_KV, r, c, v);
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P8: mlp.2
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + 256, sm_base + V_KV, sm_base + V_WML2, 80, 0, 20, umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P9: LayerNorm 2; token 0: out = view_token + LN2 -> vout0 (fp32); view rows: LN2 -> operand for the
    //      radiance head (x + LN2 is applied inside the head's GEMM: W0x.x + W0x.LN2)
    {
      float v[24];
      int nc = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 10) {
          umma::tmem_ld8(tlane + 256 + 8 * c, v + 8 * i);
          nc = i + 1;
        }
      }
      umma::tmem_ld_wait();
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 24; ++i)
        if (i < 8 * nc) {
          s += v[i];
          ss = fmaf(v[i], v[i], ss);
        }
      red[g * 128 + r] = make_float2(s, ss);
      __syncthreads();
      const float2 a0 = red[r], a1 = red[128 + r], a2 = red[256 + r], a3 = red[384 + r];
      const float mean = (a0.x + a1.x + a2.x + a3.x) * (1.f / 80.f);
      const float var = fmaxf((a0.y + a1.y + a2.y + a3.y) * (1.f / 80.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 10) {
          float o[8];
  #pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * prm.n2w[8 * c + k] + prm.n2b[8 * c + k];
          st_chunk<BF16>(smem + V_M, r, c, o);
          if (row_ok && l == 0 && my_p < P) {
            float4* dst = reinterpret_cast<float4*>(vout0 + my_slot * kDView + 8 * c);
            dst[0] = make_float4(prm.vtok[8 * c] + o[0], prm.vtok[8 * c + 1] + o[1], prm.vtok[8 * c + 2] + o[2], prm.vtok[8 * c + 3] + o[3]);
            dst[1] = make_float4(prm.vtok[8 * c + 4] + o[4], prm.vtok[8 * c + 5] + o[5], prm.vtok[8 * c + 6] + o[6], prm.vtok[8 * c + 7] + o[7]);
          }
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P10: radiance head layer 0 on [x | LN2]                  (ray_transformer.py:159-163,313)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + 384, sm_base + V_X, sm_base + V_WRAD, 16, 0, 20, umma::make_idesc(128, 16, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P11: head tail, masked softmax over views, colour blend  (ray_transformer.py:313-320)
    if (g == 0) {
      float h[16];
      umma::tmem_ld16(tlane + 384, h);
      umma::tmem_ld_wait();
      float w = -1e9f;
      if (view_row && my_mask != 0.f) {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          h[o] = fmaxf(h[o] + prm.rb0[o] + prm.rw0d[o][0] * my_dir.x + prm.rw0d[o][1] * my_dir.y + prm.rw0d[o][2] * my_dir.z, 0.f);
        float acc = prm.rb4;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          float a = prm.rb2[o];
#pragma unroll
          for (int i = 0; i < 16; ++i) a = fmaf(h[i], prm.rw2[o][i], a);
          acc = fmaf(fmaxf(a, 0.f), prm.rw4[o], acc);
        }
        w = acc;
      }
      omg[r] = w;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (tid < PPT && pbase + tid < P) {
      const size_t p = (size_t)tc_slot(pbase + tid, half);
      float om[NV];
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        om[n] = omg[tid * L + 1 + n];
        mx = fmaxf(mx, om[n]);
      }
      float den = 0.f;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        om[n] = __expf(om[n] - mx);
        den += om[n];
      }
      float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float4 c = __ldg(rgbm + p * NV + n);
        const float pw = om[n] / den;
        cr = fmaf(c.x, pw, cr);
        cg = fmaf(c.y, pw, cg);
        cb = fmaf(c.z, pw, cb);
      }
      radiance[p] = make_float4(cr, cg, cb, 0.f);
    }
    // the next tile's P0 writes X only after this tile's last MMA (P10) has completed: guaranteed by the wait above
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}


// =================================================================================================
// ray stage: density_ray_transformer (d = 88) + DensityMLP, tokens = samples of one ray
// =================================================================================================
// SN = 128: one ray per tile; SN = 64: two rays per tile.  vout0 [P][80] fp32 (token-0 output of the view
// stage), pe_table [128][8], srdf [P] out, ray_out [P][88] optional tap.
template <int SN, bool BF16>
__global__ void __launch_bounds__(tc::kThreads, 1)
k_ray_tc(const uint8_t* __restrict__ wimg, const __grid_constant__ RayParams prm, const float* __restrict__ vout0,
         const float* __restrict__ pe_table, const uint8_t* __restrict__ perm, long long P, float* __restrict__ srdf,
         float* __restrict__ ray_out) {
  using namespace tc;
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  constexpr int NSEQ = 128 / SN;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + R_BAR);        // MMA completion
  uint64_t* barA = reinterpret_cast<uint64_t*>(smem + R_BAR + 8);   // slot A filled
  uint64_t* barB = reinterpret_cast<uint64_t*>(smem + R_BAR + 16);  // slot B filled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + R_BAR + 24);
  float2* red = reinterpret_cast<float2*>(smem + R_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2, r = q * 32 + lane;
  const long long n_tiles = (P + 127) / 128;

  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(barA, 1);
    umma::mbar_init(barB, 1);
    umma::fence_barrier_init();
  }
  // zero all activation regions once (stale-but-finite invariant of the padded K chunks), constant columns
  for (uint32_t i = tid; i < R_SLOTA / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (uint32_t i = tid; i < (2 * 32 * 96 * 2) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem + R_WDEN)[i] = __ldg(reinterpret_cast<const uint4*>(wimg + RW_DEN) + i);
  __syncthreads();
  if (g == 0) {
    float pe[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pe[k] = __ldg(pe_table + (r % SN) * 8 + k);      // ray_transformer.py:301-303
    st_chunk<BF16>(smem + R_X, r, 10, pe);
    float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};                    // ones columns: D rows 88..95 = sum_s K'_s
    st_chunk<BF16>(smem + R_V, r, 11, one);
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t sm_base = umma::smem_u32(smem);
  uint32_t ph = 0, phA = 0, phB = 0;
  if (tid == 0 && (long long)blockIdx.x < n_tiles) {
    bulk_load(smem + R_SLOTA, wimg + RW_QKV, 272 * 96 * 2, barA);
    bulk_load(smem + R_SLOTB, wimg + RW_MRG, 96 * 96 * 2, barB);
  }
  // TMEM columns
  constexpr uint32_t D_QKV = 0, D_KV = 272, D_MSG = 0, D_MRG = 96, D_ML0 = 192, D_ML2 = 0, D_DEN = 96;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long prow = tile * 128 + r;           // this thread's token: ray prow/SN, sorted sample prow%SN
    const bool row_ok = prow < P;
    // its view-stage result lives at slot(ray, evaluation index): coarse pass = sample index, fine pass = perm
    size_t in_row = 0;
    if (row_ok) in_row = (SN == kNC) ? (size_t)tc_slot(prow, 0) : (size_t)(tile * 128 + (perm ? (int)perm[prow] : r));
    // ---- R0: x = token-0 output of the view stage -> 16-bit A operand (columns 80..87 = order PE, constant)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = g + 4 * i;
      if (c < 10) {
        float v[8];
        if (row_ok) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(vout0 + in_row * kDView + 8 * c));
          const float4 b = __ldg(reinterpret_cast<const float4*>(vout0 + in_row * kDView + 8 * c + 4));
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
  #pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = 0.f;
        }
        st_chunk<BF16>(smem + R_X, r, c, v);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R1: q|k|v = x . Wqkv^T   (K = 96: columns 88..95 hit zero weight columns)
    if (tid == 0) {
      umma::mbar_wait(barA, phA);
      phA ^= 1;
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_QKV, sm_base + R_X, sm_base + R_SLOTA, 272, 0, 12, umma::make_idesc(128, 176, FMT, false, false), 0);
      issue_gemm_sub(tmem + D_QKV + 176, sm_base + R_X, sm_base + R_SLOTA, 272, 176, 12, umma::make_idesc(128, 96, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    if (tid == 0) bulk_load(smem + R_SLOTA, wimg + RW_ML0, 176 * 176 * 2, barA);   // slot A is free again
    // ---- R2: Q' = elu(q)+1, K' = elu(k)+1, V' = v  -> 16-bit operand tiles    (linear_attention.py:36-41)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = g + 4 * i;
      if (c < 11) {
        float a[8], b[8], d[8];
        umma::tmem_ld8(tlane + D_QKV + 8 * c, a);
        umma::tmem_ld8(tlane + D_QKV + 88 + 8 * c, b);
        umma::tmem_ld8(tlane + D_QKV + 176 + 8 * c, d);
        umma::tmem_ld_wait();
  #pragma unroll
        for (int k = 0; k < 8; ++k) {
          a[k] = elu1_fast(a[k]);
          b[k] = elu1_fast(b[k]);
        }
        st_chunk<BF16>(smem + R_Q, r, c, a);
        st_chunk<BF16>(smem + R_K, r, c, b);
        st_chunk<BF16>(smem + R_V, r, c, d);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R3: per sequence  D[b][a] = sum_s V'[s][b] K'[s][a]   (rows 88..95 = sum_s K'[s][a]);  both operands MN-major
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t idesc = umma::make_idesc(128, 96, FMT, true, true);
#pragma unroll
      for (int sq = 0; sq < NSEQ; ++sq) {
        for (int ks = 0; ks < SN / 16; ++ks) {
          const uint32_t off = (uint32_t)(sq * (SN / 16) + ks) * 256u;
          const uint64_t ad = umma::make_smem_desc(sm_base + R_V + off, 128, kChunk);
          const uint64_t bd = umma::make_smem_desc(sm_base + R_K + off, 128, kChunk);
          umma::mma_f16(tmem + D_KV + 96 * sq, ad, bd, idesc, ks > 0);
        }
      }
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- R4: block-diagonal KV (per head 11x11) + the K-sum row as the B operand of the message GEMM
    if (r < 96) {
#pragma unroll
      for (int sq = 0; sq < NSEQ; ++sq) {
        uint8_t* kvbd = smem + (sq == 0 ? R_K : R_V);        // [96 rows b][96 cols a], chunk stride 96*16
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = g + 4 * i;
          if (c < 12) {
            float v[8];
            umma::tmem_ld8(tlane + D_KV + 96 * sq + 8 * c, v);
            umma::tmem_ld_wait();
  #pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int a = 8 * c + k;
              // rows 0..87: KV_h of the row's head; row 88+h: the K-sum of head h (per-head normaliser)
            const bool keep = (a < 88) && ((a / 11) == (r < 88 ? r / 11 : r - 88));
              v[k] = keep ? v[k] : 0.f;
            }
            *reinterpret_cast<uint4*>(kvbd + c * (96 * 16) + r * 16) = pack8<BF16>(v);
          }
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R5: message numerator Q'.KV_h (columns 0..87) and per-head normalisers Q'_h.Ksum_h (columns 88..95)
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t idesc = umma::make_idesc(128, 96, FMT, false, false);
#pragma unroll
      for (int sq = 0; sq < NSEQ; ++sq)
        issue_gemm_sub(tmem + D_MSG + 96 * sq, sm_base + R_Q, sm_base + (sq == 0 ? R_K : R_V), 96, 0, 12, idesc, 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- R6: msg = numerator / (normaliser + 1e-6)                (linear_attention.py:44-45)
    {
      const uint32_t dm = tlane + D_MSG + 96 * (NSEQ == 1 ? 0 : (r / SN));
      float zr[8];
      umma::tmem_ld8(dm + 88, zr);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) zr[j] = 1.f / (zr[j] + 1e-6f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 11) {
          float v[8];
          umma::tmem_ld8(dm + 8 * c, v);
          umma::tmem_ld_wait();
          // the 8 columns of a chunk belong to at most two heads (11 channels each)
          const int h0 = (8 * c) / 11, h1 = (8 * c + 7) / 11, split = 11 * h1;
          float z0 = zr[0], z1 = zr[0];
#pragma unroll
          for (int j = 1; j < 8; ++j) {
            z0 = (j == h0) ? zr[j] : z0;
            z1 = (j == h1) ? zr[j] : z1;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] *= ((8 * c + k) < split) ? z0 : z1;
          st_chunk<BF16>(smem + R_M, r, c, v);
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R7: merge  (A = message chunks 11..22 of the concat buffer; chunk 22 is the zero pad)
    if (tid == 0) {
      umma::mbar_wait(barB, phB);
      phB ^= 1;
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_MRG, sm_base + R_M, sm_base + R_SLOTB, 96, 0, 12, umma::make_idesc(128, 96, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    if (tid == 0) bulk_load(smem + R_SLOTB, wimg + RW_ML2, 96 * 176 * 2, barB);
    // ---- R8: LayerNorm 1 -> second half of the concat operand
    {
      float v[24];
      int nc = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 11) {
          umma::tmem_ld8(tlane + D_MRG + 8 * c, v + 8 * i);
          nc = i + 1;
        }
      }
      umma::tmem_ld_wait();
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 24; ++i)
        if (i < 8 * nc) {
          s += v[i];
          ss = fmaf(v[i], v[i], ss);
        }
      red[g * 128 + r] = make_float2(s, ss);
      __syncthreads();
      const float2 a0 = red[r], a1 = red[128 + r], a2 = red[256 + r], a3 = red[384 + r];
      const float mean = (a0.x + a1.x + a2.x + a3.x) * (1.f / 88.f);
      const float var = fmaxf((a0.y + a1.y + a2.y + a3.y) * (1.f / 88.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 11) {
          float o[8];
  #pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[8 * i + k] - mean) * rstd * prm.n1w[8 * c + k] + prm.n1b[8 * c + k];
          st_chunk<BF16>(smem + R_M, r, c, o);
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R9: mlp.0 on [x | msg]  (K = 176)
    if (tid == 0) {
      umma::mbar_wait(barA, phA);
      phA ^= 1;
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_ML0, sm_base + R_X, sm_base + R_SLOTA, 176, 0, 22, umma::make_idesc(128, 176, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    if (tid == 0 && tile + (long long)gridDim.x < n_tiles) bulk_load(smem + R_SLOTA, wimg + RW_QKV, 272 * 96 * 2, barA);
    // ---- R10: ReLU -> H1 operand (aliases K'/V' chunks 0..21)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = g + 4 * i;
      if (c < 22) {
        float v[8];
        umma::tmem_ld8(tlane + D_ML0 + 8 * c, v);
        umma::tmem_ld_wait();
  #pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
        st_chunk<BF16>(smem + R_K, r, c, v);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R11: mlp.2
    if (tid == 0) {
      umma::mbar_wait(barB, phB);
      phB ^= 1;
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_ML2, sm_base + R_K, sm_base + R_SLOTB, 96, 0, 22, umma::make_idesc(128, 96, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    if (tid == 0 && tile + (long long)gridDim.x < n_tiles) bulk_load(smem + R_SLOTB, wimg + RW_MRG, 96 * 96 * 2, barB);
    // ---- R12: LayerNorm 2, residual in fp32 from the fp32 input, split hi/lo for the SRDF head
    {
      float v[24];
      int nc = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 11) {
          umma::tmem_ld8(tlane + D_ML2 + 8 * c, v + 8 * i);
          nc = i + 1;
        }
      }
      umma::tmem_ld_wait();
      float s = 0.f, ss = 0.f;
#pragma unroll
      for (int i = 0; i < 24; ++i)
        if (i < 8 * nc) {
          s += v[i];
          ss = fmaf(v[i], v[i], ss);
        }
      red[g * 128 + r] = make_float2(s, ss);
      __syncthreads();
      const float2 a0 = red[r], a1 = red[128 + r], a2 = red[256 + r], a3 = red[384 + r];
      const float mean = (a0.x + a1.x + a2.x + a3.x) * (1.f / 88.f);
      const float var = fmaxf((a0.y + a1.y + a2.y + a3.y) * (1.f / 88.f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int c = g + 4 * i;
        if (c < 12) {
          float hi[8], lo[8];
          if (c < 11) {
            float x[8];
            if (c < 10) {
              if (row_ok) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(vout0 + in_row * kDView + 8 * c));
                const float4 b = __ldg(reinterpret_cast<const float4*>(vout0 + in_row * kDView + 8 * c + 4));
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
              } else {
  #pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = 0.f;
              }
            } else {
  #pragma unroll
              for (int k = 0; k < 8; ++k) x[k] = __ldg(pe_table + (r % SN) * 8 + k);
            }
            float o[8];
  #pragma unroll
            for (int k = 0; k < 8; ++k) {
              o[k] = x[k] + ((v[8 * i + k] - mean) * rstd * prm.n2w[8 * c + k] + prm.n2b[8 * c + k]);
              split_hi_lo<BF16>(o[k], hi[k], lo[k]);
            }
            if (ray_out != nullptr && row_ok) {
              float4* dst = reinterpret_cast<float4*>(ray_out + (size_t)prow * kDRay + 8 * c);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
          } else {
  #pragma unroll
            for (int k = 0; k < 8; ++k) hi[k] = lo[k] = 0.f;
          }
          st_chunk<BF16>(smem + R_Q, r, c, hi);
          st_chunk<BF16>(smem + R_RLO, r, c, lo);
        }
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- R13: DensityMLP layer 0 in split precision: r_hi.W_hi + r_lo.W_hi + r_hi.W_lo   (ray_transformer.py:147-150)
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t idesc = umma::make_idesc(128, 32, FMT, false, false);
      issue_gemm_sub(tmem + D_DEN, sm_base + R_Q, sm_base + R_WDEN, 32, 0, 12, idesc, 0);
      issue_gemm_sub(tmem + D_DEN, sm_base + R_RLO, sm_base + R_WDEN, 32, 0, 12, idesc, 1);
      issue_gemm_sub(tmem + D_DEN, sm_base + R_Q, sm_base + R_WDEN + 32 * 96 * 2, 32, 0, 12, idesc, 1);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- R14: DensityMLP tail 32 -> 16 -> 1 in fp32
    if (g == 0) {
      float h[32];
      umma::tmem_ld16(tlane + D_DEN, h);
      umma::tmem_ld16(tlane + D_DEN + 16, h + 16);
      umma::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = fmaxf(h[i] + prm.db0[i], 0.f);
      float acc = prm.db4;
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        float a = prm.db2[o];
#pragma unroll
        for (int i = 0; i < 32; ++i) a = fmaf(h[i], prm.dw2[o][i], a);
        acc = fmaf(fmaxf(a, 0.f), prm.dw4[o], acc);
      }
      if (row_ok) srdf[prow] = acc;
    }
    umma::tc_fence_before();
    __syncthreads();   // Q'/r_hi region and V' chunk 0.. are rewritten by the next tile's R2
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace ufo
