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

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float* v) {
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
  v[0] = __uint_as_float(r[0]);
  v[1] = __uint_as_float(r[1]);
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// elu(x) + 1 on a pair (packed fp32x2 multiply/add of sm_100, one MUFU.EX2 per element)
__device__ __forceinline__ float2 elu1_2(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 p = __fadd2_rn(x, make_float2(1.f, 1.f));
  return make_float2(x.x > 0.f ? p.x : ex2_ftz(t.x), x.y > 0.f ? p.y : ex2_ftz(t.y));
}
__device__ __forceinline__ float elu1_fast(float x) { return x > 0.f ? x + 1.f : ex2_ftz(x * 1.4426950408889634f); }

template <int N>
struct IC {
  static constexpr int value = N;
};
// call fn(IC<g>) with the warp-uniform column group as a compile-time constant (all chunk indices become static)
#define UFO_G_DISPATCH(fn)  \
  switch (g) {              \
    case 0: fn(IC<0>{}); break; \
    case 1: fn(IC<1>{}); break; \
    case 2: fn(IC<2>{}); break; \
    default: fn(IC<3>{}); break; \
  }

// 8 consecutive fp32 TMEM columns as 4 pairs
__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, float2* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2v(float2 a) { return umma::pack2<BF16>(a.x, a.y); }

template <bool BF16>
__device__ __forceinline__ void st_chunk2(uint8_t* tile, int row, int chunk, const float2* v) {
  uint4 u;
  u.x = pack2v<BF16>(v[0]); u.y = pack2v<BF16>(v[1]); u.z = pack2v<BF16>(v[2]); u.w = pack2v<BF16>(v[3]);
  *reinterpret_cast<uint4*>(tile_ptr(tile, row, chunk)) = u;
}

// relu on a pair after rounding to the operand format (relu commutes with the rounding)
template <bool BF16>
__device__ __forceinline__ uint32_t relu_pack2(float2 a) {
  uint32_t u = pack2v<BF16>(a);
  if (BF16) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    h = __hmax2(h, __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = *reinterpret_cast<__half2*>(&u);
  h = __hmax2(h, __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

// LayerNorm over a row whose columns are split over the 4 column groups: NCH chunks of 8 columns, group GG owns
// chunks GG, GG+4, ...  ld: load the chunks and return this thread's partial (sum, sum of squares).
template <int GG, int NCH>
__device__ __forceinline__ float2 ln_load(uint32_t tcol, float2 (*v)[4]) {
  constexpr int NI = (NCH - GG + 3) / 4;
#pragma unroll
  for (int i = 0; i < NI; ++i) tmem_ld8p(tcol + 8 * (GG + 4 * i), v[i]);
  umma::tmem_ld_wait();
  float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s = __fadd2_rn(s, v[i][k]);
      q = __ffma2_rn(v[i][k], v[i][k], q);
    }
  return make_float2(s.x + s.y, q.x + q.y);
}
// (mean, rstd) of the row from the 4 partials
__device__ __forceinline__ float2 ln_stats(const float2* red, int r, float inv_n) {
  const float2 a0 = red[r], a1 = red[128 + r], a2 = red[256 + r], a3 = red[384 + r];
  const float mean = ((a0.x + a1.x) + (a2.x + a3.x)) * inv_n;
  const float var = fmaxf(((a0.y + a1.y) + (a2.y + a3.y)) * inv_n - mean * mean, 0.f);
  return make_float2(mean, rsqrtf(var + 1e-5f));
}
// y = (v - mean) * rstd * w + b for one chunk, packed: a = w*rstd, y = v*a + (b - mean*a)
__device__ __forceinline__ void ln_apply(const float2* v, float2 st, const float* w, const float* b, float2* o) {
  const float2 rs = make_float2(st.y, st.y), nm = make_float2(-st.x, -st.x);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __fmul2_rn(make_float2(w[2 * k], w[2 * k + 1]), rs);
    const float2 c = __ffma2_rn(a, nm, make_float2(b[2 * k], b[2 * k + 1]));
    o[k] = __ffma2_rn(v[k], a, c);
  }
}

// split a float into a head that is exactly representable in the 16-bit operand format (mantissa truncated to
// 8 / 11 bits; assumes |x| inside the fp16 range in fp16 mode) and the exact remainder: hi + lo == x
template <bool BF16>
__device__ __forceinline__ void split_hi_lo(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & (BF16 ? 0xffff0000u : 0xffffe000u));
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
// tok [P][NV][80] 16-bit, channel c at tok_pos(c) (ufo_common.cuh): [8 x (feat 4 | vol 3 | depth-PE 1) | sim 16]
// =================================================================================================
// Per-point buffers of the tensor-core path are laid out [ray][128 slots]: slots 0..63 hold the coarse samples,
// 64..127 the importance samples, both in evaluation order, so that the fine pass of infer only evaluates the 64
// NEW points of a ray (the view stage is point-wise: re-evaluating the coarse points, as the reference does, gives
// identical values).  A pass over R*64 points with half = 0 | 1 addresses slot(p).
__device__ __forceinline__ long long tc_slot(long long p, int half) { return (p >> 6) * kNS + half * kNC + (p & 63); }

// Trilinear sample (align_corners=True, zeros) with one tap per lane: lane j of the 8-lane point group fetches the
// whole 8-channel voxel of tap (dx,dy,dz) = (j&1, j>>1&1, j>>2), scales it by its tap weight, and a 3-step
// reduce-scatter over the group leaves channel j of the interpolated feature in lane j (7 shuffles) and the
// interpolated weight-volume value in every lane (3 shuffles).  Same taps and weights as tri_fetch, other summation
// order.  Split in two halves so that the loads of several volumes can be in flight together.
// one 256-bit read-only load (sm_100: LDG.E.256): an 8-channel fp32 voxel is one 32-byte sector
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
struct TriTap {
  size_t idx;     // voxel index of this lane's tap (0 when the tap is out of range)
  float wgt;      // trilinear weight of the tap (0 when out of range)
};
__device__ __forceinline__ TriTap tri_tap(int D, int H, int W, float u, float v, float zn, int j) {
  const float ix = gs_unnorm<true>(u, W), iy = gs_unnorm<true>(v, H), iz = gs_unnorm<true>(zn, D);
  const bool near_vol = (ix > -1.f) && (ix < (float)W) && (iy > -1.f) && (iy < (float)H) && (iz > -1.f) && (iz < (float)D);
  const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
  const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
  const int x = (int)fx + dx, y = (int)fy + dy, z = (int)fz + dz;
  const float wx = dx ? ix - fx : (fx + 1.f) - ix, wy = dy ? iy - fy : (fy + 1.f) - iy, wz = dz ? iz - fz : (fz + 1.f) - iz;
  const bool ok = near_vol && (x >= 0) && (x < W) && (y >= 0) && (y < H) && (z >= 0) && (z < D);
  TriTap t;
  t.wgt = ok ? wx * wy * wz : 0.f;
  t.idx = ok ? ((size_t)z * H + y) * W + x : 0;
  return t;
}
__device__ __forceinline__ void tri_reduce(float4 a, float4 b, float ww, float wgt, int j, unsigned gmask, float& f_out, float& w_out) {
  ww *= wgt;
  a = f4_scale(a, wgt);
  b = f4_scale(b, wgt);
  const bool b2 = (j & 4) != 0, b1 = (j & 2) != 0, b0 = (j & 1) != 0;
  float4 keep = b2 ? b : a;
  const float4 send = b2 ? a : b;
  keep.x += __shfl_xor_sync(gmask, send.x, 4);
  keep.y += __shfl_xor_sync(gmask, send.y, 4);
  keep.z += __shfl_xor_sync(gmask, send.z, 4);
  keep.w += __shfl_xor_sync(gmask, send.w, 4);
  float kx = b1 ? keep.z : keep.x, ky = b1 ? keep.w : keep.y;
  const float sx = b1 ? keep.x : keep.z, sy = b1 ? keep.y : keep.w;
  kx += __shfl_xor_sync(gmask, sx, 2);
  ky += __shfl_xor_sync(gmask, sy, 2);
  float k1 = b0 ? ky : kx;
  const float s1 = b0 ? kx : ky;
  k1 += __shfl_xor_sync(gmask, s1, 1);
  ww += __shfl_xor_sync(gmask, ww, 4);
  ww += __shfl_xor_sync(gmask, ww, 2);
  ww += __shfl_xor_sync(gmask, ww, 1);
  f_out = k1;
  w_out = ww;
}

// bil_setup with the grid_sample convention as run-time flags (one code path for the three gather flavours)
__device__ __forceinline__ BilTaps bil_setup_rt(float u, float v, int H, int W, bool align, bool border) {
  float ix = align ? ((u + 1.f) / 2.f) * (float)(W - 1) : ((u + 1.f) * (float)W - 1.f) / 2.f;
  float iy = align ? ((v + 1.f) / 2.f) * (float)(H - 1) : ((v + 1.f) * (float)H - 1.f) / 2.f;
  if (border) {
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  }
  BilTaps t;
  const bool near_img = (ix > -1.f) && (ix < (float)W) && (iy > -1.f) && (iy < (float)H);  // false for NaN/inf
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix, wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const bool vx0 = near_img && x0 >= 0, vx1 = near_img && x1 < W, vy0 = y0 >= 0, vy1 = y1 < H;
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1), cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  t.i00 = near_img ? cy0 * W + cx0 : 0; t.i01 = near_img ? cy0 * W + cx1 : 0;
  t.i10 = near_img ? cy1 * W + cx0 : 0; t.i11 = near_img ? cy1 * W + cx1 : 0;
  t.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.f;
  t.w01 = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  t.w10 = (vx0 && vy1) ? wx0 * wy1 : 0.f;
  t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}

// shared-memory bytes of k_gather_tc<NV>: pre_sim_mlp weights, similarity staging, per-round projections and taps
// -DUFO_GATHER_SMEM2: bank-conflict-free shared-memory layout of k_gather_tc.  ncu (v38, source page) counts 5.5 wavefronts per LDS.64
// where 2 are ideal (the pre_sim weight fragments: rows of 32 floats put the eight rows a warp reads together into the same banks) and 5.6
// per STS.128 where 2.4 are ideal (the tap records: lane j writes 32 bytes at a 32-byte stride, lanes j and j + 4 collide) - 23 M of the
// kernel's 70 M shared-memory wavefronts, on the pipe that bounds it.  With the flag the weight rows are 40 floats apart and a point's tap
// records are two planes (indices | weights) that the eight lanes write as consecutive 16-byte pieces.  Same values, same order of operations.
#ifdef UFO_GATHER_SMEM2
constexpr int kGatherWRow = 40;
#else
constexpr int kGatherWRow = 32;
#endif
constexpr int kGatherWFloats = 8 * 32 + 32 + 32 * kGatherWRow + 32 + 16 * kGatherWRow + 16 + 16;   // + pad to a 16-byte multiple
template <int NV>
constexpr int gather_tc_smem() { return kGatherWFloats * 4 + 256 * 9 * 4 + 32 * NV * 16 + 32 * 3 * NV * 32; }

// Per round of 32 points the 8 lanes of a point group SHARE the per-view work that does not depend on the channel:
// lane j projects the point into view k%NV and sets up the bilinear taps of flavour k/NV (0: features, align_corners
// False / zeros on h x w; 1: colour+depth, same on H x W; 2: match maps, align_corners True / border on h x w) for
// k = j, j+8, ..; results go through shared memory.  (The exact path k_gather recomputes them in every lane and,
// for the match maps, for every pair: 3 NV + NV(NV-1) set-ups per lane instead of ceil(3 NV / 8).)
#ifndef UFO_GATHER_MINB
#define UFO_GATHER_MINB 3
#endif
template <int NV, bool BF16>
__global__ void __launch_bounds__(256, UFO_GATHER_MINB) k_gather_tc(SceneDev sc, const float* __restrict__ rayinfo,
                                                      const float* __restrict__ zbuf, int R, int half,
                                                      const float* __restrict__ freqs, const float* __restrict__ phases,
                                                      Mlp3Dev presim, uint16_t* __restrict__ tok, float4* __restrict__ rgbm,
                                                      float4* __restrict__ dirs, float* __restrict__ sim8_out) {
  constexpr int SN = kNC;
  extern __shared__ __align__(16) uint8_t gsm[];
  float* s_w = reinterpret_cast<float*>(gsm);
  float(*s_sim)[9] = reinterpret_cast<float(*)[9]>(gsm + kGatherWFloats * 4);
  float4* s_prj = reinterpret_cast<float4*>(gsm + kGatherWFloats * 4 + 256 * 9 * 4);
  uint4* s_tap = reinterpret_cast<uint4*>(gsm + kGatherWFloats * 4 + 256 * 9 * 4 + 32 * NV * 16);
  constexpr int WR = kGatherWRow;
  constexpr int o_b0 = 8 * 32, o_w2 = o_b0 + 32, o_b2 = o_w2 + 32 * WR, o_w4 = o_b2 + 32, o_b4 = o_w4 + 16 * WR;
  {  // pre_sim_mlp weights -> shared memory
    for (int i = threadIdx.x; i < 8 * 32; i += 256) s_w[i] = __ldg(presim.w0 + i);
    for (int i = threadIdx.x; i < 32; i += 256) s_w[o_b0 + i] = __ldg(presim.b0 + i);
    for (int i = threadIdx.x; i < 32 * 32; i += 256) s_w[o_w2 + (i >> 5) * WR + (i & 31)] = __ldg(presim.w2 + i);
    for (int i = threadIdx.x; i < 32; i += 256) s_w[o_b2 + i] = __ldg(presim.b2 + i);
    for (int i = threadIdx.x; i < 32 * 16; i += 256) s_w[o_w4 + (i >> 5) * WR + (i & 31)] = __ldg(presim.w4 + i);
    for (int i = threadIdx.x; i < 16; i += 256) s_w[o_b4 + i] = __ldg(presim.b4 + i);
  }
  __syncthreads();        // the weights are read after the gather rounds, which end in a warp-level barrier only (below)
  const int sub = threadIdx.x >> 3, j = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  const long long P = (long long)R * SN;
  const long long p0 = (long long)blockIdx.x * 256;
  const float fj = __ldg(freqs + j), pj = __ldg(phases + j);
  const size_t fstride = (size_t)sc.h * sc.w * kFeatC, istride = (size_t)sc.H * sc.W;
  float4* my_prj = s_prj + sub * NV;
  uint4* my_tap = s_tap + sub * (3 * NV) * 2;
#ifdef UFO_GATHER_SMEM2
  constexpr int kTapI = 1, kTapW = 3 * NV;      // record k of a point: indices at [k], weights at [3 NV + k]
#else
  constexpr int kTapI = 2, kTapW = 1;           // indices at [2 k], weights at [2 k + 1]
#endif
  for (int round = 0; round < 8; ++round) {
    const long long p = p0 + round * 32 + sub;
    if (p >= P) break;
    const int r = (int)(p / SN);
    const float* ri = rayinfo + (size_t)r * 8;
    const float zz = zbuf[p];
    const float x = __fadd_rn(sc.ray_o[0], __fmul_rn(zz, ri[0]));
    const float y = __fadd_rn(sc.ray_o[1], __fmul_rn(zz, ri[1]));
    const float z = __fadd_rn(sc.ray_o[2], __fmul_rn(zz, ri[2]));
    const size_t sl = (size_t)tc_slot(p, half);
    // ---- shared per-view work: projection + tap set-up, item k = flavour * NV + view
    __syncwarp(gmask);                                   // the previous round's readers are done
#pragma unroll
    for (int k0 = 0; k0 < 3 * NV; k0 += 8) {
      const int k = k0 + j;
      if (k < 3 * NV) {
        const int kind = k / NV, n = k - kind * NV;
        float u, v, qz;
        project_pt(sc.P[n], x, y, z, u, v, qz);
        if (kind == 0) my_prj[n] = make_float4(u, v, qz, 0.f);
        const bool big = (kind == 1);
        const BilTaps t = bil_setup_rt(u, v, big ? sc.H : sc.h, big ? sc.W : sc.w, kind == 2, kind == 2);
        my_tap[kTapI * k] = make_uint4((unsigned)t.i00, (unsigned)t.i01, (unsigned)t.i10, (unsigned)t.i11);
        my_tap[kTapI * k + kTapW] = make_uint4(__float_as_uint(t.w00), __float_as_uint(t.w01), __float_as_uint(t.w10), __float_as_uint(t.w11));
      }
    }
    __syncwarp(gmask);
    auto taps_of = [&](int k) {
      const uint4 a = my_tap[kTapI * k], b = my_tap[kTapI * k + kTapW];
      BilTaps t;
      t.i00 = (int)a.x; t.i01 = (int)a.y; t.i10 = (int)a.z; t.i11 = (int)a.w;
      t.w00 = __uint_as_float(b.x); t.w01 = __uint_as_float(b.y); t.w10 = __uint_as_float(b.z); t.w11 = __uint_as_float(b.w);
      return t;
    };
    // ---- frustum volumes (their 24 values go to every view row): blended over views with the summed weights
    float G[3] = {0.f, 0.f, 0.f}, Wsum = 0.f;
    {
      const float range = sc.far0 - sc.near0;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float4 pr = my_prj[n];
        const float zn = ((pr.z - sc.near0) / range) * 2.f - 1.f;         // camera.py:399-400
        float f[3], wl = 0.f;
        TriTap tp[3];
        float4 va[3], vb[3];
        float vwt[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) tp[s] = tri_tap(sc.vd[s], sc.vh[s], sc.vw[s], pr.x, pr.y, zn, j);
#pragma unroll
        for (int s = 0; s < 3; ++s) {          // all 9 loads of this view in flight together
          const size_t vox = (size_t)sc.vd[s] * sc.vh[s] * sc.vw[s];
          const float* vf = sc.vol_feat_cl[s] + (n * vox + tp[s].idx) * kVolC;
          ldg8(vf, va[s], vb[s]);            // the whole 32-byte voxel in one 256-bit load
          vwt[s] = __ldg(sc.vol_w[s] + n * vox + tp[s].idx);
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          float ws;
          tri_reduce(va[s], vb[s], vwt[s], tp[s].wgt, j, gmask, f[s], ws);
          wl += ws;                                                        // model.py:375-378
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) G[s] = fmaf(f[s], wl, G[s]);           // model.py:381-386
        Wsum += wl;
      }
      const float inv = 1.f / (Wsum + 1e-8f);                              // model.py:388
#pragma unroll
      for (int s = 0; s < 3; ++s) G[s] *= inv;
    }
    const uint16_t g0 = (uint16_t)(umma::pack2<BF16>(G[0], 0.f) & 0xffffu), g1 = (uint16_t)(umma::pack2<BF16>(G[1], 0.f) & 0xffffu),
                   g2 = (uint16_t)(umma::pack2<BF16>(G[2], 0.f) & 0xffffu);
    // ---- per view: feature, colour, MVS-depth gathers, depth PE, mask, relative direction -> token row (streamed out)
    const float rx = x - sc.ref_o[0], ry = y - sc.ref_o[1], rz = z - sc.ref_o[2];
    const float rn = 1.f / sqrtf(rx * rx + ry * ry + rz * rz);
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      uint16_t* row = tok + (sl * NV + n) * kDView;
      const float4 ft = bil_fetch32(sc.feat_cl + n * fstride, taps_of(n), j);
      const float4 c = bil_fetch_rgbd(sc.rgbd_cl + n * istride, taps_of(NV + n));
      const float zc = fmaf(sc.w2c_z[n][0], x, fmaf(sc.w2c_z[n][1], y, fmaf(sc.w2c_z[n][2], z, sc.w2c_z[n][3])));
      const float pe = __sinf(fmaf(c.w - zc, fj, pj));                     // ray_transformer.py:66,245 (|arg| < ~40)
      uint4 f;                                                             // this lane's 16 bytes of the row (tok_pos)
      f.x = umma::pack2<BF16>(ft.x, ft.y);
      f.y = umma::pack2<BF16>(ft.z, ft.w);
      f.z = (uint32_t)g0 | ((uint32_t)g1 << 16);
      f.w = (uint32_t)g2 | (umma::pack2<BF16>(pe, 0.f) << 16);
      *reinterpret_cast<uint4*>(row + 8 * j) = f;
      if (j == 0) {
        const float4 pr = my_prj[n];
        const bool inb = (pr.x <= 1.f) && (pr.x >= -1.f) && (pr.y <= 1.f) && (pr.y >= -1.f);
        rgbm[sl * NV + n] = make_float4(c.x, c.y, c.z, (inb && pr.z > 0.f) ? 1.f : 0.f);
        const float sx = x - sc.cam_o[n][0], sy = y - sc.cam_o[n][1], sz = z - sc.cam_o[n][2];
        const float sn = 1.f / sqrtf(sx * sx + sy * sy + sz * sz);
        dirs[sl * NV + n] = make_float4(rx * rn - sx * sn, ry * rn - sy * sn, rz * rn - sz * sn, 0.f);
      }
    }
    // ---- pairwise similarity prior (one map per pair, stored in both views' slots)
    {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < NV - 1; ++a) {
        const BilTaps ta = taps_of(2 * NV + a);
#pragma unroll
        for (int b = a + 1; b < NV; ++b) {
          const BilTaps tb = taps_of(2 * NV + b);
          const float4 fa = bil_fetch32(sc.match_cl + (size_t)sc.match_slot[a][b] * fstride, ta, j);
          const float4 fb = bil_fetch32(sc.match_cl + (size_t)sc.match_slot[b][a] * fstride, tb, j);
          acc += cos4(fa, fb);
        }
      }
      const float sim = acc / (float)(NV * (NV - 1) / 2);
      s_sim[sub * 8 + round][j] = sim;                  // row sub * 8 + round: the 32 points of a warp are the 32 rows of ITS two MMA tiles
      if (sim8_out != nullptr) sim8_out[sl * 8 + j] = sim;
    }
  }
  // A warp's MMA rows 32 w .. 32 w + 31 are the points it gathered itself (sub = 4 w .. 4 w + 3, eight rounds), so a warp-level barrier
  // is enough here; with rows in point order (round * 32 + sub) this was a block-wide barrier at which the warps of a block waited for
  // the slowest one's last gather round (8 % of the kernel's stall samples).
  __syncwarp();
  // ---- pre_sim_mlp 8 -> 32 -> 32 -> 16 (ray_transformer.py:128-132) for the block's 256 points on the tensor cores: warp-level
  //      mma.sync m16n8k16 (16-bit operands, fp32 accumulate, biases as the initial accumulator).  A warp owns 32 points = two
  //      16-row tiles; the accumulator fragment of one layer IS the A fragment of the next (rows g / g+8, column pairs 2t), so the
  //      hidden layers never leave registers.  This was one thread per point on the FMA pipe: 1400 instructions per thread, 14 %
  //      of the kernel's instructions and 31 % of its stall samples (dependent FMA chains) for 3.6 kFLOP per point.
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    auto pk = [&](float a, float b) { return umma::pack2<BF16>(a, b); };
    auto mma = [&](float* c, const uint32_t* af, uint32_t b0, uint32_t b1) {
      if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(b0), "r"(b1));
    };
    // B fragments (col-major k x n = the [out][in] weight rows): b0 = in 2t, 2t+1, b1 = in 2t+8, 2t+9 of out column g + 8 nt
    uint32_t w0f[4], w2f[2][4][2], w4f[2][2][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int n = gq + 8 * nt;
      w0f[nt] = pk(s_w[n * 8 + 2 * tq], s_w[n * 8 + 2 * tq + 1]);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float* w = s_w + o_w2 + n * WR + 16 * ks + 2 * tq;
        w2f[ks][nt][0] = pk(w[0], w[1]);
        w2f[ks][nt][1] = pk(w[8], w[9]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float* w = s_w + o_w4 + (gq + 8 * nt) * WR + 16 * ks + 2 * tq;
        w4f[ks][nt][0] = pk(w[0], w[1]);
        w4f[ks][nt][1] = pk(w[8], w[9]);
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = warp * 32 + mt * 16 + gq, r1 = r0 + 8;          // the two point rows of this thread's fragments
      uint32_t a[4] = {pk(s_sim[r0][2 * tq], s_sim[r0][2 * tq + 1]), pk(s_sim[r1][2 * tq], s_sim[r1][2 * tq + 1]), 0u, 0u};
      float c1[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        c1[nt][0] = c1[nt][2] = s_w[o_b0 + 8 * nt + 2 * tq];
        c1[nt][1] = c1[nt][3] = s_w[o_b0 + 8 * nt + 2 * tq + 1];
        mma(c1[nt], a, w0f[nt], 0u);
      }
      uint32_t h[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        h[ks][0] = pk(fmaxf(c1[2 * ks][0], 0.f), fmaxf(c1[2 * ks][1], 0.f));
        h[ks][1] = pk(fmaxf(c1[2 * ks][2], 0.f), fmaxf(c1[2 * ks][3], 0.f));
        h[ks][2] = pk(fmaxf(c1[2 * ks + 1][0], 0.f), fmaxf(c1[2 * ks + 1][1], 0.f));
        h[ks][3] = pk(fmaxf(c1[2 * ks + 1][2], 0.f), fmaxf(c1[2 * ks + 1][3], 0.f));
      }
      float c2[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        c2[nt][0] = c2[nt][2] = s_w[o_b2 + 8 * nt + 2 * tq];
        c2[nt][1] = c2[nt][3] = s_w[o_b2 + 8 * nt + 2 * tq + 1];
        mma(c2[nt], h[0], w2f[0][nt][0], w2f[0][nt][1]);
        mma(c2[nt], h[1], w2f[1][nt][0], w2f[1][nt][1]);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        h[ks][0] = pk(fmaxf(c2[2 * ks][0], 0.f), fmaxf(c2[2 * ks][1], 0.f));
        h[ks][1] = pk(fmaxf(c2[2 * ks][2], 0.f), fmaxf(c2[2 * ks][3], 0.f));
        h[ks][2] = pk(fmaxf(c2[2 * ks + 1][0], 0.f), fmaxf(c2[2 * ks + 1][1], 0.f));
        h[ks][3] = pk(fmaxf(c2[2 * ks + 1][2], 0.f), fmaxf(c2[2 * ks + 1][3], 0.f));
      }
      float c3[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        c3[nt][0] = c3[nt][2] = s_w[o_b4 + 8 * nt + 2 * tq];
        c3[nt][1] = c3[nt][3] = s_w[o_b4 + 8 * nt + 2 * tq + 1];
        mma(c3[nt], h[0], w4f[0][nt][0], w4f[0][nt][1]);
        mma(c3[nt], h[1], w4f[1][nt][0], w4f[1][nt][1]);
      }
      // token columns 56..71, the same for every view row of the point: the four lanes of a quad hold 4 bytes each of a row's
      // two 16-byte halves; they exchange them by shuffle and lane t writes the rows of views t, t+4, .. as 16-byte stores
      // (4-byte stores, 4 NV per thread, cost 25 % of the kernel at NV = 10)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const uint32_t v0 = pk(c3[0][2 * hr], c3[0][2 * hr + 1]), v1 = pk(c3[1][2 * hr], c3[1][2 * hr + 1]);
        const int qb = lane & ~3;
        uint4 lo, hi;
        lo.x = __shfl_sync(0xffffffffu, v0, qb);     lo.y = __shfl_sync(0xffffffffu, v0, qb + 1);
        lo.z = __shfl_sync(0xffffffffu, v0, qb + 2); lo.w = __shfl_sync(0xffffffffu, v0, qb + 3);
        hi.x = __shfl_sync(0xffffffffu, v1, qb);     hi.y = __shfl_sync(0xffffffffu, v1, qb + 1);
        hi.z = __shfl_sync(0xffffffffu, v1, qb + 2); hi.w = __shfl_sync(0xffffffffu, v1, qb + 3);
        const int row = hr ? r1 : r0;                                    // row = sub * 8 + round of point p0 + round * 32 + sub
        const long long p = p0 + (row & 7) * 32 + (row >> 3);
        if (p < P) {
          const size_t sl = (size_t)tc_slot(p, half);
#pragma unroll
          for (int n0 = 0; n0 < NV; n0 += 4) {
            const int n = n0 + tq;
            if (n < NV) {
              uint16_t* row = tok + (sl * NV + n) * kDView;
              *reinterpret_cast<uint4*>(row + 64) = lo;                   // tok_pos(56..71)
              *reinterpret_cast<uint4*>(row + 72) = hi;
            }
          }
        }
      }
    }
  }
}

// =================================================================================================
// view stage: density_view_transformer (d = 80) + radiance-weight head, tokens = views of one point
// =================================================================================================
template <int NV, bool BF16>
__global__ void __launch_bounds__(tc::kThreads, 1)
k_view_tc(const uint8_t* __restrict__ wimg, const __grid_constant__ ViewParams prm, const uint16_t* __restrict__ tok,
          const float4* __restrict__ rgbm, const float4* __restrict__ dirs, int P, int half,
          float* __restrict__ vout0, float4* __restrict__ radiance) {
  using namespace tc;
  static_assert(kGroups == 4, "epilogues are written for 4 column groups (2 heads per thread)");
  constexpr int L = NV + 1, PPT = 128 / L, ROWS = PPT * L;
  constexpr int NPF = (1280 + kThreads - 1) / kThreads;   // 16-byte token pieces per thread per tile
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  constexpr uint32_t D_QKV = 0, D_RAD0 = 240, D_RAD1 = 496, D_ML0 = 256, D_MRG = 416, D_ML2 = 0;   // TMEM columns
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + V_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + V_BAR + 16);
  float2* red = reinterpret_cast<float2*>(smem + V_RED);
  float* omg = reinterpret_cast<float*>(smem + V_OMG);
  float4* s_rgbm = reinterpret_cast<float4*>(smem + V_RGBM);
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
  for (int i = tid; i < 128 * 22; i += kThreads) {
    const int rr = i & 127, c = i >> 7;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (c < 10 && rr < ROWS && (rr % L) == 0) ? prm.vtok_x[c * 8 + k] : 0.f;
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
  const int n_tiles = (P + PPT - 1) / PPT;

  // 16-byte pieces of the token rows, prefetched one tile ahead into registers; the (row, chunk) -> (point, view)
  // mapping of a piece does not depend on the tile
  uint4 pf[NPF];
  int pf_pt[NPF], pf_off[NPF];       // point within the tile (-1: no load), element offset inside the point's rows
  uint32_t pf_dst[NPF];
#pragma unroll
  for (int k = 0; k < NPF; ++k) {
    const int i = tid + k * kThreads;
    const int rr = i & 127, c = i >> 7;
    const int pr = rr / L, ll = rr - pr * L;
    const bool on = (i < 1280) && (rr < ROWS) && (ll > 0);
    pf_pt[k] = on ? pr : -1;
    pf_off[k] = (ll - 1) * kDView + c * 8;
    pf_dst[k] = (uint32_t)(c * kChunk + rr * 16);
  }
  auto slot_of = [&](int p) -> int { return (p >> 6) * kNS + half * kNC + (p & 63); };
  auto prefetch = [&](int tile) {
    const int pbase = tile * PPT;
#pragma unroll
    for (int k = 0; k < NPF; ++k) {
      pf[k] = make_uint4(0, 0, 0, 0);
      const int p = pbase + pf_pt[k];
      if (pf_pt[k] >= 0 && tile < n_tiles && p < P)
        pf[k] = __ldg(reinterpret_cast<const uint4*>(tok + (size_t)slot_of(p) * (NV * kDView) + pf_off[k]));
    }
  };
  prefetch(blockIdx.x);

  // P11: radiance-head tail 16 -> 8 -> 1 (hidden units g and g+4 per thread), masked softmax over views, colour blend.
  // Reads the accumulator D_RAD[parity] of a finished tile; called while the next tile's QKV GEMM is in flight.
  auto rad_tail = [&](int pb, uint32_t drad, uint32_t cpar) {
  {
    float h[16];
    umma::tmem_ld16(tlane + drad, h);
    umma::tmem_ld_wait();
    auto tail = [&](auto GGc) {
      constexpr int GG = decltype(GGc)::value;
#pragma unroll
      for (int o = 0; o < 16; ++o) h[o] = fmaxf(h[o], 0.f);      // bias and direction terms came through the GEMM
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int o = GG + 4 * j;
        float a0 = prm.rb2[o], a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          a0 = fmaf(h[i], prm.rw2[o][i], a0);
          a1 = fmaf(h[i + 1], prm.rw2[o][i + 1], a1);
        }
        part = fmaf(fmaxf(a0 + a1, 0.f), prm.rw4[o], part);
      }
      omg[GG * 128 + r] = part;
    };
    UFO_G_DISPATCH(tail)
  }
  umma::tc_fence_before();
  __syncthreads();
  if (tid < PPT && pb + tid < P) {
    const size_t p = (size_t)slot_of(pb + tid);
    float om[NV];
    float4 col[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      col[n] = s_rgbm[cpar * 128 + tid * NV + n];          // staged by the tile itself (P5), no global latency here
      const int rr = tid * L + 1 + n;
      const float w = prm.rb4 + ((omg[rr] + omg[128 + rr]) + (omg[256 + rr] + omg[384 + rr]));
      om[n] = (col[n].w == 0.f) ? -1e9f : w;                     // ray_transformer.py:316
      mx = fmaxf(mx, om[n]);
    }
    float den = 0.f;
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      om[n] = ex2_ftz((om[n] - mx) * 1.4426950408889634f);
      den += om[n];
    }
    float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
    for (int n = 0; n < NV; ++n) {
      const float pw = om[n] / den;
      cr = fmaf(col[n].x, pw, cr);
      cg = fmaf(col[n].y, pw, cg);
      cb = fmaf(col[n].z, pw, cb);
    }
    radiance[p] = make_float4(cr, cg, cb, 0.f);
  }
  };
  bool have_prev = false;
  int prev_pbase = 0;
  uint32_t prev_drad = D_RAD0, par = 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int pbase = tile * PPT;
    const uint32_t D_RAD = par ? D_RAD1 : D_RAD0;
    // ---- P0: token rows of this tile (prefetched) -> X (A operand, K = 80)
#pragma unroll
    for (int k = 0; k < NPF; ++k)
      if (pf_pt[k] >= 0) *reinterpret_cast<uint4*>(smem + V_X + pf_dst[k]) = pf[k];
    const int my_p = pbase + pl;
    const size_t my_slot = (size_t)slot_of(my_p);
    const bool view_row = row_ok && l > 0 && my_p < P;
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P1: q|k|v = X . Wqkv^T (transformer.py:47); the x halves of mlp.0 and of the radiance head are issued
    //      right behind it so that they run under the attention phase
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_QKV, sm_base + V_X, sm_base + V_WQKV, 240, 0, 10, umma::make_idesc(128, 240, FMT, false, false), 0);
      umma::commit(bar);
      issue_gemm_sub(tmem + D_ML0, sm_base + V_X, sm_base + V_WML0, 160, 0, 10, umma::make_idesc(128, 160, FMT, false, false), 0);
      issue_gemm_sub(tmem + D_RAD, sm_base + V_X, sm_base + V_WRAD, 16, 0, 10, umma::make_idesc(128, 16, FMT, false, false), 0);
    }
    prefetch(tile + gridDim.x);
    // relative direction of this row's (point, view): loaded as float2 + float - a float4 load would leave the unused .w
    // register pending, and the next writer of that register would wait for the whole load (WAW on the scoreboard)
    float4 my_dir = make_float4(0.f, 0.f, 0.f, 0.f);
    if (view_row) {
      const float* dp = reinterpret_cast<const float*>(dirs + my_slot * NV + (l - 1));
      const float2 dxy = __ldg(reinterpret_cast<const float2*>(dp));
      my_dir.x = dxy.x;
      my_dir.y = dxy.y;
      my_dir.z = __ldg(dp + 2);
    }
    // colour + mask of this tile's (point, view) pairs: loaded now, parked in shared memory at P5, used by rad_tail
    float4 my_col = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < PPT * NV) {
      const int pp = pbase + tid / NV;
      if (pp < P) my_col = __ldg(rgbm + (size_t)slot_of(pp) * NV + (tid % NV));
    }
    if (have_prev) rad_tail(prev_pbase, prev_drad, par ^ 1);   // the previous tile's head tail, under this tile's QKV GEMM
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P2: elu+1 on q, k; stage K', V' (16-bit) per head for the per-point attention   (linear_attention.py:36-41)
    //      thread (row, g) owns heads 2g, 2g+1 = columns 20g .. 20g+19 of q, k and v
    float2 qv[10];
    {
      float t[20];
      umma::tmem_ld16(tlane + D_QKV + 20 * g, t);
      tmem_ld4(tlane + D_QKV + 20 * g + 16, t + 16);
      float kk[20];
      umma::tmem_ld16(tlane + D_QKV + 80 + 20 * g, kk);
      tmem_ld4(tlane + D_QKV + 80 + 20 * g + 16, kk + 16);
      float vv[20];
      umma::tmem_ld16(tlane + D_QKV + 160 + 20 * g, vv);
      tmem_ld4(tlane + D_QKV + 160 + 20 * g + 16, vv + 16);
      umma::tmem_ld_wait();
      uint32_t pk[10], pv[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        qv[i] = elu1_2(make_float2(t[2 * i], t[2 * i + 1]));
        pk[i] = pack2v<BF16>(elu1_2(make_float2(kk[2 * i], kk[2 * i + 1])));
        pv[i] = umma::pack2<BF16>(vv[2 * i], vv[2 * i + 1]);
      }
      // staging [8 heads][128 rows][k 10 | v 10] 16-bit = 40 B per (head, row)
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        uint2* dst = reinterpret_cast<uint2*>(smem + V_KV + (size_t)((2 * g + hp) * 128 + r) * 40);
        dst[0] = make_uint2(pk[5 * hp + 0], pk[5 * hp + 1]);
        dst[1] = make_uint2(pk[5 * hp + 2], pk[5 * hp + 3]);
        dst[2] = make_uint2(pk[5 * hp + 4], pv[5 * hp + 0]);
        dst[3] = make_uint2(pv[5 * hp + 1], pv[5 * hp + 2]);
        dst[4] = make_uint2(pv[5 * hp + 3], pv[5 * hp + 4]);
      }
    }
    __syncthreads();
    // ---- P3: msg_l = sum_s (Q_l.K_s) V_s / (sum_s Q_l.K_s + 1e-6)  per head   (== Q (K^T V) Z, linear_attention.py:43-45)
    {
      uint32_t mo[10];
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        float2 msg[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) msg[i] = make_float2(0.f, 0.f);
        if (row_ok) {
          float den = 0.f;
#pragma unroll
          for (int s = 0; s < L; ++s) {
            const uint2* src = reinterpret_cast<const uint2*>(smem + V_KV + (size_t)((2 * g + hp) * 128 + pl * L + s) * 40);
            float2 kv[10];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const uint2 u = src[i];
              kv[2 * i] = unpack2<BF16>(u.x);
              kv[2 * i + 1] = unpack2<BF16>(u.y);
            }
            float2 acc = __fmul2_rn(qv[5 * hp], kv[0]);
#pragma unroll
            for (int a = 1; a < 5; ++a) acc = __ffma2_rn(qv[5 * hp + a], kv[a], acc);
            const float sc = acc.x + acc.y;
            den += sc;
            const float2 sc2 = make_float2(sc, sc);
#pragma unroll
            for (int b = 0; b < 5; ++b) msg[b] = __ffma2_rn(sc2, kv[5 + b], msg[b]);
          }
          const float zi = 1.f / (den + 1e-6f);
          const float2 z2 = make_float2(zi, zi);
#pragma unroll
          for (int b = 0; b < 5; ++b) msg[b] = __fmul2_rn(msg[b], z2);
        }
#pragma unroll
        for (int b = 0; b < 5; ++b) mo[5 * hp + b] = pack2v<BF16>(msg[b]);
      }
      // columns 20g .. 20g+19 of the message tile: two full chunks and one half chunk
      const int c0 = (20 * g) >> 3;
      if ((g & 1) == 0) {
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_M, r, c0)) = make_uint4(mo[0], mo[1], mo[2], mo[3]);
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_M, r, c0 + 1)) = make_uint4(mo[4], mo[5], mo[6], mo[7]);
        *reinterpret_cast<uint2*>(tile_ptr(smem + V_M, r, c0 + 2)) = make_uint2(mo[8], mo[9]);
      } else {
        *reinterpret_cast<uint2*>(tile_ptr(smem + V_M, r, c0) + 8) = make_uint2(mo[0], mo[1]);
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_M, r, c0 + 1)) = make_uint4(mo[2], mo[3], mo[4], mo[5]);
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_M, r, c0 + 2)) = make_uint4(mo[6], mo[7], mo[8], mo[9]);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P4: merge                                                (transformer.py:55)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_MRG, sm_base + V_M, sm_base + V_WMRG, 80, 0, 10, umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P5: LayerNorm 1 -> second half of the concat operand     (transformer.py:56)
    if (tid < PPT * NV) s_rgbm[par * 128 + tid] = my_col;
    {
      auto ln1 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (10 - GG + 3) / 4;
        float2 v[NI][4];
        red[GG * 128 + r] = ln_load<GG, 10>(tlane + D_MRG, v);
        __syncthreads();
        const float2 st = ln_stats(red, r, 1.f / 80.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
          float2 o[4];
          ln_apply(v[i], st, prm.n1w + 8 * c, prm.n1b + 8 * c, o);
          st_chunk2<BF16>(smem + V_M, r, c, o);
        }
      };
      UFO_G_DISPATCH(ln1)
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P6: message half of mlp.0 on top of the x half issued in P1   (transformer.py:57)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_ML0, sm_base + V_M, sm_base + V_WML0 + 10 * (160 * 16), 160, 0, 10, umma::make_idesc(128, 160, FMT, false, false), 1);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P7: ReLU -> H1 operand (aliases the K'/V' staging); group g owns chunks g, g+4, .., g+16
    {
      float2 v[5][4];
#pragma unroll
      for (int i = 0; i < 5; ++i) tmem_ld8p(tlane + D_ML0 + 8 * (g + 4 * i), v[i]);
      umma::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 5; ++i)
        *reinterpret_cast<uint4*>(tile_ptr(smem + V_KV, r, g + 4 * i)) =
            make_uint4(relu_pack2<BF16>(v[i][0]), relu_pack2<BF16>(v[i][1]), relu_pack2<BF16>(v[i][2]), relu_pack2<BF16>(v[i][3]));
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P8: mlp.2
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_ML2, sm_base + V_KV, sm_base + V_WML2, 80, 0, 20, umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // ---- P9: LayerNorm 2; token 0: out = view_token + LN2 -> vout0 (fp32); view rows: LN2 -> operand for the
    //      radiance head (x + LN2 is applied inside the head's GEMM: W0x.x + W0x.LN2)
    {
      auto ln2 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (10 - GG + 3) / 4;
        float2 v[NI][4];
        red[GG * 128 + r] = ln_load<GG, 10>(tlane + D_ML2, v);
        __syncthreads();
        const float2 st = ln_stats(red, r, 1.f / 80.f);
        const bool tok0 = row_ok && l == 0 && my_p < P;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
          float2 o[4];
          ln_apply(v[i], st, prm.n2w + 8 * c, prm.n2b + 8 * c, o);
          st_chunk2<BF16>(smem + V_M, r, c, o);
          if (tok0) {
            float4* dst = reinterpret_cast<float4*>(vout0 + my_slot * kDView + 8 * c);
            dst[0] = make_float4(prm.vtok[8 * c] + o[0].x, prm.vtok[8 * c + 1] + o[0].y, prm.vtok[8 * c + 2] + o[1].x, prm.vtok[8 * c + 3] + o[1].y);
            dst[1] = make_float4(prm.vtok[8 * c + 4] + o[2].x, prm.vtok[8 * c + 5] + o[2].y, prm.vtok[8 * c + 6] + o[3].x, prm.vtok[8 * c + 7] + o[3].y);
          }
        }
      };
      UFO_G_DISPATCH(ln2)
      if (g == 0) {   // radiance-head side inputs of this row: relative direction and the constant 1 that carries the bias
        const float e[8] = {my_dir.x, my_dir.y, my_dir.z, 1.f, 0.f, 0.f, 0.f, 0.f};
        st_chunk<BF16>(smem + V_M, r, 10, e);
      }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    // ---- P10: LN2 half of the radiance head's first layer + direction/bias chunks   (ray_transformer.py:159-163,313)
    if (tid == 0) {
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_RAD, sm_base + V_M, sm_base + V_WRAD + 10 * (16 * 16), 16, 0, 12, umma::make_idesc(128, 16, FMT, false, false), 1);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    // P11 (head tail, softmax, colour blend) of this tile is deferred: it runs under the QKV GEMM of the next tile
    have_prev = true;
    prev_pbase = pbase;
    prev_drad = D_RAD;
    par ^= 1;
    // the next tile's P0 writes X only after this tile's last MMA (P10) has completed: guaranteed by the wait above;
    // omg is next written after several more block-wide barriers
  }
  if (have_prev) rad_tail(prev_pbase, prev_drad, par ^ 1);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// =================================================================================================
// ray stage: density_ray_transformer (d = 88) + DensityMLP, tokens = samples of one ray
// =================================================================================================
// SN = 128: one ray per tile; SN = 64: two rays per tile.  vout0 [slots][80] fp32 (token-0 output of the view
// stage), pe_table [128][8], perm [R][128] (fine pass), srdf [P] out, ray_out [P][88] optional tap.
template <int SN, bool BF16>
__global__ void __launch_bounds__(tc::kThreads, 1)
k_ray_tc(const uint8_t* __restrict__ wimg, const __grid_constant__ RayParams prm, const float* __restrict__ vout0,
         const float* __restrict__ pe_table, const uint8_t* __restrict__ perm, long long P, float* __restrict__ srdf,
         float* __restrict__ ray_out) {
  using namespace tc;
  constexpr int G = kGroups;
  static_assert(kGroups == 4, "epilogues are written for 4 column groups");
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  constexpr int NSEQ = 128 / SN;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + R_BAR);        // MMA completion
  uint64_t* barA = reinterpret_cast<uint64_t*>(smem + R_BAR + 8);   // slot A filled
  uint64_t* barB = reinterpret_cast<uint64_t*>(smem + R_BAR + 16);  // slot B filled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + R_BAR + 24);
  float2* red = reinterpret_cast<float2*>(smem + R_SCR);            // [G][128] partials (inside dead V' chunks)
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
  // x of a tile: token-0 output of the view stage at slot(ray, evaluation index) -> 16-bit A operand chunks 0..9
  auto in_row_of = [&](long long tile) -> long long {
    const long long prow = tile * 128 + r;
    if (prow >= P) return -1;
    return (SN == kNC) ? tc_slot(prow, 0) : tile * 128 + (perm ? (long long)perm[prow] : (long long)r);
  };
  // chunks g, g+4, g+8 (< 10) of the row: raw loads (issued early) and conversion + store (after the loads' latency)
  float4 xr[6];
  auto x_issue = [&](long long ir) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = g + 4 * i;
      xr[2 * i] = xr[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < 10 && ir >= 0) {
        xr[2 * i] = __ldg(reinterpret_cast<const float4*>(vout0 + (size_t)ir * kDView + 8 * c));
        xr[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(vout0 + (size_t)ir * kDView + 8 * c + 4));
      }
    }
  };
  auto x_store = [&]() {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = g + 4 * i;
      if (c < 10) {
        const float2 v[4] = {make_float2(xr[2 * i].x, xr[2 * i].y), make_float2(xr[2 * i].z, xr[2 * i].w),
                             make_float2(xr[2 * i + 1].x, xr[2 * i + 1].y), make_float2(xr[2 * i + 1].z, xr[2 * i + 1].w)};
        st_chunk2<BF16>(smem + R_X, r, c, v);
      }
    }
  };
  if ((long long)blockIdx.x < n_tiles) {
    x_issue(in_row_of(blockIdx.x));
    x_store();
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
  constexpr uint32_t D_QKV = 0, D_KV = 272, D_MSG = 0, D_MRG = 96, D_ML0 = 192, D_ML2 = 0, D_DEN = 464;

  // R14: DensityMLP tail 32 -> 16 -> 1 in fp32 (hidden units g, g+4, g+8, g+12 per thread) of a finished tile, read from
  // its accumulator D_DEN; called while the next tile's QKV GEMM is in flight (D_DEN has its own TMEM columns).
  auto srdf_tail = [&](long long pr, bool ok) {
  {
    float h[32];
    umma::tmem_ld16(tlane + D_DEN, h);
    umma::tmem_ld16(tlane + D_DEN + 16, h + 16);
    umma::tmem_ld_wait();
    float* part_s = reinterpret_cast<float*>(smem + R_SCR);     // the split operands are dead (MMA completed)
    auto tail = [&](auto GGc) {
      constexpr int GG = decltype(GGc)::value;
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = fmaxf(h[i] + prm.db0[i], 0.f);
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = GG + 4 * j;
        float2 acc = make_float2(prm.db2[o], 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          acc = __ffma2_rn(make_float2(h[i], h[i + 1]), make_float2(prm.dw2[o][i], prm.dw2[o][i + 1]), acc);
        part = fmaf(fmaxf(acc.x + acc.y, 0.f), prm.dw4[o], part);
      }
      part_s[GG * 128 + r] = part;
    };
    UFO_G_DISPATCH(tail)
    umma::tc_fence_before();
    __syncthreads();
    if (g == 0 && ok) srdf[pr] = prm.db4 + ((part_s[r] + part_s[128 + r]) + (part_s[256 + r] + part_s[384 + r]));
  }
    __syncthreads();   // the partial-sum scratch lives in V' chunks that the next R2 rewrites
  };
  bool have_prev = false, prev_ok = false;
  long long prev_prow = 0;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long prow = tile * 128 + r;           // this thread's token: ray prow/SN, sorted sample prow%SN
    const bool row_ok = prow < P;
    const long long in_row = in_row_of(tile);
    const bool has_next = tile + (long long)gridDim.x < n_tiles;
    const long long in_row_nx = has_next ? in_row_of(tile + gridDim.x) : -1;     // its perm lookup completes under R1..R9
    // ---- R1: q|k|v = x . Wqkv^T   (K = 96: columns 88..95 hit zero weight columns); x was staged by the previous
    //      iteration (or the prologue) and fenced there
    if (tid == 0) {
      umma::mbar_wait(barA, phA);
      phA ^= 1;
      umma::tc_fence_after();
      issue_gemm_sub(tmem + D_QKV, sm_base + R_X, sm_base + R_SLOTA, 272, 0, 12, umma::make_idesc(128, 176, FMT, false, false), 0);
      issue_gemm_sub(tmem + D_QKV + 176, sm_base + R_X, sm_base + R_SLOTA, 272, 176, 12, umma::make_idesc(128, 96, FMT, false, false), 0);
      umma::commit(bar);
    }
    if (have_prev) srdf_tail(prev_prow, prev_ok);       // the previous tile's SRDF tail, under this tile's QKV GEMM
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
    if (tid == 0) bulk_load(smem + R_SLOTA, wimg + RW_ML0, 176 * 176 * 2, barA);   // slot A is free again
    // ---- R2: Q' = elu(q)+1, K' = elu(k)+1, V' = v  -> 16-bit operand tiles    (linear_attention.py:36-41)
    {
      auto r2 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (11 - GG + 3) / 4;
        float2 a[NI][4], b[NI][4], d[NI][4];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
          tmem_ld8p(tlane + D_QKV + 8 * c, a[i]);
          tmem_ld8p(tlane + D_QKV + 88 + 8 * c, b[i]);
          tmem_ld8p(tlane + D_QKV + 176 + 8 * c, d[i]);
        }
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            a[i][k] = elu1_2(a[i][k]);
            b[i][k] = elu1_2(b[i][k]);
          }
          st_chunk2<BF16>(smem + R_Q, r, c, a[i]);
          st_chunk2<BF16>(smem + R_K, r, c, b[i]);
          st_chunk2<BF16>(smem + R_V, r, c, d[i]);
        }
      };
      UFO_G_DISPATCH(r2)
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
    // ---- R4: block-diagonal KV (per head 11x11) + per-head K-sum rows as the B operand of the message GEMM
    if (r < 96) {
      // rows 0..87: KV_h of the row's head; row 88+h: the K-sum of head h (per-head normaliser)
      const int hr = r < 88 ? r / 11 : r - 88;
      auto r4 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
#pragma unroll
        for (int sq = 0; sq < NSEQ; ++sq) {
          uint8_t* kvbd = smem + (sq == 0 ? R_K : R_V);        // [96 rows b][96 cols a], chunk stride 96*16
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c = GG + 4 * i;
            float v[8];
            umma::tmem_ld8(tlane + D_KV + 96 * sq + 8 * c, v);
            umma::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ((8 * c + k) < 88 && ((8 * c + k) / 11) == hr) ? v[k] : 0.f;
            *reinterpret_cast<uint4*>(kvbd + c * (96 * 16) + r * 16) = pack8<BF16>(v);
          }
        }
      };
      UFO_G_DISPATCH(r4)
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
      auto r6 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (11 - GG + 3) / 4;
        float v[NI][8];
#pragma unroll
        for (int i = 0; i < NI; ++i) umma::tmem_ld8(dm + 8 * (GG + 4 * i), v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
#pragma unroll
          for (int k = 0; k < 8; ++k) v[i][k] *= zr[(8 * c + k) / 11];      // static index: head of column 8c+k
          st_chunk<BF16>(smem + R_M, r, c, v[i]);
        }
      };
      UFO_G_DISPATCH(r6)
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
      auto ln1 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (11 - GG + 3) / 4;
        float2 v[NI][4];
        red[GG * 128 + r] = ln_load<GG, 11>(tlane + D_MRG, v);
        __syncthreads();
        const float2 st = ln_stats(red, r, 1.f / 88.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = GG + 4 * i;
          float2 o[4];
          ln_apply(v[i], st, prm.n1w + 8 * c, prm.n1b + 8 * c, o);
          st_chunk2<BF16>(smem + R_M, r, c, o);
        }
      };
      UFO_G_DISPATCH(ln1)
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
    if (tid == 0 && has_next) bulk_load(smem + R_SLOTA, wimg + RW_QKV, 272 * 96 * 2, barA);
    // x of the next tile: the concat operand is free from here on; the loads fly under the ReLU epilogue
    if (has_next) x_issue(in_row_nx);
    // ---- R10: ReLU -> H1 operand (aliases K'/V' chunks 0..21)
    {
      auto r10 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (22 - GG + 3) / 4;
        float2 v[NI][4];
#pragma unroll
        for (int i = 0; i < NI; ++i) tmem_ld8p(tlane + D_ML0 + 8 * (GG + 4 * i), v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NI; ++i)
          *reinterpret_cast<uint4*>(tile_ptr(smem + R_K, r, GG + 4 * i)) =
              make_uint4(relu_pack2<BF16>(v[i][0]), relu_pack2<BF16>(v[i][1]), relu_pack2<BF16>(v[i][2]), relu_pack2<BF16>(v[i][3]));
      };
      UFO_G_DISPATCH(r10)
    }
    if (has_next) x_store();
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
    if (tid == 0 && has_next) bulk_load(smem + R_SLOTB, wimg + RW_MRG, 96 * 96 * 2, barB);
    // ---- R12: LayerNorm 2, residual in fp32 from the fp32 input, split hi/lo for the SRDF head
    {
      auto ln2 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int NI = (11 - GG + 3) / 4;
        // fp32 residual input of this row (chunks < 10 from the view stage, chunk 10 = order encoding): issued first
        float4 xa[3], xb[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = GG + 4 * i;
          xa[i] = xb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < 10) {
            if (row_ok) {
              xa[i] = __ldg(reinterpret_cast<const float4*>(vout0 + (size_t)in_row * kDView + 8 * c));
              xb[i] = __ldg(reinterpret_cast<const float4*>(vout0 + (size_t)in_row * kDView + 8 * c + 4));
            }
          } else if (c == 10) {
            xa[i] = __ldg(reinterpret_cast<const float4*>(pe_table + (r % SN) * 8));
            xb[i] = __ldg(reinterpret_cast<const float4*>(pe_table + (r % SN) * 8 + 4));
          }
        }
        float2 v[NI][4];
        red[GG * 128 + r] = ln_load<GG, 11>(tlane + D_ML2, v);
        __syncthreads();
        const float2 st = ln_stats(red, r, 1.f / 88.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = GG + 4 * i;
          float hi[8], lo[8];
          if (c < 11) {
            float2 o2[4];
            ln_apply(v[i < NI ? i : 0], st, prm.n2w + 8 * (c < 11 ? c : 0), prm.n2b + 8 * (c < 11 ? c : 0), o2);
            const float o[8] = {xa[i].x + o2[0].x, xa[i].y + o2[0].y, xa[i].z + o2[1].x, xa[i].w + o2[1].y,
                                xb[i].x + o2[2].x, xb[i].y + o2[2].y, xb[i].z + o2[3].x, xb[i].w + o2[3].y};
#pragma unroll
            for (int k = 0; k < 8; ++k) split_hi_lo<BF16>(o[k], hi[k], lo[k]);
            if (ray_out != nullptr && row_ok) {
              float4* dst = reinterpret_cast<float4*>(ray_out + (size_t)prow * kDRay + 8 * c);
              dst[0] = make_float4(o[0], o[1], o[2], o[3]);
              dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) hi[k] = lo[k] = 0.f;
          }
          if (c < 12) {
            st_chunk<BF16>(smem + R_Q, r, c, hi);
            st_chunk<BF16>(smem + R_RLO, r, c, lo);
          }
        }
      };
      UFO_G_DISPATCH(ln2)
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
    // R14 (SRDF tail) of this tile is deferred: it runs under the QKV GEMM of the next tile
    have_prev = true;
    prev_prow = prow;
    prev_ok = row_ok;
    __syncthreads();   // scratch / Q' / V' regions are rewritten by the next tile's R2
  }
  if (have_prev) srdf_tail(prev_prow, prev_ok);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace ufo
