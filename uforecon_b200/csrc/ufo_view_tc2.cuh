// View stage, second generation: TWO TILES IN FLIGHT per CTA.
//
//   LoFTREncoderLayer   code1/attention/transformer.py:35-58
//   LinearAttention     code1/attention/linear_attention.py:20-47
//   radiance head       code1/ray_transformer.py:159-163,310-320
//
// k_view_tc (ufo_xfmr_tc.cuh) runs one 128-row tile per CTA in lock step: MMA -> mbarrier wait -> fp32 epilogue ->
// __syncthreads, seven times per tile; ncu showed the tensor pipe 17 % active and 39 % of the issue slots used.
// Here the 512 threads are two independent halves of 256 threads.  Each half runs the whole per-tile program on its
// own tile with its own 256 TMEM columns, its own mbarrier, its own issuing thread and a named barrier, so that one
// half's MMAs and barrier waits are covered by the other half's epilogue arithmetic.  What makes two tiles fit:
//   * every operand an epilogue produces (message, LayerNorm outputs, ReLU hidden layer, head side inputs) is written
//     by its row owner straight into TMEM (tcgen05.st) and consumed as the A operand of a TS-form tcgen05.mma - no
//     shared-memory staging, no fence.proxy.async, half the shared-memory bandwidth per MMA;
//   * the per-point attention exchanges K'/V' between the L = NV+1 token rows of a point inside one warp (points are
//     aligned to warps: floor(32/L) points per warp; to warp pairs with a 64-thread named barrier for L = 7, 9, 11, where a
//     single warp would leave up to 10 of 32 rows idle) through a 2.5 KB per-warp buffer, in fp32, one head at a time -
//     no 16-bit staging tile, no CTA-wide barrier between the elu phase and the attention (warp shuffles were
//     measured first: 2560 SHFL per tile, 19 % of the stall samples on the MIO queue);
//   * the weights (138 KB) stay resident and are shared by both halves; the only activation operand in shared memory
//     is the token tile X (A of the QKV GEMM, 20 KB), refilled by cp.async for the next tile as soon as the last MMA
//     that reads it (mlp.0) has completed.
// A thread owns one token row and one of TWO column groups (half of every accumulator row), so the fixed costs of a
// phase (barriers, LayerNorm statistics, addressing) are paid by 8 warps per tile instead of 16.
//
// TMEM columns of a half (256): QKV [0,240) as [g0: q40 k40 v40 | g1: q40 k40 v40], radiance-head accumulator
// [240,256).  Once QKV is consumed: message operand g0 [0,24) g1 [120,144) (40 values + 8 zeros = 3 K-steps each),
// merge accumulator [144,224), LN1 operand [0,24) [24,48), mlp.0 accumulator [80,240), hidden operand [0,80),
// mlp.2 accumulator [80,160), LN2 + direction/bias operand [0,24) [24,48).
#pragma once
#include "ufo_xfmr_tc.cuh"
#include <cstdio>

namespace ufo {
namespace tc {
// shared-memory map of k_view_tc2 (bytes); the weight block is also the layout of the global image
constexpr uint32_t V2_WQKV = 0;                              // [240][80]   rows: g0 (q40 k40 v40) | g1 (q40 k40 v40)
constexpr uint32_t V2_WMRG = V2_WQKV + 240 * 80 * 2;         // [80][96]    K: g0 40 | 0 x8 | g1 40 | 0 x8
constexpr uint32_t V2_WML0 = V2_WMRG + 80 * 96 * 2;          // [160][176]  K: x 80 | message, padded like merge (96)
constexpr uint32_t V2_WML2 = V2_WML0 + 160 * 176 * 2;        // [80][160]
constexpr uint32_t V2_WRAD = V2_WML2 + 80 * 160 * 2;         // [16][176]   K: x 80 | LN2 g0 40 | dir 3, 1, 0 x4 | LN2 g1 40 | 0 x8
constexpr uint32_t V2_WEND = V2_WRAD + 16 * 176 * 2;         // 141,312
constexpr uint32_t V2H_X = 0;                                // per half: token operand, 10 chunks
// Chunk stride (= LBO of the descriptor) of the token tile.  -DUFO_VIEW_TOK_CONTIG pads it by one 16-byte piece and fills the tile with
// consecutive lanes on consecutive 16-byte pieces of the tile's CONTIGUOUS token block in global memory (4-5 cache lines per cp.async
// instead of 24, bank-conflict-free writes thanks to the padding).  ncu counts 24 L1 wavefronts per cp.async in the default mapping
// (lanes on consecutive rows of one chunk, 160 bytes apart in global memory), 22 % of the kernel's shared-memory/L1 wavefronts - but
// the copies run a whole tile ahead and nothing waits for them: measured 230.7 ms per map against 229.4 ms for the default, not kept.
#ifdef UFO_VIEW_TOK_CONTIG
constexpr uint32_t V2_XLBO = kChunk + 16;
#else
constexpr uint32_t V2_XLBO = kChunk;
#endif
constexpr uint32_t V2H_XCH = 10 * V2_XLBO;                   // per warp [32 rows][K' 10 | V 10] fp32: attention exchange of one head
constexpr uint32_t V2H_RED = V2H_XCH + 8 * 2560;             // float2 [2][128] LayerNorm partials / float [2][128] head partials
constexpr uint32_t V2H_RGBM = V2H_RED + 2 * 128 * 8;         // float4 [112]: (r,g,b,mask) of the tile's (point, view) pairs
constexpr uint32_t V2H_SIZE = V2H_RGBM + 112 * 16;
constexpr uint32_t V2_HALF = V2_WEND;
constexpr uint32_t V2_BAR = V2_HALF + 2 * V2H_SIZE;
#ifdef UFO_PHASE_TIMING                                     // debug build: per-phase clock64 sums of one warp of half 0, printed by block 0
constexpr uint32_t V2_SMEM = V2_BAR + 64 + 256;
#else
constexpr uint32_t V2_SMEM = V2_BAR + 64;
#endif
static_assert(V2_SMEM <= 232448, "view-stage (v2) shared memory exceeds the 227 KB opt-in limit");

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Rows over which the L = NV+1 token rows of a point may spread: one warp (32 TMEM lanes) when that wastes < 5 % more rows than
// a warp pair, else two warps (64 rows) - L = 7, 9, 11 (NV = 6, 8, 10) leave 4, 5, 10 of 32 rows idle in a single warp, 1, 1, 9 of 64 in
// a pair.  The K'/V' exchange of the attention is then synchronised by a 64-thread named barrier instead of __syncwarp.
__host__ __device__ constexpr int view2_group_rows(int nv) {
  const int L = nv + 1;
  return ((64 / L) * L * 32 - (32 / L) * L * 64) * 20 > 64 * 32 ? 64 : 32;      // util64 - util32 > 0.05
}
__host__ __device__ constexpr int view2_points_per_tile(int nv) { return (128 / view2_group_rows(nv)) * (view2_group_rows(nv) / (nv + 1)); }

// K loop of an SS-form GEMM (A: the 128-row K-major token tile at a_base, chunk stride V2_XLBO; B: K-major weight tile with b_rows rows at b_base) with the descriptors
// advanced by one 32-bit add per K step (ufo_umma.cuh: desc_lo / desc_hi)
__device__ __forceinline__ void issue_gemm_lh(uint32_t tmem_d, uint32_t a_base, uint32_t b_base, uint32_t b_rows, uint32_t k_chunks,
                                              uint32_t idesc, uint32_t acc_first) {
  const uint32_t a0 = umma::desc_lo(a_base, V2_XLBO), b0 = umma::desc_lo(b_base, b_rows * 16u);
#pragma unroll
  for (uint32_t c = 0; c < k_chunks; c += 2)
    umma::mma_f16_lh(tmem_d, a0 + c * (V2_XLBO >> 4), umma::desc_hi(128u), b0 + c * b_rows, umma::desc_hi(128u), idesc, (c > 0) ? 1u : acc_first);
}
// K loop of a TS-form GEMM: A chunk pairs from the TMEM columns a_col(ks), B K-major weight tile (b_rows rows) from K step k0 on
template <typename ACol>
__device__ __forceinline__ void issue_ts_lh(uint32_t tmem_d, ACol a_col, uint32_t b_base, uint32_t b_rows, int k0, int ksteps, uint32_t idesc,
                                            uint32_t acc_first) {
  const uint32_t b0 = umma::desc_lo(b_base, b_rows * 16u);
#pragma unroll
  for (int ks = 0; ks < ksteps; ++ks)
    umma::mma_f16_ts_lh(tmem_d, a_col(ks), b0 + (uint32_t)(k0 + ks) * 2u * b_rows, umma::desc_hi(128u), idesc, ks > 0 ? 1u : acc_first);
}

#define UFO_G2_DISPATCH(fn)     \
  if (g == 0) fn(IC<0>{});      \
  else fn(IC<1>{});

// LayerNorm partial (sum, sum of squares) of 40 accumulator columns starting at tcol
__device__ __forceinline__ float2 ln40_load(uint32_t tcol, float2 (*v)[4]) {
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_ld8p(tcol + 8 * i, v[i]);
  umma::tmem_ld_wait();
  // four independent accumulation chains per statistic: with two warps per scheduler the 4-cycle FMA latency of a single
  // chain over 20 pairs would be exposed
  float2 s[4], q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s[k] = v[0][k];
    q2[k] = __fmul2_rn(v[0][k], v[0][k]);
  }
#pragma unroll
  for (int i = 1; i < 5; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] = __fadd2_rn(s[k], v[i][k]);
      q2[k] = __ffma2_rn(v[i][k], v[i][k], q2[k]);
    }
  const float2 st = __fadd2_rn(__fadd2_rn(s[0], s[1]), __fadd2_rn(s[2], s[3]));
  const float2 qt = __fadd2_rn(__fadd2_rn(q2[0], q2[1]), __fadd2_rn(q2[2], q2[3]));
  return make_float2(st.x + st.y, qt.x + qt.y);
}
__device__ __forceinline__ float2 ln2_stats(const float2* red, int r, float inv_n) {
  const float2 a0 = red[r], a1 = red[128 + r];
  const float mean = (a0.x + a1.x) * inv_n;
  const float var = fmaxf((a0.y + a1.y) * inv_n - mean * mean, 0.f);
  return make_float2(mean, rsqrtf(var + 1e-5f));
}
}  // namespace tc

template <int NV, bool BF16>
__global__ void __launch_bounds__(512, 1)
k_view_tc2(const uint8_t* __restrict__ wimg, const __grid_constant__ ViewParams prm, const uint16_t* __restrict__ tok,
           const float4* __restrict__ rgbm, const float4* __restrict__ dirs, int P, int half, float* __restrict__ vout0,
           float4* __restrict__ radiance) {
  using namespace tc;
  constexpr int L = NV + 1, GR = tc::view2_group_rows(NV), PPW = GR / L, PPT = (128 / GR) * PPW, RPW = PPW * L;   // rows of an exchange
                                                                        // group (warp or warp pair), points per group / tile, rows in use per group
  static_assert(PPT * NV <= 112, "colour staging too small");
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  constexpr uint32_t D_QKV = 0, D_RAD = 240, D_MRG = 144, D_ML0 = 80, D_ML2 = 80;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  // The warp index goes through a shuffle so that the compiler can prove it warp-uniform: the half (hf), the TMEM base and the
  // token-tile address then live in uniform registers, and the elected thread's tcgen05.mma operands (descriptors, TMEM addresses)
  // are built on the uniform datapath.  With hf = tid >> 8 as plain thread arithmetic every MMA was wrapped in an
  // ELECT / R2UR.BROADCAST / BRA.U.ANY loop: 19 dependent instructions per MMA in the issuing thread (-DUFO_NO_UNIFORM_ISSUE: old form).
#ifndef UFO_NO_UNIFORM_ISSUE
  const int tid = threadIdx.x, lane = tid & 31, warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int hf = warp_u >> 3, t = tid & 255, wl = warp_u & 7;
#else
  const int tid = threadIdx.x, hf = tid >> 8, t = tid & 255, lane = tid & 31, wl = t >> 5;
#endif
  const int q = wl & 3, g = wl >> 2, r = q * 32 + lane;
  // the thread of a half that issues its MMAs: one elected lane of the half's warp 0 (-DUFO_NO_ELECT_ISSUE: thread 0 of the half)
#if !defined(UFO_NO_ELECT_ISSUE) && !defined(UFO_NO_UNIFORM_ISSUE)
#define UFO_VIEW_ISSUER (wl == 0 && umma::elect_one())
#else
#define UFO_VIEW_ISSUER (t == 0)
#endif
  const int gr = r % GR;                                // row inside its exchange group
  const int pl = (r / GR) * PPW + gr / L, l = gr % L;
  const bool row_ok = gr < RPW;
  const int sl0 = (gr / L) * L;                         // first group row of this row's point
  uint8_t* const hs = smem + V2_HALF + hf * V2H_SIZE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + V2_BAR) + hf;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + V2_BAR + 32);
  float2* red = reinterpret_cast<float2*>(hs + V2H_RED);
  float* omg = reinterpret_cast<float*>(hs + V2H_RED);
  float4* s_rgbm = reinterpret_cast<float4*>(hs + V2H_RGBM);
  const uint32_t bar_id = 1 + hf;

  if (tid < 32) umma::tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    umma::mbar_init(reinterpret_cast<uint64_t*>(smem + V2_BAR), 1);
    umma::mbar_init(reinterpret_cast<uint64_t*>(smem + V2_BAR) + 1, 1);
    umma::fence_barrier_init();
  }
  for (uint32_t i = tid; i < V2_WEND / 16; i += 512)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
  // token operand buffers: row l == 0 of every point is the learnable view token (constant), unused rows are zero
  for (int i = tid; i < 2 * 1280; i += 512) {
    const int hb = i / 1280, j = i - hb * 1280;          // half, piece
    const int rr = j & 127, c = j >> 7;
    const int ln = rr % GR;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (ln < RPW && (ln % L) == 0) ? prm.vtok_x[c * 8 + k] : 0.f;   // token-row order (tok_pos)
    *reinterpret_cast<uint4*>(smem + V2_HALF + hb * V2H_SIZE + V2H_X + c * V2_XLBO + rr * 16) = pack8<BF16>(v);
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
#ifndef UFO_NO_UNIFORM_ISSUE
  const uint32_t tm = __shfl_sync(0xffffffffu, *tmem_slot, 0) + 256u * hf;
#else
  const uint32_t tm = *tmem_slot + 256u * hf;
#endif
  const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);
  const uint32_t sm_base = umma::smem_u32(smem);
  const uint32_t x_base = umma::smem_u32(hs + V2H_X);
  const uint32_t G = 120u * g;
  uint32_t ph = 0;
  const int n_tiles = (P + PPT - 1) / PPT;
  const int tstep = 2 * gridDim.x;
  auto slot_of = [&](int p) -> int { return (p >> 6) * kNS + half * kNC + (p & 63); };

  // token rows of a tile -> X by cp.async; rows of the view token stay as initialised
  auto load_tokens = [&](int tile) {
    const int pbase = tile * PPT;
#ifdef UFO_VIEW_TOK_CONTIG
    constexpr int PIECES = PPT * NV * 10;                  // 16-byte pieces of the tile's token block: point-major, then view, then chunk
#pragma unroll
    for (int k = 0; k < (PIECES + 255) / 256; ++k) {
      const int j = t + k * 256;
      const int pp = j / (NV * 10), w = j - pp * (NV * 10);
      const int v = w / 10, c = w - v * 10;
      const int p = pbase + pp;
      const int rr = (pp / PPW) * GR + (pp % PPW) * L + v + 1;
      if (j < PIECES && p < P)
        cp_async16(x_base + c * V2_XLBO + rr * 16, tok + (size_t)slot_of(p) * (NV * kDView) + w * 8);
    }
#else
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int i = t + k * 256;
      const int rr = i & 127, c = i >> 7;
      const int ln = rr % GR;
      const int ll = ln % L;
      const int p = pbase + (rr / GR) * PPW + ln / L;
      if (ln < RPW && ll > 0 && p < P)
        cp_async16(x_base + c * kChunk + rr * 16,
                   tok + (size_t)slot_of(p) * (NV * kDView) + (ll - 1) * kDView + c * 8);
    }
#endif
    cp_async_commit();
  };
  // wait for this half's last commit: with UFO_VIEW_POLL1 one warp polls the mbarrier and releases the others through the
  // named barrier (the other seven warps then block without using issue slots)
  auto half_wait = [&]() {
#ifdef UFO_VIEW_POLL1
    if (wl == 0) umma::mbar_wait(bar, ph);
    umma::bar_sync(bar_id, 256);
#else
    umma::mbar_wait(bar, ph);
#endif
    ph ^= 1;
    umma::tc_fence_after();
  };
  int tile = 2 * blockIdx.x + hf;
  if (tile < n_tiles) load_tokens(tile);
  // exchange buffer of this (column group, row group): [GR rows][5] float4; same 20 KB per half for either group size
  float4* const xw = reinterpret_cast<float4*>(hs + V2H_XCH + (g * (128 / GR) + r / GR) * (GR * 80));
  const uint32_t xbar_id = 3 + hf * 4 + g * 2 + (q >> 1);                    // named barrier of a warp pair (GR == 64), ids 3..10
  auto group_sync = [&]() {
    if (GR == 32) __syncwarp();
    else umma::bar_sync(xbar_id, 64);
  };

  // masked softmax over the views and colour blend of one tile (ray_transformer.py:316-320): one thread per point, warp 0 of the half.
  // It runs DEFERRED, under the next tile's QKV GEMM (or after the loop for the last tile): at the end of its own tile it was serial work
  // of one warp that the other seven waited for at the next barrier (-DUFO_VIEW_BLEND_INLINE: the old place).
  auto blend = [&](int pb) {
    if (t < PPT && pb + t < P) {
      const size_t p = (size_t)slot_of(pb + t);
      const int rr0 = (t / PPW) * GR + (t % PPW) * L + 1;        // row of (point t, view 0)
      float om[NV];
      float4 col[NV];
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        col[n] = s_rgbm[t * NV + n];
        const float w = prm.rb4 + (omg[rr0 + n] + omg[128 + rr0 + n]);
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
  int prev_pbase = -1;
#ifdef UFO_PHASE_TIMING
  unsigned long long* tim = reinterpret_cast<unsigned long long*>(smem + V2_BAR + 64);
  if (tid == 32) for (int i = 0; i < 32; ++i) tim[i] = 0;
  long long tim_prev = clock64();
  int tim_idx = 0, tim_tiles = 0;
#define UFO_TIM() do { if (tid == 32) { const long long now_ = clock64(); tim[tim_idx] += (unsigned long long)(now_ - tim_prev); tim_prev = now_; } ++tim_idx; } while (0)
#else
#define UFO_TIM() do { } while (0)
#endif
  for (; tile < n_tiles; tile += tstep) {
#ifdef UFO_PHASE_TIMING
    tim_idx = 0;
    ++tim_tiles;
#endif
    const int pbase = tile * PPT;
    const int my_p = pbase + pl;
    const size_t my_slot = (size_t)slot_of(my_p);
    const bool live = row_ok && my_p < P;
    const uint32_t xb = x_base;
    // ---- P0: this tile's token rows have landed
    cp_async_wait_all();
    umma::fence_async_smem();
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    // ---- P1: q|k|v = X . Wqkv^T (transformer.py:47) and the x part of the radiance head's first layer
    if (UFO_VIEW_ISSUER) {
      umma::tc_fence_after();
      issue_gemm_lh(tm + D_QKV, xb, sm_base + V2_WQKV, 240, 10, umma::make_idesc(128, 240, FMT, false, false), 0);
      umma::commit(bar);
      issue_gemm_lh(tm + D_RAD, xb, sm_base + V2_WRAD, 16, 10, umma::make_idesc(128, 16, FMT, false, false), 0);
    }
    float3 my_dir = make_float3(0.f, 0.f, 0.f);
    if (live && l > 0 && g == 0) {
      const float* dp = reinterpret_cast<const float*>(dirs + my_slot * NV + (l - 1));
      const float2 dxy = __ldg(reinterpret_cast<const float2*>(dp));
      my_dir.x = dxy.x;
      my_dir.y = dxy.y;
      my_dir.z = __ldg(dp + 2);
    }
    float4 my_col = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < PPT * NV) {
      const int pp = pbase + t / NV;
      if (pp < P) my_col = __ldg(rgbm + (size_t)slot_of(pp) * NV + (t % NV));
    }
#ifndef UFO_VIEW_BLEND_INLINE
    if (prev_pbase >= 0) blend(prev_pbase);                      // the previous tile's colours, under this tile's QKV GEMM
#endif
    half_wait();
    UFO_TIM();
    // ---- P2+P3: elu+1 on q, k; msg_l = sum_s (Q_l.K_s) V_s / (sum_s Q_l.K_s + 1e-6) per head over the L token rows of
    //      the point (== Q (K^T V) Z, linear_attention.py:36-45), K'/V' of the other rows by warp shuffle.
    //      Thread (row, g) owns heads 4g .. 4g+3 = columns [120g, 120g+120) of the accumulator.
    //      The view-token row (s = 0) of every point is the same constant input, so its K' and V are constants of the weights: they come
    //      from the parameter block (ViewParams::k0 / v0, computed on the host from the same 16-bit operands), not through the exchange
    //      buffer - a quarter of the phase's LDS.128 at NV = 3.  The phase is bound by the shared-memory pipe (58 % busy over the whole
    //      kernel, 6400 of the 14 900 cycles of a tile by per-phase clocks).  -DUFO_VIEW_TOK0_SMEM: every row through the buffer.
    {
      auto attn = [&](auto GGc) {
      constexpr int GG = decltype(GGc)::value;
      uint32_t mo[24];
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        float qf[20], kf[20], vf[20];
        const uint32_t c0 = tl + D_QKV + G + 20 * hp;
        umma::tmem_ld16(c0, qf);
        tmem_ld4(c0 + 16, qf + 16);
        umma::tmem_ld16(c0 + 40, kf);
        tmem_ld4(c0 + 56, kf + 16);
        umma::tmem_ld16(c0 + 80, vf);
        tmem_ld4(c0 + 96, vf + 16);
        umma::tmem_ld_wait();
        float2 q2[10], k2[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          q2[i] = elu1_2(make_float2(qf[2 * i], qf[2 * i + 1]));
          k2[i] = elu1_2(make_float2(kf[2 * i], kf[2 * i + 1]));
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          // this row's K' | V of the head -> the warp's exchange buffer (80 B rows: conflict-free 16-byte stores)
          group_sync();                                         // the previous head's readers are done
          xw[gr * 5 + 0] = make_float4(k2[5 * hh].x, k2[5 * hh].y, k2[5 * hh + 1].x, k2[5 * hh + 1].y);
          xw[gr * 5 + 1] = make_float4(k2[5 * hh + 2].x, k2[5 * hh + 2].y, k2[5 * hh + 3].x, k2[5 * hh + 3].y);
          xw[gr * 5 + 2] = make_float4(k2[5 * hh + 4].x, k2[5 * hh + 4].y, vf[10 * hh], vf[10 * hh + 1]);
          xw[gr * 5 + 3] = make_float4(vf[10 * hh + 2], vf[10 * hh + 3], vf[10 * hh + 4], vf[10 * hh + 5]);
          xw[gr * 5 + 4] = make_float4(vf[10 * hh + 6], vf[10 * hh + 7], vf[10 * hh + 8], vf[10 * hh + 9]);
          group_sync();
          float2 msg[5];
#pragma unroll
          for (int b = 0; b < 5; ++b) msg[b] = make_float2(0.f, 0.f);
          float den = 0.f;
#pragma unroll
          for (int s = 0; s < L; ++s) {
            float4 a0, a1, a2, a3, a4;
#ifndef UFO_VIEW_TOK0_SMEM
            if (s == 0) {                                          // the view-token row: constants
              const float* kc = prm.k0 + 40 * GG + 20 * hp + 10 * hh;
              const float* vc = prm.v0 + 40 * GG + 20 * hp + 10 * hh;
              a0 = make_float4(kc[0], kc[1], kc[2], kc[3]);
              a1 = make_float4(kc[4], kc[5], kc[6], kc[7]);
              a2 = make_float4(kc[8], kc[9], vc[0], vc[1]);
              a3 = make_float4(vc[2], vc[3], vc[4], vc[5]);
              a4 = make_float4(vc[6], vc[7], vc[8], vc[9]);
            } else
#endif
            {
              const float4* src = xw + ((sl0 + s) & (GR - 1)) * 5;   // the same address for the L rows of a point: broadcast
              a0 = src[0]; a1 = src[1]; a2 = src[2]; a3 = src[3]; a4 = src[4];
            }
            float2 acc = __fmul2_rn(q2[5 * hh], make_float2(a0.x, a0.y));
            acc = __ffma2_rn(q2[5 * hh + 1], make_float2(a0.z, a0.w), acc);
            acc = __ffma2_rn(q2[5 * hh + 2], make_float2(a1.x, a1.y), acc);
            acc = __ffma2_rn(q2[5 * hh + 3], make_float2(a1.z, a1.w), acc);
            acc = __ffma2_rn(q2[5 * hh + 4], make_float2(a2.x, a2.y), acc);
            const float sc = acc.x + acc.y;
            den += sc;
            const float2 sc2 = make_float2(sc, sc);
            msg[0] = __ffma2_rn(sc2, make_float2(a2.z, a2.w), msg[0]);
            msg[1] = __ffma2_rn(sc2, make_float2(a3.x, a3.y), msg[1]);
            msg[2] = __ffma2_rn(sc2, make_float2(a3.z, a3.w), msg[2]);
            msg[3] = __ffma2_rn(sc2, make_float2(a4.x, a4.y), msg[3]);
            msg[4] = __ffma2_rn(sc2, make_float2(a4.z, a4.w), msg[4]);
          }
          const float zi = 1.f / (den + 1e-6f);
          const float2 z2 = make_float2(zi, zi);
#pragma unroll
          for (int b = 0; b < 5; ++b) mo[10 * hp + 5 * hh + b] = pack2v<BF16>(__fmul2_rn(msg[b], z2));
        }
      }
      mo[20] = mo[21] = mo[22] = mo[23] = 0u;
      // message operand of this group: its own (consumed) q columns [120g, 120g+24)
      umma::tmem_st8(tl + G, mo);
      umma::tmem_st8(tl + G + 8, mo + 8);
      umma::tmem_st8(tl + G + 16, mo + 16);
      umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(attn)
    }
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    // ---- P4: merge (transformer.py:55), A from TMEM
    if (UFO_VIEW_ISSUER) {
      umma::tc_fence_after();
      issue_ts_lh(tm + D_MRG, [&](int ks) { return tm + 120 * (ks / 3) + 8 * (ks % 3); }, sm_base + V2_WMRG, 80, 0, 6,
                  umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    if (t < PPT * NV) s_rgbm[t] = my_col;
    half_wait();
    UFO_TIM();
    // ---- P5: LayerNorm 1 (transformer.py:56) -> message half of the mlp.0 operand, columns [24g, 24g+24)
    {
      auto ln1 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        float2 v[5][4];
        red[GG * 128 + r] = ln40_load(tl + D_MRG + 40 * GG, v);
        umma::tc_fence_before();
        umma::bar_sync(bar_id, 256);
        if (GG == 0 && UFO_VIEW_ISSUER) {   // the merge accumulator is consumed: the x half of mlp.0 runs under the rest of this phase
          umma::tc_fence_after();
          issue_gemm_lh(tm + D_ML0, xb, sm_base + V2_WML0, 160, 10, umma::make_idesc(128, 160, FMT, false, false), 0);
        }
        const float2 st = ln2_stats(red, r, 1.f / 80.f);
        uint32_t o[24];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          float2 y[4];
          ln_apply(v[i], st, prm.n1w + 40 * GG + 8 * i, prm.n1b + 40 * GG + 8 * i, y);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[4 * i + k] = pack2v<BF16>(y[k]);
        }
        o[20] = o[21] = o[22] = o[23] = 0u;
        umma::tmem_st8(tl + 24 * GG, o);
        umma::tmem_st8(tl + 24 * GG + 8, o + 8);
        umma::tmem_st8(tl + 24 * GG + 16, o + 16);
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(ln1)
    }
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    // ---- P6: mlp.0 on [x | LN1]  (transformer.py:57): message half from TMEM on top of the x half issued inside P5
    if (UFO_VIEW_ISSUER) {
      umma::tc_fence_after();
      issue_ts_lh(tm + D_ML0, [&](int ks) { return tm + 24 * (ks / 3) + 8 * (ks % 3); }, sm_base + V2_WML0, 160, 5, 6,
                  umma::make_idesc(128, 160, FMT, false, false), 1u);
      umma::commit(bar);
    }
    half_wait();
    UFO_TIM();
    if (tile + tstep < n_tiles) load_tokens(tile + tstep);     // X is free: mlp.0 was the last MMA reading it
    // ---- P7: ReLU -> hidden operand [40g, 40g+40)
    {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float2 v[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i) tmem_ld8p(tl + D_ML0 + 80 * g + 40 * b + 8 * i, v[i]);
        umma::tmem_ld_wait();
        uint32_t o[20];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) o[4 * i + k] = relu_pack2<BF16>(v[i][k]);
        umma::tmem_st8(tl + 40 * g + 20 * b, o);
        umma::tmem_st8(tl + 40 * g + 20 * b + 8, o + 8);
        umma::tmem_st4(tl + 40 * g + 20 * b + 16, o[16], o[17], o[18], o[19]);
      }
      umma::tmem_st_wait();
    }
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    // ---- P8: mlp.2
    if (UFO_VIEW_ISSUER) {
      umma::tc_fence_after();
      issue_ts_lh(tm + D_ML2, [&](int ks) { return tm + 8 * ks; }, sm_base + V2_WML2, 80, 0, 10, umma::make_idesc(128, 80, FMT, false, false), 0);
      umma::commit(bar);
    }
    half_wait();
    UFO_TIM();
    // ---- P9: LayerNorm 2; token 0: out = view_token + LN2 -> vout0 (fp32); view rows: LN2 (+ direction, 1) -> operand of the
    //      radiance head (x + LN2 is applied inside the head's GEMM: W0x.x + W0x.LN2)
    {
      auto ln2 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        float2 v[5][4];
        red[GG * 128 + r] = ln40_load(tl + D_ML2 + 40 * GG, v);
        umma::bar_sync(bar_id, 256);
        const float2 st = ln2_stats(red, r, 1.f / 80.f);
        const bool tok0 = live && l == 0;
        uint32_t o[24];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          float2 y[4];
          const int c = 40 * GG + 8 * i;
          ln_apply(v[i], st, prm.n2w + c, prm.n2b + c, y);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[4 * i + k] = pack2v<BF16>(y[k]);
          if (tok0) {
            float4* dst = reinterpret_cast<float4*>(vout0 + my_slot * kDView + c);
            dst[0] = make_float4(prm.vtok[c] + y[0].x, prm.vtok[c + 1] + y[0].y, prm.vtok[c + 2] + y[1].x, prm.vtok[c + 3] + y[1].y);
            dst[1] = make_float4(prm.vtok[c + 4] + y[2].x, prm.vtok[c + 5] + y[2].y, prm.vtok[c + 6] + y[3].x, prm.vtok[c + 7] + y[3].y);
          }
        }
        if (GG == 0) {        // side inputs of the head: relative direction and the constant 1 that carries the bias
          o[20] = umma::pack2<BF16>(my_dir.x, my_dir.y);
          o[21] = umma::pack2<BF16>(my_dir.z, 1.f);
          o[22] = o[23] = 0u;
        } else {
          o[20] = o[21] = o[22] = o[23] = 0u;
        }
        umma::tmem_st8(tl + 24 * GG, o);
        umma::tmem_st8(tl + 24 * GG + 8, o + 8);
        umma::tmem_st8(tl + 24 * GG + 16, o + 16);
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(ln2)
    }
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    // ---- P10: LN2 / direction / bias part of the radiance head's first layer   (ray_transformer.py:159-163,313)
    if (UFO_VIEW_ISSUER) {
      umma::tc_fence_after();
      issue_ts_lh(tm + D_RAD, [&](int ks) { return tm + 24 * (ks / 3) + 8 * (ks % 3); }, sm_base + V2_WRAD, 16, 5, 6,
                  umma::make_idesc(128, 16, FMT, false, false), 1u);
      umma::commit(bar);
    }
    half_wait();
    UFO_TIM();
    // ---- P11: head tail 16 -> 8 -> 1 (hidden units 4g .. 4g+3 per thread), masked softmax over views, colour blend
    {
      float h[16];
      umma::tmem_ld16(tl + D_RAD, h);
      umma::tmem_ld_wait();
      auto tail = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
#pragma unroll
        for (int o = 0; o < 16; ++o) h[o] = fmaxf(h[o], 0.f);      // bias and direction terms came through the GEMM
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = 4 * GG + j;
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
      UFO_G2_DISPATCH(tail)
    }
#ifdef UFO_VIEW_BLEND_INLINE
    umma::tc_fence_before();
    umma::bar_sync(bar_id, 256);
    UFO_TIM();
    blend(pbase);
#else
    prev_pbase = pbase;         // blended after the next P0 barrier, which also publishes omg
#endif
    // omg / s_rgbm are next written after several more barriers of this half; the next tile's QKV MMA overwrites TMEM only
    // after its P0 barrier, which every thread reaches after its last TMEM read above
  }
#ifdef UFO_PHASE_TIMING
  if (blockIdx.x == 0 && tid == 32 && n_tiles >= 4096) {
    printf("VIEWTIM NV=%d tiles=%d :", NV, tim_tiles);
    for (int i = 0; i < 12; ++i) printf(" %llu", tim[i] / (unsigned long long)tim_tiles);
    printf("\n");
  }
#endif
#undef UFO_TIM
  cp_async_wait_all();
#ifndef UFO_VIEW_BLEND_INLINE
  if (prev_pbase >= 0) {                                         // the last tile of this half
    umma::bar_sync(bar_id, 256);
    blend(prev_pbase);
  }
#endif
  umma::tc_fence_before();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(*tmem_slot, 512);
#undef UFO_VIEW_ISSUER
}

}  // namespace ufo
