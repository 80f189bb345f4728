// Ray stage, second generation: TWO CTAs PER SM.
//
//   LoFTREncoderLayer   code1/attention/transformer.py:35-58         (d = 88, tokens = the samples of one ray)
//   LinearAttention     code1/attention/linear_attention.py:20-47
//   order encoding      code1/ray_transformer.py:165-173,301-305
//   DensityMLP          code1/ray_transformer.py:147-150,307
//
// k_ray_tc (ufo_xfmr_tc.cuh) is one 512-thread CTA per SM with 222 KB of shared memory, 14 lock-step phases per
// 128-token tile; ncu showed 24 % of the issue slots used and the tensor pipe 15 % active - nothing runs while a CTA
// waits for an MMA or a weight slot.  This kernel is built to be resident TWICE per SM (256 threads, <= 113 KB of shared
// memory, 256 TMEM columns), so that the hardware interleaves two independent tiles:
//   * every A operand an epilogue produces (x, Q', message, LayerNorm outputs, hidden layer, the split SRDF-head input)
//     is written by its row owner into TMEM (tcgen05.st) and consumed by TS-form MMAs; shared memory holds only what must
//     be there: the MN-major K'/V' tiles of the K^T.V product (later the block-diagonal KV operand) and the weights;
//   * the 184 KB of weights stream through two 33 KB slots in seven pieces per tile (cp.async.bulk + mbarrier), each load
//     issued as soon as the MMA that last read its slot has completed;
//   * a thread owns one token row and one of two column halves.
// TMEM columns (256): x [208,256) | k,v accumulator [0,176) | q accumulator [0,96) | Q' [96,144) | K^T.V per sequence
// [0,96) [144,240) | message accumulator same | message operand [96,144) | merge accumulator [0,96) | [LN1 | x] operand
// [168,256) (one sequence per tile: [164,252), x still in place from R0) | mlp.0 accumulator, two passes [0,96) [0,80) | hidden operand [96,184) | mlp.2 accumulator [0,96) |
// r_hi [96,144) r_lo [144,192) | SRDF-head accumulator [0,32).
#pragma once
#include "ufo_view_tc2.cuh"
#include <cstdio>

namespace ufo {
namespace tc {
// global image of the ray-stage weights for k_ray_tc2: seven pieces, each the exact shared-memory operand image
constexpr uint32_t R2W_KV = 0;                             // [176][96]   rows: k 0..87 | v 88..175
constexpr uint32_t R2W_Q = R2W_KV + 176 * 96 * 2;          // [96][96]    rows 88..95 zero
constexpr uint32_t R2W_MRG = R2W_Q + 96 * 96 * 2;          // [96][96]
constexpr uint32_t R2W_ML0A = R2W_MRG + 96 * 96 * 2;       // [96][176]   rows 0..95 of mlp.0
constexpr uint32_t R2W_ML0B = R2W_ML0A + 96 * 176 * 2;     // [80][176]   rows 96..175
constexpr uint32_t R2W_ML2 = R2W_ML0B + 80 * 176 * 2;      // [96][176]   rows 88..95 zero
constexpr uint32_t R2W_DEN = R2W_ML2 + 96 * 176 * 2;       // [32][96] hi, [32][96] lo
constexpr uint32_t R2W_END = R2W_DEN + 2 * 32 * 96 * 2;    // 178,688
// shared-memory map (bytes)
constexpr uint32_t R2_K = 0;                               // K' 11 chunks (MN-major B of the K^T.V product); later KV block-diagonal of sequence 0
constexpr uint32_t R2_V = R2_K + 11 * kChunk;              // V' 12 chunks, chunk 11 = ones block; later sequence 1; the M = 128 read of the
                                                           // K^T.V product runs 8 KB past it into slot 0 (finite weights, rows never used)
constexpr uint32_t R2_XROW = 336;                           // stash of the fp32 input rows: 320 bytes + 16 of padding (conflict-free 16-byte reads)
constexpr uint32_t R2_SCR = R2_K + 128 * R2_XROW;          // LayerNorm / SRDF partials: inside the V' tile (chunk 10), behind the KV operand and the stash
static_assert(R2_SCR >= R2_V + 96 * 96 * 2 && R2_SCR + 2 * 128 * 8 <= R2_V + 11 * kChunk, "scratch must lie between the KV operand and the ones block of V'");
constexpr uint32_t R2_SLOT = R2_V + 12 * kChunk;           // two weight slots
constexpr uint32_t R2_SLOT_BYTES = 176 * 96 * 2;           // 33,792: the largest piece
constexpr uint32_t R2_BAR = R2_SLOT + 2 * R2_SLOT_BYTES;
#ifdef UFO_PHASE_TIMING                                     // debug build: per-phase clock64 sums of one warp, printed by block 0
constexpr uint32_t R2_SMEM = R2_BAR + 64 + 256;
#else
constexpr uint32_t R2_SMEM = R2_BAR + 64;
#endif
static_assert(R2_SMEM <= 115712, "ray-stage (v2) shared memory: two CTAs must fit one SM");
static_assert(R2_SCR + 2 * 128 * 8 <= R2_SLOT, "scratch overlaps the weight slots");
}  // namespace tc

namespace tc {
// LayerNorm partial (sum, sum of squares) over NC chunks of 8 values, four independent accumulation chains per statistic
template <int NC>
__device__ __forceinline__ float2 ln_partial(const float2 (*v)[4]) {
  float2 s[4], q2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s[k] = v[0][k];
    q2[k] = __fmul2_rn(v[0][k], v[0][k]);
  }
#pragma unroll
  for (int i = 1; i < NC; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] = __fadd2_rn(s[k], v[i][k]);
      q2[k] = __ffma2_rn(v[i][k], v[i][k], q2[k]);
    }
  const float2 st = __fadd2_rn(__fadd2_rn(s[0], s[1]), __fadd2_rn(s[2], s[3]));
  const float2 qt = __fadd2_rn(__fadd2_rn(q2[0], q2[1]), __fadd2_rn(q2[2], q2[3]));
  return make_float2(st.x + st.y, qt.x + qt.y);
}
}  // namespace tc

// SN = 128: one ray per tile; SN = 64: two rays per tile.
template <int SN, bool BF16>
__global__ void __launch_bounds__(256, 2)
k_ray_tc2(const uint8_t* __restrict__ wimg, const __grid_constant__ RayParams prm, const float* __restrict__ vout0,
          const float* __restrict__ pe_table, const uint8_t* __restrict__ perm, long long P, float* __restrict__ srdf,
          float* __restrict__ ray_out) {
  using namespace tc;
  constexpr uint32_t FMT = BF16 ? umma::kFmtBF16 : umma::kFmtF16;
  constexpr int NSEQ = 128 / SN;
  // [LN1 | x] operand of mlp.0: one sequence per tile (SN = 128) leaves the x operand of R0 untouched until mlp.0, so LN1 is
  // written right below it and x is used in place; with two sequences the second K^T.V accumulator overwrites x and it is re-staged
  constexpr uint32_t C_X = 208, C_QP = 96, C_M = 96, C_XL = (NSEQ == 1 ? 164 : 168), C_H1 = 96, C_RHI = 96, C_RLO = 144;      // operands
  constexpr uint32_t D_KV = 0, D_Q = 0, D_S0 = 0, D_S1 = 144, D_MRG = 0, D_ML0 = 0, D_ML2 = 0, D_DEN = 0;  // accumulators
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + R2_BAR);          // MMA completion
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + R2_BAR + 8);     // [2] weight slot filled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + R2_BAR + 32);
  float2* red = reinterpret_cast<float2*>(smem + R2_SCR);
  float* part_s = reinterpret_cast<float*>(smem + R2_SCR);
  // warp index through a shuffle: provably warp-uniform, so that TMEM addresses and the column-group dispatch run on the uniform datapath
  const int t = threadIdx.x, lane = t & 31, wl = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int q = wl & 3, g = wl >> 2, r = q * 32 + lane;
  // the thread that issues the MMAs / the weight copies: one elected lane of warp 0 / warp 4 (-DUFO_NO_ELECT_ISSUE: threads 0 / 128)
#ifndef UFO_NO_ELECT_ISSUE
#define UFO_RAY_ISSUER (wl == 0 && umma::elect_one())
#define UFO_RAY_ISSUE_WARP (wl == 0)          // GEMMs that wait for a weight piece: the whole warp waits, then one lane is elected
#define UFO_RAY_ELECT (umma::elect_one())
#define UFO_RAY_LOADER (wl == kLoadThread / 32 && umma::elect_one())
#else
#define UFO_RAY_ISSUER (t == 0)
#define UFO_RAY_ISSUE_WARP (t == 0)
#define UFO_RAY_ELECT (true)
#define UFO_RAY_LOADER (t == kLoadThread)
#endif
  const long long n_tiles = (P + 127) / 128;

  if (wl == 0) umma::tmem_alloc(tmem_slot, 256);
  if (t == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_init(full, 1);
    umma::mbar_init(full + 1, 1);
    umma::mbar_init(reinterpret_cast<uint64_t*>(smem + R2_BAR + 24), 1);
    umma::fence_barrier_init();
  }
  // the ones block of V' (rows 88..95 of the K^T.V product = sum_s K'_s) never changes
  if (g == 0) {
    float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    st_chunk<BF16>(smem + R2_V, r, 11, one);
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);
  const uint32_t sm_base = umma::smem_u32(smem);
  uint32_t ph = 0;

  // ---- weight streaming: piece j of the running sequence goes to slot (pc & 1); the copies and the waits are thread 0's, the
  //      BOOKKEEPING (piece counters, mbarrier phases) is kept by every thread at uniform program points: slot addresses and barrier
  //      parities are then warp-uniform values and the issuing thread's descriptors are built on the uniform datapath.  (Kept by thread 0
  //      alone they were thread-private data: every tcgen05.mma / cp.async.bulk was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop,
  //      14 dependent instructions per MMA in the one thread the other 255 wait for.)
#ifdef UFO_RAY_LOAD_T0
  constexpr int kLoadThread = 0;
#else
  constexpr int kLoadThread = 128;
#endif
  uint32_t pc_load = 0, pc_use = 0, fph = 0;                  // fph bit s: phase parity of slot s
  auto load_piece = [&](int j) {
    const uint32_t off[7] = {R2W_KV, R2W_Q, R2W_MRG, R2W_ML0A, R2W_ML0B, R2W_ML2, R2W_DEN};
    const uint32_t len[7] = {176 * 96 * 2, 96 * 96 * 2, 96 * 96 * 2, 96 * 176 * 2, 80 * 176 * 2, 96 * 176 * 2, 2 * 32 * 96 * 2};
    const uint32_t s = pc_load & 1;
    ++pc_load;
    // the copies are issued by a thread of warp 4: the g = 1 warps own 5 of the 11 chunks of a row, so they reach the end of an
    // epilogue first, and warp 0 (which issues the MMAs) starts its epilogue without this detour (-DUFO_RAY_LOAD_T0: thread 0)
    if (UFO_RAY_LOADER) bulk_load(smem + R2_SLOT + s * R2_SLOT_BYTES, wimg + off[j], len[j], full + s);
  };
  // every thread: slot and parity of the piece about to be used; thread 0 waits until it has landed (use_wait) before it issues
  uint32_t use_s = 0, use_par = 0;
  auto use_piece = [&]() -> uint32_t {
    use_s = pc_use & 1;
    use_par = (fph >> use_s) & 1u;
    fph ^= 1u << use_s;
    ++pc_use;
    return sm_base + R2_SLOT + use_s * R2_SLOT_BYTES;
  };
  auto use_wait = [&]() { umma::mbar_wait(full + use_s, use_par); };
  // -DUFO_RAY_PREWAIT: the issuer waits for the piece of the NEXT GEMM right after the commit of the current one (while its MMAs run)
  // instead of in front of the next issue.  Measured on a B200: no gain (ray stage 167.9 vs 166.3 ms per map without it) - the wait on a
  // completed phase is cheap next to the phase itself, and the extra wait delays the issuing warp's own epilogue.  Not the default.
  auto commit_and_prewait = [&](bool next_exists) {
    umma::commit(bar);
#ifdef UFO_RAY_PREWAIT
    if (next_exists) {
      const uint32_t ns = pc_use & 1;
      umma::mbar_wait(full + ns, (fph >> ns) & 1u);
    }
#endif
  };
  bool first_gemm = true;
  // the two warps that share the token rows 32 q .. 32 q + 31 (column halves g = 0, 1): a 64-thread named barrier where only they
  // exchange data (LayerNorm / SRDF partials); -DUFO_RAY_CTA_SYNC restores the CTA-wide barrier
  auto pair_sync = [&]() {
#ifdef UFO_RAY_CTA_SYNC
    __syncthreads();
#else
    umma::bar_sync(1 + q, 64);
#endif
  };
  auto mma_wait = [&]() {
    umma::mbar_wait(bar, ph);
    ph ^= 1;
    umma::tc_fence_after();
  };
  // TS-form K loop: A chunks from TMEM column a_col (8 columns per K step), B from a K-major weight tile with b_rows rows
  auto issue_ts = [&](uint32_t d_col, uint32_t a_col, uint32_t b_addr, uint32_t b_rows, uint32_t n, int ksteps, uint32_t acc_first) {
    const uint32_t idesc = umma::make_idesc(128, n, FMT, false, false);
    const uint32_t lbo = b_rows * 16u;
    const uint32_t lo0 = umma::desc_lo(b_addr, lbo);             // one K step = two chunks = 2 * lbo bytes = 2 * b_rows descriptor units
    for (int ks = 0; ks < ksteps; ++ks)
      umma::mma_f16_ts_lh(tm + d_col, tm + a_col + 8 * ks, lo0 + (uint32_t)ks * 2u * b_rows, umma::desc_hi(128u), idesc, ks > 0 ? 1u : acc_first);
  };

  if ((long long)blockIdx.x < n_tiles) {
    load_piece(0);
    load_piece(1);
  }
  // Input row of this thread's token in a tile.  With sorted samples it goes through the sort permutation.  The byte load is issued by
  // inline asm ONE TILE AHEAD (perm_load, at the top of the previous tile) and its first use is pinned to the place where the row is
  // needed (in_row_from, in front of the x loads of R13) by an empty asm: as plain C++ the compiler sank the load next to its first use
  // (10 % of the stall samples on the perm -> address -> x chain), and with only the load pinned it put the address add right behind
  // the load (3 % of the samples at the top of every tile).  Rows past P read the last valid byte (branch-free) and are dropped later.
  auto perm_load = [&](long long tile) -> unsigned {
    unsigned pv = 0;
    if (SN != kNC && perm != nullptr) {                      // uniform
      const long long prow = tile * 128 + r;
      asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(pv) : "l"(perm + (prow < P ? prow : P - 1)));
    }
    return pv;
  };
  auto in_row_from = [&](long long tile, unsigned pv) -> long long {
    const long long prow = tile * 128 + r;
    if (SN == kNC) return prow < P ? tc_slot(prow, 0) : -1;
    if (perm == nullptr) return prow < P ? prow : -1;
    asm volatile("" : "+r"(pv));
    return prow < P ? tile * 128 + (long long)pv : -1;
  };
  // This thread's half of the fp32 input row, chunks 6 g .. 6 g + 5: view-stage output (chunks 0..9), order encoding (chunk 10,
  // ray_transformer.py:301-303), zero (chunk 11).  A tile needs it twice (three times with two sequences): as the 16-bit operand of
  // R0 (and of mlp.0 when x was overwritten in TMEM) and in fp32 as the residual of R12.  The 128 rows of a tile (320 bytes each) go
  // global -> SHARED memory as 128 bulk copies (cp.async.bulk, one per row, issued by the row's g = 0 thread, completing on the mbarrier
  // xbar) and are read from there (x_get).  History, measured with per-phase clocks (profiles/r02_phase_clocks_*.txt):
  //   * register loads issued "under" an MMA wait: the warp cannot leave the wait before its loads have landed (shared scoreboards) -
  //     those two waits took 2180 and 2990 cycles where the others take 850-1050;
  //   * cp.async (LDGSTS), 12 per thread: no scoreboard, but 3072 LSU operations per pass and CTA at ~8 cycles each, and the next
  //     barrier drains them - the phases that issued them took 3270 and 2520 cycles (ray stage 166 -> 151 ms all the same);
  //   * bulk copies run on the copy engine: 128 per pass, nothing for the LSU to drain.
  // The stash is the K' / V' operand area, dead between the message GEMM (R5) and the next tile's R2a: row r at R2_K + r * 336.
  uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + R2_BAR + 24);
  uint32_t xph = 0;                                                          // parity of the next stash completion (every thread)
  auto x_copy = [&](long long ir, long long tile_of_rows) {
    const long long left = P - tile_of_rows * 128;                           // rows of that tile inside P: the prefix r < left
    const uint32_t bytes = (uint32_t)(left < 128 ? left : 128) * 320u;
    if (wl == 1 && umma::elect_one())
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(xbar)), "r"(bytes) : "memory");
    if (g == 0 && ir >= 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_base + R2_K + (uint32_t)r * R2_XROW),
                   "l"(vout0 + (size_t)ir * kDView), "r"(320u), "r"(umma::smem_u32(xbar))
                   : "memory");
  };
  auto x_wait = [&]() {
    umma::mbar_wait(xbar, xph);
    xph ^= 1;
  };
  const float4* const xs_ptr = reinterpret_cast<const float4*>(smem + R2_K + r * R2_XROW + g * 192);
  float4 xr[12];
  auto x_get = [&](bool ok) {                                                // after x_wait(); rows past P read as zeros
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[j] = ok ? xs_ptr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (g == 0) {
#pragma unroll
      for (int j = 8; j < 12; ++j) xr[j] = ok ? xs_ptr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float4* pe = reinterpret_cast<const float4*>(pe_table + (r % SN) * 8);
      xr[8] = __ldg(pe);
      xr[9] = __ldg(pe + 1);
      xr[10] = xr[11] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // the next tile's rows of vout0 -> L2, one tile ahead of the copy that reads them (they come from DRAM: the view stage wrote 1.5 GB)
  auto x_prefetch = [&](long long tile_nx) {
    if (SN == kNC) {                                                         // two rays per tile: 64 rows of 320 bytes each
      const long long ray = tile_nx * 2 + (t >> 7);
      if (ray * kNC < P) {
        const char* base = reinterpret_cast<const char*>(vout0 + (size_t)ray * kNS * kDView);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (t & 127) * 128));
        if ((t & 127) < 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (128 + (t & 127)) * 128));
      }
    } else {                                                                 // one ray: 128 contiguous rows
      const char* base = reinterpret_cast<const char*>(vout0 + (size_t)tile_nx * kNS * kDView);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + t * 128));
      if (t < 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (256 + t) * 128));
    }
  };
  // the loaded chunks as 16-bit operand chunks -> TMEM columns col0 + 4 c   (n_chunks: 12 for the QKV operand, 11 for [LN1 | x])
  auto x_store = [&](uint32_t col0, int n_chunks) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = 6 * g + i;
      if (c < n_chunks)
        umma::tmem_st4(tl + col0 + 4 * c, umma::pack2<BF16>(xr[2 * i].x, xr[2 * i].y), umma::pack2<BF16>(xr[2 * i].z, xr[2 * i].w),
                       umma::pack2<BF16>(xr[2 * i + 1].x, xr[2 * i + 1].y), umma::pack2<BF16>(xr[2 * i + 1].z, xr[2 * i + 1].w));
    }
  };
  long long in_row_cur = (long long)blockIdx.x < n_tiles ? in_row_from(blockIdx.x, perm_load(blockIdx.x)) : -1;
  if ((long long)blockIdx.x < n_tiles) x_copy(in_row_cur, blockIdx.x);

#ifdef UFO_PHASE_TIMING
  // interval k of a tile ends at the k-th UFO_TIM(): even k = a barrier in front of an MMA issue (epilogue + barrier), odd k = the return
  // of the MMA wait (issue + MMA + commit + wake-up); timed by lane 0 of warp 1 (not the issuing warp, column half g = 0)
  unsigned long long* tim = reinterpret_cast<unsigned long long*>(smem + R2_BAR + 64);
  if (t == 32) for (int i = 0; i < 32; ++i) tim[i] = 0;
  long long tim_prev = clock64();
  int tim_idx = 0;
#define UFO_TIM() do { if (t == 32) { const long long now_ = clock64(); tim[tim_idx] += (unsigned long long)(now_ - tim_prev); tim_prev = now_; } ++tim_idx; } while (0)
#else
#define UFO_TIM() do { } while (0)
#endif
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#ifdef UFO_PHASE_TIMING
    tim_idx = 0;
#endif
    const long long prow = tile * 128 + r;           // this thread's token: ray prow / SN, sorted sample prow % SN
    const bool row_ok = prow < P;
    const long long in_row = in_row_cur;             // looked up one tile ahead (in_row_nx of the previous iteration)
    const bool has_next = tile + (long long)gridDim.x < n_tiles;
    const unsigned pv_nx = has_next ? perm_load(tile + gridDim.x) : 0u;         // the next tile's perm byte: used in R13
    if (has_next) x_prefetch(tile + gridDim.x);
    // ---- R0: x = [view-stage token 0 output | order encoding | 0] -> 16-bit A operand in TMEM (chunks 6 g .. 6 g + 5)
    x_wait();                                                    // copy issued by the previous tile's R13 (or the prologue)
    x_get(row_ok);
    x_store(C_X, 12);
    umma::tmem_st_wait();
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R1a: k | v = x . Wkv^T
    const uint32_t b1 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifdef UFO_RAY_PREWAIT
      if (first_gemm) use_wait();
#else
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_KV, C_X, b1, 176, 176, 6, 0);
        commit_and_prewait(true);
      }
    }
    first_gemm = false;
    mma_wait();
    UFO_TIM();
    load_piece(2);                                   // merge weights -> the slot Wkv leaves
    // ---- R2a: K' = elu(k)+1, V' = v -> MN-major operand tiles in shared memory   (linear_attention.py:36-41)
    {
      auto r2a = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int C0 = GG ? 6 : 0, NC = GG ? 5 : 6;          // chunks of this column half: 0..5 | 6..10
        float2 kk[NC][4], vv[NC][4];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          tmem_ld8p(tl + D_KV + 8 * (C0 + i), kk[i]);
          tmem_ld8p(tl + D_KV + 88 + 8 * (C0 + i), vv[i]);
        }
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) kk[i][k] = elu1_2(kk[i][k]);
          st_chunk2<BF16>(smem + R2_K, r, C0 + i, kk[i]);
          st_chunk2<BF16>(smem + R2_V, r, C0 + i, vv[i]);
        }
      };
      UFO_G2_DISPATCH(r2a)
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R1b: q = x . Wq^T   (the k | v accumulator is consumed)
    const uint32_t b2 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_Q, C_X, b2, 96, 96, 6, 0);
        commit_and_prewait(true);
      }
    }
    mma_wait();
    UFO_TIM();
    load_piece(3);                                   // mlp.0 rows 0..95
    // ---- R2b: Q' = elu(q)+1 -> A operand of the message GEMM (chunk 11 = 0)
    {
      auto r2b = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int C0 = GG ? 6 : 0, NC = GG ? 5 : 6;
        float2 a[NC][4];
#pragma unroll
        for (int i = 0; i < NC; ++i) tmem_ld8p(tl + D_Q + 8 * (C0 + i), a[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NC; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) a[i][k] = elu1_2(a[i][k]);
          umma::tmem_st4(tl + C_QP + 4 * (C0 + i), pack2v<BF16>(a[i][0]), pack2v<BF16>(a[i][1]), pack2v<BF16>(a[i][2]), pack2v<BF16>(a[i][3]));
        }
        if (GG == 1) umma::tmem_st4(tl + C_QP + 44, 0u, 0u, 0u, 0u);
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r2b)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R3: per sequence  D[b][a] = sum_s V'[s][b] K'[s][a]   (rows 88..95 = sum_s K'[s][a]);  both operands MN-major
    if (UFO_RAY_ISSUER) {
      umma::tc_fence_after();
      const uint32_t idesc = umma::make_idesc(128, 96, FMT, true, true);
      const uint32_t alo = umma::desc_lo(sm_base + R2_V, 128), blo = umma::desc_lo(sm_base + R2_K, 128);
#pragma unroll
      for (int sq = 0; sq < NSEQ; ++sq) {
#pragma unroll
        for (int ks = 0; ks < SN / 16; ++ks) {
          const uint32_t off = (uint32_t)(sq * (SN / 16) + ks) * (256u >> 4);      // 16 token rows = 256 bytes of an MN-major tile
          umma::mma_f16_lh(tm + (sq == 0 ? D_S0 : D_S1), alo + off, umma::desc_hi(kChunk), blo + off, umma::desc_hi(kChunk), idesc, ks > 0);
        }
      }
      umma::commit(bar);
    }
    mma_wait();
    UFO_TIM();
    // ---- R4: block-diagonal KV (per head 11x11) + per-head K-sum rows as the B operand of the message GEMM
    if (r < 96) {
      const int hr = r < 88 ? r / 11 : r - 88;       // rows 0..87: KV_h of the row's head; row 88+h: the K-sum of head h
      auto r4 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
#pragma unroll
        for (int sq = 0; sq < NSEQ; ++sq) {
          uint8_t* kvbd = smem + (sq == 0 ? R2_K : R2_V);        // [96 rows b][96 cols a], chunk stride 96*16
          float v[6][8];
#pragma unroll
          for (int i = 0; i < 6; ++i) umma::tmem_ld8(tl + (sq == 0 ? D_S0 : D_S1) + 8 * (6 * GG + i), v[i]);
          umma::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int c = 6 * GG + i;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i][k] = ((8 * c + k) < 88 && ((8 * c + k) / 11) == hr) ? v[i][k] : 0.f;
            *reinterpret_cast<uint4*>(kvbd + c * (96 * 16) + r * 16) = pack8<BF16>(v[i]);
          }
        }
      };
      UFO_G2_DISPATCH(r4)
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R5: message numerator Q'.KV_h (columns 0..87) and per-head normalisers Q'_h.Ksum_h (columns 88..95)
    if (UFO_RAY_ISSUER) {
      umma::tc_fence_after();
#pragma unroll
      for (int sq = 0; sq < NSEQ; ++sq) issue_ts(sq == 0 ? D_S0 : D_S1, C_QP, sm_base + (sq == 0 ? R2_K : R2_V), 96, 96, 6, 0);
      umma::commit(bar);
    }
    mma_wait();
    UFO_TIM();
    x_copy(in_row, tile);                                        // K' / V' / KV are dead: this tile's x again -> stash (for R8 / R12), an L2 hit
    // ---- R6: msg = numerator / (normaliser + 1e-6)                (linear_attention.py:44-45) -> A operand of the merge
    {
      const uint32_t dm = tl + ((NSEQ == 1 || r < SN) ? D_S0 : D_S1);
      float zr[8];
      umma::tmem_ld8(dm + 88, zr);
      auto r6 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int C0 = GG ? 6 : 0, NC = GG ? 5 : 6;
        float v[NC][8];
#pragma unroll
        for (int i = 0; i < NC; ++i) umma::tmem_ld8(dm + 8 * (C0 + i), v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) zr[j] = 1.f / (zr[j] + 1e-6f);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int c = C0 + i;
#pragma unroll
          for (int k = 0; k < 8; ++k) v[i][k] *= zr[(8 * c + k) / 11];      // static index: head of column 8c+k
          const uint4 u = pack8<BF16>(v[i]);
          umma::tmem_st4(tl + C_M + 4 * c, u.x, u.y, u.z, u.w);
        }
        if (GG == 1) umma::tmem_st4(tl + C_M + 44, 0u, 0u, 0u, 0u);
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r6)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R7: merge
    const uint32_t b3 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_MRG, C_M, b3, 96, 96, 6, 0);
        commit_and_prewait(true);
      }
    }
    mma_wait();
    UFO_TIM();
    load_piece(4);                                   // mlp.0 rows 96..175
    // ---- R8: LayerNorm 1 -> first half of the concat operand [LN1 | x] (chunks 0..10; the weight image has the same K order)
    {
      auto r8 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int C0 = GG ? 6 : 0, NC = GG ? 5 : 6;
        float2 v[NC][4];
#pragma unroll
        for (int i = 0; i < NC; ++i) tmem_ld8p(tl + D_MRG + 8 * (C0 + i), v[i]);
        umma::tmem_ld_wait();
        red[GG * 128 + r] = ln_partial<NC>(v);
        if (NSEQ > 1) {                                           // x again, for the [LN1 | x] operand
          x_wait();
          x_get(row_ok);
          x_store(C_XL + 44, 11);
        }
        umma::tc_fence_before();
        pair_sync();                                               // the partials of a row come from the warps q and q + 4 only
        const float2 st = ln2_stats(red, r, 1.f / 88.f);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int c = C0 + i;
          float2 o[4];
          ln_apply(v[i], st, prm.n1w + 8 * c, prm.n1b + 8 * c, o);
          umma::tmem_st4(tl + C_XL + 4 * c, pack2v<BF16>(o[0]), pack2v<BF16>(o[1]), pack2v<BF16>(o[2]), pack2v<BF16>(o[3]));
        }
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r8)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R9a: mlp.0 on [LN1 | x]  (K = 176), output rows 0..95
    const uint32_t b4 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_ML0, C_XL, b4, 96, 96, 11, 0);
        commit_and_prewait(true);
      }
    }
    mma_wait();
    UFO_TIM();
    load_piece(5);                                   // mlp.2
    // ---- R10a: ReLU -> hidden operand chunks 0..11
    {
      auto r10a = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        float2 v[6][4];
#pragma unroll
        for (int i = 0; i < 6; ++i) tmem_ld8p(tl + D_ML0 + 8 * (6 * GG + i), v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 6; ++i)
          umma::tmem_st4(tl + C_H1 + 4 * (6 * GG + i), relu_pack2<BF16>(v[i][0]), relu_pack2<BF16>(v[i][1]), relu_pack2<BF16>(v[i][2]),
                         relu_pack2<BF16>(v[i][3]));
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r10a)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R9b: mlp.0 output rows 96..175 (the first accumulator is consumed)
    const uint32_t b5 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_ML0, C_XL, b5, 80, 80, 11, 0);
        commit_and_prewait(true);
      }
    }
    mma_wait();
    UFO_TIM();
    load_piece(6);                                   // SRDF head layer 0 (hi | lo)
    // ---- R10b: ReLU -> hidden operand chunks 12..21 (over the dead [LN1 | x] operand)
    {
      auto r10b = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        float2 v[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i) tmem_ld8p(tl + D_ML0 + 8 * (5 * GG + i), v[i]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 5; ++i)
          umma::tmem_st4(tl + C_H1 + 4 * (12 + 5 * GG + i), relu_pack2<BF16>(v[i][0]), relu_pack2<BF16>(v[i][1]), relu_pack2<BF16>(v[i][2]),
                         relu_pack2<BF16>(v[i][3]));
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r10b)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R11: mlp.2
    const uint32_t b6 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_ML2, C_H1, b6, 96, 96, 11, 0);
        commit_and_prewait(true);
      }
    }
    mma_wait();
    UFO_TIM();
    if (has_next) load_piece(0);                       // the next tile's Wkv
    // ---- R12: LayerNorm 2, residual in fp32 from the fp32 input, split hi/lo for the SRDF head
    {
      auto r12 = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
        constexpr int C0 = GG ? 6 : 0, NC = GG ? 5 : 6;
        float2 v[NC][4];
#pragma unroll
        for (int i = 0; i < NC; ++i) tmem_ld8p(tl + D_ML2 + 8 * (C0 + i), v[i]);
        umma::tmem_ld_wait();
        red[GG * 128 + r] = ln_partial<NC>(v);
        umma::tc_fence_before();
        pair_sync();
        if (NSEQ == 1) x_wait();                                   // (with two sequences R8 has waited for this copy)
        x_get(row_ok);                                             // fp32 input row from the stash (copied at R6)
        const float2 st = ln2_stats(red, r, 1.f / 88.f);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int c = C0 + i;
          const float4 xa = xr[2 * i], xb = xr[2 * i + 1];        // chunk c = 6 g + i of the fp32 input row
          float2 o2[4];
          ln_apply(v[i], st, prm.n2w + 8 * c, prm.n2b + 8 * c, o2);
          const float o[8] = {xa.x + o2[0].x, xa.y + o2[0].y, xa.z + o2[1].x, xa.w + o2[1].y,
                              xb.x + o2[2].x, xb.y + o2[2].y, xb.z + o2[3].x, xb.w + o2[3].y};
          float hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) split_hi_lo<BF16>(o[k], hi[k], lo[k]);
          if (ray_out != nullptr && row_ok) {
            float4* dst = reinterpret_cast<float4*>(ray_out + (size_t)prow * kDRay + 8 * c);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
          }
          const uint4 uh = pack8<BF16>(hi), ul = pack8<BF16>(lo);
          umma::tmem_st4(tl + C_RHI + 4 * c, uh.x, uh.y, uh.z, uh.w);
          umma::tmem_st4(tl + C_RLO + 4 * c, ul.x, ul.y, ul.z, ul.w);
        }
        if (GG == 1) {
          umma::tmem_st4(tl + C_RHI + 44, 0u, 0u, 0u, 0u);
          umma::tmem_st4(tl + C_RLO + 44, 0u, 0u, 0u, 0u);
        }
        umma::tmem_st_wait();
      };
      UFO_G2_DISPATCH(r12)
    }
    umma::tc_fence_before();
    __syncthreads();
    UFO_TIM();
    // ---- R13: DensityMLP layer 0 in split precision: r_hi.W_hi + r_lo.W_hi + r_hi.W_lo   (ray_transformer.py:147-150)
    const uint32_t b7 = use_piece();
    if (UFO_RAY_ISSUE_WARP) {
#ifndef UFO_RAY_PREWAIT
      use_wait();
#endif
      if (UFO_RAY_ELECT) {
        umma::tc_fence_after();
        issue_ts(D_DEN, C_RHI, b7, 32, 32, 6, 0);
        issue_ts(D_DEN, C_RLO, b7, 32, 32, 6, 1);
        issue_ts(D_DEN, C_RHI, b7 + 32 * 96 * 2, 32, 32, 6, 1);
        commit_and_prewait(has_next);
      }
    }
    const long long in_row_nx = has_next ? in_row_from(tile + gridDim.x, pv_nx) : -1;
    if (has_next) x_copy(in_row_nx, tile + gridDim.x);           // the next tile's x: under the SRDF-head GEMM and its tail
    in_row_cur = in_row_nx;
    mma_wait();
    UFO_TIM();
    if (has_next) load_piece(1);                       // the next tile's Wq
    // ---- R14: DensityMLP tail 32 -> 16 -> 1 in fp32 (hidden units 8 g .. 8 g + 7 per thread)
    {
      float h[32];
      umma::tmem_ld16(tl + D_DEN, h);
      umma::tmem_ld16(tl + D_DEN + 16, h + 16);
      umma::tmem_ld_wait();
      auto tail = [&](auto GGc) {
        constexpr int GG = decltype(GGc)::value;
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = fmaxf(h[i] + prm.db0[i], 0.f);
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int o = 8 * GG + j;
          float2 acc = make_float2(prm.db2[o], 0.f);
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            acc = __ffma2_rn(make_float2(h[i], h[i + 1]), make_float2(prm.dw2[o][i], prm.dw2[o][i + 1]), acc);
          part = fmaf(fmaxf(acc.x + acc.y, 0.f), prm.dw4[o], part);
        }
        part_s[GG * 128 + r] = part;
      };
      UFO_G2_DISPATCH(tail)
      umma::tc_fence_before();
      pair_sync();
      if (g == 0 && row_ok) srdf[prow] = prm.db4 + (part_s[r] + part_s[128 + r]);
    }
    UFO_TIM();                                                   // interval 18: the SRDF-head tail (R14)
    // the scratch is rewritten by the next tile's R2a only after its R0 barrier; TMEM is rewritten after that barrier too
  }
#ifdef UFO_PHASE_TIMING
  if (blockIdx.x == 0 && t == 32 && n_tiles >= 4096) {
    const unsigned long long nt = (unsigned long long)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    printf("RAYTIM SN=%d tiles=%llu :", SN, nt);
    for (int i = 0; i < 19; ++i) printf(" %llu", tim[i] / nt);
    printf("\n");
  }
#endif
#undef UFO_TIM
  umma::tc_fence_before();
  __syncthreads();
  if (wl == 0) umma::tmem_dealloc(tm, 256);
#undef UFO_RAY_ISSUER
#undef UFO_RAY_ISSUE_WARP
#undef UFO_RAY_ELECT
#undef UFO_RAY_LOADER
}

}  // namespace ufo
