// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; no CUTLASS dependency).
//
// Operand layouts used throughout (no swizzle, "interleaved" canonical UMMA layout): a 16-bit operand
// tile is a grid of 8x8 core matrices, each 128 contiguous bytes (8 rows of 16 bytes).
//   K-major  operand [R rows][K]: core (r/8, k/8) at  (k/8)*LBO + (r/8)*SBO, row r%8 at +16*(r%8)
//   MN-major operand [K][R]     : core (r/8, k/8) at  (k/8)*LBO + (r/8)*SBO, k-row k%8 at +16*(k%8)
// With the tile stored chunk-major (all row groups of k-chunk 0, then k-chunk 1, ...):
//   SBO = 128 B, LBO = (R/8)*128 B, and a thread that owns row r writes its 8 consecutive K values of
//   chunk c as ONE 16-byte store at  c*LBO + (r/8)*128 + (r%8)*16  (a warp covers 512 contiguous bytes).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ufo {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One try_wait.  -DUFO_MBAR_HINT adds a suspend-time hint (the warp sleeps in hardware until the phase completes or the
// hint expires); measured on a B200 it wakes later than the plain form: k_ray_tc +2.5 %, k_view_tc2 +1.4 % - not the default.
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
#ifndef UFO_MBAR_HINT
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
#endif
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error (a phase that never completes) traps after ~2 s instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t n = 1;; ++n) {
    if (mbar_try(bar, parity)) return;
    if ((n & 63u) == 0u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
// named barrier over `nthreads` threads (a multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; ncols power of two >= 32.  The allocated base address is written to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4)..+31).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 4 / 8 consecutive 32-bit TMEM columns of this thread's lane (16-bit A operands: column j = K elements 2j, 2j+1)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE (layout_type 0), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1 };

// Instruction descriptor of tcgen05.mma.kind::f16 with fp32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t fmt, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                      // c_format = F32
         | (fmt << 7) | (fmt << 10)     // a_format, b_format
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]   (one K=16 step).  Issued by ONE thread.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TS form: A [128 x 16] from TMEM (lane = row, 8 columns of packed 16-bit pairs), B from shared memory.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The same two instructions with the shared-memory descriptor passed as its two 32-bit halves: the K loop of a GEMM then advances
// the descriptor by ONE 32-bit add on the start-address field (bits 0-13 of the low word, in 16-byte units; the field cannot carry:
// shared-memory addresses stay below 2^18), and the high word is a compile-time constant.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void mma_f16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of a CONVERGED warp (all 32 lanes must execute this).  The single-thread instructions (tcgen05.mma, tcgen05.commit,
// cp.async.bulk) issued under `if (warp_uniform_condition && elect_one())` compile to straight-line code; under `if (threadIdx.x == 0)`
// the compiler cannot know that one thread is active and wraps each of them in an ELECT / BRA.U.ANY loop over the active threads.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Arrive on `bar` when all previously issued MMAs of this thread have completed (implies
// tcgen05.fence::before_thread_sync).  Issued by ONE thread.
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand staging helpers ------------------------------------------------------------------
template <bool kBF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (kBF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    uint32_t d;   // saturating: an activation beyond the fp16 range becomes +-65504, never inf (bf16 needs no clamp)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
  }
}

// Byte offset of the 16-byte piece (row r, k-chunk c) inside a chunk-major tile with `rows` rows.
__device__ __forceinline__ uint32_t tile_off(uint32_t rows, uint32_t r, uint32_t c) {
  return c * (rows * 16u) + (r >> 3) * 128u + (r & 7u) * 16u;
}

// Issue the K-loop of one GEMM: D[128 x N] = A[128 x K] . B[N x K]^T, both K-major chunk-major tiles.
// a_base/b_base are shared-memory byte addresses of the tiles; a_rows/b_rows their row counts.
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_base, uint32_t a_rows, uint32_t b_base, uint32_t b_rows,
                                           uint32_t k_chunks /* K/8 */, uint32_t idesc, uint32_t accumulate_first) {
  const uint32_t a_lbo = a_rows * 16u, b_lbo = b_rows * 16u;
  for (uint32_t c = 0; c < k_chunks; c += 2) {
    const uint64_t ad = make_smem_desc(a_base + c * a_lbo, a_lbo, 128u);
    const uint64_t bd = make_smem_desc(b_base + c * b_lbo, b_lbo, 128u);
    mma_f16(tmem_d, ad, bd, idesc, (c > 0) ? 1u : accumulate_first);
  }
}

}  // namespace umma
}  // namespace ufo
