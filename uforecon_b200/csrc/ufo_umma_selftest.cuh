// Hardware self-test of the UMMA descriptor / layout conventions in ufo_umma.cuh: one CTA computes
// D[128][N] = A . B^T on tcgen05 from operands staged by the same thread-per-row writers the fused
// kernels use.  mode 0: A [128][K], B [N][K] (both K-major).  mode 1: A given as At [K][128] and B as
// Bt [K][N] (both MN-major; the K^T.V product of the ray-stage linear attention).
#pragma once
#include "ufo_common.cuh"
#include "ufo_umma.cuh"

namespace ufo {

template <bool kBF16>
__global__ void __launch_bounds__(128) k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ D, int N, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_u8[];
  uint8_t* smem = smem_u8;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;                                  // 128 x K  (or K x 128)
  uint8_t* sB = smem + 128 * 256 * 2;                  // up to 256 x 256 halves
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_barrier_init();
  }
  if (mode == 0) {
    for (int c = 0; c < K / 8; ++c) {
      const float* a = A + (size_t)tid * K + c * 8;
      uint4 v;
      v.x = umma::pack2<kBF16>(a[0], a[1]); v.y = umma::pack2<kBF16>(a[2], a[3]);
      v.z = umma::pack2<kBF16>(a[4], a[5]); v.w = umma::pack2<kBF16>(a[6], a[7]);
      *reinterpret_cast<uint4*>(sA + umma::tile_off(128, tid, c)) = v;
      for (int n = tid; n < N; n += 128) {
        const float* b = B + (size_t)n * K + c * 8;
        uint4 w;
        w.x = umma::pack2<kBF16>(b[0], b[1]); w.y = umma::pack2<kBF16>(b[2], b[3]);
        w.z = umma::pack2<kBF16>(b[4], b[5]); w.w = umma::pack2<kBF16>(b[6], b[7]);
        *reinterpret_cast<uint4*>(sB + umma::tile_off(N, n, c)) = w;
      }
    }
  } else {
    // thread k owns row k of At [K][128] and Bt [K][N]   (K <= 128)
    if (tid < K) {
      for (int c = 0; c < 16; ++c) {
        const float* a = A + (size_t)tid * 128 + c * 8;
        uint4 v;
        v.x = umma::pack2<kBF16>(a[0], a[1]); v.y = umma::pack2<kBF16>(a[2], a[3]);
        v.z = umma::pack2<kBF16>(a[4], a[5]); v.w = umma::pack2<kBF16>(a[6], a[7]);
        *reinterpret_cast<uint4*>(sA + umma::tile_off(K, tid, c)) = v;
      }
      for (int c = 0; c < N / 8; ++c) {
        const float* b = B + (size_t)tid * N + c * 8;
        uint4 w;
        w.x = umma::pack2<kBF16>(b[0], b[1]); w.y = umma::pack2<kBF16>(b[2], b[3]);
        w.z = umma::pack2<kBF16>(b[4], b[5]); w.w = umma::pack2<kBF16>(b[6], b[7]);
        *reinterpret_cast<uint4*>(sB + umma::tile_off(K, tid, c)) = w;
      }
    }
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t fmt = kBF16 ? umma::kFmtBF16 : umma::kFmtF16;
  if (tid == 0) {
    if (mode == 0) {
      umma::issue_gemm(tmem_base, umma::smem_u32(sA), 128, umma::smem_u32(sB), N, K / 8, umma::make_idesc(128, N, fmt, false, false), 0);
    } else {
      const uint32_t idesc = umma::make_idesc(128, N, fmt, true, true);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = umma::make_smem_desc(umma::smem_u32(sA) + ks * 256, 128, K * 16);
        const uint64_t bd = umma::make_smem_desc(umma::smem_u32(sB) + ks * 256, 128, K * 16);
        umma::mma_f16(tmem_base, ad, bd, idesc, ks > 0);
      }
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    umma::tmem_ld16(tmem_base + lane_base + n0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (n0 + i < N) D[(size_t)tid * N + n0 + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 256);
}


// mode 2: TS form with TWO independent halves in one CTA (the schedule of the two-tiles-in-flight kernels): half h
// (threads 128h..128h+127) computes D_h[128][N] = A_h[128][K] . B[N][K]^T with A_h written to TMEM by its row-owner
// threads (tcgen05.st, packed 16-bit pairs: column j of the operand = K elements 2j, 2j+1), B shared by both halves in
// shared memory, its own accumulator columns, its own mbarrier, its own issuing thread and a named barrier.
template <bool kBF16>
__global__ void __launch_bounds__(256) k_umma_selftest_ts(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem_u8[];
  uint8_t* sB = smem_u8;                               // up to 256 x 256 halves
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, h = tid >> 7, r = tid & 127;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    umma::mbar_init(&bar[0], 1);
    umma::mbar_init(&bar[1], 1);
    umma::fence_barrier_init();
  }
  for (int c = 0; c < K / 8; ++c)
    for (int n = tid; n < N; n += 256) {
      const float* b = B + (size_t)n * K + c * 8;
      uint4 w;
      w.x = umma::pack2<kBF16>(b[0], b[1]); w.y = umma::pack2<kBF16>(b[2], b[3]);
      w.z = umma::pack2<kBF16>(b[4], b[5]); w.w = umma::pack2<kBF16>(b[6], b[7]);
      *reinterpret_cast<uint4*>(sB + umma::tile_off(N, n, c)) = w;
    }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s + 256u * h;                    // this half's 256 columns
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t a_col = 0, d_col = 96;                              // A: K/2 <= 88 columns, D: N <= 160 columns
  {
    const float* a = A + ((size_t)h * 128 + r) * K;
    for (int c = 0; c < K / 8; ++c)
      umma::tmem_st4(tlane + a_col + 4 * c, umma::pack2<kBF16>(a[8 * c], a[8 * c + 1]), umma::pack2<kBF16>(a[8 * c + 2], a[8 * c + 3]),
                     umma::pack2<kBF16>(a[8 * c + 4], a[8 * c + 5]), umma::pack2<kBF16>(a[8 * c + 6], a[8 * c + 7]));
    umma::tmem_st_wait();
  }
  umma::tc_fence_before();
  umma::bar_sync(1 + h, 128);
  if (r == 0) {
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc(128, N, kBF16 ? umma::kFmtBF16 : umma::kFmtF16, false, false);
    const uint32_t b_lbo = (uint32_t)N * 16u;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bd = umma::make_smem_desc(umma::smem_u32(sB) + 2 * ks * b_lbo, b_lbo, 128u);
      umma::mma_f16_ts(tmem + d_col, tmem + a_col + 8 * ks, bd, idesc, ks > 0);
    }
    umma::commit(&bar[h]);
  }
  umma::mbar_wait(&bar[h], 0);
  umma::tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    umma::tmem_ld16(tlane + d_col + n0, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (n0 + i < N) D[((size_t)h * 128 + r) * N + n0 + i] = v[i];
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace ufo
