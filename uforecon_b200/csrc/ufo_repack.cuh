// Layout kernels: reference NCHW / NCDHW tensors -> channel-last texel arrays for the fused gathers.
// (The reference keeps NCHW and lets F.grid_sample stride over channels; a 32-channel texel here
//  is one 128-byte line, an 8-channel voxel one 32-byte sector.)
#pragma once
#include "ufo_common.cuh"

namespace ufo {

// in [N][C][S] -> out [N][S][C].  One block transposes a 32(spatial) x C tile through shared memory:
// reads are coalesced along S, writes along C.
template <int C>
__global__ void __launch_bounds__(256) k_nchw_to_nhwc(const float* __restrict__ in, float* __restrict__ out,
                                                     long long S) {
  __shared__ float tile[C][33];
  const int n = blockIdx.y;
  const long long s0 = (long long)blockIdx.x * 32;
  const float* src = in + (long long)n * C * S;
  float* dst = out + (long long)n * S * C;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int c = ty; c < C; c += 8) {
    long long s = s0 + tx;
    tile[c][tx] = (s < S) ? __ldg(src + (long long)c * S + s) : 0.f;
  }
  __syncthreads();
  // write: consecutive threads -> consecutive channels of consecutive pixels
  for (int i = threadIdx.x; i < 32 * C; i += 256) {
    int p = i / C, c = i % C;
    long long s = s0 + p;
    if (s < S) dst[s * C + c] = tile[c][p];
  }
}

// (r,g,b) planes + MVS depth plane -> float4 texels.  imgs [N][3][S], depth [N][S] -> out [N][S].
static __global__ void __launch_bounds__(256) k_pack_rgbd(const float* __restrict__ imgs, const float* __restrict__ depth,
                                                  float4* __restrict__ out, long long S, int N) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * N) return;
  long long n = i / S, s = i % S;
  const float* im = imgs + n * 3 * S;
  out[i] = make_float4(__ldg(im + s), __ldg(im + S + s), __ldg(im + 2 * S + s), __ldg(depth + n * S + s));
}

// Are the two copies of every pair map identical?  match [NV][(NV-1)*32][S] (reference layout): slot (a, b-1) vs slot
// (b, a) for a < b, compared bit for bit; *differ is set to 1 on the first mismatch.
static __global__ void __launch_bounds__(256) k_match_sym_check(const float* __restrict__ match, int NV, long long S, int* __restrict__ differ) {
  const long long per = 32 * S;                       // floats of one pair map
  const int a = blockIdx.y, b = blockIdx.z;
  if (a >= b) return;
  const unsigned* pa = reinterpret_cast<const unsigned*>(match) + ((long long)a * (NV - 1) + (b - 1)) * per;
  const unsigned* pb = reinterpret_cast<const unsigned*>(match) + ((long long)b * (NV - 1) + a) * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x)
    if (__ldg(pa + i) != __ldg(pb + i)) {
      *differ = 1;
      return;
    }
}

}  // namespace ufo
