// Shared definitions of the UFORecon B200 hot-path library (device structs, error plumbing).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/uforecon_b200.h"

namespace ufo {

constexpr int kFeatC = 32;   // FPN feature / match-map channels   (ray_transformer.py:91)
constexpr int kVolC = 8;     // CostRegNetWeight feature channels  (feature_volume.py:110)
constexpr int kDView = 80;   // 32 + 24 + 16 + 8                   (ray_transformer.py:135)
constexpr int kDRay = 88;    // kDView + 8 order PE                (ray_transformer.py:138)
// Position of token channel c (reference order: feat 32 | vol 24 | sim 16 | depth PE 8, ray_transformer.py:258-288) inside the 16-bit
// token rows of the tensor-core path.  The gather's lane j of a point owns feat 4j..4j+3, channel j of the three frustum features and
// depth-PE component j: the row keeps those eight values adjacent ([lane 0: f f f f v0 v1 v2 pe | lane 1: ... ] | sim 16), so a lane
// writes its part of a view row as ONE 16-byte store (it was one 8-byte and four 2-byte stores).  The K columns of every weight that
// multiplies a token (QKV, the x half of mlp.0, the x part of the radiance head) are permuted the same way on the host.
__host__ __device__ constexpr int tok_pos(int c) {
  return c < 32 ? 8 * (c / 4) + c % 4 : (c < 56 ? 8 * ((c - 32) % 8) + 4 + (c - 32) / 8 : (c < 72 ? 64 + (c - 56) : 8 * (c - 72) + 7));
}
constexpr int kHeads = 8;
constexpr int kNC = UFO_N_COARSE;
constexpr int kNS = UFO_N_SAMPLES;
constexpr int kMaxV = UFO_MAX_VIEWS;

// Per-view-set constants, passed to kernels by value (fits the 4 KB parameter space).
struct SceneDev {
  int nv, H, W, h, w;
  int vd[3], vh[3], vw[3];
  const float* feat_cl;         // [NV][h][w][32]
  const float4* rgbd_cl;        // [NV][H][W] (r,g,b,mvs_depth)
  const float* match_cl;        // [slots][h][w][32]: NV*(NV-1) slots in the reference's layout, NV*(NV-1)/2 for compact pair maps
  unsigned char match_slot[kMaxV][kMaxV];  // [v][o]: slot of the map sampled at view v's projection for the pair {v, o}.  Reference
                                // layout: slot v*(NV-1) + (o > v ? o-1 : o); when the two copies of every pair map are bit-identical
                                // (verified at scene creation) or the maps were given compact, both views share one slot
  int match_sym;                // 0: two slots per pair, 1: copies identical (one slot read), 2: compact input
  const float* vol_feat_cl[3];  // [NV][D][hs][ws][8]
  const float* vol_w[3];        // [NV][D][hs][ws]
  const float* ray_d;           // [3][H*W]
  const float* cam_ray_d;       // [3][H*W]
  float P[kMaxV][12];           // rows 0..2 of the world->NDC matrices (source_poses)
  float w2c_z[kMaxV][4];        // third row of the scaled w2c (camera-space z of a point)
  float cam_o[kMaxV][3];        // source camera centres  (source_poses_inv[:, :3, 3])
  float ref_o[3];               // render camera centre   (ref_pose_inv[:3, 3])
  float ray_o[3];
  float near0, far0;            // near_fars[0]  (model.py:328,416-421 use view 0 for every view)
};

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define UFO_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::ufo::fail(UFO_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define UFO_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    ::ufo::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ::ufo::fail(UFO_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

// Opt-in dynamic shared memory size of a kernel, set once per (call site, device): function attributes are per device,
// and one process may drive several.
#define UFO_SMEM_ATTR(kernel, bytes)                                                                     \
  do {                                                                                                   \
    static std::atomic<unsigned long long> _done{0};                                                     \
    int _dev = 0;                                                                                        \
    UFO_CUDA(cudaGetDevice(&_dev));                                                                      \
    const unsigned long long _bit = 1ull << (_dev & 63);                                                 \
    if (!(_done.load(std::memory_order_relaxed) & _bit)) {                                               \
      UFO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _done.fetch_or(_bit, std::memory_order_relaxed);                                                   \
    }                                                                                                    \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Per-kernel device-time accounting (ufo_profile_begin / ufo_profile_end): when enabled, a
// ProfScope brackets one launch with CUDA events on the launching stream.
struct ProfState;
extern std::atomic<int> g_prof_on;
long long prof_open(const char* name, cudaStream_t st);   // returns the record the matching prof_close completes
void prof_close(long long rec, cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  long long rec;
  ProfScope(const char* name, cudaStream_t s) : st(s), rec(-1) {
    if (g_prof_on.load(std::memory_order_relaxed) != 0) rec = prof_open(name, st);
  }
  ~ProfScope() {
    if (rec >= 0) prof_close(rec, st);
  }
};
// Launch one kernel: optional event bracket, launch counter, launch-error check (returns on failure).
#define UFO_KERNEL(name, st, ...)                 \
  do {                                            \
    {                                             \
      ::ufo::ProfScope _ufo_prof_scope_(name, st); \
      __VA_ARGS__;                                \
    }                                             \
    UFO_LAUNCH_CHECK();                           \
  } while (0)

}  // namespace ufo
