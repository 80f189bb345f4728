// C-ABI entry points of libuforecon_b200.so (declared in include/uforecon_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ufo_common.cuh"
#include "ufo_costvol.cuh"
#include "ufo_gather.cuh"
#include "ufo_repack.cuh"
#include "ufo_sampler_render.cuh"
#include "ufo_xfmr_fp32.cuh"
#include "ufo_handles.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ufo_umma_selftest.cuh"
#include "ufo_tsdf.cuh"
#include "ufo_fgrid.cuh"

namespace ufo {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// ---- per-kernel device-time accounting -----------------------------------------------------------
std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
long long g_prof_epoch = 0;             // bumped whenever the record list is cleared: a handle from an earlier epoch is ignored
std::vector<cudaEvent_t> g_prof_pool;   // recycled events
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
// prof_open returns the index of ITS record and prof_close takes it back: with several host threads launching concurrently the
// "last record" is not necessarily the caller's
long long prof_open(const char* name, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  ProfRec r{name, prof_event(), prof_event()};
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
  return (g_prof_epoch << 32) | ((long long)g_prof_recs.size() - 1);
}
void prof_close(long long rec, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  const long long idx = rec & 0xffffffffll;
  if (rec >= 0 && (rec >> 32) == g_prof_epoch && idx < (long long)g_prof_recs.size()) cudaEventRecord(g_prof_recs[(size_t)idx].e1, st);
}
}  // namespace ufo

using namespace ufo;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
static int fp32_chunk_rays();
static int tc_chunk_rays(int sms, int nv);

static int check_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(UFO_ENODEVICE, "no CUDA device visible: libuforecon_b200 has no CPU fallback");
  }
  return UFO_OK;
}

// Stream-ordered temporaries of one call: released (cudaFreeAsync) on every return path.
namespace {
struct AsyncTemps {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit AsyncTemps(cudaStream_t s) : st(s) {
    // The default pool hands freed memory back to the driver at the next synchronisation (release threshold 0), so every call
    // re-mapped its gigabytes of temporaries: kernel 1 inside the reference cascade was SLOWER than the PyTorch loop it replaces
    // at NV >= 5.  Keep the pool's memory cached, once per device (ufo_trim_pool() returns it).
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
      std::lock_guard<std::mutex> lk(mu);
      if (!done[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
          unsigned long long keep = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        done[dev] = true;
      }
    }
  }
  ~AsyncTemps() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  int alloc(void** out, size_t bytes) {
    UFO_CUDA(cudaMallocAsync(out, bytes, st));
    ptrs.push_back(*out);
    return UFO_OK;
  }
};
}  // namespace

extern "C" int ufo_abi_version(void) { return UFO_ABI_VERSION; }
extern "C" int ufo_trim_pool(void) {
  if (int e = check_device()) return e;
  int dev = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool;
  UFO_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  UFO_CUDA(cudaDeviceSynchronize());
  UFO_CUDA(cudaMemPoolTrimTo(pool, 0));
  return UFO_OK;
}
extern "C" const char* ufo_last_error(void) { return g_err; }
extern "C" int64_t ufo_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int ufo_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  if (int e = check_device()) return e;
  int dev = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  UFO_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return UFO_OK;
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
namespace {
struct BlobBuilder {
  std::vector<float> host;
  size_t add(const float* p, size_t n) {
    size_t off = (host.size() + 3) & ~size_t(3);  // 16-byte aligned tensors
    host.resize(off + n);
    memcpy(host.data() + off, p, n * sizeof(float));
    return off;
  }
};
}  // namespace

// ---- tensor-core operand images ------------------------------------------------------------------
static uint16_t cvt16(float v, bool bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));   // saturate like the device-side packing
  return *reinterpret_cast<uint16_t*>(&h);
}
static float back16(uint16_t u, bool bf16) {
  if (bf16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}
// W [n_real][k_real] (row stride ldw) -> K-major chunk-major B operand image with n_pad rows, k_pad columns.
// part 0: rounded value; part 1: rounded remainder (split precision).
static float g_pack_max_abs = 0.f;   // running max |w| of the values pack_b has seen (read by tc_weights_build under its caller's lock-free, single-threaded build)
static void pack_b(uint8_t* img, const float* W, int n_real, int k_real, int ldw, int n_pad, int k_pad, bool bf16, int part = 0) {
  uint16_t* o = reinterpret_cast<uint16_t*>(img);
  for (int n = 0; n < n_pad; ++n)
    for (int k = 0; k < k_pad; ++k) {
      const float v = (n < n_real && k < k_real) ? W[(size_t)n * ldw + k] : 0.f;
      if (!(fabsf(v) <= g_pack_max_abs)) g_pack_max_abs = std::isfinite(v) ? fabsf(v) : INFINITY;
      uint16_t h = cvt16(v, bf16);
      if (part == 1) h = cvt16(v - back16(h, bf16), bf16);
      o[(size_t)(k / 8) * (n_pad * 8) + (size_t)n * 8 + (k % 8)] = h;
    }
}

// K columns k0 .. k0+79 of W [rows][ldw] moved to the channel order of the 16-bit token rows: out[., k0 + tok_pos(c)] = W[., k0 + c]
static void permute_tok_cols(float* W, int rows, int ldw, int k0) {
  float tmp[80];
  for (int o = 0; o < rows; ++o) {
    float* w = W + (size_t)o * ldw + k0;
    for (int c = 0; c < 80; ++c) tmp[tok_pos(c)] = w[c];
    memcpy(w, tmp, sizeof(tmp));
  }
}

// byte offsets of the k_view_tc2 weight image (mirrors tc::V2_W* in ufo_view_tc2.cuh, which only device TUs include)
constexpr size_t kV2WQkv = 0, kV2WMrg = kV2WQkv + 240 * 80 * 2, kV2WMl0 = kV2WMrg + 80 * 96 * 2, kV2WMl2 = kV2WMl0 + 160 * 176 * 2,
                 kV2WRad = kV2WMl2 + 80 * 160 * 2, kV2WEnd = kV2WRad + 16 * 176 * 2;
// byte offsets of the k_ray_tc2 weight pieces (mirrors tc::R2W_* in ufo_ray_tc2.cuh)
constexpr size_t kR2WKv = 0, kR2WQ = kR2WKv + 176 * 96 * 2, kR2WMrg = kR2WQ + 96 * 96 * 2, kR2WMl0a = kR2WMrg + 96 * 96 * 2,
                 kR2WMl0b = kR2WMl0a + 96 * 176 * 2, kR2WMl2 = kR2WMl0b + 80 * 176 * 2, kR2WDen = kR2WMl2 + 96 * 176 * 2,
                 kR2WEnd = kR2WDen + 2 * 32 * 96 * 2;
static std::mutex g_tc_build_mu;
static int tc_weights_build(const UfoWeightsDesc* d, TcWeights* t, cudaStream_t st) {
  std::lock_guard<std::mutex> build_lock(g_tc_build_mu);
  g_pack_max_abs = 0.f;
  struct Record { TcWeights* t; ~Record() { t->max_abs_weight = g_pack_max_abs; } } record{t};
  for (int f = 0; f < 2; ++f) {
    const bool bf16 = (f == 0);
    std::vector<uint8_t> vi(tc::V_WEND, 0), ri(tc::RW_END, 0);
    {  // view stage
      std::vector<float> qkv((size_t)240 * 80);
      memcpy(qkv.data(), d->view.q, sizeof(float) * 6400);
      memcpy(qkv.data() + 6400, d->view.k, sizeof(float) * 6400);
      memcpy(qkv.data() + 12800, d->view.v, sizeof(float) * 6400);
      permute_tok_cols(qkv.data(), 240, 80, 0);
      pack_b(vi.data() + tc::V_WQKV, qkv.data(), 240, 80, 80, 240, 80, bf16);
      pack_b(vi.data() + tc::V_WMRG, d->view.merge, 80, 80, 80, 80, 80, bf16);
      std::vector<float> ml0p(d->view.mlp0, d->view.mlp0 + 160 * 160);
      permute_tok_cols(ml0p.data(), 160, 160, 0);
      pack_b(vi.data() + tc::V_WML0, ml0p.data(), 160, 160, 160, 160, 160, bf16);
      pack_b(vi.data() + tc::V_WML2, d->view.mlp2, 80, 160, 160, 80, 160, bf16);
      // [W0x | W0x | W0dir b0 0..]: the head sees x + LN2 without forming the sum; direction and bias ride in two
      // extra K chunks of the operand (columns 160..162 = relative direction, 163 = 1)
      std::vector<float> rad((size_t)16 * 176, 0.f);
      for (int o = 0; o < 16; ++o) {
        for (int k = 0; k < 80; ++k) rad[o * 176 + k] = rad[o * 176 + 80 + k] = d->radiance.w0[o * 83 + k];
        for (int k = 0; k < 3; ++k) rad[o * 176 + 160 + k] = d->radiance.w0[o * 83 + 80 + k];
        rad[o * 176 + 163] = d->radiance.b0[o];
      }
      permute_tok_cols(rad.data(), 16, 176, 0);
      pack_b(vi.data() + tc::V_WRAD, rad.data(), 16, 176, 176, 16, 176, bf16);
    }
    {  // ray stage
      std::vector<float> qkv((size_t)264 * 88);
      memcpy(qkv.data(), d->ray.q, sizeof(float) * 7744);
      memcpy(qkv.data() + 7744, d->ray.k, sizeof(float) * 7744);
      memcpy(qkv.data() + 15488, d->ray.v, sizeof(float) * 7744);
      pack_b(ri.data() + tc::RW_QKV, qkv.data(), 264, 88, 88, 272, 96, bf16);
      pack_b(ri.data() + tc::RW_MRG, d->ray.merge, 88, 88, 88, 96, 96, bf16);
      pack_b(ri.data() + tc::RW_ML0, d->ray.mlp0, 176, 176, 176, 176, 176, bf16);
      pack_b(ri.data() + tc::RW_ML2, d->ray.mlp2, 88, 176, 176, 96, 176, bf16);
      pack_b(ri.data() + tc::RW_DEN, d->density.w0, 32, 88, 88, 32, 96, bf16, 0);
      pack_b(ri.data() + tc::RW_DEN + 32 * 96 * 2, d->density.w0, 32, 88, 88, 32, 96, bf16, 1);
    }
    {  // view stage, two-tiles-in-flight kernel (ufo_view_tc2.cuh): rows / K columns regrouped per column group
      std::vector<uint8_t> v2(kV2WEnd, 0);
      std::vector<float> qkv((size_t)240 * 80);
      const float* part[3] = {d->view.q, d->view.k, d->view.v};
      for (int g = 0; g < 2; ++g)
        for (int pt = 0; pt < 3; ++pt)
          for (int j = 0; j < 40; ++j) memcpy(qkv.data() + (size_t)(120 * g + 40 * pt + j) * 80, part[pt] + (size_t)(40 * g + j) * 80, sizeof(float) * 80);
      permute_tok_cols(qkv.data(), 240, 80, 0);
      pack_b(v2.data() + kV2WQkv, qkv.data(), 240, 80, 80, 240, 80, bf16);
      // K = 80 message / LayerNorm channels as [g0: 40 | 8 zeros | g1: 40 | 8 zeros]
      auto pad96 = [](const float* W, int ldw, int col0, float* out, int ldo, int ocol0, int rows) {
        for (int o = 0; o < rows; ++o)
          for (int k = 0; k < 96; ++k) {
            const int src = k < 40 ? k : (k >= 48 && k < 88 ? k - 8 : -1);
            out[(size_t)o * ldo + ocol0 + k] = src < 0 ? 0.f : W[(size_t)o * ldw + col0 + src];
          }
      };
      std::vector<float> mrg((size_t)80 * 96);
      pad96(d->view.merge, 80, 0, mrg.data(), 96, 0, 80);
      pack_b(v2.data() + kV2WMrg, mrg.data(), 80, 96, 96, 80, 96, bf16);
      std::vector<float> ml0((size_t)160 * 176);
      for (int o = 0; o < 160; ++o) memcpy(ml0.data() + (size_t)o * 176, d->view.mlp0 + (size_t)o * 160, sizeof(float) * 80);
      pad96(d->view.mlp0, 160, 80, ml0.data(), 176, 80, 160);
      permute_tok_cols(ml0.data(), 160, 176, 0);
      pack_b(v2.data() + kV2WMl0, ml0.data(), 160, 176, 176, 160, 176, bf16);
      pack_b(v2.data() + kV2WMl2, d->view.mlp2, 80, 160, 160, 80, 160, bf16);
      // radiance head layer 0 on [x | LN2 g0 | dir 3, 1, 0.. | LN2 g1 | 0..]: x + LN2 without forming the sum, bias via the 1
      std::vector<float> rad((size_t)16 * 176, 0.f);
      for (int o = 0; o < 16; ++o) {
        memcpy(rad.data() + (size_t)o * 176, d->radiance.w0 + (size_t)o * 83, sizeof(float) * 80);
        pad96(d->radiance.w0 + (size_t)o * 83, 83, 0, rad.data() + (size_t)o * 176, 176, 80, 1);
        for (int k = 0; k < 3; ++k) rad[(size_t)o * 176 + 80 + 40 + k] = d->radiance.w0[o * 83 + 80 + k];
        rad[(size_t)o * 176 + 80 + 43] = d->radiance.b0[o];
      }
      permute_tok_cols(rad.data(), 16, 176, 0);
      pack_b(v2.data() + kV2WRad, rad.data(), 16, 176, 176, 16, 176, bf16);
      UFO_CUDA(cudaMalloc(&t->view_img2[f], v2.size()));
      UFO_CUDA(cudaMemcpyAsync(t->view_img2[f], v2.data(), v2.size(), cudaMemcpyHostToDevice, st));
      UFO_CUDA(cudaStreamSynchronize(st));
    }
    {  // ray stage, two-CTAs-per-SM kernel (ufo_ray_tc2.cuh): seven streamed pieces
      std::vector<uint8_t> r2(kR2WEnd, 0);
      std::vector<float> kv((size_t)176 * 88);
      memcpy(kv.data(), d->ray.k, sizeof(float) * 7744);
      memcpy(kv.data() + 7744, d->ray.v, sizeof(float) * 7744);
      pack_b(r2.data() + kR2WKv, kv.data(), 176, 88, 88, 176, 96, bf16);
      pack_b(r2.data() + kR2WQ, d->ray.q, 88, 88, 88, 96, 96, bf16);
      pack_b(r2.data() + kR2WMrg, d->ray.merge, 88, 88, 88, 96, 96, bf16);
      // mlp.0 consumes [LN1 | x] (the kernel keeps the x operand where it is and writes LN1 below it): swap the K halves
      std::vector<float> ml0((size_t)176 * 176);
      for (int o = 0; o < 176; ++o)
        for (int k = 0; k < 88; ++k) {
          ml0[(size_t)o * 176 + k] = d->ray.mlp0[(size_t)o * 176 + 88 + k];
          ml0[(size_t)o * 176 + 88 + k] = d->ray.mlp0[(size_t)o * 176 + k];
        }
      pack_b(r2.data() + kR2WMl0a, ml0.data(), 96, 176, 176, 96, 176, bf16);
      pack_b(r2.data() + kR2WMl0b, ml0.data() + (size_t)96 * 176, 80, 176, 176, 80, 176, bf16);
      pack_b(r2.data() + kR2WMl2, d->ray.mlp2, 88, 176, 176, 96, 176, bf16);
      pack_b(r2.data() + kR2WDen, d->density.w0, 32, 88, 88, 32, 96, bf16, 0);
      pack_b(r2.data() + kR2WDen + 32 * 96 * 2, d->density.w0, 32, 88, 88, 32, 96, bf16, 1);
      UFO_CUDA(cudaMalloc(&t->ray_img2[f], r2.size()));
      UFO_CUDA(cudaMemcpyAsync(t->ray_img2[f], r2.data(), r2.size(), cudaMemcpyHostToDevice, st));
      UFO_CUDA(cudaStreamSynchronize(st));
    }
    UFO_CUDA(cudaMalloc(&t->view_img[f], vi.size()));
    UFO_CUDA(cudaMalloc(&t->ray_img[f], ri.size()));
    UFO_CUDA(cudaMemcpyAsync(t->view_img[f], vi.data(), vi.size(), cudaMemcpyHostToDevice, st));
    UFO_CUDA(cudaMemcpyAsync(t->ray_img[f], ri.data(), ri.size(), cudaMemcpyHostToDevice, st));
    UFO_CUDA(cudaStreamSynchronize(st));
  }
  ViewParams& vp = t->vp;
  memcpy(vp.n1w, d->view.norm1_w, 320); memcpy(vp.n1b, d->view.norm1_b, 320);
  memcpy(vp.n2w, d->view.norm2_w, 320); memcpy(vp.n2b, d->view.norm2_b, 320);
  memcpy(vp.vtok, d->view_token, 320);
  for (int c = 0; c < 80; ++c) vp.vtok_x[tok_pos(c)] = d->view_token[c];
  memcpy(vp.rb0, d->radiance.b0, 64);
  for (int o = 0; o < 16; ++o)
    for (int i = 0; i < 3; ++i) vp.rw0d[o][i] = d->radiance.w0[o * 83 + 80 + i];
  memcpy(vp.rw2, d->radiance.w2, sizeof(vp.rw2));
  memcpy(vp.rb2, d->radiance.b2, 32);
  memcpy(vp.rw4, d->radiance.w4, 32);
  vp.rb4 = d->radiance.b4[0];
  for (int j = 0; j < 80; ++j) {      // view-token row of the QKV GEMM with the fp16 operands of the tensor-core path, accumulated in double
    double k = 0.0, v = 0.0;
    for (int c = 0; c < 80; ++c) {
      const double x = back16(cvt16(d->view_token[c], false), false);
      k += x * back16(cvt16(d->view.k[j * 80 + c], false), false);
      v += x * back16(cvt16(d->view.v[j * 80 + c], false), false);
    }
    vp.k0[j] = (float)(k > 0.0 ? k + 1.0 : exp(k));
    vp.v0[j] = (float)v;
  }
  RayParams& rp = t->rp;
  memcpy(rp.n1w, d->ray.norm1_w, 352); memcpy(rp.n1b, d->ray.norm1_b, 352);
  memcpy(rp.n2w, d->ray.norm2_w, 352); memcpy(rp.n2b, d->ray.norm2_b, 352);
  memcpy(rp.db0, d->density.b0, 128);
  memcpy(rp.dw2, d->density.w2, sizeof(rp.dw2));
  memcpy(rp.db2, d->density.b2, 64);
  memcpy(rp.dw4, d->density.w4, 64);
  rp.db4 = d->density.b4[0];
  return UFO_OK;
}

extern "C" int ufo_weights_create(const UfoWeightsDesc* d, UfoWeights** out, void* stream_) {
  if (!d || !out) return fail(UFO_EINVAL, "ufo_weights_create: null argument");
  if (int e = check_device()) return e;
  cudaStream_t stream = (cudaStream_t)stream_;
  const UfoLoftrLayer* ll[2] = {&d->view, &d->ray};
  for (int i = 0; i < 2; ++i) {
    const UfoLoftrLayer* l = ll[i];
    if (!l->q || !l->k || !l->v || !l->merge || !l->mlp0 || !l->mlp2 || !l->norm1_w || !l->norm1_b || !l->norm2_w ||
        !l->norm2_b)
      return fail(UFO_EINVAL, "ufo_weights_create: missing transformer tensor");
  }
  const UfoMlp3* mm[3] = {&d->pre_sim, &d->density, &d->radiance};
  for (int i = 0; i < 3; ++i)
    if (!mm[i]->w0 || !mm[i]->b0 || !mm[i]->w2 || !mm[i]->b2 || !mm[i]->w4 || !mm[i]->b4)
      return fail(UFO_EINVAL, "ufo_weights_create: missing MLP tensor");
  if (!d->view_token || !d->depth_freqs || !d->depth_phases) return fail(UFO_EINVAL, "ufo_weights_create: missing tensor");

  UfoWeights* w = new UfoWeights();
  struct Guard {                 // frees the handle on every early return below
    UfoWeights* w;
    ~Guard() { if (w) ufo_weights_destroy(w); }
  } guard{w};
  UFO_CUDA(cudaGetDevice(&w->device));
  BlobBuilder b;
  size_t off[64];
  int k = 0;
  const int dims[2] = {kDView, kDRay};
  for (int i = 0; i < 2; ++i) {
    const int dm = dims[i];
    const UfoLoftrLayer* l = ll[i];
    std::vector<float> qkv((size_t)3 * dm * dm);
    memcpy(qkv.data(), l->q, sizeof(float) * dm * dm);
    memcpy(qkv.data() + (size_t)dm * dm, l->k, sizeof(float) * dm * dm);
    memcpy(qkv.data() + (size_t)2 * dm * dm, l->v, sizeof(float) * dm * dm);
    off[k++] = b.add(qkv.data(), qkv.size());
    off[k++] = b.add(l->merge, (size_t)dm * dm);
    off[k++] = b.add(l->mlp0, (size_t)4 * dm * dm);
    off[k++] = b.add(l->mlp2, (size_t)2 * dm * dm);
    off[k++] = b.add(l->norm1_w, dm);
    off[k++] = b.add(l->norm1_b, dm);
    off[k++] = b.add(l->norm2_w, dm);
    off[k++] = b.add(l->norm2_b, dm);
  }
  const int mdim[3][4] = {{8, 32, 32, 16}, {kDRay, 32, 16, 1}, {kDView + 3, 16, 8, 1}};
  for (int i = 0; i < 3; ++i) {
    off[k++] = b.add(mm[i]->w0, (size_t)mdim[i][0] * mdim[i][1]);
    off[k++] = b.add(mm[i]->b0, mdim[i][1]);
    off[k++] = b.add(mm[i]->w2, (size_t)mdim[i][1] * mdim[i][2]);
    off[k++] = b.add(mm[i]->b2, mdim[i][2]);
    off[k++] = b.add(mm[i]->w4, (size_t)mdim[i][2] * mdim[i][3]);
    off[k++] = b.add(mm[i]->b4, mdim[i][3]);
  }
  off[k++] = b.add(d->view_token, kDView);
  off[k++] = b.add(d->depth_freqs, 8);
  off[k++] = b.add(d->depth_phases, 8);
  // sample-order sinusoid table (ray_transformer.py:165-173), float64 then cast like the reference
  std::vector<float> pe((size_t)kNS * 8);
  for (int i = 0; i < kNS; ++i)
    for (int j = 0; j < 8; ++j) {
      const double ang = (double)i / std::pow(10000.0, 2.0 * (double)(j / 2) / 8.0);
      pe[(size_t)i * 8 + j] = (float)((j % 2 == 0) ? std::sin(ang) : std::cos(ang));
    }
  off[k++] = b.add(pe.data(), pe.size());

  w->blob_floats = b.host.size();
  UFO_CUDA(cudaMalloc(&w->blob, w->blob_floats * sizeof(float)));
  UFO_CUDA(cudaMemcpyAsync(w->blob, b.host.data(), w->blob_floats * sizeof(float), cudaMemcpyHostToDevice, stream));
  UFO_CUDA(cudaStreamSynchronize(stream));
  k = 0;
  LoftrDev* ld[2] = {&w->view, &w->ray};
  for (int i = 0; i < 2; ++i) {
    ld[i]->qkv = w->blob + off[k++];
    ld[i]->merge = w->blob + off[k++];
    ld[i]->mlp0 = w->blob + off[k++];
    ld[i]->mlp2 = w->blob + off[k++];
    ld[i]->n1w = w->blob + off[k++];
    ld[i]->n1b = w->blob + off[k++];
    ld[i]->n2w = w->blob + off[k++];
    ld[i]->n2b = w->blob + off[k++];
  }
  Mlp3Dev* md[3] = {&w->pre_sim, &w->density, &w->radiance};
  for (int i = 0; i < 3; ++i) {
    md[i]->w0 = w->blob + off[k++];
    md[i]->b0 = w->blob + off[k++];
    md[i]->w2 = w->blob + off[k++];
    md[i]->b2 = w->blob + off[k++];
    md[i]->w4 = w->blob + off[k++];
    md[i]->b4 = w->blob + off[k++];
  }
  w->view_token = w->blob + off[k++];
  w->freqs = w->blob + off[k++];
  w->phases = w->blob + off[k++];
  w->pe_table = w->blob + off[k++];
  // SingleVarianceNetwork: exp(10*variance) clipped to [1e-6, 1e6] (single_variance_network.py:11, renderer.py:25)
  w->inv_s = fminf(fmaxf(expf(d->variance * 10.0f), 1e-6f), 1e6f);
  if (int e = tc_weights_build(d, &w->tc, stream)) return e;
  guard.w = nullptr;
  *out = w;
  return UFO_OK;
}

extern "C" void ufo_weights_destroy(UfoWeights* w) {
  if (!w) return;
  for (int f = 0; f < 2; ++f) { cudaFree(w->tc.view_img[f]); cudaFree(w->tc.ray_img[f]); cudaFree(w->tc.view_img2[f]); cudaFree(w->tc.ray_img2[f]); }
  cudaFree(w->blob);
  delete w;
}

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
template <int C>
static int repack_cl(const float* in, float* out, long long S, int N, cudaStream_t st) {
  dim3 grid(cdiv(S, 32), N);
  UFO_KERNEL("k_nchw_to_nhwc<C>", st, k_nchw_to_nhwc<C><<<grid, 256, 0, st>>>(in, out, S));
  return UFO_OK;
}

extern "C" int ufo_scene_create(const UfoSceneDesc* d, UfoScene** out, void* stream_) {
  if (!d || !out) return fail(UFO_EINVAL, "ufo_scene_create: null argument");
  if (int e = check_device()) return e;
  if (d->n_views < 2 || d->n_views > UFO_MAX_VIEWS)
    return fail(UFO_EINVAL, "ufo_scene_create: n_views=%d outside [2,%d]", d->n_views, UFO_MAX_VIEWS);
  if (d->img_h <= 0 || d->img_w <= 0 || d->feat_h <= 0 || d->feat_w <= 0) return fail(UFO_EINVAL, "ufo_scene_create: bad size");
  if ((d->match_feats != nullptr) == (d->match_pairs != nullptr))
    return fail(UFO_EINVAL, "ufo_scene_create: exactly one of match_feats (reference layout) and match_pairs (compact) must be given");
  if (!d->source_imgs || !d->img_feats || !d->depth_info || !d->source_poses || !d->source_poses_inv ||
      !d->ref_pose_inv || !d->w2cs || !d->near_fars || !d->ray_o || !d->ray_d || !d->cam_ray_d)
    return fail(UFO_EINVAL, "ufo_scene_create: null tensor pointer");
  for (int s = 0; s < 3; ++s)
    if (!d->vol_feat[s] || !d->vol_weight[s] || d->vol_d[s] <= 0 || d->vol_h[s] <= 0 || d->vol_w[s] <= 0)
      return fail(UFO_EINVAL, "ufo_scene_create: bad volume %d", s);
  cudaStream_t st = (cudaStream_t)stream_;
  UfoScene* sc = new UfoScene();
  struct Guard {                 // every early return below (the UFO_CUDA / UFO_KERNEL macros included) frees the scene
    UfoScene* s;
    ~Guard() { if (s) ufo_scene_destroy(s); }
  } guard{sc};
  UFO_CUDA(cudaGetDevice(&sc->device));
  SceneDev& D = sc->d;
  const int nv = d->n_views;
  D.nv = nv; D.H = d->img_h; D.W = d->img_w; D.h = d->feat_h; D.w = d->feat_w;
  auto dalloc = [&](size_t bytes, void** p) -> int {
    UFO_CUDA(cudaMalloc(p, bytes));
    sc->owned.push_back(*p);
    sc->bytes += (int64_t)bytes;
    return UFO_OK;
  };
  int e;
  const long long hw = (long long)D.h * D.w, HW = (long long)D.H * D.W;
  float* feat_cl; float4* rgbd; float* match_cl;
  if ((e = dalloc(sizeof(float) * nv * hw * kFeatC, (void**)&feat_cl))) return e;
  if ((e = dalloc(sizeof(float4) * nv * HW, (void**)&rgbd))) return e;
  const int n_pairs = nv * (nv - 1) / 2;
  const bool compact = d->match_pairs != nullptr;
  if ((e = dalloc(sizeof(float) * (compact ? n_pairs : nv * (nv - 1)) * hw * kFeatC, (void**)&match_cl))) return e;
  if ((e = repack_cl<kFeatC>(d->img_feats, feat_cl, hw, nv, st))) return e;
  if ((e = repack_cl<kFeatC>(compact ? d->match_pairs : d->match_feats, match_cl, hw, compact ? n_pairs : nv * (nv - 1), st))) return e;
  UFO_KERNEL("k_pack_rgbd", st, k_pack_rgbd<<<cdiv(HW * nv, 256), 256, 0, st>>>(d->source_imgs, d->depth_info, rgbd, HW, nv));
  D.feat_cl = feat_cl; D.rgbd_cl = rgbd; D.match_cl = match_cl;
  if (compact) {
    D.match_sym = 2;
  } else {  // the reference stores every pair map twice; when the two copies are bit-identical both samples of a
            // pair read the same copy, which halves the working set of the dominant gather at large NV
    int* flag = nullptr;           // owned by the scene (4 bytes), so that no return path can leak it
    if ((e = dalloc(sizeof(int), (void**)&flag))) return e;
    UFO_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    UFO_KERNEL("k_match_sym_check", st, k_match_sym_check<<<dim3(64, nv, nv), 256, 0, st>>>(d->match_feats, nv, hw, flag));
    int differ = 1;
    cudaError_t ce = cudaMemcpyAsync(&differ, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { return fail(UFO_ECUDA, "ufo_scene_create: %s", cudaGetErrorString(ce)); }
    D.match_sym = differ ? 0 : 1;
  }
  {  // slot table: pairs in the reference's enumeration order (model.py:273-276)
    int pi = 0;
    for (int a2 = 0; a2 < nv; ++a2)
      for (int b2 = a2 + 1; b2 < nv; ++b2, ++pi) {
        const int sa = a2 * (nv - 1) + (b2 - 1), sb = b2 * (nv - 1) + a2;
        D.match_slot[a2][b2] = (unsigned char)(compact ? pi : sa);
        D.match_slot[b2][a2] = (unsigned char)(compact ? pi : (D.match_sym ? sa : sb));
      }
  }
  for (int s = 0; s < 3; ++s) {
    D.vd[s] = d->vol_d[s]; D.vh[s] = d->vol_h[s]; D.vw[s] = d->vol_w[s];
    const long long vox = (long long)D.vd[s] * D.vh[s] * D.vw[s];
    float* vf;
    if ((e = dalloc(sizeof(float) * nv * vox * kVolC, (void**)&vf))) return e;
    if ((e = repack_cl<kVolC>(d->vol_feat[s], vf, vox, nv, st))) return e;
    D.vol_feat_cl[s] = vf;
    float* vw;  // single channel: layout already [NV][D][h][w]; copied so the scene owns its inputs
    if ((e = dalloc(sizeof(float) * nv * vox, (void**)&vw))) return e;
    UFO_CUDA(cudaMemcpyAsync(vw, d->vol_weight[s], sizeof(float) * nv * vox, cudaMemcpyDeviceToDevice, st));
    D.vol_w[s] = vw;
  }
  float *rd, *crd;
  if ((e = dalloc(sizeof(float) * 3 * HW, (void**)&rd))) return e;
  if ((e = dalloc(sizeof(float) * 3 * HW, (void**)&crd))) return e;
  UFO_CUDA(cudaMemcpyAsync(rd, d->ray_d, sizeof(float) * 3 * HW, cudaMemcpyDeviceToDevice, st));
  UFO_CUDA(cudaMemcpyAsync(crd, d->cam_ray_d, sizeof(float) * 3 * HW, cudaMemcpyDeviceToDevice, st));
  D.ray_d = rd; D.cam_ray_d = crd;
  for (int v = 0; v < nv; ++v) {
    memcpy(D.P[v], d->source_poses + 16 * v, sizeof(float) * 12);
    memcpy(D.w2c_z[v], d->w2cs + 16 * v + 8, sizeof(float) * 4);
    for (int i = 0; i < 3; ++i) D.cam_o[v][i] = d->source_poses_inv[16 * v + 4 * i + 3];
  }
  for (int i = 0; i < 3; ++i) {
    D.ref_o[i] = d->ref_pose_inv[4 * i + 3];
    D.ray_o[i] = d->ray_o[i];
  }
  D.near0 = d->near_fars[0];
  D.far0 = d->near_fars[1];
  guard.s = nullptr;
  *out = sc;
  return UFO_OK;
}

extern "C" void ufo_scene_destroy(UfoScene* s) {
  if (!s) return;
  for (void* p : s->owned) cudaFree(p);
  cudaFree(s->ws.base);
  cudaFree(s->tws.base);
  cudaFree(s->u_dev);
  cudaFree(s->out_dev);
  if (s->copy_st) {
    cudaEventDestroy(s->copy_ev[0]);
    cudaEventDestroy(s->copy_ev[1]);
    cudaStreamDestroy(s->copy_st);
  }
  delete s;
}

extern "C" int64_t ufo_scene_device_bytes(const UfoScene* s) { return s ? s->bytes : 0; }

// ------------------------------------------------------------------------------------------------
// exact fp32 pipeline
// ------------------------------------------------------------------------------------------------
static int fp32_chunk_rays() {
  const char* e = getenv("UFO_FP32_CHUNK");
  int v = e ? atoi(e) : 2048;
  return v < 32 ? 32 : v;
}

static int ws_ensure(const UfoScene* sc, int rays) {
  Workspace& w = sc->ws;
  const int nv = sc->d.nv, L = nv + 1;
  if (w.base && w.cap_rays >= rays && w.nv == nv) return UFO_OK;
  if (w.base) { cudaFree(w.base); w.base = nullptr; }
  const size_t P = (size_t)rays * kNS;
  auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
  struct Item { float** p; size_t n; };
  float *rgbm_f, *dirs_f;
  Item items[] = {
      {&w.rayinfo, (size_t)rays * 8}, {&w.z_c, (size_t)rays * kNC}, {&w.z_all, P}, {&w.z_fine, (size_t)rays * kNC},
      {&w.XV, P * L * 160}, {&w.QKV, mx(P * L * 240, P * 264)}, {&w.MSG, mx(P * L * 80, P * 88)},
      {&w.MRG, mx(P * L * 80, P * 88)}, {&w.H1, mx(P * L * 160, P * 176)}, {&w.Y2, mx(P * L * 80, P * 88)},
      {&w.VOUT, P * L * 80}, {&w.XR, P * 176}, {&w.ROUT, P * 88}, {&w.sim8, P * 8}, {&rgbm_f, P * nv * 4},
      {&dirs_f, P * nv * 4}, {&w.radiance, P * 4}, {&w.srdf, P}, {&w.weight, P}, {&w.pts, P * 3}};
  size_t total = 0;
  for (auto& it : items) total += (it.n + 63) & ~size_t(63);
  UFO_CUDA(cudaMalloc(&w.base, total * sizeof(float)));
  size_t off = 0;
  for (auto& it : items) { *it.p = w.base + off; off += (it.n + 63) & ~size_t(63); }
  w.rgbm = reinterpret_cast<float4*>(rgbm_f);
  w.dirs = reinterpret_cast<float4*>(dirs_f);
  w.floats = total; w.cap_rays = rays; w.nv = nv;
  return UFO_OK;
}

template <int K, int N, bool R>
static int launch_linear(const float* X, int ldx, const float* W, float* Y, int ldy, long long M, int sms, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)K * N + 64 * K);
  UFO_SMEM_ATTR((k_linear<K, N, R>), (int)smem);
  const long long tiles = (M + 63) / 64;
  const int grid = (int)(tiles < sms ? tiles : sms);
  struct Name {            // built once, thread-safely (function-local static initialisation)
    char s[40];
    Name() { snprintf(s, sizeof(s), "k_linear<%d,%d>", K, N); }
  };
  static const Name name_holder;
  const char* name = name_holder.s;
  UFO_KERNEL(name, st, k_linear<K, N, R><<<grid, 256, smem, st>>>(X, ldx, W, Y, ldy, M));
  return UFO_OK;
}

template <int NV>
static int launch_gather(const UfoScene* sc, const UfoWeights* w, int R, int SN, const float* z, float* pts, cudaStream_t st) {
  const Workspace& ws = sc->ws;
  const long long P = (long long)R * SN;
  UFO_KERNEL("k_gather<NV>", st, k_gather<NV><<<cdiv(P, 32), 256, 0, st>>>(sc->d, ws.rayinfo, z, R, SN, w->freqs, w->phases, ws.XV, ws.sim8, ws.rgbm,
                                           ws.dirs, pts));
  return UFO_OK;
}

static int dispatch_gather(const UfoScene* sc, const UfoWeights* w, int R, int SN, const float* z, float* pts, cudaStream_t st) {
  switch (sc->d.nv) {
    case 2: return launch_gather<2>(sc, w, R, SN, z, pts, st);
    case 3: return launch_gather<3>(sc, w, R, SN, z, pts, st);
    case 4: return launch_gather<4>(sc, w, R, SN, z, pts, st);
    case 5: return launch_gather<5>(sc, w, R, SN, z, pts, st);
    case 6: return launch_gather<6>(sc, w, R, SN, z, pts, st);
    case 7: return launch_gather<7>(sc, w, R, SN, z, pts, st);
    case 8: return launch_gather<8>(sc, w, R, SN, z, pts, st);
    case 9: return launch_gather<9>(sc, w, R, SN, z, pts, st);
    case 10: return launch_gather<10>(sc, w, R, SN, z, pts, st);
  }
  return fail(UFO_EINVAL, "unsupported n_views");
}

// One sample2rgb pass (code1/model.py:308-348) over R rays x SN samples, fp32 CUDA-core arithmetic.
static int pass_fp32(const UfoScene* sc, const UfoWeights* w, int R, int SN, const float* z, int sms, float* pts,
                     cudaStream_t st) {
  const Workspace& ws = sc->ws;
  const int nv = sc->d.nv, L = nv + 1;
  const long long P = (long long)R * SN, MV = P * L;
  int e;
  if ((e = dispatch_gather(sc, w, R, SN, z, pts, st))) return e;
  UFO_KERNEL("k_presim", st, k_presim<<<cdiv(P, 128), 128, 0, st>>>(ws.sim8, w->pre_sim, w->view_token, L, P, ws.XV));
  // ---- view transformer (tokens = views of one sample point)
  if ((e = launch_linear<80, 240, false>(ws.XV, 160, w->view.qkv, ws.QKV, 240, MV, sms, st))) return e;
  UFO_KERNEL("k_linattn<10>", st, k_linattn<10><<<cdiv(P * kHeads, 128), 128, 0, st>>>(ws.QKV, ws.MSG, P, L));
  if ((e = launch_linear<80, 80, false>(ws.MSG, 80, w->view.merge, ws.MRG, 80, MV, sms, st))) return e;
  UFO_KERNEL("k_layernorm<80>", st, k_layernorm<80><<<cdiv(MV, 8), 256, 0, st>>>(ws.MRG, 80, w->view.n1w, w->view.n1b, nullptr, 0, ws.XV + 80, 160, MV));
  if ((e = launch_linear<160, 160, true>(ws.XV, 160, w->view.mlp0, ws.H1, 160, MV, sms, st))) return e;
  if ((e = launch_linear<160, 80, false>(ws.H1, 160, w->view.mlp2, ws.Y2, 80, MV, sms, st))) return e;
  UFO_KERNEL("k_layernorm<80>", st, k_layernorm<80><<<cdiv(MV, 8), 256, 0, st>>>(ws.Y2, 80, w->view.n2w, w->view.n2b, ws.XV, 160, ws.VOUT, 80, MV));
  // ---- ray transformer (tokens = samples of one ray)
  UFO_KERNEL("k_ray_tokens", st, k_ray_tokens<<<cdiv(P * kDRay, 256), 256, 0, st>>>(ws.VOUT, L, SN, P, w->pe_table, ws.XR));
  if ((e = launch_linear<88, 264, false>(ws.XR, 176, w->ray.qkv, ws.QKV, 264, P, sms, st))) return e;
  UFO_KERNEL("k_linattn<11>", st, k_linattn<11><<<cdiv((long long)R * kHeads, 128), 128, 0, st>>>(ws.QKV, ws.MSG, R, SN));
  if ((e = launch_linear<88, 88, false>(ws.MSG, 88, w->ray.merge, ws.MRG, 88, P, sms, st))) return e;
  UFO_KERNEL("k_layernorm<88>", st, k_layernorm<88><<<cdiv(P, 8), 256, 0, st>>>(ws.MRG, 88, w->ray.n1w, w->ray.n1b, nullptr, 0, ws.XR + 88, 176, P));
  if ((e = launch_linear<176, 176, true>(ws.XR, 176, w->ray.mlp0, ws.H1, 176, P, sms, st))) return e;
  if ((e = launch_linear<176, 88, false>(ws.H1, 176, w->ray.mlp2, ws.Y2, 88, P, sms, st))) return e;
  UFO_KERNEL("k_layernorm<88>", st, k_layernorm<88><<<cdiv(P, 8), 256, 0, st>>>(ws.Y2, 88, w->ray.n2w, w->ray.n2b, ws.XR, 176, ws.ROUT, 88, P));
  // ---- heads
  UFO_KERNEL("k_density", st, k_density<<<cdiv(P, 128), 128, 0, st>>>(ws.ROUT, w->density, P, ws.srdf));
  UFO_KERNEL("k_radiance", st, k_radiance<<<cdiv(P, 128), 128, 0, st>>>(ws.VOUT, ws.dirs, ws.rgbm, w->radiance, nv, P, reinterpret_cast<float4*>(ws.radiance)));
  return UFO_OK;
}

__global__ void k_copy_strided(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int cols, long long rows) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * cols) return;
  const long long r = t / cols;
  const int c = (int)(t % cols);
  dst[r * ldd + c] = src[r * lds + c];
}

static int copy_rows(const float* src, int lds, float* dst, int ldd, int cols, long long rows, cudaStream_t st) {
  UFO_KERNEL("k_copy_strided", st, k_copy_strided<<<cdiv(rows * cols, 256), 256, 0, st>>>(src, lds, dst, ldd, cols, rows));
  return UFO_OK;
}

static int render_chunk_fp32(const UfoScene* sc, const UfoWeights* w, const int64_t* ray_idx, int64_t ray_begin, int R,
                             const float* u_c, const float* u_f, int64_t u_stride, int64_t off, const UfoRenderOut* out,
                             const UfoDebugTaps* taps, int sms, cudaStream_t st) {
  const Workspace& ws = sc->ws;
  const int nv = sc->d.nv, L = nv + 1;
  int e;
  UFO_KERNEL("k_ray_setup", st, k_ray_setup<<<cdiv(R, 256), 256, 0, st>>>(sc->d, (const long long*)(ray_idx ? ray_idx + off : nullptr), ray_begin + off, R, ws.rayinfo));
  UFO_KERNEL("k_coarse_z", st, k_coarse_z<<<cdiv((long long)R * kNC, 256), 256, 0, st>>>(ws.rayinfo, u_c + off, u_stride, R, ws.z_c));
  if ((e = pass_fp32(sc, w, R, kNC, ws.z_c, sms, nullptr, st))) return e;
  UFO_KERNEL("k_render<kNC>", st, k_render<kNC><<<cdiv(R, 8), 256, 0, st>>>(ws.z_c, ws.srdf, reinterpret_cast<const float4*>(ws.radiance), w->inv_s, R, ws.weight,
                                          nullptr, nullptr, nullptr, ws.rayinfo));
  if (taps) {
    if (taps->z_coarse) UFO_CUDA(cudaMemcpyAsync(taps->z_coarse + off * kNC, ws.z_c, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->weight_coarse) UFO_CUDA(cudaMemcpyAsync(taps->weight_coarse + off * kNC, ws.weight, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->srdf_coarse) UFO_CUDA(cudaMemcpyAsync(taps->srdf_coarse + off * kNC, ws.srdf, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
  }
  UFO_KERNEL("k_importance", st, k_importance<<<cdiv(R, 8), 256, 0, st>>>(ws.weight, ws.z_c, u_f + off, u_stride, R, ws.z_fine, ws.z_all));
  float* pts = out->points ? out->points + off * kNS * 3 : nullptr;
  if ((e = pass_fp32(sc, w, R, kNS, ws.z_all, sms, pts, st))) return e;
  UFO_KERNEL("k_render<kNS>", st, k_render<kNS><<<cdiv(R, 8), 256, 0, st>>>(ws.z_all, ws.srdf, reinterpret_cast<const float4*>(ws.radiance), w->inv_s, R, ws.weight,
                                          out->depth ? out->depth + off : nullptr, out->rgb ? out->rgb + off * 3 : nullptr,
                                          out->depth_z ? out->depth_z + off : nullptr, ws.rayinfo));
  const long long P = (long long)R * kNS;
  if (out->srdf) UFO_CUDA(cudaMemcpyAsync(out->srdf + off * kNS, ws.srdf, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  if (out->z) UFO_CUDA(cudaMemcpyAsync(out->z + off * kNS, ws.z_all, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  if (taps) {
    const long long po = off * kNS;
    if (taps->z_fine) UFO_CUDA(cudaMemcpyAsync(taps->z_fine + off * kNC, ws.z_fine, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->sim8) UFO_CUDA(cudaMemcpyAsync(taps->sim8 + po * 8, ws.sim8, sizeof(float) * P * 8, cudaMemcpyDeviceToDevice, st));
    if (taps->vol24 && (e = copy_rows(ws.XV + 160 + 32, L * 160, taps->vol24 + po * 24, 24, 24, P, st))) return e;
    if (taps->tokens)
      for (int n = 0; n < nv; ++n)
        if ((e = copy_rows(ws.XV + (size_t)(n + 1) * 160, L * 160, taps->tokens + po * nv * 80 + (size_t)n * 80, nv * 80, 80, P, st))) return e;
    if (taps->view_tok0 && (e = copy_rows(ws.VOUT, L * 80, taps->view_tok0 + po * 80, 80, 80, P, st))) return e;
    if (taps->ray_out) UFO_CUDA(cudaMemcpyAsync(taps->ray_out + po * 88, ws.ROUT, sizeof(float) * P * 88, cudaMemcpyDeviceToDevice, st));
    if (taps->radiance && (e = copy_rows(ws.radiance, 4, taps->radiance + po * 3, 3, 3, P, st))) return e;
    if (taps->weight) UFO_CUDA(cudaMemcpyAsync(taps->weight + po, ws.weight, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  }
  return UFO_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor-core pipeline
// ------------------------------------------------------------------------------------------------
static int tc_chunk_rays(int sms, int nv) {
  // Rays per pass of the tensor-core pipeline.  Larger chunks mean fewer launches and shorter kernel tails: measured on a
  // B200 at 1600x1216, NV=3 (ms per depth map) 64 tiles per CTA 891, 96: 875, 128: 873, 192: 870, 256: 865.  The per-chunk
  // workspace (tws_ensure: ~123 KB per ray at NV=3, ~294 KB at NV=10) is capped at 8 GiB.
  if (const char* e = getenv("UFO_TC_CHUNK")) {
    const int v = atoi(e);
    return v < 2 ? 2 : v;
  }
  const size_t per_ray = (size_t)kNS * (kDView * 4 + 8 + (size_t)nv * (kDView * 2 + 32) + 16 + 32 + 1) + 32 + 2 * kNC * 4;
  const long long cap = (long long)((8ull << 30) / per_ray);
  long long v = (long long)sms * 256;
  if (v > cap) v = cap / sms * sms;
  return v < sms ? sms : (int)v;
}

static int tws_ensure(const UfoScene* sc, int rays) {
  TcWorkspace& w = sc->tws;
  const int nv = sc->d.nv;
  if (w.base && w.cap_rays >= rays && w.nv == nv) return UFO_OK;
  if (w.base) { cudaFree(w.base); w.base = nullptr; }
  const size_t P = (size_t)rays * kNS;
  struct Item { void** p; size_t n; };
  Item items[] = {
      {(void**)&w.rayinfo, (size_t)rays * 8 * 4}, {(void**)&w.z_c, (size_t)rays * kNC * 4}, {(void**)&w.z_all, P * 4},
      {(void**)&w.z_fine, (size_t)rays * kNC * 4}, {(void**)&w.vout0, P * kDView * 4}, {(void**)&w.srdf, P * 4},
      {(void**)&w.weight, P * 4}, {(void**)&w.tok, P * nv * kDView * 2}, {(void**)&w.rgbm, P * nv * 16},
      {(void**)&w.dirs, P * nv * 16}, {(void**)&w.radiance, P * 16}, {(void**)&w.sim8, P * 8 * 4}, {(void**)&w.perm, P}};
  size_t total = 0;
  for (auto& it : items) total += (it.n + 255) & ~size_t(255);
  UFO_CUDA(cudaMalloc(&w.base, total));
  size_t off = 0;
  for (auto& it : items) { *it.p = w.base + off; off += (it.n + 255) & ~size_t(255); }
  w.bytes = total; w.cap_rays = rays; w.nv = nv;
  return UFO_OK;
}

namespace ufo {
int tc_pass(bool bf16, const UfoScene* sc, const UfoWeights* w, int R, int half, const float* z, bool want_sim8, float* ray_out_tap,
            int sms, cudaStream_t st) {
  const bool lo = sc->d.nv <= 5;
  if (bf16) return fail(UFO_EINVAL, "bf16 operands are retired");
  return lo ? tc_pass_f16_lo(sc, w, R, half, z, want_sim8, ray_out_tap, sms, st) : tc_pass_f16_hi(sc, w, R, half, z, want_sim8, ray_out_tap, sms, st);
}
}  // namespace ufo

// debug taps (sorted sample order) of the per-point buffers, which live in evaluation order [ray][128 slots]
template <bool BF16>
__global__ void k_tap_tokens(const uint16_t* __restrict__ tok, const uint8_t* __restrict__ perm, int NV, long long P,
                             float* __restrict__ tokens, float* __restrict__ vol24) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * NV * kDView) return;
  const int c = (int)(t % kDView);
  const long long pn = t / kDView;
  const long long p = pn / NV;
  const int n = (int)(pn % NV);
  const long long src = (p / kNS) * kNS + perm[p];
  const uint32_t u = tok[(src * NV + n) * kDView + tok_pos(c)];
  const float v = BF16 ? __uint_as_float(u << 16) : __half2float(__ushort_as_half((unsigned short)u));
  if (tokens) tokens[t] = v;
  if (vol24 && n == 0 && c >= 32 && c < 56) vol24[p * 24 + (c - 32)] = v;
}

__global__ void k_tap_rows(const float* __restrict__ src, int lds, const uint8_t* __restrict__ perm, int cols, long long P,
                           float* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P * cols) return;
  const long long p = t / cols;
  const int c = (int)(t % cols);
  dst[t] = src[((p / kNS) * kNS + perm[p]) * lds + c];
}

static int render_chunk_tc(const UfoScene* sc, const UfoWeights* w, bool bf16, const int64_t* ray_idx, int64_t ray_begin, int R,
                           const float* u_c, const float* u_f, int64_t u_stride, int64_t off, const UfoRenderOut* out,
                           const UfoDebugTaps* taps, int sms, cudaStream_t st) {
  const TcWorkspace& ws = sc->tws;
  const int nv = sc->d.nv;
  int e;
  const bool want_sim8 = taps && taps->sim8;
  UFO_KERNEL("k_ray_setup", st, k_ray_setup<<<cdiv(R, 256), 256, 0, st>>>(sc->d, (const long long*)(ray_idx ? ray_idx + off : nullptr), ray_begin + off, R, ws.rayinfo));
  UFO_KERNEL("k_coarse_z", st, k_coarse_z<<<cdiv((long long)R * kNC, 256), 256, 0, st>>>(ws.rayinfo, u_c + off, u_stride, R, ws.z_c));
  // coarse pass: 64 samples per ray -> slots 0..63
  if ((e = tc_pass(bf16, sc, w, R, 0, ws.z_c, want_sim8, nullptr, sms, st))) return e;
  UFO_KERNEL("k_render<kNC>", st, k_render<kNC><<<cdiv(R, 8), 256, 0, st>>>(ws.z_c, ws.srdf, ws.radiance, w->inv_s, R, ws.weight, nullptr, nullptr, nullptr,
                                                                          ws.rayinfo, kNS, nullptr));
  if (taps) {
    if (taps->z_coarse) UFO_CUDA(cudaMemcpyAsync(taps->z_coarse + off * kNC, ws.z_c, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->weight_coarse) UFO_CUDA(cudaMemcpyAsync(taps->weight_coarse + off * kNC, ws.weight, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->srdf_coarse) UFO_CUDA(cudaMemcpyAsync(taps->srdf_coarse + off * kNC, ws.srdf, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
  }
  UFO_KERNEL("k_importance", st, k_importance<<<cdiv(R, 8), 256, 0, st>>>(ws.weight, ws.z_c, u_f + off, u_stride, R, ws.z_fine, ws.z_all, ws.perm));
  // fine pass: only the 64 new samples go through the gathers and the view stage (slots 64..127); the ray stage
  // runs over all 128 samples in sorted order
  const long long po = off * kNS;
  const long long P = (long long)R * kNS;
  float* ray_tap = (taps && taps->ray_out) ? taps->ray_out + po * kDRay : nullptr;
  if ((e = tc_pass(bf16, sc, w, R, 1, ws.z_fine, want_sim8, ray_tap, sms, st))) return e;
  UFO_KERNEL("k_render<kNS>", st, k_render<kNS><<<cdiv(R, 8), 256, 0, st>>>(ws.z_all, ws.srdf, ws.radiance, w->inv_s, R, ws.weight,
                                          out->depth ? out->depth + off : nullptr, out->rgb ? out->rgb + off * 3 : nullptr,
                                          out->depth_z ? out->depth_z + off : nullptr, ws.rayinfo, kNS, ws.perm));
  if (out->points) UFO_KERNEL("k_points", st, k_points<<<cdiv(P, 256), 256, 0, st>>>(sc->d, ws.rayinfo, ws.z_all, P, kNS, out->points + po * 3));
  if (out->srdf) UFO_CUDA(cudaMemcpyAsync(out->srdf + off * kNS, ws.srdf, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  if (out->z) UFO_CUDA(cudaMemcpyAsync(out->z + off * kNS, ws.z_all, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  if (taps) {
    if (taps->z_fine) UFO_CUDA(cudaMemcpyAsync(taps->z_fine + off * kNC, ws.z_fine, sizeof(float) * R * kNC, cudaMemcpyDeviceToDevice, st));
    if (taps->tokens || taps->vol24) {
      float* tk = taps->tokens ? taps->tokens + po * nv * kDView : nullptr;
      float* vl = taps->vol24 ? taps->vol24 + po * 24 : nullptr;
      if (bf16) UFO_KERNEL("k_tap_tokens", st, k_tap_tokens<true><<<cdiv(P * nv * kDView, 256), 256, 0, st>>>(ws.tok, ws.perm, nv, P, tk, vl));
      else UFO_KERNEL("k_tap_tokens", st, k_tap_tokens<false><<<cdiv(P * nv * kDView, 256), 256, 0, st>>>(ws.tok, ws.perm, nv, P, tk, vl));
    }
    if (taps->sim8) UFO_KERNEL("k_tap_rows", st, k_tap_rows<<<cdiv(P * 8, 256), 256, 0, st>>>(ws.sim8, 8, ws.perm, 8, P, taps->sim8 + po * 8));
    if (taps->view_tok0) UFO_KERNEL("k_tap_rows", st, k_tap_rows<<<cdiv(P * kDView, 256), 256, 0, st>>>(ws.vout0, kDView, ws.perm, kDView, P, taps->view_tok0 + po * kDView));
    if (taps->radiance) UFO_KERNEL("k_tap_rows", st, k_tap_rows<<<cdiv(P * 3, 256), 256, 0, st>>>(reinterpret_cast<const float*>(ws.radiance), 4, ws.perm, 3, P, taps->radiance + po * 3));
    if (taps->weight) UFO_CUDA(cudaMemcpyAsync(taps->weight + po, ws.weight, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  }
  return UFO_OK;
}

extern "C" int ufo_render_rays(const UfoScene* sc, const UfoWeights* w, const int64_t* ray_idx, int64_t ray_begin,
                               int32_t n_rays, const float* u_coarse, const float* u_fine, int64_t u_stride, int32_t mode,
                               const UfoRenderOut* out, const UfoDebugTaps* taps, void* stream_) {
  if (!sc || !w || !out || !u_coarse || !u_fine) return fail(UFO_EINVAL, "ufo_render_rays: null argument");
  if (n_rays < 0 || u_stride < n_rays) return fail(UFO_EINVAL, "ufo_render_rays: bad n_rays/u_stride");
  if (mode == UFO_MODE_TC)
    return fail(UFO_EINVAL, "ufo_render_rays: UFO_MODE_TC (bf16 operands) is retired - it misses the depth/colour tolerance at 1600x1216 "
                            "(p99 5.6e-3 of the interval, 46 dB); use UFO_MODE_TC_F16");
  if (mode != UFO_MODE_FP32 && mode != UFO_MODE_TC_F16) return fail(UFO_EINVAL, "ufo_render_rays: unknown mode %d", mode);
  // fp16 operands saturate at +-65504: a checkpoint whose GEMM weights leave that range would be clipped silently - refuse instead
  // (activations saturate too, cvt.rn.satfinite, but cannot be checked ahead of time; UFO_MODE_FP32 has no such limit)
  if (mode == UFO_MODE_TC_F16 && !(w->tc.max_abs_weight <= 65504.f))
    return fail(UFO_EINVAL, "ufo_render_rays: a transformer / head weight of magnitude %g is outside the fp16 range of UFO_MODE_TC_F16; use UFO_MODE_FP32",
                (double)w->tc.max_abs_weight);
  if (!ray_idx && (ray_begin < 0 || ray_begin + n_rays > (int64_t)sc->d.H * sc->d.W))
    return fail(UFO_EINVAL, "ufo_render_rays: ray range [%lld,%lld) outside the %dx%d grid", (long long)ray_begin,
                (long long)(ray_begin + n_rays), sc->d.H, sc->d.W);
  if (n_rays == 0) return UFO_OK;
  if (int e = check_device()) return e;
  int dev = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  if (dev != sc->device || dev != w->device) return fail(UFO_EINVAL, "ufo_render_rays: handles belong to another device");
  cudaStream_t st = (cudaStream_t)stream_;
  int sms = 0;
  UFO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  std::lock_guard<std::mutex> lock(sc->mu);
  if (mode == UFO_MODE_TC || mode == UFO_MODE_TC_F16) {
    int cc_major = 0;
    UFO_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_major != 10) return fail(UFO_ENODEVICE, "ufo_render_rays: the tensor-core modes need an sm_100 device (tcgen05), found sm_%d", cc_major * 10);
    const int chunk = tc_chunk_rays(sms, sc->d.nv);
    if (int e = tws_ensure(sc, n_rays < chunk ? n_rays : chunk)) return e;
    for (int64_t off = 0; off < n_rays; off += chunk) {
      const int R = (int)((n_rays - off) < chunk ? (n_rays - off) : chunk);
      if (int e = render_chunk_tc(sc, w, mode == UFO_MODE_TC, ray_idx, ray_begin, R, u_coarse, u_fine, u_stride, off, out, taps, sms, st)) return e;
    }
    return UFO_OK;
  }
  const int chunk = fp32_chunk_rays();
  if (int e = ws_ensure(sc, n_rays < chunk ? n_rays : chunk)) return e;
  for (int64_t off = 0; off < n_rays; off += chunk) {
    const int R = (int)((n_rays - off) < chunk ? (n_rays - off) : chunk);
    if (int e = render_chunk_fp32(sc, w, ray_idx, ray_begin, R, u_coarse, u_fine, u_stride, off, out, taps, sms, st)) return e;
  }
  return UFO_OK;
}

extern "C" int ufo_render_rays_host(const UfoScene* sc, const UfoWeights* w, int64_t ray_begin, int32_t n_rays,
                                    const float* u_c_host, const float* u_f_host, int32_t mode, float* depth_z_host,
                                    float* rgb_host, void* stream_) {
  if (!sc || !w || !u_c_host || !u_f_host || !depth_z_host || !rgb_host) return fail(UFO_EINVAL, "ufo_render_rays_host: null argument");
  if (n_rays <= 0) return n_rays == 0 ? UFO_OK : fail(UFO_EINVAL, "ufo_render_rays_host: n_rays < 0");
  if (int e = check_device()) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  {
    std::lock_guard<std::mutex> lock(sc->mu);
    if (sc->u_cap < (size_t)n_rays) {
      cudaFree(sc->u_dev); cudaFree(sc->out_dev);
      sc->u_dev = nullptr; sc->out_dev = nullptr; sc->u_cap = 0;
      UFO_CUDA(cudaMalloc(&sc->u_dev, sizeof(float) * 2 * kNC * (size_t)n_rays));
      UFO_CUDA(cudaMalloc(&sc->out_dev, sizeof(float) * 5 * (size_t)n_rays));
      sc->u_cap = (size_t)n_rays;
    }
    if (!sc->copy_st) {
      UFO_CUDA(cudaStreamCreateWithFlags(&sc->copy_st, cudaStreamNonBlocking));
      UFO_CUDA(cudaEventCreateWithFlags(&sc->copy_ev[0], cudaEventDisableTiming));
      UFO_CUDA(cudaEventCreateWithFlags(&sc->copy_ev[1], cudaEventDisableTiming));
    }
  }
  float* u_c = sc->u_dev;
  float* u_f = sc->u_dev + (size_t)kNC * n_rays;
  UfoRenderOut o{};
  o.depth_z = sc->out_dev;
  o.rgb = sc->out_dev + n_rays;
  o.depth = sc->out_dev + 4 * (size_t)n_rays;
  // The sampler uniforms ([64][n_rays] per sampler, 512 B per ray) are the bulk of the transfer: they go up in column
  // blocks on a private copy stream, and the render of block k overlaps the upload of block k+1.
  int sms = 0, dev = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  UFO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t block = (int64_t)(mode == UFO_MODE_FP32 ? fp32_chunk_rays() : tc_chunk_rays(sms, sc->d.nv)) * 4;
  const size_t pitch = sizeof(float) * (size_t)n_rays;
  int k = 0;
  for (int64_t off = 0; off < n_rays; off += block, ++k) {
    const int nb = (int)std::min<int64_t>(block, n_rays - off);
    UFO_CUDA(cudaMemcpy2DAsync(u_c + off, pitch, u_c_host + off, pitch, sizeof(float) * nb, kNC, cudaMemcpyHostToDevice, sc->copy_st));
    UFO_CUDA(cudaMemcpy2DAsync(u_f + off, pitch, u_f_host + off, pitch, sizeof(float) * nb, kNC, cudaMemcpyHostToDevice, sc->copy_st));
    UFO_CUDA(cudaEventRecord(sc->copy_ev[k & 1], sc->copy_st));
    UFO_CUDA(cudaStreamWaitEvent(st, sc->copy_ev[k & 1], 0));
    UfoRenderOut ob = o;
    ob.depth_z = o.depth_z + off;
    ob.rgb = o.rgb + 3 * off;
    ob.depth = o.depth + off;
    if (int e = ufo_render_rays(sc, w, nullptr, ray_begin + off, nb, u_c + off, u_f + off, n_rays, mode, &ob, nullptr, stream_)) return e;
  }
  // cudaMemcpyDefault: the destination is normally (pinned) host memory; a rank of a row-sharded render passes device buffers so
  // that the NCCL gather can follow without a round trip through the host
  UFO_CUDA(cudaMemcpyAsync(depth_z_host, o.depth_z, sizeof(float) * n_rays, cudaMemcpyDefault, st));
  UFO_CUDA(cudaMemcpyAsync(rgb_host, o.rgb, sizeof(float) * 3 * (size_t)n_rays, cudaMemcpyDefault, st));
  UFO_CUDA(cudaStreamSynchronize(st));
  return UFO_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel 1: cost volume
// ------------------------------------------------------------------------------------------------
namespace {
void mat4_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = s;
    }
}
bool mat4_inv(const double* m, double* inv) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][j + 4] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (std::fabs(a[piv][c]) < 1e-300) return false;
    if (piv != c) for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
    const double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int r = 0; r < 4; ++r) if (r != c) { const double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][j + 4];
  return true;
}
// proj [2][4][4] -> K[:3,:3] @ E[:3,:4] in the top rows of E  (TransMVSNet.py:77-80)
void fold_proj(const float* p, double* out) {
  const float* E = p;
  const float* K = p + 16;
  for (int i = 0; i < 16; ++i) out[i] = E[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)K[i * 4 + k] * (double)E[k * 4 + j];
      out[i * 4 + j] = s;
    }
}
}  // namespace

template <int C, int D>
static int launch_costvol(const float* ref, const float* const* src_dev, const WarpMats* mats, const float* hyp, const float* vw_in,
                          const PixelwiseDev& pw, int N, int V, int h, int w, float* sim, float* vw_out, cudaStream_t st) {
  const long long threads = (long long)N * h * w * (C / 4);
  if (vw_in)
    UFO_KERNEL("k_costvol<C, D, false>", st, k_costvol<C, D, false><<<cdiv(threads, 256), 256, 0, st>>>(ref, src_dev, mats, hyp, vw_in, pw, N, V, h, w, sim, vw_out));
  else
    UFO_KERNEL("k_costvol<C, D, true>", st, k_costvol<C, D, true><<<cdiv(threads, 256), 256, 0, st>>>(ref, src_dev, mats, hyp, vw_in, pw, N, V, h, w, sim, vw_out));
  return UFO_OK;
}

static int costvolume_impl(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t h, int32_t w, int32_t D,
                           const std::vector<WarpMats>& mats, const float* hyp, const float* vw_in, const UfoPixelwiseNet* pwn,
                           float* sim, float* vw_out, cudaStream_t st);

static int costvolume_check(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t D, const void* proj, const float* hyp,
                            const float* vw_in, const UfoPixelwiseNet* pwn, float* sim, float* vw_out) {
  if (!feats || !proj || !hyp || !sim) return fail(UFO_EINVAL, "ufo_costvolume_stage: null argument");
  if (!vw_in && (!pwn || !vw_out)) return fail(UFO_EINVAL, "ufo_costvolume_stage: stage 1 needs the pixel-wise net and view_w_out");
  if (V < 2 || V > UFO_MAX_VIEWS || N < 1) return fail(UFO_EINVAL, "ufo_costvolume_stage: bad N/V");
  if (!((C == 32 && D == 48) || (C == 16 && D == 32) || (C == 8 && D == 8)))
    return fail(UFO_EINVAL, "ufo_costvolume_stage: unsupported (C=%d, D=%d); cascade is (32,48),(16,32),(8,8)", C, D);
  return check_device();
}

// Homographies supplied by the caller: rot_trans [N][V-1][12] = rows of rot (9) then trans (3) of src_proj_new . ref_proj_new^-1,
// built by the host exactly as the reference builds them (fp32 torch.matmul / torch.inverse, TransMVSNet.py:77-81, module.py:340-342)
extern "C" int ufo_costvolume_stage_rt(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t h, int32_t w, int32_t D,
                                       const float* rot_trans, const float* hyp, const float* vw_in, const UfoPixelwiseNet* pwn,
                                       float* sim, float* vw_out, void* stream_) {
  if (int e = costvolume_check(feats, N, V, C, D, rot_trans, hyp, vw_in, pwn, sim, vw_out)) return e;
  std::vector<WarpMats> mats((size_t)N * (V - 1));
  for (size_t k = 0; k < mats.size(); ++k) {
    memcpy(mats[k].r, rot_trans + k * 12, sizeof(float) * 9);
    memcpy(mats[k].t, rot_trans + k * 12 + 9, sizeof(float) * 3);
  }
  return costvolume_impl(feats, N, V, C, h, w, D, mats, hyp, vw_in, pwn, sim, vw_out, (cudaStream_t)stream_);
}

// Homographies from the projection matrices, in double on the host (more accurate than the reference's fp32 inverse; use
// ufo_costvolume_stage_rt for bit-level agreement with the reference's own matrices)
extern "C" int ufo_costvolume_stage(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t h, int32_t w, int32_t D,
                                    const float* proj, const float* hyp, const float* vw_in, const UfoPixelwiseNet* pwn,
                                    float* sim, float* vw_out, void* stream_) {
  if (int e = costvolume_check(feats, N, V, C, D, proj, hyp, vw_in, pwn, sim, vw_out)) return e;
  std::vector<WarpMats> mats((size_t)N * (V - 1));
  for (int n = 0; n < N; ++n) {
    double ref[16], ref_inv[16];
    fold_proj(proj + ((size_t)n * V + 0) * 32, ref);
    if (!mat4_inv(ref, ref_inv)) return fail(UFO_EINVAL, "ufo_costvolume_stage: singular reference projection");
    for (int i = 1; i < V; ++i) {
      double src[16], T[16];
      fold_proj(proj + ((size_t)n * V + i) * 32, src);
      mat4_mul(src, ref_inv, T);
      WarpMats& m = mats[(size_t)n * (V - 1) + (i - 1)];
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) m.r[a * 3 + b] = (float)T[a * 4 + b];
        m.t[a] = (float)T[a * 4 + 3];
      }
    }
  }
  return costvolume_impl(feats, N, V, C, h, w, D, mats, hyp, vw_in, pwn, sim, vw_out, (cudaStream_t)stream_);
}

static int costvolume_impl(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t h, int32_t w, int32_t D,
                           const std::vector<WarpMats>& mats, const float* hyp, const float* vw_in, const UfoPixelwiseNet* pwn,
                           float* sim, float* vw_out, cudaStream_t st) {
  PixelwiseDev pw{};
  if (!vw_in) {
    for (int c = 0; c < 16; ++c) {
      const double s = (double)pwn->bn0_w[c] / std::sqrt((double)pwn->bn0_var[c] + 1e-5);
      pw.s0[c] = (float)(s * pwn->conv0_w[c]);
      pw.t0[c] = (float)(pwn->bn0_b[c] - s * pwn->bn0_mean[c]);
    }
    for (int o = 0; o < 8; ++o) {
      const double s = (double)pwn->bn1_w[o] / std::sqrt((double)pwn->bn1_var[o] + 1e-5);
      for (int c = 0; c < 16; ++c) pw.w1[o][c] = (float)(s * pwn->conv1_w[o * 16 + c]);
      pw.t1[o] = (float)(pwn->bn1_b[o] - s * pwn->bn1_mean[o]);
      pw.w2[o] = pwn->conv2_w[o];
    }
    pw.b2 = pwn->conv2_b;
  }
  const size_t fl = (size_t)N * C * h * w;
  float* cl = nullptr;          // V channel-last copies
  WarpMats* mats_dev = nullptr;
  const float** src_ptrs_dev = nullptr;
  AsyncTemps tmp(st);
  int e = UFO_OK;
  if ((e = tmp.alloc((void**)&cl, sizeof(float) * fl * V))) return e;
  if ((e = tmp.alloc((void**)&mats_dev, sizeof(WarpMats) * mats.size()))) return e;
  if ((e = tmp.alloc((void**)&src_ptrs_dev, sizeof(float*) * (V - 1)))) return e;
  for (int i = 0; i < V && !e; ++i) {
    if (C == 32) e = repack_cl<32>(feats[i], cl + fl * i, (long long)h * w, N, st);
    else if (C == 16) e = repack_cl<16>(feats[i], cl + fl * i, (long long)h * w, N, st);
    else e = repack_cl<8>(feats[i], cl + fl * i, (long long)h * w, N, st);
  }
  std::vector<const float*> src_ptrs(V - 1);
  for (int i = 1; i < V; ++i) src_ptrs[i - 1] = cl + fl * i;
  if (!e) {
    cudaError_t ce = cudaMemcpyAsync(mats_dev, mats.data(), sizeof(WarpMats) * mats.size(), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(src_ptrs_dev, src_ptrs.data(), sizeof(float*) * (V - 1), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // host vectors go out of scope below
    if (ce != cudaSuccess) e = fail(UFO_ECUDA, "ufo_costvolume_stage: %s", cudaGetErrorString(ce));
  }
  if (!e) {
    if (C == 32) e = launch_costvol<32, 48>(cl, src_ptrs_dev, mats_dev, hyp, vw_in, pw, N, V, h, w, sim, vw_out, st);
    else if (C == 16) e = launch_costvol<16, 32>(cl, src_ptrs_dev, mats_dev, hyp, vw_in, pw, N, V, h, w, sim, vw_out, st);
    else e = launch_costvol<8, 8>(cl, src_ptrs_dev, mats_dev, hyp, vw_in, pw, N, V, h, w, sim, vw_out, st);
  }
  return e;
}

// ------------------------------------------------------------------------------------------------
// alternative feature grid (row a19)
// ------------------------------------------------------------------------------------------------
extern "C" int ufo_feature_grid(const float* feats, int32_t nv, int32_t h, int32_t w, const float* poses, int32_t reso,
                                const UfoMlp3* lin, float* out, void* stream_) {
  if (!feats || !poses || !lin || !out) return fail(UFO_EINVAL, "ufo_feature_grid: null argument");
  if (!lin->w0 || !lin->b0 || !lin->w2 || !lin->b2 || !lin->w4 || !lin->b4) return fail(UFO_EINVAL, "ufo_feature_grid: missing MLP tensor");
  if (nv < 1 || nv > UFO_MAX_VIEWS || h <= 0 || w <= 0 || reso < 2) return fail(UFO_EINVAL, "ufo_feature_grid: bad size");
  if (int e = check_device()) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  FGridViews V{};
  V.nv = nv;
  for (int v = 0; v < nv; ++v) memcpy(V.P[v], poses + 16 * v, sizeof(float) * 12);
  std::vector<float> wts;
  wts.insert(wts.end(), lin->w0, lin->w0 + 32 * 32);
  wts.insert(wts.end(), lin->b0, lin->b0 + 32);
  wts.insert(wts.end(), lin->w2, lin->w2 + 16 * 32);
  wts.insert(wts.end(), lin->b2, lin->b2 + 16);
  wts.insert(wts.end(), lin->w4, lin->w4 + 8 * 16);
  wts.insert(wts.end(), lin->b4, lin->b4 + 8);
  float *cl = nullptr, *wd = nullptr;
  const size_t fl = (size_t)nv * kFeatC * h * w;
  AsyncTemps tmp(st);
  int e = UFO_OK;
  if ((e = tmp.alloc((void**)&cl, sizeof(float) * fl))) return e;
  if ((e = tmp.alloc((void**)&wd, sizeof(float) * wts.size()))) return e;
  e = repack_cl<kFeatC>(feats, cl, (long long)h * w, nv, st);
  if (!e) {
    cudaError_t ce = cudaMemcpyAsync(wd, wts.data(), sizeof(float) * wts.size(), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);     // host vector goes out of scope below
    if (ce != cudaSuccess) e = fail(UFO_ECUDA, "ufo_feature_grid: %s", cudaGetErrorString(ce));
  }
  if (!e) {
    const long long total = (long long)reso * reso * reso;
    e = [&]() -> int {
      UFO_KERNEL("k_feature_grid", st, k_feature_grid<<<cdiv(total, 128), 128, 0, st>>>(cl, h, w, reso, V, wd, out));
      return UFO_OK;
    }();
  }
  return e;
}

// ------------------------------------------------------------------------------------------------
// TSDF integration (next row N3)
// ------------------------------------------------------------------------------------------------
extern "C" int ufo_tsdf_integrate(const UfoTsdfGrid* g, float* tsdf, float* weight, const UfoTsdfView* views, int32_t n_views,
                                  float obs_weight, void* stream_) {
  if (!g || !tsdf || !weight || (n_views > 0 && !views)) return fail(UFO_EINVAL, "ufo_tsdf_integrate: null argument");
  if (g->dim[0] <= 0 || g->dim[1] <= 0 || g->dim[2] <= 0 || !(g->voxel_size > 0.f) || !(g->trunc_margin > 0.f))
    return fail(UFO_EINVAL, "ufo_tsdf_integrate: bad grid");
  if (n_views < 0) return fail(UFO_EINVAL, "ufo_tsdf_integrate: n_views < 0");
  if (int e = check_device()) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  UFO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = (long long)g->dim[0] * g->dim[1] * g->dim[2];
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
  for (int v0 = 0; v0 < n_views; v0 += kTsdfMaxViews) {
    TsdfLaunch L{};
    L.n_views = std::min(kTsdfMaxViews, n_views - v0);
    for (int i = 0; i < L.n_views; ++i) {
      const UfoTsdfView& s = views[v0 + i];
      if (!s.depth || s.im_h <= 0 || s.im_w <= 0) return fail(UFO_EINVAL, "ufo_tsdf_integrate: bad view %d", v0 + i);
      TsdfViewDev& d = L.v[i];
      d.depth = s.depth; d.im_h = s.im_h; d.im_w = s.im_w;
      d.fx = s.intr[0]; d.cx = s.intr[2]; d.fy = s.intr[4]; d.cy = s.intr[5];
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) d.r[a * 3 + b] = s.pose[a * 4 + b];
        d.t[a] = s.pose[a * 4 + 3];
      }
    }
    UFO_KERNEL("k_tsdf_integrate", st, k_tsdf_integrate<<<grid, 256, 0, st>>>(tsdf, weight, g->dim[0], g->dim[1], g->dim[2], g->origin[0],
                                                                            g->origin[1], g->origin[2], g->voxel_size, g->trunc_margin,
                                                                            obs_weight, L));
  }
  return UFO_OK;
}

// ------------------------------------------------------------------------------------------------
// profiling
// ------------------------------------------------------------------------------------------------
extern "C" int ufo_profile_begin(void) {
  if (int e = check_device()) return e;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof_recs.clear();
  ++g_prof_epoch;
  g_prof_on.store(1);
  return UFO_OK;
}

extern "C" int ufo_profile_end(UfoProfileEntry* out, int32_t cap, int32_t* n_out) {
  if (!n_out || (cap > 0 && !out)) return fail(UFO_EINVAL, "ufo_profile_end: null argument");
  g_prof_on.store(0);
  UFO_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof_mu);
  std::map<std::string, std::pair<long long, double>> acc;
  std::vector<std::string> order;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    auto it = acc.find(r.name);
    if (it == acc.end()) { order.push_back(r.name); it = acc.emplace(r.name, std::make_pair(0LL, 0.0)).first; }
    it->second.first += 1;
    it->second.second += (double)ms;
    g_prof_pool.push_back(r.e0);
    g_prof_pool.push_back(r.e1);
  }
  g_prof_recs.clear();
  ++g_prof_epoch;
  int n = 0;
  for (auto& name : order) {
    if (n >= cap) break;
    UfoProfileEntry& e = out[n++];
    memset(&e, 0, sizeof(e));
    strncpy(e.name, name.c_str(), sizeof(e.name) - 1);
    e.launches = acc[name].first;
    e.ms = acc[name].second;
  }
  *n_out = n;
  return UFO_OK;
}

// ------------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------------
extern "C" int ufo_debug_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode, int32_t bf16,
                                       void* stream_) {
  if (!A || !B || !D) return fail(UFO_EINVAL, "ufo_debug_umma_selftest: null argument");
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 256 || K % 16 || (mode == 1 && K > 128) || mode < 0 || mode > 2 ||
      (mode == 2 && (N > 160 || K > 176)))
    return fail(UFO_EINVAL, "ufo_debug_umma_selftest: need 16<=N<=256, 16<=K<=256 (<=128 in mode 1; N<=160, K<=176 in mode 2), multiples of 16");
  if (int e = check_device()) return e;
  cudaStream_t st = (cudaStream_t)stream_;
  const size_t smem = (size_t)128 * 256 * 2 + (size_t)256 * 256 * 2;
  if (mode == 2) {   // TS form, two independent halves: A [2][128][K], D [2][128][N]
    if (bf16) {
      UFO_CUDA(cudaFuncSetAttribute(k_umma_selftest_ts<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      UFO_KERNEL("k_umma_selftest_ts<true>", st, k_umma_selftest_ts<true><<<1, 256, smem, st>>>(A, B, D, N, K));
    } else {
      UFO_CUDA(cudaFuncSetAttribute(k_umma_selftest_ts<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      UFO_KERNEL("k_umma_selftest_ts<false>", st, k_umma_selftest_ts<false><<<1, 256, smem, st>>>(A, B, D, N, K));
    }
    return UFO_OK;
  }
  if (bf16) {
    UFO_CUDA(cudaFuncSetAttribute(k_umma_selftest<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UFO_KERNEL("k_umma_selftest<true>", st, k_umma_selftest<true><<<1, 128, smem, st>>>(A, B, D, N, K, mode));
  } else {
    UFO_CUDA(cudaFuncSetAttribute(k_umma_selftest<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UFO_KERNEL("k_umma_selftest<false>", st, k_umma_selftest<false><<<1, 128, smem, st>>>(A, B, D, N, K, mode));
  }
  return UFO_OK;
}
