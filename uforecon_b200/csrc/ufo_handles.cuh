// Handle structs of the C ABI (shared by the translation units of the library).
#pragma once
#include <mutex>
#include <vector>

#include "ufo_common.cuh"
#include "ufo_xfmr_fp32.cuh"
#include "ufo_tc_params.cuh"

struct LoftrDev {
  const float *qkv, *merge, *mlp0, *mlp2, *n1w, *n1b, *n2w, *n2b;
};
// 16-bit operand images + fp32 head parameters of the tensor-core path (one set per operand format)
struct TcWeights {
  uint8_t* view_img[2] = {nullptr, nullptr};   // [0] bf16, [1] fp16   (tc::V_WEND bytes)
  uint8_t* ray_img[2] = {nullptr, nullptr};    //                       (tc::RW_END bytes)
  uint8_t* view_img2[2] = {nullptr, nullptr};  // k_view_tc2 layout     (tc::V2_WEND bytes, ufo_view_tc2.cuh)
  uint8_t* ray_img2[2] = {nullptr, nullptr};   // k_ray_tc2 pieces      (tc::R2W_END bytes, ufo_ray_tc2.cuh)
  ufo::ViewParams vp;
  ufo::RayParams rp;
  float max_abs_weight = 0.f;                   // largest |w| among the GEMM weights packed as 16-bit operands
};

struct UfoWeights {
  int device = -1;
  float* blob = nullptr;  // all fp32 tensors, one allocation
  size_t blob_floats = 0;
  LoftrDev view{}, ray{};
  ufo::Mlp3Dev pre_sim{}, density{}, radiance{};
  const float *view_token = nullptr, *freqs = nullptr, *phases = nullptr, *pe_table = nullptr;  // pe_table [128][8]
  float inv_s = 1.f;
  TcWeights tc;  // bf16 operand images for the tensor-core path
};

struct TcWorkspace {
  uint8_t* base = nullptr;
  size_t bytes = 0;
  int cap_rays = 0, nv = 0;
  float *rayinfo, *z_c, *z_all, *z_fine, *vout0, *srdf, *weight, *sim8;
  uint16_t* tok;
  uint8_t* perm;    // [rays][128] evaluation-order index of each sorted sample (k_importance)
  float4 *rgbm, *dirs, *radiance;
};

struct Workspace {
  float* base = nullptr;
  size_t floats = 0;
  int cap_rays = 0, nv = 0;
  float *rayinfo, *z_c, *z_all, *z_fine, *XV, *QKV, *MSG, *MRG, *H1, *Y2, *VOUT, *XR, *ROUT, *sim8, *radiance, *srdf,
      *weight, *pts;
  float4 *rgbm, *dirs;
};

struct UfoScene {
  int device = -1;
  ufo::SceneDev d{};
  std::vector<void*> owned;
  int64_t bytes = 0;
  mutable Workspace ws;
  mutable TcWorkspace tws;
  mutable float* u_dev = nullptr;      // staging for ufo_render_rays_host
  mutable float* out_dev = nullptr;
  mutable size_t u_cap = 0;
  mutable cudaStream_t copy_st = nullptr;   // uploads of ufo_render_rays_host, overlapped with the render
  mutable cudaEvent_t copy_ev[2] = {nullptr, nullptr};
  mutable std::mutex mu;
};


namespace ufo {
// One sample2rgb pass of the tensor-core pipeline (gather -> view stage -> ray stage); defined per operand
// format and view-count group in ufo_tc_inst_*.cu.
int tc_pass(bool bf16, const UfoScene* sc, const UfoWeights* w, int R, int half, const float* z, bool want_sim8,
            float* ray_out_tap, int sms, cudaStream_t st);
int tc_pass_f16_lo(const UfoScene*, const UfoWeights*, int, int, const float*, bool, float*, int, cudaStream_t);
int tc_pass_f16_hi(const UfoScene*, const UfoWeights*, int, int, const float*, bool, float*, int, cudaStream_t);
}  // namespace ufo
