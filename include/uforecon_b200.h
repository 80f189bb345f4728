/*
 * uforecon_b200.h - C ABI of the B200-native per-ray rendering hot path of UFORecon.
 *
 * The reference (Youngju-Na/UFORecon) is pure Python/PyTorch and has no FFI: the boundary this
 * library replaces is the Python method
 *
 *     UFORecon.infer(batch, ray_idx, source_imgs_feat, feature_volume, extract_geometry=True,
 *                    match_feature, ...)  ->  (srdf, points_x_all, depth, rgb)
 *                                                          code1/model.py:393-478
 *
 * as called per ray chunk by UFORecon.extract_geometry (code1/model.py:814-823), plus the
 * cost-volume build inside DepthNet.forward (code1/encoder_utils/fmt/TransMVSNet.py:61-100).
 * Each entry point below cites the reference code whose work it performs.  INTEGRATION.md shows
 * the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative UFO_E* code on failure; ufo_last_error()
 *     returns a thread-local message.  No exceptions or torch types cross this boundary.
 *   - pointers marked [dev] are device pointers (fp32 unless stated), [host] are host pointers.
 *   - all tensors are contiguous in the reference's own layouts (NCHW / NCDHW), batch dim B=1
 *     squeezed; the library repacks what it needs into its own buffers (owned by the handle).
 *   - calls are asynchronous on the supplied cudaStream_t (passed as void*) except where noted;
 *     handles are bound to the CUDA device that was current at creation.
 *   - there is no CPU fallback: without a CUDA device every call fails with UFO_ENODEVICE.
 */
#ifndef UFORECON_B200_H_
#define UFORECON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UFO_ABI_VERSION 2
#define UFO_MAX_VIEWS 10
#define UFO_N_STAGES 3
#define UFO_N_COARSE 64 /* --test_sample_coarse, main.py:73 */
#define UFO_N_FINE 64   /* --test_sample_fine,   main.py:74 */
#define UFO_N_SAMPLES (UFO_N_COARSE + UFO_N_FINE)

enum {
  UFO_OK = 0,
  UFO_EINVAL = -1,    /* bad argument / unsupported configuration */
  UFO_ENODEVICE = -2, /* no CUDA device / wrong architecture      */
  UFO_ECUDA = -3,     /* CUDA runtime error (message in ufo_last_error) */
  UFO_ENOMEM = -4
};

/* Transformer arithmetic of ufo_render_rays. */
enum {
  UFO_MODE_FP32 = 0, /* CUDA-core fp32 everywhere (parity: 1e-5 relative)                        */
  UFO_MODE_TC = 1,   /* RETIRED (round 2): BF16 operands on the tensor cores.  Measured p99 5.6e-3 of the interval /
                      * 46 dB at 1600x1216 - outside the tolerance - so ufo_render_rays returns UFO_EINVAL for it.  */
  UFO_MODE_TC_F16 = 2 /* tcgen05 tensor cores, FP16 operands / FP32 accumulate (packing saturates at +-65504):
                       * inside the tolerance (p99 depth error <= 0.5 % of the interval, PSNR >= 50 dB) at
                       * 1600x1216.  The throughput mode.                                              */
};

typedef struct UfoScene UfoScene;     /* one view set: repacked source tensors + cameras        */
typedef struct UfoWeights UfoWeights; /* packed ray_transformer.* / deviation_network weights   */

/* Inputs of infer() that describe one view set.  Spec: DtuFitSparse.__getitem__
 * (code1/dataset/dtu_test_sparse.py:382-436) + the encoder outputs assembled in
 * extract_geometry (code1/model.py:777-808). */
typedef struct {
  int32_t n_views;        /* NV, 2..UFO_MAX_VIEWS                                   */
  int32_t img_h, img_w;   /* H, W of source images == ray grid                      */
  int32_t feat_h, feat_w; /* h, w of the stage-1 feature / match maps (H/4, W/4)     */
  const float* source_imgs;  /* [dev] batch['source_imgs'][0]      [NV,3,H,W]           */
  const float* img_feats;    /* [dev] source_imgs_feat[0]          [NV,32,h,w]          */
  const float* depth_info;   /* [dev] batch['depth_info'][0]       [NV,H,W]             */
  const float* match_feats;  /* [dev] match_feature[0][0]          [NV,(NV-1)*32,h,w]  (NULL when match_pairs is given) */
  const float* vol_feat[UFO_N_STAGES];   /* [dev] feature_volume[stage]['feature_volume'] [NV,8,D,hs,ws] */
  const float* vol_weight[UFO_N_STAGES]; /* [dev] feature_volume[stage]['weight_volume']  [NV,1,D,hs,ws] */
  int32_t vol_d[UFO_N_STAGES], vol_h[UFO_N_STAGES], vol_w[UFO_N_STAGES];
  const float* source_poses;     /* [host] batch['source_poses'][0]      [NV,4,4] world->NDC */
  const float* source_poses_inv; /* [host] batch['source_poses_inv'][0]  [NV,4,4]            */
  const float* ref_pose_inv;     /* [host] batch['ref_pose_inv'][0]      [4,4]               */
  const float* w2cs;             /* [host] batch['w2cs'][0]              [NV,4,4]            */
  const float* near_fars;        /* [host] batch['near_fars'][0]         [NV,2]              */
  const float* ray_o;            /* [host] batch['ray_o'][0]             [3]                 */
  const float* ray_d;            /* [dev]  batch['ray_d'][0]             [3,H*W]             */
  const float* cam_ray_d;        /* [dev]  batch['cam_ray_d'][0]         [3,H*W]             */
  /* ABI 2: the compact pair maps, either/or with match_feats (exactly one of the two is non-NULL).  The reference's
   * get_match_feat stores the map of every unordered view pair twice (FMT.py:197,308-309,
   * TransMVSNet.py:362-366); an encoder that emits each pair once passes [NV(NV-1)/2, 32, h, w] here, pairs in the
   * reference's enumeration order (a, b) for a in 0..NV-2, b in a+1..NV-1 (model.py:273-276). */
  const float* match_pairs;      /* [dev]  [NV(NV-1)/2,32,h,w] or NULL */
} UfoSceneDesc;

/* Hot-path block of the checkpoint state dict; all [host] fp32, torch layout
 * ([out,in] for Linear weights).  Reference modules: code1/ray_transformer.py:127-163,
 * code1/attention/transformer.py:17-33, code1/encoder_utils/single_variance_network.py:8. */
typedef struct {
  const float *q, *k, *v, *merge; /* [d,d]   */
  const float *mlp0;              /* [2d,2d] */
  const float *mlp2;              /* [d,2d]  */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b; /* [d] */
} UfoLoftrLayer;

typedef struct {
  const float *w0, *b0, *w2, *b2, *w4, *b4;
} UfoMlp3;

typedef struct {
  UfoLoftrLayer view;    /* density_view_transformer.layers.0, d = 80 */
  UfoLoftrLayer ray;     /* density_ray_transformer.layers.0,  d = 88 */
  UfoMlp3 pre_sim;       /* pre_sim_mlp            8 -> 32 -> 32 -> 16 */
  UfoMlp3 density;       /* DensityMLP            88 -> 32 -> 16 -> 1  */
  UfoMlp3 radiance;      /* linear_radianceweight_1_softmax 83 -> 16 -> 8 -> 1 */
  const float* view_token;      /* viewToken.view_token [80]       */
  const float* depth_freqs;     /* depthcode._freqs  [8]           */
  const float* depth_phases;    /* depthcode._phases [8]           */
  float variance;               /* deviation_network.variance      */
} UfoWeightsDesc;

/* Optional device buffers that receive intermediate tensors of the FINE pass (second sample2rgb
 * call, 128 samples) and the coarse pass; any pointer may be NULL.  Used by the parity tests at the
 * reference's sub-boundaries. */
typedef struct {
  float* z_coarse;      /* [n,64]        FixedSampler z                 sampler.py:15-50   */
  float* weight_coarse; /* [n,64]        renderer weights, coarse pass  renderer.py:41     */
  float* srdf_coarse;   /* [n,64]                                                          */
  float* z_fine;        /* [n,64]        ImportanceSampler z (sorted)   sampler.py:74-108  */
  float* sim8;          /* [n,128,8]     query_cond_info feat_info      model.py:218-305   */
  float* vol24;         /* [n,128,24]    query_depth_from_volume        model.py:350-390   */
  float* tokens;        /* [n,128,NV,80] view tokens before the view transformer           */
  float* view_tok0;     /* [n,128,80]    view-transformer output, token 0                  */
  float* ray_out;       /* [n,128,88]    ray-transformer output                            */
  float* radiance;      /* [n,128,3]     blended colour per sample                         */
  float* weight;        /* [n,128]       renderer weights, fine pass                       */
} UfoDebugTaps;

typedef struct {
  float* depth;   /* [dev] [n]      ray-distance depth  (infer's depth_2, model.py:478)        */
  float* depth_z; /* [dev] [n]      depth * cam_ray_d.z (extract_geometry, model.py:818-821)   */
  float* rgb;     /* [dev] [n,3]                                                               */
  float* srdf;    /* [dev] [n,128]  may be NULL                                                */
  float* z;       /* [dev] [n,128]  sorted sample distances; may be NULL                       */
  float* points;  /* [dev] [n,128,3] may be NULL                                               */
} UfoRenderOut;

int ufo_abi_version(void);
const char* ufo_last_error(void);

/* Device properties the host side needs (SM count for sharding heuristics). */
int ufo_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* Pack the hot-path weights (replaces nn.Module state of RayTransformer / SingleVarianceNetwork). */
int ufo_weights_create(const UfoWeightsDesc* desc, UfoWeights** out, void* stream);
void ufo_weights_destroy(UfoWeights* w);

/* Repack one view set (replaces nothing in the reference - it keeps NCHW tensors; the fused
 * gathers want channel-last texels).  Synchronous w.r.t. `stream` ordering only. */
int ufo_scene_create(const UfoSceneDesc* desc, UfoScene** out, void* stream);
void ufo_scene_destroy(UfoScene* s);
/* Bytes of device memory owned by the scene (for DESIGN.md's layout accounting). */
int64_t ufo_scene_device_bytes(const UfoScene* s);

/* UFORecon.infer(extract_geometry=True) for `n_rays` rays (code1/model.py:393-478):
 * coarse sampling (sampler.py:15-50) -> sample2rgb (model.py:308-348) -> importance sampling
 * (sampler.py:74-108) -> merge+sort (model.py:466-470) -> sample2rgb on all 128 samples.
 *   ray_idx   [dev] int64 [n_rays] pixel indices h*W+w, or NULL for the range
 *             [ray_begin, ray_begin+n_rays)
 *   u_coarse  [dev] [64, u_stride]  uniforms of FixedSampler's jitter, column i <-> ray i
 *   u_fine    [dev] [64, u_stride]  uniforms of ImportanceSampler (reference draws [64,RN] and
 *             transposes, sampler.py:86); u_stride >= n_rays is the row pitch in floats
 *   mode      UFO_MODE_FP32 | UFO_MODE_TC_F16
 */
int ufo_render_rays(const UfoScene* scene, const UfoWeights* weights, const int64_t* ray_idx,
                    int64_t ray_begin, int32_t n_rays, const float* u_coarse, const float* u_fine,
                    int64_t u_stride, int32_t mode, const UfoRenderOut* out,
                    const UfoDebugTaps* taps, void* stream);

/* Same call with HOST buffers: copies uniforms host->device, renders, copies depth_z/rgb back and
 * synchronises the stream.  u_* are [64,n_rays] pinned or pageable host memory; depth_z [n_rays],
 * rgb [n_rays,3] host (device pointers are accepted too: the copies use cudaMemcpyDefault - a rank of a row-sharded render
 * keeps its block on the device for the NCCL gather).  Rays are the range [ray_begin, ray_begin+n_rays). */
int ufo_render_rays_host(const UfoScene* scene, const UfoWeights* weights, int64_t ray_begin,
                         int32_t n_rays, const float* u_coarse_host, const float* u_fine_host,
                         int32_t mode, float* depth_z_host, float* rgb_host, void* stream);

/* The library's temporaries (repacked feature maps of ufo_costvolume_stage*, staging of ufo_scene_create) come from the
 * device's default stream-ordered memory pool, which the library keeps cached between calls (release threshold raised on
 * first use; the default threshold of 0 re-maps gigabytes on every call).  ufo_trim_pool() synchronises the device and
 * returns the cached memory to the driver. */
int ufo_trim_pool(void);

/* Number of kernel launches issued by this library since process start (bench.py's gpu_launches). */
int64_t ufo_launch_count(void);

/* Per-kernel device-time accounting for bench.py's roofline: between ufo_profile_begin() and
 * ufo_profile_end() every kernel the library launches is bracketed by CUDA events on its launching
 * stream; ufo_profile_end() synchronises the device and returns, per kernel name, the launch count and
 * the summed event time.  Process-global; not meant to be left on in production. */
typedef struct {
  char name[48];
  int64_t launches;
  double ms;
} UfoProfileEntry;
int ufo_profile_begin(void);
int ufo_profile_end(UfoProfileEntry* out, int32_t cap, int32_t* n_out);

/* Cost-volume build of one cascade stage for all N reference rotations (replaces the loop of
 * DepthNet.forward, TransMVSNet.py:76-100, with homo_warping_trans, fmt/module.py:329-367 and
 * PixelwiseNet, TransMVSNet.py:23-41; the 3-D CNN regulariser stays in PyTorch).
 *   feats       [dev] [V][N,C,h,w]   V = views per rotation; feats[0] = reference features
 *   proj        [host] [N,V,2,4,4]   proj_matrices of this stage
 *   depth_hyp   [dev] [N,D,h,w]
 *   view_w_in   [dev] [N,V-1,h,w] or NULL (stage 1: computed from pixelwise net)
 *   pw          [host] folded PixelwiseNet parameters, see UfoPixelwiseNet; ignored if view_w_in
 *   similarity  [dev] [N,1,D,h,w]  out
 *   view_w_out  [dev] [N,V-1,h,w]  out (written only when view_w_in == NULL)
 */
typedef struct {
  const float *conv0_w;                              /* [16]  (1->16, 1x1x1, no bias)            */
  const float *bn0_w, *bn0_b, *bn0_mean, *bn0_var;   /* [16]                                     */
  const float *conv1_w;                              /* [8,16]                                   */
  const float *bn1_w, *bn1_b, *bn1_mean, *bn1_var;   /* [8]                                      */
  const float *conv2_w;                              /* [8]                                      */
  float conv2_b;
} UfoPixelwiseNet;

int ufo_costvolume_stage(const float* const* feats, int32_t n_rot, int32_t n_views, int32_t channels,
                         int32_t h, int32_t w, int32_t n_depth, const float* proj,
                         const float* depth_hyp, const float* view_w_in, const UfoPixelwiseNet* pw,
                         float* similarity, float* view_w_out, void* stream);

/* The same stage with the homographies supplied by the caller: rot_trans [host] [N][V-1][12] = rot (3x3, row-major) then
 * trans (3) of src_proj_new . inverse(ref_proj_new), built the way the reference builds them (fp32 torch.matmul /
 * torch.inverse, TransMVSNet.py:77-81, fmt/module.py:340-342) so that the warp coordinates agree with the reference's to
 * rounding.  ufo_costvolume_stage derives them from proj in double instead. */
int ufo_costvolume_stage_rt(const float* const* feats, int32_t N, int32_t V, int32_t C, int32_t h, int32_t w,
                            int32_t D, const float* rot_trans, const float* depth_hyp, const float* view_w_in,
                            const UfoPixelwiseNet* pwn, float* sim_out, float* view_w_out, void* stream);

/* Alternative feature grid of --volume_type featuregrid (row a19): FeatureVolume.forward up to its 3-D regulariser
 * (code1/feature_volume.py:40-92).  feats [dev] [NV,32,h,w]; source_poses [host] [NV,4,4] world->NDC; linear = the
 * module's nn.Sequential (32->32 ReLU, 32->16 ReLU, 16->8; torch [out,in] layout, host); out [dev] [16,reso,reso,reso]
 * in the (C, Z, Y, X) order the reference hands to VolumeRegularization: channels 0..7 masked mean over views,
 * 8..15 masked variance. */
int ufo_feature_grid(const float* feats, int32_t n_views, int32_t h, int32_t w, const float* source_poses, int32_t reso,
                     const UfoMlp3* linear, float* out, void* stream);

/* TSDF integration of depth maps ("next" row N3; replaces the PyCUDA kernel `integrate`, tsdf_fusion.py:77-152, and
 * the per-view loop of save_tsdf, tsdf_fusion.py:486-502).  tsdf / weight are [dev] fp32 volumes [X,Y,Z] (Z fastest,
 * like the reference's numpy arrays), updated in place; views are integrated in array order with the reference's
 * running average.  intr = cam_intr row-major 3x3, pose = camera-to-world 4x4 row-major (what the reference passes
 * as cam_pose = inv(extrinsic)).  The reference's colour volume is never written by its kernel and has no entry here. */
typedef struct {
  int32_t dim[3];      /* voxels along x, y, z                                   */
  float origin[3];     /* world position of voxel (0,0,0)  (TSDFVolume._vol_origin) */
  float voxel_size;
  float trunc_margin;  /* margin * voxel_size                                    */
} UfoTsdfGrid;
typedef struct {
  const float* depth;  /* [dev] [im_h, im_w], 0 = invalid */
  int32_t im_h, im_w;
  float intr[9];
  float pose[16];
} UfoTsdfView;
int ufo_tsdf_integrate(const UfoTsdfGrid* grid, float* tsdf, float* weight, const UfoTsdfView* views, int32_t n_views,
                       float obs_weight, void* stream);

/* Iso-surface extraction from a TSDF volume on the device ("next" row N3, second half; replaces
 * skimage.measure.marching_cubes_lewiner(tsdf_vol, level=0) in TSDFVolume.get_mesh / get_point_cloud,
 * tsdf_fusion.py:319-356).  Marching cubes on the cell grid: one vertex per grid edge whose end points lie on different
 * sides of `level` (f < level = inside), placed by linear interpolation, in VOXEL coordinates like skimage's output (the
 * caller applies verts * voxel_size + vol_origin, tsdf_fusion.py:347); unit normals from the interpolated central-
 * difference gradient, pointing towards larger f; indexed triangles wound so that their normal points the same way.
 * Order: vertices by owner voxel in C order then edge axis, faces by cell in C order - independent of the launch.
 * Two calls because the sizes are data dependent:
 *   ufo_tsdf_mesh_begin  classifies the volume, returns the counts (synchronises `stream`) and a handle that BORROWS
 *                        `tsdf` - the volume must stay valid and unchanged until the handle is destroyed;
 *   ufo_tsdf_mesh_emit   fills verts [n_verts,3] f32, normals [n_verts,3] f32 (may be NULL), faces [n_faces,3] i32
 *                        (may be NULL; verts may be NULL when only faces are wanted), all [dev]; asynchronous on `stream`.
 * scikit-image is not vendored by the reference: parity with its triangulation of ambiguous cells is unpinned; the
 * vertex set is the same by construction (oracle/mc_oracle.py). */
typedef struct UfoMesh UfoMesh;
int ufo_tsdf_mesh_begin(const UfoTsdfGrid* grid, const float* tsdf, float level, UfoMesh** mesh, int64_t* n_verts,
                        int64_t* n_faces, void* stream);
int ufo_tsdf_mesh_emit(UfoMesh* mesh, float* verts, float* normals, int32_t* faces, void* stream);
void ufo_tsdf_mesh_destroy(UfoMesh* mesh);

/* Diagnostics: one-CTA tcgen05 GEMM through the library's own operand-staging and descriptor helpers
 * (csrc/ufo_umma.cuh).  mode 0: D[128,N] = A[128,K] . B[N,K]^T; mode 1: D = At[K,128]^T . Bt[K,N];
 * mode 2: two independent halves of one CTA, A operand in TMEM: D[2,128,N] = A[2,128,K] . B[N,K]^T (N<=160, K<=176).
 * All [dev] fp32; operands are rounded to fp16 (bf16 != 0: bf16) on the way to shared memory. */
int ufo_debug_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode,
                            int32_t bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UFORECON_B200_H_ */
