// "Next" row N3 of SURVEY.md section 8f, second half: iso-surface extraction from the fused TSDF volume on the device.
//
// Replaces `measure.marching_cubes_lewiner(tsdf_vol, level=0)` in TSDFVolume.get_mesh / get_point_cloud
// (tsdf_fusion.py:319-356), which copies the three volumes to the host and runs scikit-image's CPU implementation.
// scikit-image is a third-party dependency that is not in this image: parity with ITS triangulation is unpinned; the
// kernels below are bit-identical to oracle/mc_oracle.py (same generated case table, same fp32 operations, same order).
//
// Two passes over the volume, one WARP per z-row of voxels, 32 consecutive voxels (a "group") per step - rows keep x and y
// warp-uniform, so no per-voxel index division is needed and a group's 32 voxels are one coalesced 128-byte line:
//   k_mc_classify  voxel -> crossing flags of its three owned edges (+x,+y,+z) and the case index of the cell it anchors.
//                  Only 1 byte per voxel (the case) and 24 bytes per GROUP leave the kernel: the flags as three 32-bit
//                  ballot masks, and the group's vertex / triangle counts (warp reductions).  cub exclusive scans over
//                  the n/32 group counts give the output offsets; offsets inside a group are popcounts of the masks.
//   k_mc_emit      lane k of a warp inspects group k of the row (one 16-byte record): empty groups end there.  Voxels that
//                  own vertices or triangles go to a per-warp shared-memory queue and are processed 32 at a time with
//                  all lanes busy: vertices (one per sign-changing grid edge, linear interpolation; voxel coordinates
//                  like skimage), normals (interpolated central-difference gradient, unit length, towards larger f)
//                  and indexed faces - a face's vertex id is its owner group's offset plus popcounts of its masks.
// Output order is the oracle's: vertices by owner voxel in C order then axis, faces by cell in C order then table order,
// so results do not depend on the launch geometry.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdint>

#include "../../include/uforecon_b200.h"
#include "ufo_common.cuh"
#include "ufo_mc_table.cuh"

struct UfoMesh {
  int dev = 0;
  int X = 0, Y = 0, Z = 0;
  float level = 0.f;
  const float* tsdf = nullptr;   // borrowed
  long long n_groups = 0;      // X * Y * GZ, GZ = ceil(Z / 32): a group never straddles two z-rows
  int GZ = 0;
  uint8_t* cases = nullptr;      // [n] case index of the cell anchored at the voxel (0 when it anchors none)
  uint4* gmask = nullptr;        // [n_groups] x,y,z: ballot masks "owned edge along axis 0/1/2 crosses"; w: vertex offset
  int32_t* gv = nullptr;         // [n_groups + 1] vertex counts -> exclusive scan (last = total)
  int32_t* gt = nullptr;         // [n_groups + 1] triangle counts -> exclusive scan (last = total)
  void* scan_tmp = nullptr;
  size_t scan_bytes = 0;
  long long n_verts = 0, n_faces = 0;
};

namespace ufo {

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// inside bits of the four corners (x,y) (x+1,y) (x,y+1) (x+1,y+1) at height z of the row starting at `base`; the +x / +y
// neighbours collapse onto the voxel itself on the last x / y plane (offx / offy = 0), which yields "no crossing" there
__device__ __forceinline__ unsigned mc_bits(const float* __restrict__ f, unsigned base, int z, int Z, unsigned offx, unsigned offy,
                                            float level) {
  const unsigned i = base + (unsigned)min(z, Z - 1);
  return (f[i] < level ? 1u : 0u) | (f[i + offx] < level ? 2u : 0u) | (f[i + offy] < level ? 4u : 0u) |
         (f[i + offx + offy] < level ? 8u : 0u);
}

__global__ void __launch_bounds__(256) k_mc_classify(const float* __restrict__ f, int X, int Y, int Z, int GZ, float level,
                                                     uint8_t* __restrict__ cases, uint4* __restrict__ gmask,
                                                     int32_t* __restrict__ gv, int32_t* __restrict__ gt) {
  __shared__ uint8_t s_ntri[256];
  s_ntri[threadIdx.x] = kMcNtri[threadIdx.x];                          // blockDim.x == 256
  __syncthreads();
  const unsigned yz = (unsigned)Y * Z, uz = (unsigned)Z;                // X*Y*Z < 2^31 / 3 (checked by the host)
  const int lane = threadIdx.x & 31;
  const int rows = X * Y, warps = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const int x = row / Y, y = row - x * Y;
    const bool hx = x + 1 < X, hy = y + 1 < Y;
    const unsigned offx = hx ? yz : 0u, offy = hy ? uz : 0u, base = (unsigned)row * uz;
    unsigned b = mc_bits(f, base, lane, Z, offx, offy, level);
    for (int k = 0; k < GZ; ++k) {
      // the next group's bits are loaded one step ahead: its lane 0 is the z+1 neighbour of this group's lane 31
      const unsigned nb = (k + 1 < GZ) ? mc_bits(f, base, 32 * (k + 1) + lane, Z, offx, offy, level) : 0u;
      const int z = 32 * k + lane;
      unsigned up = __shfl_down_sync(0xffffffffu, b, 1);
      const unsigned up31 = __shfl_sync(0xffffffffu, nb, 0);
      if (lane == 31) up = up31;
      if (z + 1 >= Z) up = b;                                             // top plane: no +z neighbour, no crossing
      unsigned fl = 0, cs = 0;
      if (z < Z) {
        fl = ((b ^ (b >> 1)) & 1u) | (((b ^ (b >> 2)) & 1u) << 1) | (((b ^ up) & 1u) << 2);
        cs = (hx && hy && z + 1 < Z) ? (b | (up << 4)) : 0u;             // corner c = dx + 2 dy + 4 dz
        cases[base + (unsigned)z] = (uint8_t)cs;
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, fl & 1u), m1 = __ballot_sync(0xffffffffu, fl & 2u),
                     m2 = __ballot_sync(0xffffffffu, fl & 4u);
      const int nt = __reduce_add_sync(0xffffffffu, (int)s_ntri[cs]);
      if (lane == 0) {
        const long long g = (long long)row * GZ + k;
        gmask[g] = make_uint4(m0, m1, m2, 0u);
        gv[g] = __popc(m0) + __popc(m1) + __popc(m2);
        gt[g] = nt;
      }
      b = nb;
    }
  }
}

// after the scans: the group's vertex offset moves next to its masks so that a face needs ONE 16-byte load per vertex id
__global__ void __launch_bounds__(256) k_mc_pack(uint4* __restrict__ gmask, const int32_t* __restrict__ gv, long long n_groups) {
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (long long)gridDim.x * blockDim.x)
    gmask[g].w = (unsigned)gv[g];
}

// d f / d axis at grid point (x,y,z): central difference, one-sided at the borders (oracle/mc_oracle.py:gradient)
__device__ __forceinline__ float mc_diff(const float* __restrict__ f, unsigned idx, int i, int n, unsigned stride) {
  if (n < 2) return 0.f;
  if (i == 0) return __fsub_rn(f[idx + stride], f[idx]);
  if (i == n - 1) return __fsub_rn(f[idx], f[idx - stride]);
  return __fmul_rn(__fsub_rn(f[idx + stride], f[idx - stride]), 0.5f);
}

// One voxel that owns vertices and / or anchors a cell with triangles, queued by k_mc_emit's scan for dense processing.
struct McItem {
  int row, z;            // row = x * Y + y
  unsigned meta;         // bits 0-2 crossing flags of the owned edges, bits 8-15 case index
  int vbase, tbase;      // first vertex id / first triangle id of this voxel
};

__device__ __forceinline__ void mc_emit_item(const McItem it, const float* __restrict__ f, int X, int Y, int Z, int GZ, float level,
                                             const uint4* __restrict__ gmask, const uint8_t (*s_tri)[16], const uint8_t* s_ntri,
                                             float* __restrict__ verts, float* __restrict__ normals, int32_t* __restrict__ faces) {
  const unsigned yz = (unsigned)Y * Z, uz = (unsigned)Z;
  const int x = it.row / Y, y = it.row - x * Y, z = it.z;
  const unsigned idx = (unsigned)it.row * uz + (unsigned)z;
  const unsigned fl = it.meta & 7u, cs = it.meta >> 8;
  if (fl && verts) {
    long long v = it.vbase;
    const float f0 = f[idx];
    const float g0x = mc_diff(f, idx, x, X, yz), g0y = mc_diff(f, idx, y, Y, uz), g0z = mc_diff(f, idx, z, Z, 1u);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!(fl & (1u << a))) continue;
      const unsigned j = idx + (a == 0 ? yz : (a == 1 ? uz : 1u));
      const float t = __fdiv_rn(__fsub_rn(level, f0), __fsub_rn(f[j], f0));
      const float px = a == 0 ? __fadd_rn((float)x, t) : (float)x, py = a == 1 ? __fadd_rn((float)y, t) : (float)y,
                  pz = a == 2 ? __fadd_rn((float)z, t) : (float)z;
      const float g1x = mc_diff(f, j, x + (a == 0), X, yz), g1y = mc_diff(f, j, y + (a == 1), Y, uz),
                  g1z = mc_diff(f, j, z + (a == 2), Z, 1u);
      const float nx = __fadd_rn(g0x, __fmul_rn(t, __fsub_rn(g1x, g0x))), ny = __fadd_rn(g0y, __fmul_rn(t, __fsub_rn(g1y, g0y))),
                  nz = __fadd_rn(g0z, __fmul_rn(t, __fsub_rn(g1z, g0z)));
      const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
      const float inv = fmaxf(len, 1e-20f);
      verts[3 * v + 0] = px; verts[3 * v + 1] = py; verts[3 * v + 2] = pz;
      if (normals) {
        normals[3 * v + 0] = __fdiv_rn(nx, inv);
        normals[3 * v + 1] = __fdiv_rn(ny, inv);
        normals[3 * v + 2] = __fdiv_rn(nz, inv);
      }
      ++v;
    }
  }
  if (cs) {                                                              // only queued with a case when faces != nullptr
    const unsigned nt = s_ntri[cs];
    const long long o = 3LL * it.tbase;
    for (unsigned i = 0; i < 3u * nt; ++i) {
      const unsigned w = s_tri[cs][i];
      const unsigned a = w & 3u;
      const int zo = z + (int)((w >> 4) & 1u);
      const long long og = ((long long)it.row + ((w >> 2) & 1u) * Y + ((w >> 3) & 1u)) * GZ + (zo >> 5);
      const unsigned ol = (unsigned)zo & 31u;
      const uint4 om = __ldg(gmask + og);
      const unsigned olt = (1u << ol) - 1u;
      const unsigned before = (a > 0 ? ((om.x >> ol) & 1u) : 0u) + (a > 1 ? ((om.y >> ol) & 1u) : 0u);
      faces[o + i] = (int32_t)(om.w + __popc(om.x & olt) + __popc(om.y & olt) + __popc(om.z & olt) + before);
    }
  }
}

// Scan + dense processing.  Per z-row, lane k inspects GROUP k (16-byte mask record + its triangle count from the scanned
// array): empty groups - nearly all of them - cost nothing more.  The voxels of the others that own vertices or
// triangles are appended to a per-warp shared-memory queue (ballot prefix); whenever 32 are waiting they are processed
// with every lane busy.  Output slots were fixed by the scans, so the processing order does not matter.
__global__ void __launch_bounds__(256) k_mc_emit(const float* __restrict__ f, int X, int Y, int Z, int GZ, float level,
                                                 const uint8_t* __restrict__ cases, const uint4* __restrict__ gmask,
                                                 const int32_t* __restrict__ gt,
                                                 float* __restrict__ verts, float* __restrict__ normals, int32_t* __restrict__ faces) {
  // triangle table with each edge id replaced by its owner: bits 0-1 axis, bit 2 dx, bit 3 dy, bit 4 dz of the lower corner
  __shared__ uint8_t s_tri[256][16];
  __shared__ uint8_t s_ntri[256];
  __shared__ McItem s_q[8][64];
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
    const int e = (&kMcTri[0][0])[i];
    const int a = e >> 2, b = e & 3;
    const int dx = (a != 0) ? (b & 1) : 0, dy = (a == 0) ? (b & 1) : ((a == 2) ? (b >> 1) : 0), dz = (a != 2) ? (b >> 1) : 0;
    (&s_tri[0][0])[i] = e < 0 ? 0 : (uint8_t)(a | (dx << 2) | (dy << 3) | (dz << 4));
  }
  s_ntri[threadIdx.x] = kMcNtri[threadIdx.x];
  __syncthreads();
  const unsigned uz = (unsigned)Z;
  const int lane = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  McItem* q = s_q[threadIdx.x >> 5];
  int qn = 0;                                                            // warp-uniform
  const int rows = X * Y, warps = gridDim.x * (blockDim.x >> 5);
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    for (int kc = 0; kc < GZ; kc += 32) {
      const int kk = kc + lane;
      uint4 gm = make_uint4(0u, 0u, 0u, 0u);
      int t0 = 0, tcnt = 0;
      if (kk < GZ) {
        const long long g = (long long)row * GZ + kk;
        gm = __ldg(gmask + g);
        t0 = gt[g];
        tcnt = faces ? gt[g + 1] - t0 : 0;
      }
      unsigned act = __ballot_sync(0xffffffffu, ((gm.x | gm.y | gm.z) && verts) || tcnt);
      while (act) {
        const int j = __ffs(act) - 1;
        act &= act - 1;
        const unsigned m0 = __shfl_sync(0xffffffffu, gm.x, j), m1 = __shfl_sync(0xffffffffu, gm.y, j),
                       m2 = __shfl_sync(0xffffffffu, gm.z, j), vb = __shfl_sync(0xffffffffu, gm.w, j);
        const int tb = __shfl_sync(0xffffffffu, t0, j), tc = __shfl_sync(0xffffffffu, tcnt, j);
        const int z = 32 * (kc + j) + lane;
        const unsigned fl = verts ? (((m0 >> lane) & 1u) | (((m1 >> lane) & 1u) << 1) | (((m2 >> lane) & 1u) << 2)) : 0u;
        const unsigned cs = (tc && z < Z) ? cases[(unsigned)row * uz + (unsigned)z] : 0u;
        const unsigned nt = s_ntri[cs];
        const unsigned b0 = __ballot_sync(0xffffffffu, nt & 1u), b1 = __ballot_sync(0xffffffffu, nt & 2u),
                       b2 = __ballot_sync(0xffffffffu, nt & 4u);
        const bool mine = fl || cs;
        const unsigned am = __ballot_sync(0xffffffffu, mine);
        if (mine) {
          McItem it;
          it.row = row; it.z = z; it.meta = fl | (cs << 8);
          it.vbase = (int)(vb + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt));
          it.tbase = tb + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
          q[qn + __popc(am & lt)] = it;
        }
        qn += __popc(am);
        __syncwarp();
        if (qn >= 32) {
          mc_emit_item(q[lane], f, X, Y, Z, GZ, level, gmask, s_tri, s_ntri, verts, normals, faces);
          qn -= 32;
          McItem keep;
          if (lane < qn) keep = q[32 + lane];
          __syncwarp();
          if (lane < qn) q[lane] = keep;
          __syncwarp();
        }
      }
    }
  }
  if (lane < qn) mc_emit_item(q[lane], f, X, Y, Z, GZ, level, gmask, s_tri, s_ntri, verts, normals, faces);
}

}  // namespace ufo

using namespace ufo;

static void mesh_free(UfoMesh* m) {
  if (!m) return;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(m->dev);
  cudaFree(m->cases); cudaFree(m->gmask); cudaFree(m->gv); cudaFree(m->gt); cudaFree(m->scan_tmp);
  cudaSetDevice(cur);
  delete m;
}

static int mesh_begin(UfoMesh* m, cudaStream_t st) {
  const long long n = (long long)m->X * m->Y * m->Z;
  m->GZ = (m->Z + 31) / 32;
  const long long ng = m->n_groups = (long long)m->X * m->Y * m->GZ;
  UFO_CUDA(cudaMalloc(&m->cases, n));
  UFO_CUDA(cudaMalloc(&m->gmask, ng * sizeof(uint4)));
  UFO_CUDA(cudaMalloc(&m->gv, (ng + 1) * sizeof(int32_t)));
  UFO_CUDA(cudaMalloc(&m->gt, (ng + 1) * sizeof(int32_t)));
  UFO_CUDA(cudaMemsetAsync(m->gv + ng, 0, sizeof(int32_t), st));      // scanning ng + 1 items leaves the totals in the last slot
  UFO_CUDA(cudaMemsetAsync(m->gt + ng, 0, sizeof(int32_t), st));
  UFO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, m->scan_bytes, m->gv, m->gv, (int)(ng + 1), st));
  UFO_CUDA(cudaMalloc(&m->scan_tmp, std::max<size_t>(m->scan_bytes, 16)));
  int sms = 0;
  UFO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->dev));
  const int grid = (int)std::min<long long>(((long long)m->X * m->Y + 7) / 8, (long long)sms * 16);
  UFO_KERNEL("k_mc_classify", st, k_mc_classify<<<grid, 256, 0, st>>>(m->tsdf, m->X, m->Y, m->Z, m->GZ, m->level, m->cases, m->gmask,
                                                                      m->gv, m->gt));
  {
    ProfScope ps("cub_exclusive_sum", st);
    UFO_CUDA(cub::DeviceScan::ExclusiveSum(m->scan_tmp, m->scan_bytes, m->gv, m->gv, (int)(ng + 1), st));
    UFO_CUDA(cub::DeviceScan::ExclusiveSum(m->scan_tmp, m->scan_bytes, m->gt, m->gt, (int)(ng + 1), st));
  }
  UFO_KERNEL("k_mc_pack", st, k_mc_pack<<<(int)std::min<long long>((ng + 255) / 256, (long long)sms * 16), 256, 0, st>>>(m->gmask, m->gv, ng));
  int32_t host[2] = {0, 0};
  UFO_CUDA(cudaMemcpyAsync(&host[0], m->gv + ng, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  UFO_CUDA(cudaMemcpyAsync(&host[1], m->gt + ng, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  UFO_CUDA(cudaStreamSynchronize(st));          // the caller needs the counts to size its output buffers
  m->n_verts = host[0];
  m->n_faces = host[1];
  return UFO_OK;
}

extern "C" int ufo_tsdf_mesh_begin(const UfoTsdfGrid* g, const float* tsdf, float level, UfoMesh** mesh, int64_t* n_verts,
                                   int64_t* n_faces, void* stream) {
  if (!g || !tsdf || !mesh || !n_verts || !n_faces) return fail(UFO_EINVAL, "ufo_tsdf_mesh_begin: null argument");
  *mesh = nullptr;
  if (g->dim[0] <= 0 || g->dim[1] <= 0 || g->dim[2] <= 0) return fail(UFO_EINVAL, "ufo_tsdf_mesh_begin: bad grid");
  const long long n = (long long)g->dim[0] * g->dim[1] * g->dim[2];
  if (n >= (1LL << 31) / 3) return fail(UFO_EINVAL, "ufo_tsdf_mesh_begin: %lld voxels exceed the int32 vertex index range", n);
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    return fail(UFO_ENODEVICE, "no CUDA device visible: libuforecon_b200 has no CPU fallback");
  }
  UfoMesh* m = new UfoMesh();
  UFO_CUDA(cudaGetDevice(&m->dev));
  m->X = g->dim[0]; m->Y = g->dim[1]; m->Z = g->dim[2];
  m->level = level;
  m->tsdf = tsdf;
  if (int e = mesh_begin(m, (cudaStream_t)stream)) {
    mesh_free(m);
    return e;
  }
  *mesh = m;
  *n_verts = m->n_verts;
  *n_faces = m->n_faces;
  return UFO_OK;
}

extern "C" int ufo_tsdf_mesh_emit(UfoMesh* m, float* verts, float* normals, int32_t* faces, void* stream) {
  if (!m) return fail(UFO_EINVAL, "ufo_tsdf_mesh_emit: null mesh handle");
  if (normals && !verts) return fail(UFO_EINVAL, "ufo_tsdf_mesh_emit: normals need verts");
  if ((m->n_verts > 0 && !verts && !faces) || (m->n_faces > 0 && !faces && !verts))
    return fail(UFO_EINVAL, "ufo_tsdf_mesh_emit: no output buffer");
  int dev = 0;
  UFO_CUDA(cudaGetDevice(&dev));
  if (dev != m->dev) return fail(UFO_EINVAL, "ufo_tsdf_mesh_emit: mesh handle belongs to device %d, current device is %d", m->dev, dev);
  if (m->n_verts == 0 && m->n_faces == 0) return UFO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  UFO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)std::min<long long>(((long long)m->X * m->Y + 7) / 8, (long long)sms * 16);
  UFO_KERNEL("k_mc_emit", st, k_mc_emit<<<grid, 256, 0, st>>>(m->tsdf, m->X, m->Y, m->Z, m->GZ, m->level, m->cases, m->gmask, m->gt,
                                                              verts, normals, faces));
  return UFO_OK;
}

extern "C" void ufo_tsdf_mesh_destroy(UfoMesh* m) { mesh_free(m); }
