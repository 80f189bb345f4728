// Parameter blocks and shared/global memory maps of the tensor-core kernels (host + device).
#pragma once
#include <cstdint>

namespace ufo {
namespace tc {
#ifndef UFO_TC_GROUPS
#define UFO_TC_GROUPS 4
#endif
constexpr int kGroups = UFO_TC_GROUPS;   // column groups per token row: warp w -> TMEM lanes 32*(w%4).., group w/4
constexpr int kThreads = 128 * kGroups;
constexpr uint32_t kChunk = 2048;   // bytes of one 8-column chunk of a 128-row operand tile
}  // namespace tc

struct ViewParams {
  float n1w[80], n1b[80], n2w[80], n2b[80], vtok[80];
  float rb0[16], rw0d[16][3], rw2[8][16], rb2[8], rw4[8], rb4;
  // K' = elu(k) + 1 and V of the view-token row (transformer.py:47, linear_attention.py:36-37 applied to view_token): the same for every
  // sample point, evaluated once on the host from the fp16-rounded operands the QKV GEMM sees (k_view_tc2, fp16 mode)
  float k0[80], v0[80];
  float vtok_x[80];        // view_token in the channel order of the 16-bit token rows (tok_pos): the token row of the X operand tile
};

namespace tc {
// shared-memory map of k_view_tc (bytes)
constexpr uint32_t V_WQKV = 0;                      // [240][80]
constexpr uint32_t V_WMRG = V_WQKV + 240 * 80 * 2;  // [80][80]
constexpr uint32_t V_WML0 = V_WMRG + 80 * 80 * 2;   // [160][160]
constexpr uint32_t V_WML2 = V_WML0 + 160 * 160 * 2; // [80][160]
constexpr uint32_t V_WRAD = V_WML2 + 80 * 160 * 2;  // [16][176]  = [W0x | W0x | W0dir(3) b0 0..]
constexpr uint32_t V_WEND = V_WRAD + 16 * 176 * 2;  // 133632
constexpr uint32_t V_X = V_WEND;                    // chunks 0..9   (token)
constexpr uint32_t V_M = V_X + 10 * kChunk;         // chunks 10..19 (message / LN1 / LN2 output)
constexpr uint32_t V_E = V_M + 10 * kChunk;         // chunks 20..21 (relative direction, 1, 0...) for the radiance head
constexpr uint32_t V_KV = V_E + 2 * kChunk;         // K',V' staging [8 heads][128][40 B]; later H1 (20 chunks)
constexpr uint32_t V_RED = V_KV + 20 * kChunk;      // float2 [kGroups][128]  LayerNorm partials
constexpr uint32_t V_OMG = V_RED + 8 * kGroups * 128;  // float [kGroups][128]   radiance-head partials
constexpr uint32_t V_BAR = V_OMG + 4 * kGroups * 128;
constexpr uint32_t V_RGBM = V_BAR + 64;             // float4 [2][128]: (r,g,b,mask) of the tile's (point, view) pairs, double-buffered
constexpr uint32_t V_SMEM = V_RGBM + 2 * 128 * 16;
static_assert(V_SMEM <= 232448, "view-stage shared memory exceeds the 227 KB opt-in limit");
}  // namespace tc

struct RayParams {
  float n1w[88], n1b[88], n2w[88], n2b[88];
  float db0[32], dw2[16][32], db2[16], dw4[16], db4;
};

namespace tc {
// global-memory image of the ray-stage weights (each block is the shared-memory operand image)
constexpr uint32_t RW_QKV = 0;                        // [272][96]   rows: q 0..87 | k 88..175 | v 176..263 | 0
constexpr uint32_t RW_MRG = RW_QKV + 272 * 96 * 2;    // [96][96]
constexpr uint32_t RW_ML0 = RW_MRG + 96 * 96 * 2;     // [176][176]
constexpr uint32_t RW_ML2 = RW_ML0 + 176 * 176 * 2;   // [96][176]
constexpr uint32_t RW_DEN = RW_ML2 + 96 * 176 * 2;    // [32][96] hi, then [32][96] lo
constexpr uint32_t RW_END = RW_DEN + 2 * 32 * 96 * 2;
// shared-memory map of k_ray_tc (bytes)
constexpr uint32_t R_X = 0;                           // concat [x 11 chunks | m 11 chunks | zero chunk]
constexpr uint32_t R_M = R_X + 11 * kChunk;
constexpr uint32_t R_Q = R_X + 23 * kChunk;           // Q' 11 chunks      (r_hi later: 12 chunks from here)
constexpr uint32_t R_K = R_Q + 11 * kChunk;           // K' 11 chunks      (KVbd0, H1 later)
constexpr uint32_t R_V = R_K + 11 * kChunk;           // V' 12 chunks, chunk 11 = [1,0,...,0] (KVbd1 later)
constexpr uint32_t R_RLO = R_Q + 12 * kChunk;         // r_lo 12 chunks
constexpr uint32_t R_SLOTA = R_V + 12 * kChunk;       // Wqkv / Wmlp0 (streamed)
constexpr uint32_t R_SLOTB = R_SLOTA + 176 * 176 * 2; // Wmerge / Wmlp2 (streamed)
constexpr uint32_t R_WDEN = R_SLOTB + 96 * 176 * 2;   // resident
constexpr uint32_t R_SCR = R_V + 4 * kChunk;          // LayerNorm / SRDF-tail partials: 8 KB inside V' chunks 4..7, dead then
constexpr uint32_t R_BAR = R_WDEN + 2 * 32 * 96 * 2;
constexpr uint32_t R_SMEM = R_BAR + 64;
static_assert(R_SMEM <= 232448, "ray-stage shared memory exceeds the 227 KB opt-in limit");
}  // namespace tc

}  // namespace ufo
