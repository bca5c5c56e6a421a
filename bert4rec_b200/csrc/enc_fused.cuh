// Shared pieces of the fused encoder kernels (k_enc_fused.cu forward, k_enc_fused_bwd.cu backward): tile geometry,
// swizzled-tile accessors, TMEM load wrappers, descriptor builders, the device-side layer table.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace b4r {
namespace encf {
constexpr int FT = 128;                 // rows per CTA tile
constexpr int FH = 64;                  // hidden size
constexpr int FNH = 2, FD = 32;         // heads, head dim
constexpr int TILE_B = FT * 128;        // bytes of one [128][64] bf16 tile
constexpr int NTHR = 512;               // 16 warps: TMEM lane quadrant (warp & 3) x column quarter (warp >> 2)

// shared-memory map (byte offsets from the 1024-aligned base)
constexpr int OFF_X = 0;                        // layer input / residual
constexpr int OFF_Q = OFF_X + TILE_B;           // Q ; later the attention context (A operand of the output projection)
constexpr int OFF_K = OFF_Q + TILE_B;           // K ; later y = LN1 output
constexpr int OFF_V = OFF_K + TILE_B;           // V
constexpr int OFF_P = OFF_V + TILE_B;           // probabilities: FNH x [128][128] ; later h = gelu(FFN1) [128][I]
constexpr int OFF_WA = OFF_P + FNH * 2 * TILE_B;  // Wqkv (3 x 8 KB) + Wo (8 KB)
constexpr int OFF_WB = OFF_WA + 4 * 8192;       // W1 (I/64 x 8 KB) + W2 (I x 128 B), then the small arrays (see kernel)
// per-layer parameter block, in floats (the flat layout keeps these 8 vectors contiguous: api.cu make_layout)
constexpr int PB_BQKV = 0, PB_BO = 192, PB_G1 = 256, PB_BE1 = 320, PB_B1 = 384;   // then b2, g2, be2 at 384+I, 448+I, 512+I
__host__ __device__ constexpr int par_floats(int I) { return 576 + I; }

struct LayerDev {
  const float* pblock;    // bqkv | bo | ln1 gamma | ln1 beta | b1 | b2 | ln2 gamma | ln2 beta
  bf16 *qkv, *ctx, *a_pre, *y, *h_pre, *h, *o_pre, *out;
  float *lse, *mean1, *rstd1, *mean2, *rstd2;
  unsigned long long* keep;
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t byte_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((byte_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bmn(int M, int N) {  // A K-major, B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__host__ __device__ constexpr uint32_t idesc_gen(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// One lane of a converged warp (warp-uniform control flow around the single-thread tcgen05.mma issue)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// Descriptors of any tile / k-slice inside one CTA's shared memory differ from a base descriptor only in the 14-bit
// start-address field ((byte address) >> 4, never carries out for addresses < 256 KB): desc(base + off) = desc(base) + (off >> 4).
__device__ __forceinline__ uint64_t desc_at(uint64_t base_desc, uint32_t byte_off) { return base_desc + (uint64_t)(byte_off >> 4); }

__device__ __forceinline__ void tmem_ld_f32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_f16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  umma::tmem_ld16(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// NC consecutive 16-byte chunks (chunk0..) of one row of a 128-byte-swizzled [128][64] bf16 tile
template <int NC>
__device__ __forceinline__ void st_tile(unsigned char* tile, int row, int chunk0, const uint32_t (&pk)[4 * NC]) {
  unsigned char* rp = tile + row * 128;
#pragma unroll
  for (int q = 0; q < NC; ++q)
    *reinterpret_cast<uint4*>(rp + (((chunk0 + q) ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}
template <int NC>
__device__ __forceinline__ void ld_tile(const unsigned char* tile, int row, int chunk0, float (&v)[8 * NC]) {
  const unsigned char* rp = tile + row * 128;
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(rp + (((chunk0 + q) ^ (row & 7)) << 4));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf162(w[i]); v[8 * q + 2 * i] = f.x; v[8 * q + 2 * i + 1] = f.y; }
  }
}
template <int NW>
__device__ __forceinline__ void st_global(bf16* dst, const uint32_t (&pk)[NW]) {
#pragma unroll
  for (int q = 0; q < NW / 4; ++q)
    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}
template <int N>
__device__ __forceinline__ void pack_n(const float (&v)[N], uint32_t (&pk)[N / 2]) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) pk[i] = pack_bf162(v[2 * i], v[2 * i + 1]);
}
template <int N>
__device__ __forceinline__ void round_n(float (&v)[N], uint32_t (&pk)[N / 2]) {  // v := bf16-rounded v, pk := packed
  pack_n<N>(v, pk);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) { const float2 f = unpack_bf162(pk[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <int N>
__device__ __forceinline__ void zero_if(bool z, uint32_t (&pk)[N]) {
  if (z) {
#pragma unroll
    for (int i = 0; i < N; ++i) pk[i] = 0u;
  }
}
template <int N>
__device__ __forceinline__ void add_vec(float (&v)[N], const float* sp) {   // sp: 16-byte aligned shared-memory vector
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(sp + i);
    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
  }
}
}  // namespace encf


}  // namespace b4r
