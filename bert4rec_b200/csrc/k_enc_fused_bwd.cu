// Generation 2 of the encoder stack BACKWARD (hidden 64, 2 heads): all layers, last to first, in ONE launch.  Same
// tile ownership as the fused forward (k_enc_fused.cu): a CTA owns 128 token rows = 128/slot whole sequences; the
// gradient with respect to the residual stream never leaves the SM -- it lives in registers (one row x 16 columns per
// thread, fp32) from the top of the stack down to the embedding output.  Per layer, six tcgen05 batches:
//
//   E0  LN2 backward + dropout' (registers)                           -> d_branch2 tile ; h tile (saved activation)
//   M1  dW2 = h^T d_branch2            | dh = d_branch2 W2^T
//   E1  drain dW2 ; dh * gelu'(h_pre)                                 -> dh_pre tile ; y tile
//   M2  dW1 = y^T dh_pre               | dy = dh_pre W1^T
//   E2  drain dW1 ; (dy + residual grad) -> LN1 backward + dropout'   -> d_branch1 tile ; ctx tile
//   M3  dWo = ctx^T d_branch1          | dctx = d_branch1 Wo^T
//   E3  drain dWo ; dctx -> bf16, delta = rowsum(dctx * ctx) per head  -> dctx tile ; Q, K, V tiles
//   M4  S = Q K^T, dP = dctx V^T  (both heads; probabilities are recomputed from the saved log-sum-exp)
//   E4  P = exp(S*scale + mask - lse), dropout keep bits, dS = P (dP' - delta)   -> P_drop, dS tiles (bf16)
//   M5  dV = P_drop^T dctx, dK = dS^T Q, dQ = dS K   (both heads)
//   E5  scale, bf16                                                    -> dQ | dK | dV tiles ; x_in tile
//   M6  dWqkv = x_in^T dqkv            | dx = dqkv Wqkv^T
//   E6  drain dWqkv ; d(residual) = dx + d_a_pre  (registers -> next layer)
//
// Transposed operands are never copied: the same 128-byte-swizzled tile is read K-major by one MMA and MN-major by
// another (descriptor major bits).  Weight matrices arrive by TMA into tiles that are dead at that point of the layer.
// Weight-gradient / bias-gradient partials are written per CTA (deterministic) and summed by grad_reduce_kernel.
// Replaces, for the supported shapes, 12 launches per layer of the layered path (ln_bwd, wgrad x4, gemm dgrad x4,
// attn_bwd, colsum); reference op: tape.gradient through tfm TransformerEncoderBlock (bert4rec_model.py:166-167).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "enc_fused.cuh"

namespace b4r {
using namespace encf;

namespace {
constexpr int NT_ARENA = 12;   // 16 KB tiles
constexpr int ONES_OFF = NT_ARENA * TILE_B, ONES_BYTES = 2048;   // 16 k-rows x 64 bf16 ones behind the arena (see `dmn_ones`)
// arena tile indices by phase (see the liveness table in DESIGN.md section 4b)
constexpr int T_H = 0, T_DB2 = 2, T_DH = 3, T_Y = 5, T_DB1 = 6, T_C = 7, T_W2 = 8, T_W1 = 9, T_WO = 10;
constexpr int T_Q = 0, T_K = 1, T_V = 2, T_DC = 3, T_PD = 4, T_DS = 8;       // P_drop: 4..7, dS: 8..11 (head-major, 2 tiles each)
constexpr int T_DQ = 4, T_XIN = 7, T_WQKV = 8;                                // dQ|dK|dV: 4,5,6

// Column sums over the 32 rows of a warp of a [32 rows x 16 cols] register block (recursive halving, 16 shuffles).
// On return every lane with (lane & 1) == 0 holds in `out` the sum of column `col`.
__device__ __forceinline__ void warp_colsum16(const float (&v)[16], int lane, float& out, int& col) {
  float a8[8], a4[4], a2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h16 ? v[i + 8] : v[i], send = h16 ? v[i] : v[i + 8];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h8 ? a8[i + 4] : a8[i], send = h8 ? a8[i] : a8[i + 4];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h4 ? a4[i + 2] : a4[i], send = h4 ? a4[i] : a4[i + 2];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? a2[1] : a2[0], send = h2 ? a2[0] : a2[1];
  float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  out = a1;
  col = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
}

__device__ __forceinline__ void ld_global16(const bf16* src, uint32_t (&pk)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(src)), b = __ldg(reinterpret_cast<const uint4*>(src + 8));
  pk[0] = a.x; pk[1] = a.y; pk[2] = a.z; pk[3] = a.w; pk[4] = b.x; pk[5] = b.y; pk[6] = b.z; pk[7] = b.w;
}
__device__ __forceinline__ void unpack16(const uint32_t (&pk)[8], float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 f = unpack_bf162(pk[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void st_f32x16(float* dst, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
}  // namespace

struct EncBwdDev {
  const int64_t* mask;
  const bf16* x0;
  const LayerDev* layers;
  const CUtensorMap* maps;   // [L][4]: wqkv, wo, w1, w2 (the forward's maps)
  float* dx;                 // [T][64] fp32: in = d(loss)/d(last layer output), out = d(loss)/d(embedding output)
  float* wpart;              // [L][nCTA][WP] weight-gradient partials: wqkv | wo | w1 | w2
  float* bpart;              // [L][nCTA][PF] bias / LayerNorm gradient partials, laid out like the parameter block
  // embedding stage backward (fused tail): LN backward + dropout', scatter-add into the item-table gradient
  const int64_t* ids; const bf16* table; const bf16* pos; const float* emb_g;
  float* dx_rows;            // [T][64] fp32: gradient of the gathered table rows (may alias dx: each CTA overwrites its own rows)
  float* dpos_part;          // [nCTA][S][64]
  float* embln_part;         // [nCTA][128] gamma | beta
  int V;
  int B, S, L, slot, I;
  uint32_t thr_out, thr_attn; float inv_keep_out, inv_keep_attn;
  unsigned long long seed; uint32_t step; const long long* d_step;
  unsigned long long* dbg;
};

__global__ void __launch_bounds__(NTHR, 1) enc_bwd_fused_kernel(EncBwdDev a) {
  pdl_grid_sync();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto tile = [&](int i) -> unsigned char* { return smem + i * TILE_B; };
  const int I = a.I, PF = par_floats(I), IC = I / 64;     // IC: 16-column chunks of an I-wide row per thread
  const int WP = 64 * 192 + 64 * 64 + 2 * 64 * I;
  float* sPar = reinterpret_cast<float*>(smem + ONES_OFF + ONES_BYTES);   // [2][PF]
  float* sCol = sPar + 2 * PF;                                        // [4 quads][PF] column-sum partials of this layer
  float* sMask = sCol + 4 * PF;                                       // [128]
  float* sRed = sMask + FT;                                           // [2][2][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 16 * FT);
  uint64_t* barWa = bars;       // W2 + W1 + Wo landed
  uint64_t* barWq = bars + 1;   // Wqkv landed
  uint64_t* barM = bars + 2;    // MMA batch complete
  uint64_t* barP = bars + 3;    // [2] parameter block landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int row = quad * 32 + lane;
  const int S = a.S, SLOT = a.slot, G = FT / SLOT;
  const int g = row / SLOT, pos = row - g * SLOT;
  const int seq = blockIdx.x * G + g;
  const bool valid = pos < S && seq < a.B;
  const int t = seq * S + pos;
  const int cq = part * 16;
  const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);
  const Philox ph(a.seed);
  const int nCTA = gridDim.x;

  int dbg_i = 0;
  auto stamp = [&]() {
    if (a.dbg && blockIdx.x == 0 && lane == 0 && dbg_i < 32) {   // every warp of CTA 0: dbg[warp][32]
      unsigned long long tns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
      a.dbg[warp * 32 + dbg_i++] = tns;
    }
  };

  auto load_Wa = [&](int l) {   // thread 0: W2 -> T8, W1 -> T9, Wo -> T10
    const CUtensorMap* m = a.maps + l * 4;
    umma::mbar_expect_tx(barWa, (uint32_t)(I * 128 + IC * 8192 + 8192));
    umma::tma_load_2d(tile(T_W2), m + 3, 0, 0, barWa);
    for (int n = 0; n < IC; ++n) umma::tma_load_2d(tile(T_W1) + n * 8192, m + 2, n * 64, 0, barWa);
    umma::tma_load_2d(tile(T_WO), m + 1, 0, 0, barWa);
  };
  auto load_Wq = [&](int l) {   // thread 0: Wqkv -> T8 (3 x 8 KB)
    const CUtensorMap* m = a.maps + l * 4;
    umma::mbar_expect_tx(barWq, 3 * 8192);
    for (int n = 0; n < 3; ++n) umma::tma_load_2d(tile(T_WQKV) + n * 8192, m, n * 64, 0, barWq);
  };
  auto load_P = [&](int l) {
    umma::mbar_expect_tx(barP + (l & 1), (uint32_t)PF * 4);
    umma::bulk_load_1d(sPar + (l & 1) * PF, a.layers[l].pblock, (uint32_t)PF * 4, barP + (l & 1));
  };

  if (tid == 0) {
    umma::mbar_init(barWa, 1); umma::mbar_init(barWq, 1); umma::mbar_init(barM, 4);   // 4 issuer warps commit per batch
    umma::mbar_init(barP, 1); umma::mbar_init(barP + 1, 1);
    umma::fence_barrier_init();
    load_P(a.L - 1);
    if (a.L > 1) load_P(a.L - 2);
    load_Wa(a.L - 1);
  }
  if (part == 0) sMask[row] = valid ? (a.mask[t] != 0 ? 0.f : -1e9f) : 0.f;
  if (tid < ONES_BYTES / 4) reinterpret_cast<uint32_t*>(smem + ONES_OFF)[tid] = 0x3F803F80u;   // bf16 1.0 pairs
  for (int i = tid; i < 4 * PF; i += NTHR) sCol[i] = 0.f;
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
  uint32_t parM = 0;
  int red_sel = 0;
  stamp();
  // base descriptors of arena tile 0; every operand is DK / DMN + (byte offset >> 4)
  const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
  const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), TILE_B);
  auto dk = [&](int tile_i, int off) -> uint64_t { return desc_at(DK0, (uint32_t)(tile_i * TILE_B + off)); };
  auto dmn = [&](int tile_i, int off) -> uint64_t { return desc_at(DMN0, (uint32_t)(tile_i * TILE_B + off)); };
  // A operand of a weight-gradient MMA (X^T, MN-major, M = 128 but only 64 real feature rows): the second 64-row block is
  // pointed at the tile of ONES (leading-byte-offset chosen per k-step), so accumulator lane 64 = column sums of dY = the
  // bias gradient, for free on the tensor core
  const uint64_t DMN_NOLBO = DMN0 - ((uint64_t)(TILE_B >> 4) << 16);
  auto dmn_ones = [&](int tile_i, int off) -> uint64_t {
    const uint32_t a0 = (uint32_t)(tile_i * TILE_B + off);
    return desc_at(DMN_NOLBO, a0) + ((uint64_t)((ONES_OFF - a0) >> 4) << 16);
  };
  // lane 64 of a weight-gradient accumulator (held by the quad-2 warps) -> sCol slot of quad 2 (the other quads' slots of these
  // offsets stay zero), 16 columns per call
  auto bias_lane = [&](uint32_t acc_col, int off) {
    float v[16];
    tmem_ld_f16(tlane + acc_col, v);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sCol[2 * PF + off + i] = v[i];
    }
  };
  // MMA issue: warps 0..3 (one per scheduler) each issue the accumulation chains `chain(w)` gives them and commit
  auto issue = [&](auto&& chain) {
    if (warp < 4) {
      if (elect_one()) {
        umma::fence_after_sync();
        chain(warp);
        umma::mma_commit(barM);
      }
      __syncwarp();
    }
  };

  // sums of two per-thread values over the 4 column-quarter threads of a row (one barrier)
  auto quad_sum2 = [&](float& x, float& y) {
    float* buf = sRed + red_sel * 8 * FT;
    buf[part * FT + row] = x;
    buf[4 * FT + part * FT + row] = y;
    __syncthreads();
    x = (buf[row] + buf[FT + row]) + (buf[2 * FT + row] + buf[3 * FT + row]);
    y = (buf[4 * FT + row] + buf[5 * FT + row]) + (buf[6 * FT + row] + buf[7 * FT + row]);
    red_sel ^= 1;
  };
  auto pair_other = [&](float v) -> float {
    float* buf = sRed + red_sel * 8 * FT;
    buf[part * FT + row] = v;
    __syncthreads();
    const float o = buf[(part ^ 1) * FT + row];
    red_sel ^= 1;
    return o;
  };
  auto phase_sync = [&]() {
    stamp();
    umma::fence_before_sync();
    umma::fence_proxy_async();
    __syncthreads();
  };
  auto wait_mma = [&]() {
    __syncwarp();
    umma::mbar_wait(barM, parM);
    parM ^= 1;
    umma::fence_after_sync();
    stamp();
  };
  // column sums of a [row x 16] register block over the tile rows -> sCol[quad][off + cq + col] (flushed per layer)
  auto colsum = [&](const float (&v)[16], int off) {
    float s; int c;
    warp_colsum16(v, lane, s, c);
    if (!(lane & 1)) sCol[quad * PF + off + cq + c] = s;
  };
  // keep mask of the elementwise dropout of columns cq..cq+15 of token row t, applied to a gradient
  auto drop16 = [&](float (&v)[16], uint32_t site) {
    if (a.thr_out == 0) return;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)(cq / 8 + q), site, step, a.thr_out);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * q + i] = ((bits >> i) & 1u) ? v[8 * q + i] * a.inv_keep_out : 0.f;
    }
  };
  // LayerNorm backward of one row (four 16-column quarters): dy -> d(pre), + column sums for gamma / beta
  auto ln_bwd = [&](float (&dy)[16], const uint32_t (&pre_pk)[8], float mu, float rs, const float* gamma, int off_g, int off_b) {
    float xh[16];
    unpack16(pre_pk, xh);   // padding rows were prefetched as zeros with rs = 0
#pragma unroll
    for (int i = 0; i < 16; ++i) xh[i] = (xh[i] - mu) * rs;
    float tmp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) tmp[i] = dy[i] * xh[i];
    colsum(tmp, off_g);
    colsum(dy, off_b);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 gm = *reinterpret_cast<const float4*>(gamma + cq + i);
      dy[i] *= gm.x; dy[i + 1] *= gm.y; dy[i + 2] *= gm.z; dy[i + 3] *= gm.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { s1 += dy[i]; s2 += dy[i] * xh[i]; }
    quad_sum2(s1, s2);
    s1 *= (1.0f / FH); s2 *= (1.0f / FH);
#pragma unroll
    for (int i = 0; i < 16; ++i) dy[i] = rs * (dy[i] - s1 - xh[i] * s2);
  };
  // one row x 16 columns of a saved [T][ld] bf16 activation -> two 16-byte chunks of a tile (zeros for padding rows)
  // prefetch of one row x 16 columns of a saved [T][ld] bf16 activation (zeros for padding rows); issued BEFORE the wait
  // on the tensor-core batch so that the L2 latency hides behind it
  auto pf = [&](const bf16* src, int ld, int col, uint32_t (&pk)[8]) {
    if (valid) ld_global16(src + (size_t)t * ld + col, pk);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = 0u;
    }
  };

  // gradient with respect to the output of the layer being processed: this thread's row x 16 columns
  float dO[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dO[i] = 0.f;
  if (valid) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(a.dx + (size_t)t * FH + cq + i);
      dO[i] = v.x; dO[i + 1] = v.y; dO[i + 2] = v.z; dO[i + 3] = v.w;
    }
  }
  const float scale = rsqrtf((float)FD);
  const int hd = part >> 1, kh = part & 1;
  // prefetch registers (saved activations of the NEXT epilogue phase)
  uint32_t pfA[8], pfB[8], pfC[8];
  float pf_mu = 0.f, pf_rs = 0.f;
  auto prefetch_e0 = [&](const LayerDev& Ln) {   // o_pre row chunk, LN2 statistics, h chunk(s)
    pf(Ln.o_pre, FH, cq, pfA);
    pf(Ln.h, I, part * IC * 16, pfB);
    if (IC > 1) pf(Ln.h, I, (part * IC + 1) * 16, pfC);
    pf_mu = valid ? Ln.mean2[t] : 0.f;
    pf_rs = valid ? Ln.rstd2[t] : 0.f;
  };
  prefetch_e0(a.layers[a.L - 1]);

  for (int l = a.L - 1, it = 0; l >= 0; --l, ++it) {
    const LayerDev& Ly = a.layers[l];
    const float* par = sPar + (l & 1) * PF;
    float* wp = a.wpart + ((size_t)l * nCTA + blockIdx.x) * WP;
    const int OFF_B2 = PB_B1 + I, OFF_G2 = OFF_B2 + 64, OFF_BE2 = OFF_G2 + 64;
    umma::mbar_wait(barP + (l & 1), (it >> 1) & 1);
    // ================================================================ E0: LN2 backward -> d_branch2 ; h
    float dres[16];
    {
      ln_bwd(dO, pfA, pf_mu, pf_rs, par + OFF_G2, OFF_G2, OFF_BE2);
#pragma unroll
      for (int i = 0; i < 16; ++i) dres[i] = dO[i];
      drop16(dO, site_id(SITE_FFN_OUT, l));
      uint32_t pk[8];
      round_n<16>(dO, pk);
      if (I != 64) colsum(dO, OFF_B2);   // I == 64: accumulator lane 64 of dW2 (ones block)
      st_tile<2>(tile(T_DB2), row, part * 2, pk);
      for (int i = 0; i < IC; ++i) {
        const int j = part * IC + i;
        if (i == 0) st_tile<2>(tile(T_H + (j >> 2)), row, (j & 3) * 2, pfB);
        else st_tile<2>(tile(T_H + (j >> 2)), row, (j & 3) * 2, pfC);
      }
    }
    // ================================================================ M1: dW2 = h^T db2 | dh = db2 W2^T
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma::mma_bf16_ss(tmem, I == 64 ? dmn_ones(T_H, kk * 2048) : dmn(T_H, kk * 2048), dmn(T_DB2, kk * 2048), idesc_gen(128, 64, 1, 1), kk ? 1u : 0u);
      } else if (w == 1) {
        umma::mbar_wait(barWa, it & 1);
        const uint32_t id = idesc_gen(128, I, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem + 64, dk(T_DB2, k * 32), dk(T_W2, k * 32), id, k ? 1u : 0u);
      }
    });
    pf(Ly.h_pre, I, part * IC * 16, pfB);
    if (IC > 1) pf(Ly.h_pre, I, (part * IC + 1) * 16, pfC);
    pf(Ly.y, FH, cq, pfA);
    wait_mma();
    // ================================================================ E1: drain dW2 ; dh * gelu' -> dh_pre ; y
    {
      if (quad * 32 < I) {   // rows of dW2 = inner units
        float v[16];
        tmem_ld_f16(tlane + cq, v);
        if (row < I) st_f32x16(wp + 64 * 192 + 64 * 64 + 64 * I + (size_t)row * 64 + cq, v);
      }
      if (I == 64 && quad == 2) bias_lane(cq, OFF_B2 + cq);
      for (int i = 0; i < IC; ++i) {
        const int j = part * IC + i;
        float v[16], hp[16];
        tmem_ld_f16(tlane + 64 + j * 16, v);
        if (valid) {
          if (i == 0) unpack16(pfB, hp); else unpack16(pfC, hp);
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] *= gelu_erf_grad(hp[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = 0.f;
        }
        uint32_t pk[8];
        round_n<16>(v, pk);
        st_tile<2>(tile(T_DH + (j >> 2)), row, (j & 3) * 2, pk);
      }
      st_tile<2>(tile(T_Y), row, part * 2, pfA);
    }
    // ================================================================ M2: dW1 = y^T dh_pre | dy = dh_pre W1^T
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
        const uint32_t id = idesc_gen(128, I, 1, 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma::mma_bf16_ss(tmem, dmn_ones(T_Y, kk * 2048), dmn(T_DH, kk * 2048), id, kk ? 1u : 0u);
      } else if (w == 1) {
        for (int kk = 0; kk < I / 16; ++kk)
          umma::mma_bf16_ss(tmem + 256, dk(T_DH, (kk >> 2) * TILE_B + (kk & 3) * 32), dk(T_W1, (kk >> 2) * 8192 + (kk & 3) * 32),
                            idesc_gen(128, 64, 0, 0), kk ? 1u : 0u);
      }
    });
    pf(Ly.a_pre, FH, cq, pfA);
    pf(Ly.ctx, FH, cq, pfB);
    pf_mu = valid ? Ly.mean1[t] : 0.f;
    pf_rs = valid ? Ly.rstd1[t] : 0.f;
    wait_mma();
    // ================================================================ E2: drain dW1 ; LN1 backward -> d_branch1 ; ctx
    {
      if (quad < 2) {
        for (int i = 0; i < IC; ++i) {
          const int j = part * IC + i;
          float v[16];
          tmem_ld_f16(tlane + j * 16, v);
          st_f32x16(wp + 64 * 192 + 64 * 64 + (size_t)row * I + j * 16, v);
        }
      }
      if (quad == 2) {
        for (int i = 0; i < IC; ++i) bias_lane((part * IC + i) * 16, PB_B1 + (part * IC + i) * 16);
      }
      float v[16];
      tmem_ld_f16(tlane + 256 + cq, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) dO[i] = valid ? v[i] + dres[i] : 0.f;
      ln_bwd(dO, pfA, pf_mu, pf_rs, par + PB_G1, PB_G1, PB_BE1);
#pragma unroll
      for (int i = 0; i < 16; ++i) dres[i] = dO[i];
      drop16(dO, site_id(SITE_ATTN_OUT, l));
      uint32_t pk[8];
      round_n<16>(dO, pk);
      st_tile<2>(tile(T_DB1), row, part * 2, pk);
      st_tile<2>(tile(T_C), row, part * 2, pfB);
    }
    // ================================================================ M3: dWo = ctx^T db1 | dctx = db1 Wo^T
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma::mma_bf16_ss(tmem, dmn_ones(T_C, kk * 2048), dmn(T_DB1, kk * 2048), idesc_gen(128, 64, 1, 1), kk ? 1u : 0u);
      } else if (w == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem + 64, dk(T_DB1, k * 32), dk(T_WO, k * 32), idesc_gen(128, 64, 0, 0), k ? 1u : 0u);
      }
    });
    pf(Ly.qkv, 192, cq, pfA);
    pf(Ly.qkv, 192, 64 + cq, pfB);
    pf(Ly.qkv, 192, 128 + cq, pfC);
    wait_mma();
    // ================================================================ E3: drain dWo ; dctx, delta ; Q K V
    float delta = 0.f;
    {
      if (quad < 2) {
        float v[16];
        tmem_ld_f16(tlane + cq, v);
        st_f32x16(wp + 64 * 192 + (size_t)row * 64 + cq, v);
      }
      if (quad == 2) bias_lane(cq, PB_BO + cq);
      float v[16], c[16];
      tmem_ld_f16(tlane + 64 + cq, v);
      uint32_t pk[8];
      round_n<16>(v, pk);
      zero_if<8>(!valid, pk);
      ld_tile<2>(tile(T_C), row, part * 2, c);
      float dl = 0.f;
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dl += v[i] * c[i];
      }
      delta = dl + pair_other(dl);     // the pair (part, part ^ 1) covers the 32 columns of head part >> 1
      st_tile<2>(tile(T_DC), row, part * 2, pk);
      st_tile<2>(tile(T_Q), row, part * 2, pfA);
      st_tile<2>(tile(T_K), row, part * 2, pfB);
      st_tile<2>(tile(T_V), row, part * 2, pfC);
    }
    // ================================================================ M4: S = Q K^T | dP = dctx V^T   (both heads)
    phase_sync();
    issue([&](int w) {   // chain w: head w >> 1 ; even = scores, odd = dP
      const int h = w >> 1;
      if (!(w & 1)) {
#pragma unroll
        for (int k = 0; k < FD / 16; ++k)
          umma::mma_bf16_ss(tmem + h * 128, dk(T_Q, h * FD * 2 + k * 32), dk(T_K, h * FD * 2 + k * 32), idesc_gen(128, 128, 0, 0), k ? 1u : 0u);
      } else {
#pragma unroll
        for (int k = 0; k < FD / 16; ++k)
          umma::mma_bf16_ss(tmem + 256 + h * 128, dk(T_DC, h * FD * 2 + k * 32), dk(T_V, h * FD * 2 + k * 32), idesc_gen(128, 128, 0, 0), k ? 1u : 0u);
      }
    });
    const int CPS = SLOT >> 5, colbase = g * SLOT;
    const bool drop = a.thr_attn > 0;
    const int bn = seq * FNH + hd, W = (S + 63) >> 6;
    const float lse = valid ? Ly.lse[(size_t)bn * S + pos] : 0.f;
    uint32_t kbits[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int cc = kh + 2 * c;
      if ((cc / CPS) == g && drop && valid) {
        const int ci = cc % CPS;
        kbits[c] = reinterpret_cast<const uint32_t*>(Ly.keep)[(((size_t)bn * S + pos) * W + (ci >> 1)) * 2 + (ci & 1)];
      }
    }
    wait_mma();
    // ================================================================ E4: P, dS (recomputed from lse) -> bf16 tiles
    {
      unsigned char* tpd = tile(T_PD + hd * 2);
      unsigned char* tds = tile(T_DS + hd * 2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cc = kh + 2 * c;
        const bool dat = (cc / CPS) == g;
        const int jl0 = (cc % CPS) * 32;
        const uint32_t bits = kbits[c];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {       // two 16-column halves of the 32-key chunk
          uint32_t pkp[8], pkd[8];
          if (dat) {
            float s[16], dp[16];
            tmem_ld_f16(tlane + hd * 128 + cc * 32 + hh * 16, s);
            tmem_ld_f16(tlane + 256 + hd * 128 + cc * 32 + hh * 16, dp);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int jl = jl0 + hh * 16 + j;
              float p = 0.f;
              if (valid && jl < S) p = __expf(s[j] * scale + sMask[colbase + jl] - lse);
              float pd = p, dpj = dp[j];
              if (drop) {
                const bool keep = (bits >> (hh * 16 + j)) & 1u;
                pd = keep ? p * a.inv_keep_attn : 0.f;
                dpj = keep ? dpj * a.inv_keep_attn : 0.f;
              }
              s[j] = pd;
              dp[j] = p * (dpj - delta);
            }
            pack_n<16>(s, pkp);
            pack_n<16>(dp, pkd);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { pkp[i] = 0u; pkd[i] = 0u; }
          }
          st_tile<2>(tpd + (cc >> 1) * TILE_B, row, (cc & 1) * 4 + hh * 2, pkp);
          st_tile<2>(tds + (cc >> 1) * TILE_B, row, (cc & 1) * 4 + hh * 2, pkd);
        }
      }
    }
    // ================================================================ M5: dV = Pd^T dctx | dK = dS^T Q | dQ = dS K
    phase_sync();
    issue([&](int w) {   // six chains over four issuers: w0: dV0 dK0 ; w1: dQ0 ; w2: dV1 dK1 ; w3: dQ1
      const int h = w >> 1;
      const uint32_t base = tmem + h * 192;
      if (!(w & 1)) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma::mma_bf16_ss(base, dmn(T_PD + h * 2, kk * 2048), dmn(T_DC, kk * 2048), idesc_gen(128, 64, 1, 1), kk ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma::mma_bf16_ss(base + 64, dmn(T_DS + h * 2, kk * 2048), dmn(T_Q, kk * 2048), idesc_gen(128, 64, 1, 1), kk ? 1u : 0u);
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma::mma_bf16_ss(base + 128, dk(T_DS + h * 2, (kk >> 2) * TILE_B + (kk & 3) * 32), dmn(T_K, kk * 2048), idesc_gen(128, 64, 0, 1), kk ? 1u : 0u);
      }
    });
    pf(l == 0 ? a.x0 : a.layers[l - 1].out, FH, cq, pfA);
    wait_mma();
    if (tid == 0) load_Wq(l);     // dS tiles are dead: Wqkv streams into T8.. behind the epilogue
    // ================================================================ E5: dQ | dK | dV tiles ; x_in
    {
      const uint32_t base = tlane + hd * 192 + cq;   // this thread's head owns columns cq..cq+15 of its accumulators
      float v[16];
      uint32_t pk[8];
      tmem_ld_f16(base + 128, v);                    // dQ
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = valid ? v[i] * scale : 0.f;
      round_n<16>(v, pk);
      st_tile<2>(tile(T_DQ), row, part * 2, pk);
      tmem_ld_f16(base + 64, v);                     // dK
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = valid ? v[i] * scale : 0.f;
      round_n<16>(v, pk);
      st_tile<2>(tile(T_DQ + 1), row, part * 2, pk);
      tmem_ld_f16(base, v);                          // dV
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = valid ? v[i] : 0.f;
      round_n<16>(v, pk);
      st_tile<2>(tile(T_DQ + 2), row, part * 2, pk);
      st_tile<2>(tile(T_XIN), row, part * 2, pfA);
    }
    // ================================================================ M6: dWqkv = x^T dqkv | dx = dqkv Wqkv^T
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma::mma_bf16_ss(tmem, dmn_ones(T_XIN, kk * 2048), dmn(T_DQ, kk * 2048), idesc_gen(128, 192, 1, 1), kk ? 1u : 0u);
      } else if (w == 1) {
        umma::mbar_wait(barWq, it & 1);
#pragma unroll
        for (int kk = 0; kk < 12; ++kk)
          umma::mma_bf16_ss(tmem + 256, dk(T_DQ, (kk >> 2) * TILE_B + (kk & 3) * 32), dk(T_WQKV, (kk >> 2) * 8192 + (kk & 3) * 32),
                            idesc_gen(128, 64, 0, 0), kk ? 1u : 0u);
      }
    });
    if (l > 0) prefetch_e0(a.layers[l - 1]);
    wait_mma();
    if (tid == 0 && l > 0) load_Wa(l - 1);   // next layer's W2 / W1 / Wo
    // ================================================================ E6: drain dWqkv ; d(residual) for the next layer
    {
      if (quad < 2) {
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
          const int j = part * 3 + i;
          float v[16];
          tmem_ld_f16(tlane + j * 16, v);
          st_f32x16(wp + (size_t)row * 192 + j * 16, v);
        }
      }
      if (quad == 2) {
#pragma unroll 1
        for (int i = 0; i < 3; ++i) bias_lane((part * 3 + i) * 16, PB_BQKV + (part * 3 + i) * 16);
      }
      float v[16];
      tmem_ld_f16(tlane + 256 + cq, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) dO[i] = valid ? v[i] + dres[i] : 0.f;
    }
    // ---------------------------------------------------------------- flush this layer's column sums
    umma::fence_before_sync();
    __syncthreads();
    {
      float* bp = a.bpart + ((size_t)l * nCTA + blockIdx.x) * PF;
      for (int c = tid; c < PF; c += NTHR) bp[c] = (sCol[c] + sCol[PF + c]) + (sCol[2 * PF + c] + sCol[3 * PF + c]);
    }
    if (tid == 0 && l >= 2) load_P(l - 2);   // buffer l & 1 is free again (all threads are past their last read of it)
  }
  // ================================================================== embedding stage backward (k_embed.cu semantics)
  {
    // dO = d(loss)/d(x0).  x0 = drop(LN(E[id] + pos)): recompute the LN input (gather again), LN backward, scatter.
    if (a.thr_out > 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)(cq / 8 + q), site_id(SITE_EMB, 0), step, a.thr_out);
#pragma unroll
        for (int i = 0; i < 8; ++i) dO[8 * q + i] = ((bits >> i) & 1u) ? dO[8 * q + i] * a.inv_keep_out : 0.f;
      }
    }
    float x[16];
    long long id = 0;
    if (valid) {
      id = a.ids[t];
      id = id < 0 ? 0 : (id >= a.V ? a.V - 1 : id);
      uint32_t e[8], p[8];
      ld_global16(a.table + (size_t)id * FH + cq, e);
      ld_global16(a.pos + (size_t)pos * FH + cq, p);
      float pv[16];
      unpack16(e, x);
      unpack16(p, pv);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] += pv[i];
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] = 0.f; dO[i] = 0.f; }
    }
    float sm = 0.f, zero = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sm += x[i];
    quad_sum2(sm, zero);
    const float mean = sm * (1.0f / FH);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = x[i] - mean; q2 += d * d; }
    zero = 0.f;
    quad_sum2(q2, zero);
    const float rstd = valid ? rsqrtf(q2 * (1.0f / FH) + kLnEps) : 0.f;
    float tmp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = (x[i] - mean) * rstd; tmp[i] = dO[i] * x[i]; }
    // column sums -> sCol (gamma at [0,64), beta at [64,128) of quad block)
    {
      float sv; int c;
      warp_colsum16(tmp, lane, sv, c);
      if (!(lane & 1)) sCol[quad * PF + cq + c] = sv;
      warp_colsum16(dO, lane, sv, c);
      if (!(lane & 1)) sCol[quad * PF + 64 + cq + c] = sv;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(a.emb_g + cq + i));
      dO[i] *= gm.x; dO[i + 1] *= gm.y; dO[i + 2] *= gm.z; dO[i + 3] *= gm.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { s1 += dO[i]; s2 += dO[i] * x[i]; }
    quad_sum2(s1, s2);
    s1 *= (1.0f / FH); s2 *= (1.0f / FH);
#pragma unroll
    for (int i = 0; i < 16; ++i) dO[i] = rstd * (dO[i] - s1 - x[i] * s2);
    // stage dx rows (fp32) and the item ids of the tile in the (dead) arena
    float* sDx = reinterpret_cast<float*>(smem);                 // [128 rows][64]
    umma::fence_before_sync();
    __syncthreads();                                             // every MMA / tile read of the last layer is complete
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(sDx + row * FH + cq + i) = make_float4(dO[i], dO[i + 1], dO[i + 2], dO[i + 3]);
    __syncthreads();
    // item-table gradient: the dx rows go to memory (in place over the incoming gradient of the tile's own rows); their per-item
    // sums are taken in a fixed order over the id-sorted token list (k_tablegrad.cu) -- no floating-point atomics
    if (valid) {
      float* gr = a.dx_rows + (size_t)t * FH + cq;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(gr + i) = make_float4(dO[i], dO[i + 1], dO[i + 2], dO[i + 3]);
    }
    // position gradient: sum over the G sequences of the tile (fixed order), one partial per CTA
    float* dp = a.dpos_part + (size_t)blockIdx.x * S * FH;
    for (int e = tid; e < S * FH; e += NTHR) {
      const int pp = e / FH, c = e % FH;
      float v = 0.f;
      for (int gg = 0; gg < G; ++gg) v += sDx[(gg * SLOT + pp) * FH + c];
      dp[e] = v;
    }
    if (tid < 128) a.embln_part[(size_t)blockIdx.x * 128 + tid] = (sCol[tid] + sCol[PF + tid]) + (sCol[2 * PF + tid] + sCol[3 * PF + tid]);
  }
  stamp();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ host side
size_t enc_bwd_fused_smem_bytes(int I) {
  return (size_t)NT_ARENA * TILE_B + ONES_BYTES + 2 * (size_t)par_floats(I) * 4 + 4 * (size_t)par_floats(I) * 4 + FT * 4 + 16 * FT * 4 + 64 + 1024;
}
bool enc_bwd_fused_supported(int H, int N, int S, int I) {
  if (getenv("B4R_DISABLE_FUSED") || getenv("B4R_DISABLE_FUSED_BWD")) return false;
  if (!enc_fused_supported(H, N, S, I)) return false;
  if (I != 64 && I != 128) return false;
  return enc_bwd_fused_smem_bytes(I) <= 232448;
}
int enc_fused_ctas(int B, int S) {
  const int slot = S <= 32 ? 32 : (S <= 64 ? 64 : 128);
  const int G = FT / slot;
  return (B + G - 1) / G;
}
size_t enc_bwd_wpart_floats(int I) { return (size_t)64 * 192 + 64 * 64 + 2 * 64 * (size_t)I; }
size_t enc_bwd_bpart_floats(int I) { return (size_t)par_floats(I); }

cudaError_t launch_enc_bwd_fused(const EncBwdArgs& a, cudaStream_t st) {
  EncBwdDev d;
  d.mask = a.mask; d.x0 = a.x0;
  d.layers = reinterpret_cast<const LayerDev*>(a.dev_tables);
  size_t moff = ((size_t)a.L * sizeof(LayerDev) + 127) / 128 * 128;
  d.maps = reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const char*>(a.dev_tables) + moff);
  d.dx = a.dx; d.wpart = a.wpart; d.bpart = a.bpart;
  d.ids = a.ids; d.table = a.table; d.pos = a.pos; d.emb_g = a.emb_g; d.dx_rows = a.dx_rows; d.dpos_part = a.dpos_part;
  d.embln_part = a.embln_part; d.V = a.V;
  d.B = a.B; d.S = a.S; d.L = a.L; d.slot = a.S <= 32 ? 32 : (a.S <= 64 ? 64 : 128); d.I = a.I;
  d.thr_out = drop_threshold16(a.out_drop); d.thr_attn = drop_threshold16(a.attn_drop);
  d.inv_keep_out = 1.0f / (1.0f - (float)d.thr_out / 65536.0f);
  d.inv_keep_attn = 1.0f / (1.0f - (float)d.thr_attn / 65536.0f);
  d.seed = a.seed; d.step = a.step; d.d_step = a.d_step; d.dbg = reinterpret_cast<unsigned long long*>(a.dbg);
  const size_t smem = enc_bwd_fused_smem_bytes(a.I);
  static size_t cap = 0;
  if (smem > cap) {
    cudaError_t e = cudaFuncSetAttribute(enc_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cap = smem;
  }
  return launch_pdl(enc_bwd_fused_kernel, dim3(enc_fused_ctas(a.B, a.S)), dim3(NTHR), smem, st, d);
}

}  // namespace b4r
