// Generation 2 of the encoder stack FORWARD (hidden 64, 2 heads): the whole stack -- embedding gather + position +
// LayerNorm + dropout, then per layer QKV projection, masked self-attention, output projection + residual + LN,
// FFN1 + GELU, FFN2 + residual + LN -- is ONE launch.  A CTA owns a tile of 128 token rows = G whole sequences
// (G = 128 / S, e.g. 2 sequences of 50 items); the tile never leaves the SM between phases:
//
//   operands   bf16 tiles in shared memory, K-major [128 rows][64 cols] with the 128-byte swizzle tcgen05 expects;
//              weight matrices [in][out] fetched by TMA (cp.async.bulk.tensor, SWIZZLE_128B) and consumed MN-major,
//              the next layer's weights prefetched while the current layer computes
//   GEMMs      tcgen05.mma (cta_group::1, kind::f16, M = 128) issued by one thread, fp32 accumulators in TMEM
//   epilogues  256 threads, one (row, column-half) each: tcgen05.ld -> bias / mask / softmax / dropout / residual /
//              LayerNorm / GELU in registers -> bf16 tile for the next GEMM (+ the activations backward re-reads)
//
// Attention is a 128 x 128 score tile per head (block-diagonal over the G sequences: foreign keys are excluded,
// padded keys get the additive -1e9 of Keras Softmax(mask)); the [B,S,S] mask and [B,N,S,S] probabilities of the
// reference never exist.  Reference ops replaced: bert4rec_encoder.py:198-222 (tfm OnDeviceEmbedding, PositionEmbedding,
// TransformerEncoderBlock = Keras MultiHeadAttention + EinsumDense + LayerNormalization); SURVEY.md 2b rows K1-K5.
// Numerics and saved tensors are those of the layered kernels (k_embed.cu, k_gemm.cu, k_attn.cu), including the
// Philox dropout streams, so the layered backward consumes this forward unchanged.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "enc_fused.cuh"

namespace b4r {

using namespace encf;

struct EncFusedDev {
  const int64_t* ids; const int64_t* mask;
  const bf16* table; const bf16* pos; const float* emb_g; const float* emb_b;
  bf16* x0;
  const LayerDev* layers;
  const CUtensorMap* maps;  // [L][4]: wqkv, wo, w1, w2
  int B, S, V, L, slot, I;  // slot = rows reserved per sequence (32 / 64 / 128)
  int training;
  uint32_t thr_out, thr_attn; float inv_keep_out, inv_keep_attn;
  unsigned long long seed; uint32_t step; const long long* d_step;
  unsigned long long* dbg;   // optional: phase timestamps of CTA 0 + start/end of every CTA (development aid)
};

__global__ void __launch_bounds__(NTHR, 1) enc_fwd_fused_kernel(EncFusedDev a) {
  pdl_grid_sync();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sX = smem + OFF_X;
  unsigned char* sQ = smem + OFF_Q;
  unsigned char* sK = smem + OFF_K;
  unsigned char* sV = smem + OFF_V;
  unsigned char* sP = smem + OFF_P;
  unsigned char* sCtx = sQ;
  unsigned char* sY = sK;
  unsigned char* sHh = sP;
  unsigned char* sWA = smem + OFF_WA;
  unsigned char* sWB = smem + OFF_WB;
  const int I = a.I;
  const int PF = par_floats(I);
  float* sPar = reinterpret_cast<float*>(sWB + I * 256);    // [2][PF] parameter blocks of two consecutive layers
  float* sEmb = sPar + 2 * PF;                              // [128] embedding LN gamma | beta
  float* sMask = sEmb + 128;                                // [128] additive key mask of the tile rows
  float* sRed = sMask + FT;                                 // [2][4][128] exchange buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 8 * FT);
  uint64_t* barA = bars;        // Wqkv + Wo landed
  uint64_t* barB = bars + 1;    // W1 + W2 landed
  uint64_t* barM = bars + 2;    // MMA batch complete
  uint64_t* barP = bars + 3;    // [2] parameter block landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int row = quad * 32 + lane;
  const int S = a.S, SLOT = a.slot, G = FT / SLOT;
  const int g = row / SLOT, pos = row - g * SLOT;       // g is warp-uniform (SLOT >= 32)
  const int seq = blockIdx.x * G + g;
  const bool valid = pos < S && seq < a.B;
  const int t = seq * S + pos;
  const int cq = part * 16;                             // this thread's 16 columns of a 64-wide row
  const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);
  const Philox ph(a.seed);
  const bool train = a.training != 0;

  int dbg_i = 0;
  auto stamp = [&]() {
    if (a.dbg && blockIdx.x == 0 && tid == 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
      a.dbg[dbg_i++] = tns;
    }
  };
  if (a.dbg && tid == 0 && blockIdx.x < 128) {
    unsigned long long tns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
    a.dbg[256 + 2 * blockIdx.x] = tns;
  }

  auto load_A = [&](int l) {   // thread 0: Wqkv + Wo of layer l
    const CUtensorMap* m = a.maps + l * 4;
    umma::mbar_expect_tx(barA, 4 * 8192);
    for (int n = 0; n < 3; ++n) umma::tma_load_2d(sWA + n * 8192, m, n * 64, 0, barA);
    umma::tma_load_2d(sWA + 3 * 8192, m + 1, 0, 0, barA);
  };
  auto load_B = [&](int l) {   // thread 0: W1 + W2 of layer l
    const CUtensorMap* m = a.maps + l * 4;
    umma::mbar_expect_tx(barB, (uint32_t)I * 256);
    for (int n = 0; n < I / 64; ++n) umma::tma_load_2d(sWB + n * 8192, m + 2, n * 64, 0, barB);
    umma::tma_load_2d(sWB + I * 128, m + 3, 0, 0, barB);
  };
  auto load_P = [&](int l) {   // thread 0: parameter block of layer l -> buffer l & 1
    umma::mbar_expect_tx(barP + (l & 1), (uint32_t)PF * 4);
    umma::bulk_load_1d(sPar + (l & 1) * PF, a.layers[l].pblock, (uint32_t)PF * 4, barP + (l & 1));
  };

  if (tid == 0) {
    umma::mbar_init(barA, 1); umma::mbar_init(barB, 1); umma::mbar_init(barM, 4);   // 4 issuer warps commit per batch
    umma::mbar_init(barP, 1); umma::mbar_init(barP + 1, 1);
    umma::fence_barrier_init();
    load_P(0);
    if (a.L > 1) load_P(1);
    load_A(0); load_B(0);
  }
  if (tid < 128) sEmb[tid] = tid < 64 ? a.emb_g[tid] : a.emb_b[tid - 64];
  if (part == 0) sMask[row] = valid ? (a.mask[t] != 0 ? 0.f : -1e9f) : 0.f;
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
  uint32_t parM = 0;
  int red_sel = 0;
  stamp();
  // base descriptors (smem offset 0); every operand is base + (byte offset >> 4).  MN-major weights: 8 KB between
  // 64-column blocks; V (one block) shares that base.
  const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
  const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), 8192);
  auto dk = [&](int off) -> uint64_t { return desc_at(DK0, (uint32_t)off); };
  auto dmn = [&](int off) -> uint64_t { return desc_at(DMN0, (uint32_t)off); };
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

  // exchange of one float between the 4 column-quarter threads of a row
  auto quad_sum = [&](float v) -> float {
    float* buf = sRed + red_sel * 4 * FT;
    buf[part * FT + row] = v;
    __syncthreads();
    const float r = (buf[row] + buf[FT + row]) + (buf[2 * FT + row] + buf[3 * FT + row]);
    red_sel ^= 1;
    return r;
  };
  auto pair_other = [&](float v) -> float {   // value of the thread owning the other key half of the same (row, head)
    float* buf = sRed + red_sel * 4 * FT;
    buf[part * FT + row] = v;
    __syncthreads();
    const float o = buf[(part ^ 1) * FT + row];
    red_sel ^= 1;
    return o;
  };
  // LayerNorm of a 64-wide row held as four 16-column quarters; v must already be the bf16-rounded pre-LN value
  auto layer_norm = [&](float (&v)[16], const float* gamma, const float* beta, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    mean = quad_sum(s) * (1.0f / FH);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q += d * d; }
    rstd = rsqrtf(quad_sum(q) * (1.0f / FH) + kLnEps);
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 gm = *reinterpret_cast<const float4*>(gamma + cq + i);
      const float4 bt = *reinterpret_cast<const float4*>(beta + cq + i);
      v[i] = (v[i] - mean) * rstd * gm.x + bt.x; v[i + 1] = (v[i + 1] - mean) * rstd * gm.y + bt.y;
      v[i + 2] = (v[i + 2] - mean) * rstd * gm.z + bt.z; v[i + 3] = (v[i + 3] - mean) * rstd * gm.w + bt.w;
    }
  };
  auto drop16 = [&](float (&v)[16], uint32_t site) {   // elementwise dropout of columns cq..cq+15 of token row t
    if (a.thr_out == 0 || !train) return;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)(cq / 8 + q), site, step, a.thr_out);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * q + i] = ((bits >> i) & 1u) ? v[8 * q + i] * a.inv_keep_out : 0.f;
    }
  };
  // every thread: previous epilogue's smem writes / TMEM reads are ordered before the next MMA batch
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
  // shared epilogue of the two residual + LayerNorm phases (attention output, FFN output)
  auto res_ln_epilogue = [&](uint32_t acc_col, const float* bias, uint32_t site, const unsigned char* res_tile, const float* gamma,
                             const float* beta, bf16* pre_out, unsigned char* dst_tile, bf16* y_out, bool y_always, float* mean_out,
                             float* rstd_out) {
    float v[16], res[16];
    tmem_ld_f16(tlane + acc_col + cq, v);
    add_vec<16>(v, bias + cq);
    drop16(v, site);
    ld_tile<2>(res_tile, row, part * 2, res);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += res[i];
    uint32_t pk[8];
    round_n<16>(v, pk);                            // LN statistics on the bf16 value backward re-reads
    if (valid && train) st_global<8>(pre_out + (size_t)t * FH + cq, pk);
    float mean, rstd;
    layer_norm(v, gamma, beta, mean, rstd);
    pack_n<16>(v, pk);
    zero_if<8>(!valid, pk);
    st_tile<2>(dst_tile, row, part * 2, pk);
    if (valid && (train || y_always)) st_global<8>(y_out + (size_t)t * FH + cq, pk);
    if (valid && train && part == 0) { mean_out[t] = mean; rstd_out[t] = rstd; }
  };

  // ------------------------------------------------------------------ phase 0: embedding + LN + dropout -> sX
  {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (valid) {
      long long id = a.ids[t];
      id = id < 0 ? 0 : (id >= a.V ? a.V - 1 : id);
      const bf16* e = a.table + (size_t)id * FH + cq;
      const bf16* p = a.pos + (size_t)pos * FH + cq;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 eu = __ldg(reinterpret_cast<const uint4*>(e + 8 * q));
        const uint4 pu = __ldg(reinterpret_cast<const uint4*>(p + 8 * q));
        const uint32_t ew[4] = {eu.x, eu.y, eu.z, eu.w}, pw[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = unpack_bf162(ew[i]), y = unpack_bf162(pw[i]);
          v[8 * q + 2 * i] = x.x + y.x; v[8 * q + 2 * i + 1] = x.y + y.y;
        }
      }
    }
    float mean, rstd;
    layer_norm(v, sEmb, sEmb + 64, mean, rstd);
    drop16(v, site_id(SITE_EMB, 0));
    uint32_t pk[8];
    pack_n<16>(v, pk);
    zero_if<8>(!valid, pk);
    st_tile<2>(sX, row, part * 2, pk);
    if (valid && train) st_global<8>(a.x0 + (size_t)t * FH + cq, pk);
  }

  const float scale = rsqrtf((float)FD);
  for (int l = 0; l < a.L; ++l) {
    const LayerDev& Ly = a.layers[l];
    const float* par = sPar + (l & 1) * PF;
    // ---------------------------------------------------------------- phase 1: QKV = X Wqkv + b
    phase_sync();
    if (tid == 0 && l >= 1 && l + 1 < a.L) load_P(l + 1);     // buffer (l+1)&1 was last read by layer l-1
    issue([&](int w) {
      if (w == 0) {
        umma::mbar_wait(barA, l & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem, dk(OFF_X + k * 32), dmn(OFF_WA + k * 2048), idesc_bmn(FT, 192), k ? 1u : 0u);
      }
    });
    umma::mbar_wait(barP + (l & 1), (l >> 1) & 1);  // this layer's biases / LN parameters are in shared memory
    wait_mma();
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
      const int j = part * 3 + i;            // 16-column chunk of the 192-wide row: Q Q Q Q K K K K V V V V
      float v[16];
      tmem_ld_f16(tlane + j * 16, v);
      add_vec<16>(v, par + PB_BQKV + j * 16);
      uint32_t pk[8];
      pack_n<16>(v, pk);
      zero_if<8>(!valid, pk);
      unsigned char* tile = (j >> 2) == 0 ? sQ : ((j >> 2) == 1 ? sK : sV);
      st_tile<2>(tile, row, (j & 3) * 2, pk);
      if (valid && train) st_global<8>(Ly.qkv + (size_t)t * 192 + j * 16, pk);
    }
    // ---------------------------------------------------------------- phase 2: scores per head + softmax -> P
    phase_sync();
    issue([&](int w) {
      if (w < FNH) {
#pragma unroll
        for (int k = 0; k < FD / 16; ++k)
          umma::mma_bf16_ss(tmem + w * 128, dk(OFF_Q + w * FD * 2 + k * 32), dk(OFF_K + w * FD * 2 + k * 32), umma::make_idesc_bf16(FT, 128), k ? 1u : 0u);
      }
    });
    wait_mma();
    float inv_l = 0.f;
    {
      // thread = (row, head, key half): of the four 32-key chunks of the score row it owns chunks kh and kh+2; only
      // the chunks inside the row's own slot carry data, the others are zero-filled (foreign sequences)
      const int hd = part >> 1, kh = part & 1;
      const int CPS = SLOT >> 5;                 // 32-key chunks per slot
      const int colbase = g * SLOT;
      const bool drop = train && a.thr_attn > 0;
      const int bn = seq * FNH + hd, W = (S + 63) >> 6;
      bool dat[2]; int jl0[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cc = kh + 2 * c;
        dat[c] = (cc / CPS) == g;
        jl0[c] = (cc % CPS) * 32;
      }
      const int keepc = dat[0] ? 0 : 1;
      float sk[32];
      float mloc = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (dat[c]) {
          float s[32];
          tmem_ld_f32(tlane + hd * 128 + (kh + 2 * c) * 32, s);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int jl = jl0[c] + j;
            const float v = (valid && jl < S) ? s[j] * scale + sMask[colbase + jl] : -INFINITY;
            mloc = fmaxf(mloc, v);
            if (c == keepc) sk[j] = v;
          }
        }
      }
      const float m = fmaxf(mloc, pair_other(mloc));
      float lsum = 0.f;
      unsigned char* pt = sP + hd * 2 * TILE_B;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cc = kh + 2 * c;
        uint32_t pk[16];
        if (dat[c]) {
          float v[32];
          if (c == keepc) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = sk[j];
          } else {
            tmem_ld_f32(tlane + hd * 128 + cc * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int jl = jl0[c] + j;
              v[j] = (valid && jl < S) ? v[j] * scale + sMask[colbase + jl] : -INFINITY;
            }
          }
          // attention-prob dropout keep bits of this chunk: identical stream to attn_fwd_kernel (k_attn.cu)
          uint32_t bits = 0xFFFFFFFFu;
          if (drop && valid) {
            const int ci = jl0[c] >> 5, kb = ci >> 1, hf = ci & 1;
            const uint32_t grow = (uint32_t)(bn * S + pos), site = site_id(SITE_ATTN_PROBS, l);
            bits = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int o = hf + 2 * u;
              const uint4 r = ph(grow, (uint32_t)(kb * 8 + o), site, step);
              const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int bit = i * 8 + u * 2;
                bits |= ((w[i] & 0xFFFFu) >= a.thr_attn ? 1u : 0u) << bit;
                bits |= ((w[i] >> 16) >= a.thr_attn ? 1u : 0u) << (bit + 1);
              }
            }
            reinterpret_cast<uint32_t*>(Ly.keep)[(((size_t)bn * S + pos) * W + kb) * 2 + hf] = bits;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float p = v[j] == -INFINITY ? 0.f : __expf(v[j] - m);
            lsum += p;
            if (drop) p = ((bits >> j) & 1u) ? p * a.inv_keep_attn : 0.f;
            v[j] = p;
          }
          pack_n<32>(v, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        }
        st_tile<4>(pt + (cc >> 1) * TILE_B, row, (cc & 1) * 4, pk);
      }
      const float lall = lsum + pair_other(lsum);
      if (valid) {
        inv_l = 1.0f / lall;
        if (train && kh == 0) Ly.lse[(size_t)bn * S + pos] = m + __logf(lall);
      }
    }
    // ---------------------------------------------------------------- phase 3: ctx = P V per head
    phase_sync();
    issue([&](int w) {
      if (w < FNH) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma::mma_bf16_ss(tmem + 256 + w * 64, dk(OFF_P + w * 2 * TILE_B + (kk >> 2) * TILE_B + (kk & 3) * 32), dmn(OFF_V + kk * 2048),
                            idesc_bmn(FT, 64), kk ? 1u : 0u);
      }
    });
    wait_mma();
    {
      const int hd = part >> 1;
      float v[16];
      tmem_ld_f16(tlane + 256 + hd * 64 + cq, v);   // head hd's columns of its own accumulator (cq = hd*32 + kh*16)
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= inv_l;
      uint32_t pk[8];
      pack_n<16>(v, pk);
      zero_if<8>(!valid, pk);
      st_tile<2>(sCtx, row, part * 2, pk);
      if (valid && train) st_global<8>(Ly.ctx + (size_t)t * FH + cq, pk);
    }
    // ---------------------------------------------------------------- phase 4: a = x + drop(ctx Wo + bo) ; y = LN1(a)
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem + 384, dk(OFF_Q + k * 32), dmn(OFF_WA + 3 * 8192 + k * 2048), idesc_bmn(FT, 64), k ? 1u : 0u);
      }
    });
    wait_mma();
    if (tid == 0 && l + 1 < a.L) load_A(l + 1);   // Wqkv / Wo of the next layer stream in behind the epilogue
    res_ln_epilogue(384, par + PB_BO, site_id(SITE_ATTN_OUT, l), sX, par + PB_G1, par + PB_BE1, Ly.a_pre, sY, Ly.y, false, Ly.mean1, Ly.rstd1);
    // ---------------------------------------------------------------- phase 5: h = gelu(y W1 + b1)
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
        umma::mbar_wait(barB, l & 1);
        const uint32_t id1 = idesc_bmn(FT, I);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem, dk(OFF_K + k * 32), dmn(OFF_WB + k * 2048), id1, k ? 1u : 0u);
      }
    });
    wait_mma();
#pragma unroll 1
    for (int i = 0; i < I / 64; ++i) {
      const int j = part * (I / 64) + i;      // 16-column chunk of the I-wide row
      float v[16];
      tmem_ld_f16(tlane + j * 16, v);
      add_vec<16>(v, par + PB_B1 + j * 16);
      uint32_t pk[8];
      round_n<16>(v, pk);                      // GELU of the bf16 pre-activation backward re-reads
      if (valid && train) st_global<8>(Ly.h_pre + (size_t)t * I + j * 16, pk);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = gelu_erf(v[k]);
      pack_n<16>(v, pk);
      zero_if<8>(!valid, pk);
      st_tile<2>(sHh + (j >> 2) * TILE_B, row, (j & 3) * 2, pk);
      if (valid && train) st_global<8>(Ly.h + (size_t)t * I + j * 16, pk);
    }
    // ---------------------------------------------------------------- phase 6: o = y + drop(h W2 + b2) ; out = LN2(o)
    phase_sync();
    issue([&](int w) {
      if (w == 0) {
        for (int kk = 0; kk < I / 16; ++kk)
          umma::mma_bf16_ss(tmem + 448, dk(OFF_P + (kk >> 2) * TILE_B + (kk & 3) * 32), dmn(OFF_WB + I * 128 + kk * 2048), idesc_bmn(FT, 64), kk ? 1u : 0u);
      }
    });
    wait_mma();
    if (tid == 0 && l + 1 < a.L) load_B(l + 1);
    res_ln_epilogue(448, par + PB_B1 + I, site_id(SITE_FFN_OUT, l), sY, par + PB_B1 + I + 64, par + PB_B1 + I + 128, Ly.o_pre, sX, Ly.out, true,
                    Ly.mean2, Ly.rstd2);
  }
  stamp();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
  if (a.dbg && tid == 0 && blockIdx.x < 128) {
    unsigned long long tns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
    a.dbg[257 + 2 * blockIdx.x] = tns;
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool enc_fused_supported(int H, int N, int S, int I) {
  if (getenv("B4R_DISABLE_FUSED")) return false;
  if (H != FH || N != FNH || S < 1 || S > FT) return false;
  if (I < 64 || I > 256 || (I % 64)) return false;
  return enc_fused_smem_bytes(I) <= 232448;
}
size_t enc_fused_smem_bytes(int I) { return (size_t)OFF_WB + (size_t)I * 256 + 2 * (size_t)par_floats(I) * 4 + 128 * 4 + FT * 4 + 8 * FT * 4 + 64 + 1024; }
size_t enc_fused_table_bytes(int L) { return (size_t)L * sizeof(LayerDev) + 128 + (size_t)L * 4 * sizeof(CUtensorMap); }

// Fills the device-side layer table + tensor maps (host staging buffer `host`, same size as the device block).
bool enc_fused_build_tables(const EncFusedLayerHost* layers, int L, int I, const bf16* const* w_ptrs, void* host, void* dev_base) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  LayerDev* ld = reinterpret_cast<LayerDev*>(host);
  for (int l = 0; l < L; ++l) {
    const EncFusedLayerHost& h = layers[l];
    // the kernel bulk-copies the eight per-layer vectors as ONE block: they must be contiguous in this order
    if (h.bo != h.bqkv + 192 || h.g1 != h.bo + 64 || h.be1 != h.g1 + 64 || h.b1 != h.be1 + 64 || h.b2 != h.b1 + I || h.g2 != h.b2 + 64 ||
        h.be2 != h.g2 + 64 || ((uintptr_t)h.bqkv & 15))
      return false;
    ld[l] = LayerDev{h.bqkv, h.qkv, h.ctx, h.a_pre, h.y, h.h_pre, h.h, h.o_pre, h.out,
                     h.lse, h.mean1, h.rstd1, h.mean2, h.rstd2, reinterpret_cast<unsigned long long*>(h.keep)};
  }
  size_t moff = ((size_t)L * sizeof(LayerDev) + 127) / 128 * 128;
  if (((uintptr_t)dev_base + moff) & 127) return false;
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(reinterpret_cast<char*>(host) + moff);
  for (int l = 0; l < L; ++l) {
    const bf16* wqkv = w_ptrs[l * 4 + 0]; const bf16* wo = w_ptrs[l * 4 + 1];
    const bf16* w1 = w_ptrs[l * 4 + 2]; const bf16* w2 = w_ptrs[l * 4 + 3];
    bool ok = make_tmap_bf16_sw128(maps + l * 4 + 0, wqkv, 64, 192, 192, 64);
    ok = ok && make_tmap_bf16_sw128(maps + l * 4 + 1, wo, 64, 64, 64, 64);
    ok = ok && make_tmap_bf16_sw128(maps + l * 4 + 2, w1, 64, (uint64_t)I, (uint64_t)I, 64);
    ok = ok && make_tmap_bf16_sw128(maps + l * 4 + 3, w2, (uint64_t)I, 64, 64, (uint32_t)I);
    if (!ok) return false;
  }
  return true;
}

cudaError_t launch_enc_fwd_fused(const EncFusedArgs& a, cudaStream_t st) {
  EncFusedDev d;
  d.ids = a.ids; d.mask = a.mask; d.table = a.table; d.pos = a.pos; d.emb_g = a.emb_g; d.emb_b = a.emb_b; d.x0 = a.x0;
  d.layers = reinterpret_cast<const LayerDev*>(a.dev_tables);
  size_t moff = ((size_t)a.L * sizeof(LayerDev) + 127) / 128 * 128;
  d.maps = reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const char*>(a.dev_tables) + moff);
  d.B = a.B; d.S = a.S; d.V = a.V; d.L = a.L; d.slot = a.S <= 32 ? 32 : (a.S <= 64 ? 64 : 128); d.I = a.I; d.training = a.training;
  d.thr_out = drop_threshold16(a.out_drop); d.thr_attn = drop_threshold16(a.attn_drop);
  d.inv_keep_out = 1.0f / (1.0f - (float)d.thr_out / 65536.0f);
  d.inv_keep_attn = 1.0f / (1.0f - (float)d.thr_attn / 65536.0f);
  d.seed = a.seed; d.step = a.step; d.d_step = a.d_step; d.dbg = reinterpret_cast<unsigned long long*>(a.dbg);
  const size_t smem = enc_fused_smem_bytes(a.I);
  static size_t cap = 0;
  if (smem > cap) {
    cudaError_t e = cudaFuncSetAttribute(enc_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cap = smem;
  }
  const int G = FT / d.slot;
  const int grid = (a.B + G - 1) / G;
  return launch_pdl(enc_fwd_fused_kernel, dim3(grid), dim3(NTHR), smem, st, d);
}

}  // namespace b4r
