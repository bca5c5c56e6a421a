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

namespace b4r {

namespace {
constexpr int FT = 128;                 // rows per CTA tile
constexpr int FH = 64;                  // hidden size
constexpr int FNH = 2, FD = 32;         // heads, head dim
constexpr int TILE_B = FT * 128;        // bytes of one [128][64] bf16 tile

// shared-memory map (byte offsets from the 1024-aligned base)
constexpr int OFF_X = 0;                        // layer input / residual
constexpr int OFF_Q = OFF_X + TILE_B;           // Q ; later the attention context (A operand of the output projection)
constexpr int OFF_K = OFF_Q + TILE_B;           // K ; later y = LN1 output
constexpr int OFF_V = OFF_K + TILE_B;           // V
constexpr int OFF_P = OFF_V + TILE_B;           // probabilities: FNH x [128][128] ; later h = gelu(FFN1) [128][I]
constexpr int OFF_WA = OFF_P + FNH * 2 * TILE_B;  // Wqkv (3 x 8 KB) + Wo (8 KB)
constexpr int OFF_WB = OFF_WA + 4 * 8192;       // W1 (I/64 x 8 KB) + W2 (I x 128 B)

struct LayerDev {
  const float *bqkv, *bo, *g1, *be1, *b1, *b2, *g2, *be2;
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

__device__ __forceinline__ void tmem_ld_f32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  umma::tmem_ld32(taddr, r);
  umma::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
// 32 bf16 (16 packed words) of one row -> 4 consecutive 16-byte chunks (chunk0..chunk0+3) of a swizzled tile row
__device__ __forceinline__ void st_tile(unsigned char* tile, int row, int chunk0, const uint32_t (&pk)[16]) {
  unsigned char* rp = tile + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(rp + (((chunk0 + q) ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}
__device__ __forceinline__ void ld_tile(const unsigned char* tile, int row, int chunk0, float (&v)[32]) {
  const unsigned char* rp = tile + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(rp + (((chunk0 + q) ^ (row & 7)) << 4));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf162(w[i]); v[8 * q + 2 * i] = f.x; v[8 * q + 2 * i + 1] = f.y; }
  }
}
__device__ __forceinline__ void st_global32(bf16* dst, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}
__device__ __forceinline__ void pack32(const float (&v)[32], uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = pack_bf162(v[2 * i], v[2 * i + 1]);
}
__device__ __forceinline__ void round32(float (&v)[32], uint32_t (&pk)[16]) {  // v := bf16-rounded v, pk := packed
  pack32(v, pk);
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float2 f = unpack_bf162(pk[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
}  // namespace

struct EncFusedDev {
  const int64_t* ids; const int64_t* mask;
  const bf16* table; const bf16* pos; const float* emb_g; const float* emb_b;
  bf16* x0;
  const LayerDev* layers;
  const CUtensorMap* maps;  // [L][4]: wqkv, wo, w1, w2
  int B, S, V, L, G, I;
  int training;
  uint32_t thr_out, thr_attn; float inv_keep_out, inv_keep_attn;
  unsigned long long seed; uint32_t step; const long long* d_step;
};

__global__ void __launch_bounds__(256, 1) enc_fwd_fused_kernel(EncFusedDev a) {
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
  float* sMask = reinterpret_cast<float*>(sWB + I * 256);   // [128] additive key mask of the tile rows
  float* sRed = reinterpret_cast<float*>(sP);               // [2][2][128] pair-exchange buffers: P / h are dead in every LN phase
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + FT);
  uint64_t* barA = bars;        // Wqkv + Wo landed
  uint64_t* barB = bars + 1;    // W1 + W2 landed
  uint64_t* barM = bars + 2;    // MMA batch complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const int S = a.S, T = a.B * S;
  const int R = a.G * S;                       // rows of the tile in use
  const int t0 = blockIdx.x * R;               // first token of the tile
  const int t = t0 + row;
  const bool valid = row < R && t < T;
  const int c0 = half * 32;                    // this thread's 32 columns of a 64-wide row
  const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);
  const Philox ph(a.seed);
  const bool train = a.training != 0;

  if (tid == 0) {
    umma::mbar_init(barA, 1); umma::mbar_init(barB, 1); umma::mbar_init(barM, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
  uint32_t parM = 0;
  int red_sel = 0;

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
  if (tid == 0) { load_A(0); load_B(0); }

  // pair exchange: sum of a per-thread value over the two column halves of a row
  auto pair_sum = [&](float v) -> float {
    float* buf = sRed + red_sel * 2 * FT;
    buf[half * FT + row] = v;
    __syncthreads();
    const float o = buf[(half ^ 1) * FT + row];
    red_sel ^= 1;
    return v + o;
  };
  // LayerNorm of a 64-wide row held as two 32-column halves; v must already be the bf16-rounded pre-LN value
  auto layer_norm = [&](float (&v)[32], const float* gamma, const float* beta, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    mean = pair_sum(s) * (1.0f / FH);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; q += d * d; }
    rstd = rsqrtf(pair_sum(q) * (1.0f / FH) + kLnEps);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0 + i));
      v[i] = (v[i] - mean) * rstd * g.x + b.x; v[i + 1] = (v[i + 1] - mean) * rstd * g.y + b.y;
      v[i + 2] = (v[i + 2] - mean) * rstd * g.z + b.z; v[i + 3] = (v[i + 3] - mean) * rstd * g.w + b.w;
    }
  };
  auto add_bias = [&](float (&v)[32], const float* bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  };
  auto drop32 = [&](float (&v)[32], uint32_t site) {   // elementwise dropout of columns c0..c0+31 of token row t
    if (a.thr_out == 0 || !train) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)(c0 / 8 + q), site, step, a.thr_out);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * q + i] = ((bits >> i) & 1u) ? v[8 * q + i] * a.inv_keep_out : 0.f;
    }
  };
  // every thread: previous epilogue's smem writes / TMEM reads are ordered before the next MMA batch
  auto phase_sync = [&]() {
    umma::fence_before_sync();
    umma::fence_proxy_async();
    __syncthreads();
  };
  auto wait_mma = [&]() {
    __syncwarp();
    umma::mbar_wait(barM, parM);
    parM ^= 1;
    umma::fence_after_sync();
  };

  // ------------------------------------------------------------------ phase 0: embedding + LN + dropout -> sX
  {
    sMask[row] = valid ? (a.mask[t] != 0 ? 0.f : -1e9f) : 0.f;   // (both halves write the same value)
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    if (valid) {
      long long id = a.ids[t];
      id = id < 0 ? 0 : (id >= a.V ? a.V - 1 : id);
      const bf16* e = a.table + (size_t)id * FH + c0;
      const bf16* p = a.pos + (size_t)(row % S) * FH + c0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
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
    layer_norm(v, a.emb_g, a.emb_b, mean, rstd);
    drop32(v, site_id(SITE_EMB, 0));
    uint32_t pk[16];
    pack32(v, pk);
    if (!valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    }
    st_tile(sX, row, half * 4, pk);
    if (valid && train) st_global32(a.x0 + (size_t)t * FH + c0, pk);
  }

  const float scale = rsqrtf((float)FD);
  const int R16 = (R + 15) & ~15;
  for (int l = 0; l < a.L; ++l) {
    const LayerDev& Ly = a.layers[l];
    // ---------------------------------------------------------------- phase 1: QKV = X Wqkv + b
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::mbar_wait(barA, l & 1);
      const uint32_t xa = umma::smem_addr(sX), wa = umma::smem_addr(sWA);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_bf16_ss(tmem, umma::make_desc_k_sw128(xa + k * 32), desc_mn_sw128(wa + k * 2048, 8192), idesc_bmn(FT, 192), k ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      const int c = half * 3 + j;            // 32-column chunk of the 192-wide row: Q Q K K V V
      float v[32];
      tmem_ld_f32(tlane + c * 32, v);
      add_bias(v, Ly.bqkv + c * 32);
      uint32_t pk[16];
      pack32(v, pk);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      unsigned char* tile = (c >> 1) == 0 ? sQ : ((c >> 1) == 1 ? sK : sV);
      st_tile(tile, row, (c & 1) * 4, pk);
      if (valid && train) st_global32(Ly.qkv + (size_t)t * 192 + c * 32, pk);
    }
    // ---------------------------------------------------------------- phase 2: scores per head + softmax -> P
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t qa = umma::smem_addr(sQ), ka = umma::smem_addr(sK);
#pragma unroll
      for (int hd = 0; hd < FNH; ++hd)
#pragma unroll
        for (int k = 0; k < FD / 16; ++k)
          umma::mma_bf16_ss(tmem + hd * 128, umma::make_desc_k_sw128(qa + hd * FD * 2 + k * 32),
                            umma::make_desc_k_sw128(ka + hd * FD * 2 + k * 32), umma::make_idesc_bf16(FT, 128), k ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
    float inv_l = 0.f;
    {
      const int hd = half;                       // this warpgroup's head
      const int g = valid ? row / S : 0;
      const int ks = g * S, ke = ks + S;         // this row's keys (tile columns)
      // chunks of 32 score columns any lane of this warp needs (tcgen05.ld is warp-collective)
      int c_lo = 1, c_hi = 0;
      if (quad * 32 < R) {
        const int wr_hi = min(R - 1, quad * 32 + 31);
        c_lo = ((quad * 32 / S) * S) >> 5;
        c_hi = ((wr_hi / S + 1) * S - 1) >> 5;
      }
      float m = -INFINITY;
      for (int c = c_lo; c <= c_hi; ++c) {
        float s[32];
        tmem_ld_f32(tlane + hd * 128 + c * 32, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int key = c * 32 + j;
          const bool in = valid && key >= ks && key < ke;
          m = fmaxf(m, in ? s[j] * scale + sMask[key] : -INFINITY);
        }
      }
      // attention-prob dropout keep bits, identical stream to attn_fwd_kernel (k_attn.cu)
      unsigned long long kw0 = ~0ull, kw1 = ~0ull;
      const bool drop = train && a.thr_attn > 0;
      const int bn = (t / S) * FNH + hd, qi = row % S, W = (S + 63) >> 6;
      if (drop && valid) {
        const uint32_t grow = (uint32_t)(bn * S + qi);
        const uint32_t site = site_id(SITE_ATTN_PROBS, l);
        for (int kb = 0; kb < W; ++kb) {
          unsigned long long word = 0ull;
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const uint4 r = ph(grow, (uint32_t)(kb * 8 + o), site, step);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int bit = ((o & 1) * 4 + i) * 8 + (o >> 1) * 2;
              word |= (unsigned long long)((w[i] & 0xFFFFu) >= a.thr_attn ? 1u : 0u) << bit;
              word |= (unsigned long long)((w[i] >> 16) >= a.thr_attn ? 1u : 0u) << (bit + 1);
            }
          }
          if (kb == 0) kw0 = word; else kw1 = word;
          Ly.keep[((size_t)bn * S + qi) * W + kb] = word;
        }
      }
      float lsum = 0.f;
      unsigned char* pt = sP + hd * 2 * TILE_B;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c >= c_lo && c <= c_hi) {
          float s[32];
          tmem_ld_f32(tlane + hd * 128 + c * 32, s);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = c * 32 + j;
            const bool in = valid && key >= ks && key < ke;
            float p = in ? __expf(s[j] * scale + sMask[key] - m) : 0.f;
            lsum += p;
            if (drop) {
              const int jl = key - ks;
              const bool keep = (((jl < 64 ? kw0 : kw1) >> (jl & 63)) & 1ull) != 0ull;
              p = keep ? p * a.inv_keep_attn : 0.f;
            }
            s[j] = p;
          }
          pack32(s, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        }
        st_tile(pt + (c >> 1) * TILE_B, row, (c & 1) * 4, pk);
      }
      if (valid) {
        inv_l = 1.0f / lsum;
        if (train) Ly.lse[(size_t)bn * S + qi] = m + __logf(lsum);
      }
    }
    // ---------------------------------------------------------------- phase 3: ctx = P V per head
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t pa = umma::smem_addr(sP), va = umma::smem_addr(sV);
      for (int hd = 0; hd < FNH; ++hd)
        for (int kk = 0; kk < R16 / 16; ++kk)
          umma::mma_bf16_ss(tmem + 256 + hd * 64, umma::make_desc_k_sw128(pa + hd * 2 * TILE_B + (kk >> 2) * TILE_B + (kk & 3) * 32),
                            desc_mn_sw128(va + kk * 2048, 8192), idesc_bmn(FT, 64), kk ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
    {
      const int hd = half;
      float v[32];
      tmem_ld_f32(tlane + 256 + hd * 64 + hd * FD, v);   // head hd's 32 columns of its own accumulator
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= inv_l;
      uint32_t pk[16];
      pack32(v, pk);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      st_tile(sCtx, row, hd * 4, pk);
      if (valid && train) st_global32(Ly.ctx + (size_t)t * FH + hd * FD, pk);
    }
    // ---------------------------------------------------------------- phase 4: a = x + drop(ctx Wo + bo) ; y = LN1(a)
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t ca = umma::smem_addr(sCtx), wo = umma::smem_addr(sWA + 3 * 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_bf16_ss(tmem + 384, umma::make_desc_k_sw128(ca + k * 32), desc_mn_sw128(wo + k * 2048, 8192), idesc_bmn(FT, 64), k ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
    if (tid == 0 && l + 1 < a.L) load_A(l + 1);   // Wqkv / Wo of the next layer stream in behind the epilogue
    {
      float v[32], res[32];
      tmem_ld_f32(tlane + 384 + c0, v);
      add_bias(v, Ly.bo + c0);
      drop32(v, site_id(SITE_ATTN_OUT, l));
      ld_tile(sX, row, half * 4, res);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += res[i];
      uint32_t pk[16];
      round32(v, pk);                              // LN statistics on the bf16 value backward re-reads
      if (valid && train) st_global32(Ly.a_pre + (size_t)t * FH + c0, pk);
      float mean, rstd;
      layer_norm(v, Ly.g1, Ly.be1, mean, rstd);
      pack32(v, pk);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      st_tile(sY, row, half * 4, pk);
      if (valid && train) {
        st_global32(Ly.y + (size_t)t * FH + c0, pk);
        if (half == 0) { Ly.mean1[t] = mean; Ly.rstd1[t] = rstd; }
      }
    }
    // ---------------------------------------------------------------- phase 5: h = gelu(y W1 + b1)
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      umma::mbar_wait(barB, l & 1);
      const uint32_t ya = umma::smem_addr(sY), w1 = umma::smem_addr(sWB);
      const uint32_t id1 = idesc_bmn(FT, I);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_bf16_ss(tmem, umma::make_desc_k_sw128(ya + k * 32), desc_mn_sw128(w1 + k * 2048, 8192), id1, k ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
#pragma unroll 1
    for (int j = 0; j < I / 64; ++j) {
      const int c = half * (I / 64) + j;      // 32-column chunk of the I-wide row
      float v[32];
      tmem_ld_f32(tlane + c * 32, v);
      add_bias(v, Ly.b1 + c * 32);
      uint32_t pk[16];
      round32(v, pk);                          // GELU of the bf16 pre-activation backward re-reads
      if (valid && train) st_global32(Ly.h_pre + (size_t)t * I + c * 32, pk);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      pack32(v, pk);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      st_tile(sHh + (c >> 1) * TILE_B, row, (c & 1) * 4, pk);
      if (valid && train) st_global32(Ly.h + (size_t)t * I + c * 32, pk);
    }
    // ---------------------------------------------------------------- phase 6: o = y + drop(h W2 + b2) ; out = LN2(o)
    phase_sync();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t ha = umma::smem_addr(sHh), w2 = umma::smem_addr(sWB + I * 128);
      for (int kk = 0; kk < I / 16; ++kk)
        umma::mma_bf16_ss(tmem + 448, umma::make_desc_k_sw128(ha + (kk >> 2) * TILE_B + (kk & 3) * 32), desc_mn_sw128(w2 + kk * 2048, 8192),
                          idesc_bmn(FT, 64), kk ? 1u : 0u);
      umma::mma_commit(barM);
    }
    wait_mma();
    if (tid == 0 && l + 1 < a.L) load_B(l + 1);
    {
      float v[32], res[32];
      tmem_ld_f32(tlane + 448 + c0, v);
      add_bias(v, Ly.b2 + c0);
      drop32(v, site_id(SITE_FFN_OUT, l));
      ld_tile(sY, row, half * 4, res);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += res[i];
      uint32_t pk[16];
      round32(v, pk);
      if (valid && train) st_global32(Ly.o_pre + (size_t)t * FH + c0, pk);
      float mean, rstd;
      layer_norm(v, Ly.g2, Ly.be2, mean, rstd);
      pack32(v, pk);
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      st_tile(sX, row, half * 4, pk);
      if (valid) {
        st_global32(Ly.out + (size_t)t * FH + c0, pk);
        if (train && half == 0) { Ly.mean2[t] = mean; Ly.rstd2[t] = rstd; }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool enc_fused_supported(int H, int N, int S, int I) {
  if (getenv("B4R_DISABLE_FUSED")) return false;
  if (H != FH || N != FNH || S < 1 || S > FT) return false;
  if (I < 64 || I > 256 || (I % 64)) return false;
  return enc_fused_smem_bytes(I) <= 232448;
}
size_t enc_fused_smem_bytes(int I) { return (size_t)OFF_WB + (size_t)I * 256 + FT * 4 + 64 + 1024; }
size_t enc_fused_table_bytes(int L) { return (size_t)L * sizeof(LayerDev) + 128 + (size_t)L * 4 * sizeof(CUtensorMap); }

// Fills the device-side layer table + tensor maps (host staging buffer `host`, same size as the device block).
bool enc_fused_build_tables(const EncFusedLayerHost* layers, int L, int I, const bf16* const* w_ptrs, void* host, void* dev_base) {
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  LayerDev* ld = reinterpret_cast<LayerDev*>(host);
  for (int l = 0; l < L; ++l) {
    const EncFusedLayerHost& h = layers[l];
    ld[l] = LayerDev{h.bqkv, h.bo, h.g1, h.be1, h.b1, h.b2, h.g2, h.be2, h.qkv, h.ctx, h.a_pre, h.y, h.h_pre, h.h, h.o_pre, h.out,
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
  d.B = a.B; d.S = a.S; d.V = a.V; d.L = a.L; d.G = FT / a.S; d.I = a.I; d.training = a.training;
  d.thr_out = drop_threshold16(a.out_drop); d.thr_attn = drop_threshold16(a.attn_drop);
  d.inv_keep_out = 1.0f / (1.0f - (float)d.thr_out / 65536.0f);
  d.inv_keep_attn = 1.0f / (1.0f - (float)d.thr_attn / 65536.0f);
  d.seed = a.seed; d.step = a.step; d.d_step = a.d_step;
  const size_t smem = enc_fused_smem_bytes(a.I);
  static size_t cap = 0;
  if (smem > cap) {
    cudaError_t e = cudaFuncSetAttribute(enc_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cap = smem;
  }
  const int grid = (a.B + d.G - 1) / d.G;
  enc_fwd_fused_kernel<<<grid, 256, smem, st>>>(d);
  return cudaGetLastError();
}

}  // namespace b4r
