// Generation 2 of the masked self-attention FORWARD for the layered path (seq_len <= 256, head dim 32 or 64):
// tcgen05.mma + TMEM + TMA.  One CTA per (sequence, 64-column group of the hidden dimension) = 2 heads of 32 or 1 head of
// 64; the group's Q, K, V rows (up to 256 = two 128-row tiles each) are fetched once by TMA from the fused qkv activation
// and every (head, query tile) runs
//   S = Q_h K_h^T  (128 x 256 fp32 in TMEM)  ->  softmax in registers (512 threads = row x key quarter; max / sum exchanged
//   through shared memory; key padding -1e9, rows past the sequence excluded; Philox keep bits of k_attn.cu)  ->  P (bf16,
//   swizzled K-major tiles)  ->  O = P V  (128 x 64 in TMEM)  ->  ctx = O / l, log-sum-exp, keep bits to global.
// Same numerics / saved tensors as attn_fwd_kernel (k_attn.cu), which it replaces for these shapes; the [B,N,S,S]
// probabilities of Keras MultiHeadAttention (bert4rec_encoder.py:136-147) never exist.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "enc_fused.cuh"

namespace b4r {
using namespace encf;

namespace {
constexpr int TA_THREADS = 512;
// smem tiles of 16 KB: Q0 Q1 K0 K1 V0 V1 | P (4 tiles: 128 rows x 256 keys)
constexpr int TA_Q = 0, TA_K = 2, TA_V = 4, TA_P = 6, TA_TILES = 10;
constexpr int TA_SMEM = TA_TILES * TILE_B + 256 * 4 + 2 * 4 * FT * 4 + 64 + 1024;

struct TAttnDev {
  const int64_t* mask; bf16* ctx; float* lse; unsigned long long* keep;
  int S, H, N;
  uint32_t thr16; float inv_keep; unsigned long long seed; uint32_t site; uint32_t step; const long long* d_step;
};
}  // namespace

template <int D>
__global__ void __launch_bounds__(TA_THREADS, 1) tattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, TAttnDev a) {
  pdl_grid_wait();
  constexpr int NHG = 64 / D;   // heads per 64-column group
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sMask = reinterpret_cast<float*>(smem + TA_TILES * TILE_B);   // [256] additive key mask (-inf past the sequence)
  float* sRed = sMask + 256;                                            // [2][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 2 * 4 * FT);
  uint64_t* barL = bars;       // Q, K, V landed
  uint64_t* barM = bars + 1;   // MMA batch complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int row = quad * 32 + lane;
  const int S = a.S, MT = (S + 127) >> 7;
  const int b = blockIdx.y, g = blockIdx.x;          // sequence, 64-column group
  const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);
  const Philox ph(a.seed);

  if (tid == 0) {
    umma::mbar_init(barL, 1); umma::mbar_init(barM, 2);
    umma::fence_barrier_init();
    umma::mbar_expect_tx(barL, (uint32_t)(3 * MT * TILE_B));
    for (int mt = 0; mt < MT; ++mt) {
      const int r0 = b * S + mt * 128;
      umma::tma_load_2d(smem + (TA_Q + mt) * TILE_B, &tmQKV, g * 64, r0, barL);
      umma::tma_load_2d(smem + (TA_K + mt) * TILE_B, &tmQKV, a.H + g * 64, r0, barL);
      umma::tma_load_2d(smem + (TA_V + mt) * TILE_B, &tmQKV, 2 * a.H + g * 64, r0, barL);
    }
  }
  for (int j = tid; j < 256; j += TA_THREADS) sMask[j] = j < S ? (a.mask[(size_t)b * S + j] != 0 ? 0.f : -1e9f) : -INFINITY;
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
  const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
  const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), 8192);
  uint32_t parM = 0;
  int red_sel = 0;
  auto xchg4 = [&](float v, bool is_max) -> float {   // max or sum over the 4 key-quarter threads of a row
    float* buf = sRed + red_sel * 4 * FT;
    buf[part * FT + row] = v;
    __syncthreads();
    const float v0 = buf[row], v1 = buf[FT + row], v2 = buf[2 * FT + row], v3 = buf[3 * FT + row];
    red_sel ^= 1;
    return is_max ? fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) : (v0 + v1) + (v2 + v3);
  };
  auto wait_mma = [&]() {
    __syncwarp();
    umma::mbar_wait(barM, parM);
    parM ^= 1;
    umma::fence_after_sync();
  };
  // two issuer warps (0, 1): each commits once per batch
  auto issue = [&](auto&& chain) {
    if (warp < 2) {
      if (elect_one()) {
        umma::fence_after_sync();
        chain(warp);
        umma::mma_commit(barM);
      }
      __syncwarp();
    }
  };
  umma::mbar_wait(barL, 0);

  const float scale = rsqrtf((float)D);
  const int KPT = 32 * MT;                 // keys per thread (row x key quarter): 32 (one key tile) or 64 (two)
  const int W = (S + 63) >> 6;
  const bool drop = a.thr16 > 0;
#pragma unroll 1
  for (int h = 0; h < NHG; ++h) {
    const int head = g * NHG + h;
    const int bn = b * a.N + head;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int qi = mt * 128 + row;
      const bool valid = qi < S;
      // ---------------------------------------------------------------- S = Q_h K_h^T over all key tiles
      umma::fence_before_sync();
      umma::fence_proxy_async();
      __syncthreads();
      issue([&](int w) {   // key tile w (issuer 1 idles when there is one key tile)
        if (w < MT) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma::mma_bf16_ss(tmem + w * 128, desc_at(DK0, (TA_Q + mt) * TILE_B + h * D * 2 + k * 32),
                              desc_at(DK0, (TA_K + w) * TILE_B + h * D * 2 + k * 32), idesc_gen(128, 128, 0, 0), k ? 1u : 0u);
        }
      });
      wait_mma();
      // ---------------------------------------------------------------- softmax: this thread owns keys [part*KPT, +KPT)
      float mloc = -INFINITY;
      for (int c = 0; c < KPT / 32; ++c) {
        const int key0 = part * KPT + c * 32;
        float s[32];
        tmem_ld_f32(tlane + key0, s);
#pragma unroll
        for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, (valid && key0 + j < S) ? s[j] * scale + sMask[key0 + j] : -INFINITY);
      }
      const float m = xchg4(mloc, true);
      float lsum = 0.f;
      for (int c = 0; c < KPT / 32; ++c) {
        const int key0 = part * KPT + c * 32;
        float s[32];
        tmem_ld_f32(tlane + key0, s);
        uint32_t bits = 0xFFFFFFFFu;
        if (drop && valid && key0 < S) {
          // keep bits of keys key0..key0+31: identical stream to attn_fwd_kernel (k_attn.cu)
          const int ci = key0 >> 5, kb = ci >> 1, hf = ci & 1;
          const uint32_t grow = (uint32_t)(bn * S + qi);
          bits = 0u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int o = hf + 2 * u;
            const uint4 r = ph(grow, (uint32_t)(kb * 8 + o), a.site, step);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int bit = i * 8 + u * 2;
              bits |= ((w[i] & 0xFFFFu) >= a.thr16 ? 1u : 0u) << bit;
              bits |= ((w[i] >> 16) >= a.thr16 ? 1u : 0u) << (bit + 1);
            }
          }
          reinterpret_cast<uint32_t*>(a.keep)[(((size_t)bn * S + qi) * W + kb) * 2 + hf] = bits;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = (valid && key0 + j < S) ? s[j] * scale + sMask[key0 + j] : -INFINITY;
          float p = v == -INFINITY ? 0.f : __expf(v - m);
          lsum += p;
          if (drop) p = ((bits >> j) & 1u) ? p * a.inv_keep : 0.f;
          s[j] = p;
        }
        uint32_t pk[16];
        pack_n<32>(s, pk);
        st_tile<4>(smem + (TA_P + (key0 >> 6)) * TILE_B, row, ((key0 & 63) >> 3), pk);
      }
      const float l = xchg4(lsum, false);
      // ---------------------------------------------------------------- O = P V
      umma::fence_before_sync();
      umma::fence_proxy_async();
      __syncthreads();
      issue([&](int w) {
        if (w == 0) {
          for (int kk = 0; kk < 8 * MT; ++kk)
            umma::mma_bf16_ss(tmem + 256, desc_at(DK0, (TA_P + (kk >> 2)) * TILE_B + (kk & 3) * 32),
                              desc_at(DMN0, (TA_V + (kk >> 3)) * TILE_B + (kk & 7) * 2048), idesc_gen(128, 64, 0, 1), kk ? 1u : 0u);
        }
      });
      wait_mma();
      // ---------------------------------------------------------------- ctx = O / l  (16 columns per thread; D = 32: parts 0, 1)
      if (part * 16 < D) {
        float v[16];
        tmem_ld_f16(tlane + 256 + h * D + part * 16, v);
        if (valid) {
          const float inv = 1.0f / l;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= inv;
          uint32_t pk[8];
          pack_n<16>(v, pk);
          st_global<8>(a.ctx + ((size_t)b * S + qi) * a.H + g * 64 + h * D + part * 16, pk);
          if (part == 0) a.lse[(size_t)bn * S + qi] = m + __logf(l);
        }
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

// Opt-in (B4R_ENABLE_TATTN=1).  Measured on B200 against attn_fwd_kernel (mma.sync, flash-style, several CTAs per SM):
// C4 (S 200, D 64) 676 vs 529 us per layer, C3 (S 200, D 32) 94 vs 69 us.  The 256-key score tile (320 TMEM columns,
// 160 KB of tiles) allows one CTA per SM, whose MMA -> softmax -> MMA -> epilogue chain then runs unoverlapped; a
// key-blocked (online-softmax) formulation with two CTAs per SM is the next step before this becomes the default.
bool tattn_fwd_supported(const AttnArgs& a) {
  if (!getenv("B4R_ENABLE_TATTN")) return false;
  const int D = a.H / a.N;
  if ((D != 32 && D != 64) || a.H % 64 || a.S > 256 || a.S < 16) return false;
  if (((uintptr_t)a.qkv & 15)) return false;
  return true;
}

cudaError_t launch_tattn_fwd(const AttnArgs& a, cudaStream_t st) {
  CUtensorMap tm;
  const uint64_t T = (uint64_t)a.B * a.S;
  if (!make_tmap_bf16_sw128(&tm, a.qkv, T, (uint64_t)3 * a.H, (uint64_t)3 * a.H, 128)) return cudaErrorInvalidValue;
  TAttnDev d;
  d.mask = a.mask; d.ctx = a.ctx; d.lse = a.lse; d.keep = reinterpret_cast<unsigned long long*>(a.keep_bits);
  d.S = a.S; d.H = a.H; d.N = a.N;
  d.thr16 = drop_threshold16(a.drop_rate);
  d.inv_keep = 1.0f / (1.0f - (float)d.thr16 / 65536.0f);
  d.seed = a.seed; d.site = a.site; d.step = a.step; d.d_step = a.d_step;
  const int D = a.H / a.N;
  dim3 grid(a.H / 64, a.B);
  static bool done32 = false, done64 = false;
  if (D == 32) {
    if (!done32) { cudaError_t e = cudaFuncSetAttribute(tattn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM); if (e != cudaSuccess) return e; done32 = true; }
    launch_pdl(tattn_fwd_kernel<32>, dim3(grid), dim3(TA_THREADS), (size_t)(TA_SMEM), st, tm, d);
  } else {
    if (!done64) { cudaError_t e = cudaFuncSetAttribute(tattn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM); if (e != cudaSuccess) return e; done64 = true; }
    launch_pdl(tattn_fwd_kernel<64>, dim3(grid), dim3(TA_THREADS), (size_t)(TA_SMEM), st, tm, d);
  }
  return cudaGetLastError();
}

}  // namespace b4r
