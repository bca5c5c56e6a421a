// Generation 2 of the masked self-attention of the layered path (seq_len <= 256, head dim 32 or 64): tcgen05.mma with TMEM
// accumulators, operands staged by TMA, persistent warp-specialised CTAs (one per SM).  Replaces Keras MultiHeadAttention
// inside tfm TransformerEncoderBlock (bert4rec_encoder.py:136-147,216-222) for every shape the fused
// whole-encoder kernels do not cover; same numerics, saved tensors (log-sum-exp, bit-packed dropout keep mask) and Philox
// streams as the mma.sync kernels of k_attn.cu, which stay as the parity partner (session flag 6).
//
// Work item = (sequence b, 64-column group g of the hidden dimension) = 1 head of 64 or 2 heads of 32.  The item's Q, K, V
// (and dO, O in the backward) rows are fetched once by TMA as [128 rows][64 columns] 128-byte-swizzled tiles (two row tiles
// cover seq_len <= 256; rows past the sequence belong to the next sequence or are zero-filled and are masked out).
//
// BACKWARD (fattn_bwd_kernel).  Transposed arrangement, rows = keys: per (head, key tile kt, query tile qt)
//   S^T = K_kt Q_qt^T, dP^T = V_kt dO_qt^T              (two [128 x 128] fp32 accumulators in TMEM)
//   P^T = exp2(S^T c + mask - lse), Pd^T = keep P^T / (1-r), dS^T = P^T (keep dP^T / (1-r) - delta)   (registers -> bf16 smem)
//   dV_kt += Pd^T dO_qt ; dK_kt += dS^T Q_qt ; dQ_qt += dS K_kt  (dS = the SAME smem tile read MN-major: no transpose copy)
// TMEM (512 columns): S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ (two query tiles) 2 x 64.  256 compute threads (TMEM lane
// quadrant x query-column half) + one control warp (TMA + single-thread MMA issue).  The score MMAs of iteration i+1 are issued
// as soon as the compute threads have pulled iteration i out of TMEM (they hold P / dS packed in registers), so the tensor pipe
// runs S(i+1) and the three gradient GEMMs of i while the threads do the exponentials of i+1.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "enc_fused.cuh"

namespace b4r {
using namespace encf;

namespace {
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[32]) { umma::tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[16]) { umma::tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[8]) { umma::tmem_ld8(taddr, r); }
constexpr int FA_NG = 2;                 // column groups of the compute threads: 128 * FA_NG threads = TMEM lane quadrant x column group
// 32 score columns, or (last chunk of a sequence whose padded length is an odd multiple of 16) 16 columns + zeros: columns past
// the score MMA's N were never written and may hold NaN bit patterns
__device__ __forceinline__ void ld_chunk(uint32_t taddr, bool full, uint32_t (&r)[32]) {
  if (full) {
    umma::tmem_ld32(taddr, r);
  } else {
    umma::tmem_ld16(taddr, reinterpret_cast<uint32_t(&)[16]>(r));
#pragma unroll
    for (int j = 16; j < 32; ++j) r[j] = 0u;
  }
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


#ifdef FATTN_DEBUG
// development aid: a wait that gives up after ~2^24 polls and reports which barrier / phase never completed
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(umma::smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int id) {
  for (uint32_t i = 0; i < (1u << 17); ++i)
    if (mbar_try(bar, parity)) return;
  if ((threadIdx.x & 31) == 0) printf("TIMEOUT cta %d warp %d barrier %d parity %u\n", blockIdx.x, threadIdx.x >> 5, id, parity);
}
#define MBAR_WAIT(bar, par, id) mbar_wait_dbg(bar, par, id)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ unsigned long long g_fattn_ts[128];
#define FTS(cond, id) do { if (cond) g_fattn_ts[id] = gtimer(); } while (0)
#else
#define FTS(cond, id) do { } while (0)
#define MBAR_WAIT(bar, par, id) umma::mbar_wait(bar, par)
#endif

// ------------------------------------------------------------------------------------------------ backward
constexpr int FB_CT = 128 * FA_NG;       // compute threads
constexpr int FB_THREADS = FB_CT + 32;   // + control warp
constexpr int FB_Q = 0, FB_K = 2, FB_V = 4, FB_DO = 6, FB_P = 8, FB_DS = 10, FB_TILES = 12;
constexpr int FB_OFF_LD = FB_TILES * TILE_B;            // float2 [2][256]  (lse * log2e, delta)
constexpr int FB_OFF_MASK = FB_OFF_LD + 2 * 256 * 8;    // float [256]
constexpr int FB_OFF_KEEP = FB_OFF_MASK + 256 * 4;      // u64 [2][256 * 4]
constexpr int FB_OFF_BAR = FB_OFF_KEEP + 2 * 1024 * 8;
constexpr int FB_SMEM = FB_OFF_BAR + 128 + 1024;        // + alignment slack

struct FAttnBwdDev {
  const int64_t* mask; const float* lse; const unsigned long long* keep; bf16* dqkv;
  int B, S, H, N;
  uint32_t thr16; float inv_keep;
};
}  // namespace

template <int D>
__global__ void __launch_bounds__(FB_THREADS, 1)
fattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmO, FAttnBwdDev a, uint32_t stagger_ns) {
  pdl_grid_wait_single_wave();
  constexpr int NHG = 64 / D;      // heads per 64-column group
  constexpr int CPT = D / FA_NG;   // accumulator columns drained per thread
  constexpr int QPT = 128 / FA_NG; // query columns of a score tile per thread
  constexpr int NCH = QPT / 16;    // in 16-column chunks
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float2* sLD = reinterpret_cast<float2*>(smem + FB_OFF_LD);
  float* sMask = reinterpret_cast<float*>(smem + FB_OFF_MASK);
  unsigned long long* sKeep = reinterpret_cast<unsigned long long*>(smem + FB_OFF_KEEP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FB_OFF_BAR);
  // item loads arrive in three groups: barL (Q0 K0 V0 dO0 -- with two row tiles these are prefetched while the previous item's last
  // iteration runs, see the control warp), barLb (O0 O1 dO1: what the delta pass needs) and barLc (Q1 K1 V1)
  uint64_t *barL = bars, *barS = bars + 1, *barSF = bars + 2, *barPD = bars + 3, *barG = bars + 4, *barAF = bars + 5, *barLb = bars + 6, *barLc = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  auto tile = [&](int i) -> unsigned char* { return smem + i * TILE_B; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, H = a.H, MT = (S + 127) >> 7, S16 = (S + 15) & ~15, W = (S + 63) >> 6;
  const int G = H >> 6, n_items = a.B * G;
  const int n_it = NHG * MT * MT;                 // iterations per item
  const bool drop = a.thr16 > 0;

  if (tid == 0) {
    umma::mbar_init(barL, 1); umma::mbar_init(barS, 1); umma::mbar_init(barSF, FB_CT); umma::mbar_init(barPD, FB_CT);
    umma::mbar_init(barG, 1); umma::mbar_init(barAF, FB_CT); umma::mbar_init(barLb, 1); umma::mbar_init(barLc, 1);
    umma::fence_barrier_init();
  }
  if (warp == FB_CT / 32) {
    umma::tmem_alloc<512>(tmem_holder);
    if (lane == 0) { umma::prefetch_tensormap(&tmQKV); umma::prefetch_tensormap(&tmDO); umma::prefetch_tensormap(&tmO); }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;
  constexpr uint32_t TC_S = 0, TC_DP = 128, TC_DV = 256, TC_DK = 320, TC_DQ = 384;

  if (warp == FB_CT / 32) {
    // ============================================================ control warp: TMA loads + MMA issue
    const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
    const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), TILE_B);
    auto dk = [&](int t, int off) -> uint64_t { return desc_at(DK0, (uint32_t)(t * TILE_B + off)); };
    auto dmn = [&](int t, int off) -> uint64_t { return desc_at(DMN0, (uint32_t)(t * TILE_B + off)); };
    uint32_t git = 0, gblk = 0, nitem = 0;    // global iteration / (head, key tile) block / item counters of this CTA
    auto issue_scores = [&](int i) {          // iteration i of the current item: (h, kt, qt)
      const int h = i / (MT * MT), kt = (i / MT) % MT, qt = i % MT;
      const int qext = min(128, S16 - qt * 128);
      if (elect_one()) {
        umma::fence_after_sync();
        const uint32_t id = idesc_gen(128, qext, 0, 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma::mma_bf16_ss(tmem + TC_S, dk(FB_K + kt, h * D * 2 + k * 32), dk(FB_Q + qt, h * D * 2 + k * 32), id, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma::mma_bf16_ss(tmem + TC_DP, dk(FB_V + kt, h * D * 2 + k * 32), dk(FB_DO + qt, h * D * 2 + k * 32), id, k ? 1u : 0u);
        umma::mma_commit(barS);
      }
      __syncwarp();
    };
    // every CTA takes the same time per item: without a stagger all 148 CTAs issue their item loads in the same microsecond and
    // then leave the memory system idle; a one-off start offset spreads the bursts
    if (n_items > (int)gridDim.x && stagger_ns) __nanosleep((blockIdx.x % 8) * stagger_ns);
    // first row tile of an item: Q0 K0 V0 dO0.  With two row tiles the LAST iteration of an item (key tile 1 x query tile 1) touches
    // none of them, so the next item's copies are requested at the top of that iteration -- once the gradient MMAs of the iteration
    // before it have completed -- and are in place when the next item starts: its first score MMAs are issued at once and the
    // compute threads wait only for O0 O1 dO1 (48 KB instead of 160 KB).
    auto load_first = [&](int it) {
      if (elect_one()) {
        const int bb = it / G, gg = it % G, r0 = bb * S;
        umma::mbar_expect_tx(barL, (uint32_t)(4 * TILE_B));
        umma::tma_load_2d(tile(FB_Q), &tmQKV, gg * 64, r0, barL);
        umma::tma_load_2d(tile(FB_K), &tmQKV, H + gg * 64, r0, barL);
        umma::tma_load_2d(tile(FB_V), &tmQKV, 2 * H + gg * 64, r0, barL);
        umma::tma_load_2d(tile(FB_DO), &tmDO, gg * 64, r0, barL);
      }
      __syncwarp();
    };
    bool prefetched = false;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      const int b = item / G, g = item % G;
      const bool tsk = blockIdx.x == 0 && nitem == 2 && lane == 0;
      FTS(tsk, 64);
      if (git > 0) umma::mbar_wait(barG, (git - 1) & 1);     // every MMA of the previous item has completed: tiles are free
      FTS(tsk, 65);
      if (!prefetched) load_first(item);
      if (elect_one()) {
        const int r0 = b * S;
        umma::mbar_expect_tx(barLb, (uint32_t)((MT == 2 ? 3 : 1) * TILE_B));
        umma::tma_load_2d(tile(FB_P), &tmO, g * 64, r0, barLb);
        if (MT == 2) {
          umma::tma_load_2d(tile(FB_P + 1), &tmO, g * 64, r0 + 128, barLb);
          umma::tma_load_2d(tile(FB_DO + 1), &tmDO, g * 64, r0 + 128, barLb);
          umma::mbar_expect_tx(barLc, (uint32_t)(3 * TILE_B));
          umma::tma_load_2d(tile(FB_Q + 1), &tmQKV, g * 64, r0 + 128, barLc);
          umma::tma_load_2d(tile(FB_K + 1), &tmQKV, H + g * 64, r0 + 128, barLc);
          umma::tma_load_2d(tile(FB_V + 1), &tmQKV, 2 * H + g * 64, r0 + 128, barLc);
        }
      }
      __syncwarp();
      FTS(tsk, 66);
      umma::mbar_wait(barL, nitem & 1);
      FTS(tsk, 67);
      if (git > 0) umma::mbar_wait(barSF, (git - 1) & 1);    // score accumulators of the previous iteration were read
      issue_scores(0);
      FTS(tsk, 68);
      for (int i = 0; i < n_it; ++i, ++git) {
        const int h = i / (MT * MT), kt = (i / MT) % MT, qt = i % MT;
        if (i + 1 < n_it) {
          umma::mbar_wait(barSF, git & 1);
          FTS(tsk, 70 + i * 4);
          if (i == 0 && MT == 2) { umma::mbar_wait(barLb, nitem & 1); umma::mbar_wait(barLc, nitem & 1); }   // iteration 1 is the first to read row tile 1
          issue_scores(i + 1);
        }
        if (i == n_it - 1) {
          prefetched = MT == 2 && item + (int)gridDim.x < n_items;
          if (prefetched) {
            umma::mbar_wait(barG, (git - 1) & 1);    // gradient MMAs of every earlier iteration are done: Q0 K0 V0 dO0 are free
            load_first(item + gridDim.x);
          }
        }
        umma::mbar_wait(barPD, git & 1);
        FTS(tsk, 71 + i * 4);                      // Pd^T / dS^T tiles of iteration i are in shared memory
        if (qt == 0) {
          if (gblk > 0) umma::mbar_wait(barAF, (gblk - 1) & 1);   // the accumulators of the previous block were drained
          ++gblk;
        }
        const int qext = min(128, S16 - qt * 128), kext = min(128, S16 - kt * 128);
        if (elect_one()) {
          umma::fence_after_sync();
          const uint32_t id_kv = idesc_gen(128, 64, 0, 1), id_q = idesc_gen(128, 64, 1, 1);
          for (int kk = 0; kk < qext / 16; ++kk)
            umma::mma_bf16_ss(tmem + TC_DV, dk(FB_P + (kk >> 2), (kk & 3) * 32), dmn(FB_DO + qt, kk * 2048), id_kv, (qt | kk) ? 1u : 0u);
          for (int kk = 0; kk < qext / 16; ++kk)
            umma::mma_bf16_ss(tmem + TC_DK, dk(FB_DS + (kk >> 2), (kk & 3) * 32), dmn(FB_Q + qt, kk * 2048), id_kv, (qt | kk) ? 1u : 0u);
          for (int kk = 0; kk < kext / 16; ++kk)
            umma::mma_bf16_ss(tmem + TC_DQ + qt * 64, dmn(FB_DS, kk * 2048), dmn(FB_K + kt, kk * 2048), id_q, (kt | kk) ? 1u : 0u);
          umma::mma_commit(barG);
        }
        __syncwarp();
        FTS(tsk, 72 + i * 4);
        (void)h;
      }
    }
  } else {
    // ============================================================ compute warps
    const int quad = warp & 3, wg = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const float scale = rsqrtf((float)D), c1 = scale * kLog2e;
    uint32_t git = 0, g_waited = 0, nitem = 0;
    const uint32_t sLD_a = umma::smem_addr(sLD), sKeep_a = umma::smem_addr(sKeep);
    auto need_g = [&](uint32_t upto) {     // gradient MMAs of global iterations < upto have completed
      while (g_waited < upto) { umma::mbar_wait(barG, g_waited & 1); ++g_waited; }
      umma::fence_after_sync();
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      const int b = item / G, g = item % G;
      const bool tsc = blockIdx.x == 0 && nitem == 2 && tid == 0;
      FTS(tsc, 0);
      // ---- per-item vectors that do not depend on the tiles (issued before the wait: their latency hides behind the TMA loads).
      // Every thread is past its last read of these arrays (phase A of the previous item's last iteration) by the time any
      // thread gets here: the final drain waits for the gradient MMAs, which wait for every thread's phase B.
      if (tid < 256) sMask[tid] = tid < S ? (a.mask[(size_t)b * S + tid] != 0 ? 0.f : -1e9f * kLog2e) : -INFINITY;
#pragma unroll
      for (int h = 0; h < NHG; ++h) {
        const int bn = b * a.N + g * NHG + h;
        if (tid < 256) sLD[h * 256 + tid].x = tid < S ? a.lse[(size_t)bn * S + tid] * kLog2e : INFINITY;
        if (drop) {
          // (all loads first, then the stores: one round trip instead of one per loop iteration)
          const unsigned long long* src = a.keep + (size_t)bn * S * W;
          constexpr int KPT = (1024 + FB_CT - 1) / FB_CT;
          unsigned long long kv[KPT];
#pragma unroll
          for (int q = 0; q < KPT; ++q) { const int i = tid + q * FB_CT; kv[q] = i < S * W ? src[i] : 0ull; }
#pragma unroll
          for (int q = 0; q < KPT; ++q) { const int i = tid + q * FB_CT; if (i < S * W) sKeep[h * 1024 + i] = kv[q]; }
        }
      }
      umma::mbar_wait(barL, nitem & 1);
      umma::mbar_wait(barLb, nitem & 1);
      FTS(tsc, 1);
      // ---- delta = rowsum(dO * O) per (head, query)
#pragma unroll
      for (int h = 0; h < NHG; ++h) {
        if (tid < MT * 128) {
          float dv[D], ov[D], dl = 0.f;
          ld_tile<D / 8>(tile(FB_DO + (tid >> 7)), tid & 127, h * (D / 8), dv);
          ld_tile<D / 8>(tile(FB_P + (tid >> 7)), tid & 127, h * (D / 8), ov);
#pragma unroll
          for (int i = 0; i < D; ++i) dl += dv[i] * ov[i];
          sLD[h * 256 + tid].y = dl;
        } else if (tid < 256) {
          sLD[h * 256 + tid].y = 0.f;
        }
      }
      named_bar_sync(1, FB_CT);
      FTS(tsc, 2);
      for (int i = 0; i < n_it; ++i, ++git) {
        const int h = i / (MT * MT), kt = (i / MT) % MT, qt = i % MT;
        const int qext = min(128, S16 - qt * 128);
        const int key = kt * 128 + r;
        // a warp whose 32 key rows all lie past the padded sequence contributes nothing: its rows of P^T / dS^T are never read by
        // the dQ MMA (K extent = padded keys) and only feed rows of dK / dV that are not stored
        const bool warp_live = kt * 128 + quad * 32 < S16;
        const float mk = sMask[key];
        const uint32_t ld_a = sLD_a + (uint32_t)(h * 256 + qt * 128) * 8;
        const uint32_t kp_a = sKeep_a + (uint32_t)(h * 1024 + (qt * 128) * W + (key >> 6)) * 8 + ((key & 32) ? 4u : 0u);
        const uint32_t kbit = key & 31;
        // ---- phase A: scores -> Pd^T, dS^T packed in registers
        umma::mbar_wait(barS, git & 1);
        FTS(tsc, 4 + i * 6);
        umma::fence_after_sync();
        uint32_t pkp[NCH][8], pkd[NCH][8];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int q0 = wg * QPT + c * 16;
          if (q0 < qext && warp_live) {
            float s[16], dp[16];
            {
              uint32_t rs[16], rd[16];
              umma::tmem_ld16(tlane + TC_S + q0, rs);
              umma::tmem_ld16(tlane + TC_DP + q0, rd);
              umma::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) { s[j] = __uint_as_float(rs[j]); dp[j] = __uint_as_float(rd[j]); }
            }
            if (drop) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float4 l4 = lds_f4(ld_a + (uint32_t)(q0 + j) * 8);         // (lse, delta) of two queries
                const uint32_t w0 = lds_u32(kp_a + (uint32_t)((q0 + j) * W) * 8), w1 = lds_u32(kp_a + (uint32_t)((q0 + j + 1) * W) * 8);
                const float p0 = ex2(fmaf(s[j], c1, mk) - l4.x), p1 = ex2(fmaf(s[j + 1], c1, mk) - l4.z);
                const float t0 = ((w0 >> kbit) & 1u) ? a.inv_keep : 0.f, t1 = ((w1 >> kbit) & 1u) ? a.inv_keep : 0.f;
                s[j] = p0 * t0; s[j + 1] = p1 * t1;
                dp[j] = p0 * fmaf(dp[j], t0, -l4.y); dp[j + 1] = p1 * fmaf(dp[j + 1], t1, -l4.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float4 l4 = lds_f4(ld_a + (uint32_t)(q0 + j) * 8);
                const float p0 = ex2(fmaf(s[j], c1, mk) - l4.x), p1 = ex2(fmaf(s[j + 1], c1, mk) - l4.z);
                s[j] = p0; s[j + 1] = p1;
                dp[j] = p0 * (dp[j] - l4.y); dp[j + 1] = p1 * (dp[j + 1] - l4.w);
              }
            }
            pack_n<16>(s, pkp[c]);
            pack_n<16>(dp, pkd[c]);
          }
        }
        umma::fence_before_sync();
        umma::mbar_arrive(barSF);
        FTS(tsc, 5 + i * 6);
        // ---- phase B: registers -> swizzled tiles [key][query] (free once the gradient MMAs of the previous iteration are done)
        need_g(git);
        FTS(tsc, 6 + i * 6);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int q0 = wg * QPT + c * 16;
          if (q0 < qext && warp_live) {
            st_tile<2>(tile(FB_P + (q0 >> 6)), r, (q0 & 63) >> 3, pkp[c]);
            st_tile<2>(tile(FB_DS + (q0 >> 6)), r, (q0 & 63) >> 3, pkd[c]);
          }
        }
        umma::fence_proxy_async();
        umma::mbar_arrive(barPD);
        FTS(tsc, 7 + i * 6);
        // ---- end of a (head, key tile) block: drain dK / dV (and dQ after the head's last key tile).  The accumulators go through
        // the (now idle) dS^T tiles so that the global stores are row-contiguous: 8 (4) lanes x 16 bytes cover the 128 (64) bytes
        // of one row; one thread per row writing its own 64 bytes made every store instruction touch 32 half-used sectors.
        if (qt == MT - 1) {
          need_g(git + 1);
          FTS(tsc, 8 + i * 6);
          const int col0 = h * D + wg * CPT;
          const bool last_kt = kt == MT - 1;
          auto stage = [&](uint32_t tcol, int t, float mul, bool rows_live) {
            if (rows_live) {
              uint32_t rr[CPT];
              tmem_ld_n(tlane + tcol + col0, rr);
              umma::tmem_ld_wait();
              float f[CPT];
#pragma unroll
              for (int j = 0; j < CPT; ++j) f[j] = __uint_as_float(rr[j]) * mul;
              uint32_t pk[CPT / 2];
              pack_n<CPT>(f, pk);
              st_tile<CPT / 8>(tile(FB_DS + t), r, col0 >> 3, pk);
            }
          };
          auto flush_tile = [&](int t, int row0, int col_off) {     // staged tile t -> dqkv rows [row0, row0 + 128) of this sequence
            constexpr int LPR = D / 8, RPP = FB_CT / LPR;            // lanes per row, rows per pass
            const int li = tid % LPR;
            for (int rr = tid / LPR; rr < 128; rr += RPP) {
              if (row0 + rr < S) {
                const int chunk = (h * D) / 8 + li;
                const uint4 v = *reinterpret_cast<const uint4*>(tile(FB_DS + t) + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                *reinterpret_cast<uint4*>(a.dqkv + ((size_t)b * S + row0 + rr) * 3 * H + col_off + g * 64 + h * D + li * 8) = v;
              }
            }
          };
          const bool krows = kt * 128 + quad * 32 < S;
          stage(TC_DV, 0, 1.0f, krows);
          stage(TC_DK, 1, scale, krows);
          if (!last_kt) { umma::fence_before_sync(); umma::mbar_arrive(barAF); }
          named_bar_sync(1, FB_CT);
          flush_tile(0, kt * 128, 2 * H);
          flush_tile(1, kt * 128, H);
          named_bar_sync(1, FB_CT);
          if (last_kt) {
            for (int t = 0; t < MT; ++t) stage(TC_DQ + t * 64, t, scale, t * 128 + quad * 32 < S);
            umma::fence_before_sync();
            umma::mbar_arrive(barAF);
            named_bar_sync(1, FB_CT);
            for (int t = 0; t < MT; ++t) flush_tile(t, t * 128, 0);
            named_bar_sync(1, FB_CT);
          }
          FTS(tsc, 9 + i * 6);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == FB_CT / 32) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------ forward
// Per (head, query tile): S = Q K^T over ALL keys of the sequence ([128 x S16] fp32 in TMEM, S16 <= 256: no online-softmax
// correction is needed) -> two passes over TMEM by 256 threads (row x key half; row max / sum exchanged through shared
// memory): max, then p = exp2(s c + mask - m), Philox keep bits (identical stream to k_attn.cu), bf16 P into swizzled K-major
// tiles -> O = P V ([128 x 64] accumulated over the same TMEM columns the scores occupied) -> ctx = O / l, log-sum-exp.
// TMEM is double buffered (2 x 256 columns): the control warp issues the score MMA of tile t+1 while the threads work on tile t,
// and the epilogue of tile t-1 (which waits for its PV MMA) runs between the two passes of tile t.
namespace {
constexpr int FF_NG = 3;                              // column groups of the two score passes (the phases are dependency-latency bound:
                                                      // 12 warps hide more of it than 8; 16 do not fit the register file)
constexpr int FF_CT = 128 * FF_NG, FF_THREADS = FF_CT + 32;
constexpr int FF_Q = 0, FF_K = 2, FF_V = 4, FF_P = 6, FF_TILES = 10;
constexpr int FF_OFF_MASK = FF_TILES * TILE_B;           // float [2][256]  (item parity)
constexpr int FF_OFF_REDM = FF_OFF_MASK + 2 * 256 * 4;   // float [2 tile parity][FF_NG column groups][128]
constexpr int FF_OFF_REDL = FF_OFF_REDM + 2 * FF_NG * 128 * 4;
constexpr int FF_OFF_PAD = FF_OFF_REDL + 2 * FF_NG * 128 * 4;   // int [2] (item parity): != 0 when the sequence has padded keys
constexpr int FF_OFF_BAR = FF_OFF_PAD + 16;
constexpr int FF_SMEM = FF_OFF_BAR + 128 + 1024;

struct FAttnFwdDev {
  const int64_t* mask; bf16* ctx; float* lse; unsigned long long* keep;
  int B, S, H, N;
  uint32_t thr16; float inv_keep; unsigned long long seed; uint32_t site; uint32_t step; const long long* d_step;
};
}  // namespace

template <int D>
__global__ void __launch_bounds__(FF_THREADS, 1) fattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, FAttnFwdDev a) {
  pdl_grid_wait_single_wave();
  constexpr int NHG = 64 / D, CPT = D / 2;          // the accumulator is drained by column groups 0 and 1
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sMask = reinterpret_cast<float*>(smem + FF_OFF_MASK);
  float* sRedM = reinterpret_cast<float*>(smem + FF_OFF_REDM);
  float* sRedL = reinterpret_cast<float*>(smem + FF_OFF_REDL);
  int* sPad = reinterpret_cast<int*>(smem + FF_OFF_PAD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FF_OFF_BAR);
  // barS is per TMEM buffer: the score MMAs of tiles t and t+1 need nothing from the compute threads, so a single barrier could
  // complete two phases before a thread polls the first one (and a parity wait cannot tell phase t from phase t+2)
  uint64_t *barLQ = bars, *barLV = bars + 1, *barS = bars + 2, *barP = bars + 4, *barO = bars + 5, *barOF = bars + 6;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  auto tile = [&](int i) -> unsigned char* { return smem + i * TILE_B; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, H = a.H, MT = (S + 127) >> 7, S16 = (S + 15) & ~15, W = (S + 63) >> 6;
  const int G = H >> 6, n_items = a.B * G, n_t = NHG * MT;
  const bool drop = a.thr16 > 0;

  if (tid == 0) {
    umma::mbar_init(barLQ, 1); umma::mbar_init(barLV, 1); umma::mbar_init(barS, 1); umma::mbar_init(barS + 1, 1); umma::mbar_init(barP, FF_CT);
    umma::mbar_init(barO, 1); umma::mbar_init(barOF, FF_CT);
    umma::fence_barrier_init();
    sPad[0] = 0; sPad[1] = 0;
  }
  if (warp == FF_CT / 32) {
    umma::tmem_alloc<512>(tmem_holder);
    if (lane == 0) umma::prefetch_tensormap(&tmQKV);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;

  if (warp == FF_CT / 32) {
    // ============================================================ control warp
    const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
    const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), TILE_B);
    auto dk = [&](int t, int off) -> uint64_t { return desc_at(DK0, (uint32_t)(t * TILE_B + off)); };
    auto dmn = [&](int t, int off) -> uint64_t { return desc_at(DMN0, (uint32_t)(t * TILE_B + off)); };
    auto load_qk = [&](int item) {
      const int b = item / G, g = item % G;
      if (elect_one()) {
        umma::mbar_expect_tx(barLQ, (uint32_t)(2 * MT * TILE_B));
        for (int mt = 0; mt < MT; ++mt) {
          umma::tma_load_2d(tile(FF_Q + mt), &tmQKV, g * 64, b * S + mt * 128, barLQ);
          umma::tma_load_2d(tile(FF_K + mt), &tmQKV, H + g * 64, b * S + mt * 128, barLQ);
        }
      }
      __syncwarp();
    };
    auto load_v = [&](int item) {
      const int b = item / G, g = item % G;
      if (elect_one()) {
        umma::mbar_expect_tx(barLV, (uint32_t)(MT * TILE_B));
        for (int mt = 0; mt < MT; ++mt) umma::tma_load_2d(tile(FF_V + mt), &tmQKV, 2 * H + g * 64, b * S + mt * 128, barLV);
      }
      __syncwarp();
    };
    uint32_t gt = 0, nitem = 0;
    auto issue_scores = [&](int t, uint32_t gtile) {   // tile t = (h, mt) of the current item into TMEM buffer gtile & 1
      const int h = t / MT, mt = t % MT;
      if (gtile >= 2) MBAR_WAIT(barOF, (gtile - 2) & 1, 5);   // the epilogue of the tile that used this buffer is done
      if (elect_one()) {
        umma::fence_after_sync();
        const uint32_t id = idesc_gen(128, S16, 0, 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma::mma_bf16_ss(tmem + (gtile & 1) * 256, dk(FF_Q + mt, h * D * 2 + k * 32), dk(FF_K, h * D * 2 + k * 32), id, k ? 1u : 0u);
        umma::mma_commit(barS + (gtile & 1));
      }
      __syncwarp();
    };
    if ((int)blockIdx.x < n_items) { load_qk(blockIdx.x); load_v(blockIdx.x); }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      const int next = item + gridDim.x;
      MBAR_WAIT(barLQ, nitem & 1, 0);
      issue_scores(0, gt);
      for (int t = 0; t < n_t; ++t, ++gt) {
        if (t + 1 < n_t) {
          issue_scores(t + 1, gt + 1);
        } else if (next < n_items) {
          MBAR_WAIT(barS + (gt & 1), (gt >> 1) & 1, 2);       // last score MMA of the item is complete: Q / K tiles are free
          load_qk(next);
        }
        if (t == 0) MBAR_WAIT(barLV, nitem & 1, 1);
        MBAR_WAIT(barP, gt & 1, 3);
        if (elect_one()) {
          umma::fence_after_sync();
          const uint32_t id = idesc_gen(128, 64, 0, 1);
          for (int kk = 0; kk < S16 / 16; ++kk)
            umma::mma_bf16_ss(tmem + (gt & 1) * 256, dk(FF_P + (kk >> 2), (kk & 3) * 32), dmn(FF_V + (kk >> 3), (kk & 7) * 2048), id, kk ? 1u : 0u);
          umma::mma_commit(barO);
        }
        __syncwarp();
        if (t == n_t - 1 && next < n_items) {
          MBAR_WAIT(barO, gt & 1, 4);       // last PV MMA of the item is complete: V tiles are free
          load_v(next);
        }
      }
    }
  } else {
    // ============================================================ compute warps
    const int quad = warp & 3, wg = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const float c1 = rsqrtf((float)D) * kLog2e;
    const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);
    const Philox ph(a.seed);
    const float log2_inv_keep = __log2f(a.inv_keep), keep_prob = 1.0f / a.inv_keep;
    const int NC = (S16 + 31) >> 5;                  // 32-key chunks
    const int c_lo = (wg * NC) / FF_NG, c_hi = ((wg + 1) * NC) / FF_NG;   // this thread's chunks (column group wg)
    uint32_t gt = 0, nitem = 0;
    // state of the previous tile (its epilogue runs inside the next tile's iteration)
    float prev_m = 0.f; int prev_qi = 0, prev_b = 0, prev_col = 0, prev_bn = 0; bool have_prev = false;
    auto epilogue = [&](uint32_t gtile) {
      MBAR_WAIT(barO, gtile & 1, 4);
      umma::fence_after_sync();
      const float* rl = sRedL + (gtile & 1) * (FF_NG * 128);
      float l = 0.f;
#pragma unroll
      for (int q = 0; q < FF_NG; ++q) l += rl[q * 128 + row];
      uint32_t ro[CPT];
      if (wg < 2) {
        tmem_ld_n(tlane + (gtile & 1) * 256 + (prev_col & 63), ro);
        umma::tmem_ld_wait();
      }
      umma::fence_before_sync();
      umma::mbar_arrive(barOF);
      if (wg < 2 && prev_qi < S) {
        const float inv = 1.0f / l;
        float fo[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) fo[j] = __uint_as_float(ro[j]) * inv;
        uint32_t po[CPT / 2];
        pack_n<CPT>(fo, po);
        st_global<CPT / 2>(a.ctx + ((size_t)prev_b * S + prev_qi) * H + prev_col, po);
        if (wg == 0) a.lse[(size_t)prev_bn * S + prev_qi] = prev_m * kLn2 + __logf(l);
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      const int b = item / G, g = item % G;
      float* mk = sMask + (nitem & 1) * 256;
      if (tid < 256) {
        const bool padded = tid < S && a.mask[(size_t)b * S + tid] == 0;
        mk[tid] = tid < S ? (padded ? -1e9f * kLog2e : 0.f) : -INFINITY;
        if (padded) atomicOr(sPad + (nitem & 1), 1);
      }
      // (made visible by the named barrier of the first tile's max exchange)
      bool dense = false;      // no padded key in this sequence: chunks that lie inside the sequence need no mask term
      for (int t = 0; t < n_t; ++t, ++gt) {
        const int h = t / MT, mt = t % MT;
        const int head = g * NHG + h, bn = b * a.N + head;
        const int qi = mt * 128 + row;
        const bool valid = qi < S;
        // a warp whose 32 query rows all lie past the sequence end has nothing to compute: its rows of P only feed rows of O that
        // are never stored (the MMA keeps rows independent), whatever the tile holds
        const bool warp_live = mt * 128 + quad * 32 < S;
        const uint32_t tS = tlane + (gt & 1) * 256;
        MBAR_WAIT(barS + (gt & 1), (gt >> 1) & 1, 2);
        umma::fence_after_sync();
        if (t == 0) {
          named_bar_sync(1, FF_CT);
          dense = sPad[nitem & 1] == 0;
          if (tid == 0) sPad[(nitem + 1) & 1] = 0;   // next item's flag: its writers are behind this tile's max exchange barrier
        }
        // ---- pass 1: row maximum (log2 units)
        float mloc = -INFINITY;
        if (warp_live) {
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t rs[32];
            ld_chunk(tS + c * 32, c * 32 + 32 <= S16, rs);
            umma::tmem_ld_wait();
            if (dense && c * 32 + 32 <= S) {
              float mx = __uint_as_float(rs[0]);
#pragma unroll
              for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(rs[j]));
              mloc = fmaxf(mloc, mx * c1);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, fmaf(__uint_as_float(rs[j]), c1, mk[c * 32 + j]));
            }
          }
        }
        float* rm = sRedM + (gt & 1) * (FF_NG * 128);
        rm[wg * 128 + row] = mloc;
        named_bar_sync(1, FF_CT);
        float m = rm[row];
#pragma unroll
        for (int q = 1; q < FF_NG; ++q) m = fmaxf(m, rm[q * 128 + row]);
        // ---- epilogue of the previous tile (its PV MMA ran during pass 1)
        if (have_prev) epilogue(gt - 1);
        // ---- pass 2: probabilities -> P tiles.  With dropout the kept values are p / (1 - r): the factor rides in the exponent
        // (mo = m - log2(1 / (1 - r))), the row sum is taken on the scaled values and scaled back once.
        float lsum = 0.f;
        const float mo = drop ? m - log2_inv_keep : m;
        for (int c = c_lo; c < c_hi && warp_live; ++c) {
          const int key0 = c * 32;
          uint32_t rs[32];
          ld_chunk(tS + key0, key0 + 32 <= S16, rs);
          uint32_t bits = 0xFFFFFFFFu;
          if (drop && valid && key0 < S) {
            const int kb = c >> 1, hf = c & 1;
            const uint32_t grow = (uint32_t)(bn * S + qi);
            bits = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 r4 = ph(grow, (uint32_t)(kb * 8 + hf + 2 * u), a.site, step);
              const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int bit = i * 8 + u * 2;
                bits |= ((w4[i] & 0xFFFFu) >= a.thr16 ? 1u : 0u) << bit;
                bits |= ((w4[i] >> 16) >= a.thr16 ? 1u : 0u) << (bit + 1);
              }
            }
            reinterpret_cast<uint32_t*>(a.keep)[(((size_t)bn * S + qi) * W + kb) * 2 + hf] = bits;
          }
          umma::tmem_ld_wait();
          float p[32];
          if (dense && key0 + 32 <= S) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = ex2(fmaf(__uint_as_float(rs[j]), c1, -mo));
              lsum += v;
              p[j] = ((bits >> j) & 1u) ? v : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = ex2(fmaf(__uint_as_float(rs[j]), c1, mk[key0 + j]) - mo);
              lsum += v;
              p[j] = ((bits >> j) & 1u) ? v : 0.f;
            }
          }
          uint32_t pk[16];
          pack_n<32>(p, pk);
          st_tile<4>(tile(FF_P + (key0 >> 6)), row, (key0 & 63) >> 3, pk);
        }
        if (drop) lsum *= keep_prob;
        sRedL[(gt & 1) * (FF_NG * 128) + wg * 128 + row] = lsum;
        umma::fence_before_sync();
        umma::fence_proxy_async();
        umma::mbar_arrive(barP);
        prev_m = m; prev_qi = qi; prev_b = b; prev_col = g * 64 + h * D + (wg & 1) * CPT; prev_bn = bn; have_prev = true;
      }
    }
    if (have_prev) {
      named_bar_sync(1, FF_CT);   // the other half's row sums of the last tile
      epilogue(gt - 1);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == FF_CT / 32) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
}

bool fattn_supported(const AttnArgs& a) {
  static const bool off = getenv("B4R_DISABLE_FATTN") != nullptr;
  if (off) return false;
  const int D = a.H / a.N;
  if ((D != 32 && D != 64) || a.H % 64 || a.S > 256 || a.S < 16) return false;
  if (((uintptr_t)a.qkv & 15)) return false;
  return true;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    if (const char* e = getenv("B4R_FATTN_GRID")) n = atoi(e);   // development aid: few CTAs = many items per CTA
    if (n > 0) return n;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

cudaError_t launch_fattn_bwd(const AttnArgs& a, cudaStream_t st) {
  CUtensorMap tq, tdo, to;
  if (getenv("B4R_FATTN_TRACE")) fprintf(stderr, "fattn_bwd B %d S %d H %d N %d\n", a.B, a.S, a.H, a.N);
  const uint64_t T = (uint64_t)a.B * a.S;
  if (!make_tmap_bf16_sw128(&tq, a.qkv, T, (uint64_t)3 * a.H, (uint64_t)3 * a.H, 128)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_sw128(&tdo, a.dctx, T, (uint64_t)a.H, (uint64_t)a.H, 128)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_sw128(&to, a.ctx, T, (uint64_t)a.H, (uint64_t)a.H, 128)) return cudaErrorInvalidValue;
  FAttnBwdDev d;
  d.mask = a.mask; d.lse = a.lse; d.keep = reinterpret_cast<const unsigned long long*>(a.keep_bits); d.dqkv = a.dqkv;
  d.B = a.B; d.S = a.S; d.H = a.H; d.N = a.N;
  d.thr16 = drop_threshold16(a.drop_rate);
  d.inv_keep = 1.0f / (1.0f - (float)d.thr16 / 65536.0f);
  const int D = a.H / a.N;
  const int items = a.B * (a.H / 64);
  dim3 grid(items < sm_count() ? items : sm_count());
  static const uint32_t stagger = getenv("B4R_FATTN_STAGGER_NS") ? (uint32_t)atoi(getenv("B4R_FATTN_STAGGER_NS")) : 2000u;
  static bool done32 = false, done64 = false;
  if (D == 32) {
    if (!done32) { cudaError_t e = cudaFuncSetAttribute(fattn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM); if (e != cudaSuccess) return e; done32 = true; }
    launch_pdl(fattn_bwd_kernel<32>, grid, dim3(FB_THREADS), (size_t)FB_SMEM, st, tq, tdo, to, d, stagger);
  } else {
    if (!done64) { cudaError_t e = cudaFuncSetAttribute(fattn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM); if (e != cudaSuccess) return e; done64 = true; }
    launch_pdl(fattn_bwd_kernel<64>, grid, dim3(FB_THREADS), (size_t)FB_SMEM, st, tq, tdo, to, d, stagger);
  }
  return cudaGetLastError();
}

cudaError_t launch_fattn_fwd(const AttnArgs& a, cudaStream_t st) {
  CUtensorMap tq;
  if (getenv("B4R_FATTN_TRACE")) fprintf(stderr, "fattn_fwd B %d S %d H %d N %d\n", a.B, a.S, a.H, a.N);
  const uint64_t T = (uint64_t)a.B * a.S;
  if (!make_tmap_bf16_sw128(&tq, a.qkv, T, (uint64_t)3 * a.H, (uint64_t)3 * a.H, 128)) return cudaErrorInvalidValue;
  FAttnFwdDev d;
  d.mask = a.mask; d.ctx = a.ctx; d.lse = a.lse; d.keep = reinterpret_cast<unsigned long long*>(a.keep_bits);
  d.B = a.B; d.S = a.S; d.H = a.H; d.N = a.N;
  d.thr16 = drop_threshold16(a.drop_rate);
  d.inv_keep = 1.0f / (1.0f - (float)d.thr16 / 65536.0f);
  d.seed = a.seed; d.site = a.site; d.step = a.step; d.d_step = a.d_step;
  const int D = a.H / a.N;
  const int items = a.B * (a.H / 64);
  dim3 grid(items < sm_count() ? items : sm_count());
  static bool done32 = false, done64 = false;
  if (D == 32) {
    if (!done32) { cudaError_t e = cudaFuncSetAttribute(fattn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM); if (e != cudaSuccess) return e; done32 = true; }
    launch_pdl(fattn_fwd_kernel<32>, grid, dim3(FF_THREADS), (size_t)FF_SMEM, st, tq, d);
  } else {
    if (!done64) { cudaError_t e = cudaFuncSetAttribute(fattn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM); if (e != cudaSuccess) return e; done64 = true; }
    launch_pdl(fattn_fwd_kernel<64>, grid, dim3(FF_THREADS), (size_t)FF_SMEM, st, tq, d);
  }
  return cudaGetLastError();
}

#ifdef FATTN_DEBUG
extern "C" void b4r_fattn_dump_ts() {
  unsigned long long h[128];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_fattn_ts, sizeof(h));
  unsigned long long t0 = h[0];
  for (int i = 0; i < 128; ++i) if (h[i]) printf("ts %3d %8.2f us\n", i, (double)((long long)(h[i] - t0)) / 1e3);
}
#endif

}  // namespace b4r
