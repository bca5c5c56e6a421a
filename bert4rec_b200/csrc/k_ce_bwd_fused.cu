// Generation 3 of the tied-projection / softmax-CE BACKWARD (hidden 64): ONE tcgen05 pass produces both dT and dE.
//
// The two-pass generation (k_ce_bwd_umma.cu) recomputes every [128 x 128] logits tile twice (once per operand role) and is
// bound by the per-tile epilogue (exp, one-hot, bf16 pack: ~2 us per tile per SM), not by the tensor pipe.  Here a work item
// is (vocabulary range of VR = 2 tiles) x (row chunk of <= MR_MAX = 5 tiles), all of whose operand tiles are resident in
// shared memory (TMA, 7 x 16 KB).  For every pair (row tile i, vocabulary tile j):
//
//   MMA1 :  S[128 m x 128 v]  = T_i . E_j^T                       (double-buffered in TMEM, overlaps the previous epilogue)
//   epilogue (two groups of 8 warps on alternate pairs; one TMEM lane = one row m, 64 columns per thread):
//           dl = (exp2(S*log2e + bias_v*log2e - lse_m*log2e) - [v == label_m]) * [w_m > 0]  -> bf16 tile in shared memory
//   MMA2a:  dT_i[128 m x 64] += dl     . E_j     (A = dl K-major,  B = E_j read MN-major)
//   MMA2b:  dE_j[128 v x 80] += dl^T   . [T_i|1] (A = the SAME dl tile read MN-major, B = T_i read MN-major with a second
//           64-column block pointed at a tile of ONES: accumulator column 64 = sum_m dl[m, v] = the output-bias gradient)
//
// dT_i is drained after the last j of a row tile (partial slot = vocabulary range), the dE_j after the item (partial slot =
// row chunk); head_bwd_fused sums the dT partials, ce_bwd_fused_reduce_kernel the dE / bias partials.  Deterministic.
// Gradient of the SUM loss, like the other generations (bert4rec_model.py:166-167).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace b4r {
namespace {

constexpr int FT = 128, FH = 64, MR_MAX = CF_MR_MAX, VR = CF_VR;
constexpr int TILE = FT * 128;      // [128 rows][64 bf16], 128-byte swizzled rows
constexpr int DL_BYTES = 2 * TILE;  // [128 m][128 v] as two [128][64] sub-tiles
constexpr int E_STAGE = VR * TILE;  // the vocabulary tiles of one item; two stages (the next item is prefetched)
constexpr int OFF_T = 0, OFF_E = MR_MAX * TILE, OFF_DL = OFF_E + 2 * E_STAGE, OFF_ONES = OFF_DL + 2 * DL_BYTES, ONES_BYTES = 2048;
constexpr int OFF_VEC = OFF_ONES + ONES_BYTES;   // bias [2 items][VR*128] | lse [MR_MAX*128] | label [MR_MAX*128]
constexpr int VEC_WORDS = 2 * VR * FT + 2 * MR_MAX * FT;
constexpr int OFF_BAR = OFF_VEC + VEC_WORDS * 4;
constexpr int SMEM = OFF_BAR + 256 + 1024;
static_assert(SMEM <= 232448, "shared memory");
constexpr int NTHR = 128 + 512;   // warp 0: TMA, warps 1-3: MMA issuers, warps 4-19: epilogue
constexpr uint32_t C_S = 0, C_DT = 256, C_DE = 320, DE_COLS = 80;   // TMEM columns: S0, S1 | dT | dE_0, dE_1 (64 + 16 ones columns)

// Work of one CTA: super-items (row chunk c, vocabulary group q) strided by the grid; inside, the group's vocabulary ranges in
// order (an "item").  All three warp roles walk the same sequence.
struct Walk {
  int G, nch, b, per, nvr, si, jr, jr_hi;
  bool valid, first;
  __device__ void start() {
    valid = si < nch * b;
    if (valid) { jr = (si % b) * per; jr_hi = min(nvr, jr + per); first = true; }
  }
  __device__ void init(const CfSplit& sp, int bid, int grid) {
    G = grid; nch = sp.nch; b = sp.b; per = sp.per; nvr = sp.nvr; si = bid;
    start();
  }
  __device__ void next() {
    ++jr; first = false;
    if (jr >= jr_hi) { si += G; start(); }
  }
  __device__ int c() const { return si / b; }
  __device__ int q() const { return si % b; }
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// dT slot update: the first vocabulary range of a super-item stores, the later ones ADD with a fire-and-forget reduction (no
// read round trip in the drain).  Only this thread ever touches the address and its operations apply in program order, so the
// summation order is fixed (deterministic).
__device__ __forceinline__ void put4(float* dst, bool first, float x, float y, float z, float w) {
  if (first) *reinterpret_cast<float4*>(dst) = make_float4(x, y, z, w);
  else asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(dst), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// packed fp32 pair arithmetic (sm_100 f32x2): {o0, o1} = {x0, x1} * l2 + ({b0, b1} + nra2)
__device__ __forceinline__ void fma2_bias(float& o0, float& o1, float x0, float x1, uint64_t l2, float b0, float b1, uint64_t nra2) {
  asm("{\n"
      ".reg .b64 x, b, t;\n"
      "mov.b64 x, {%2, %3};\n"
      "mov.b64 b, {%5, %6};\n"
      "add.rn.f32x2 t, b, %7;\n"
      "fma.rn.f32x2 t, x, %4, t;\n"
      "mov.b64 {%0, %1}, t;\n"
      "}\n"
      : "=f"(o0), "=f"(o1)
      : "f"(x0), "f"(x1), "l"(l2), "f"(b0), "f"(b1), "l"(nra2));
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
#define CF_STAMP(role, pair, k)                                                                    \
  do {                                                                                             \
    if (a.dbg && blockIdx.x == 0 && (pair) < 32u && lane == 0) a.dbg[(role) * 128 + (pair) * 4 + (k)] = gtime(); \
  } while (0)
// MN-major operand: [k rows of 128 B (64 mn elements)], 16 k-rows per MMA, LBO = distance to the next 64-element mn block
__device__ __forceinline__ uint64_t desc_mn(uint32_t byte_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((byte_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct Dev {
  const float* vbias; const float* lse; const float* row_w; const int* labels; const int* d_counts;
  int M_cap, V;
  float* dt_part;   // [vocabulary groups][M_cap][64]
  float* dE_part;   // [nch][V][64]
  float* db_part;   // [nch][V]
  unsigned long long* dbg;   // development aid (B4R_CF_DEBUG): %globaltimer stamps of CTA 0, [4 roles][32 pairs][4]
};

__global__ void __launch_bounds__(NTHR, 1) ce_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmE,
                                                               Dev a) {
  pdl_grid_sync();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sBias = reinterpret_cast<float*>(smem + OFF_VEC);
  float* sLse = sBias + 2 * VR * FT;   // lse * log2e of the row, +inf for rows without a gradient (their dl is exp2(-inf) = 0)
  int* sLab = reinterpret_cast<int*>(sLse + MR_MAX * FT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* e_full = bars;             // [2] TMA landed the item's vocabulary tiles (+ the row tiles for the first item of a super-item)
  uint64_t* e_free = bars + 2;         // [2] every MMA of the item has completed
  uint64_t* s_full = bars + 4;         // [2] MMA1 done
  uint64_t* s_empty = bars + 6;        // [2] epilogue read S
  uint64_t* dl_full = bars + 8;        // [2] epilogue wrote dl
  uint64_t* dl_empty = bars + 10;      // [2] MMA2a/b consumed dl
  uint64_t* acc_done = bars + 12;      // all MMAs of the item complete
  uint64_t* drained = bars + 13;       // epilogue drained the item's dE accumulators
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);

  const int n_valid = min(a.M_cap, a.d_counts[0]);
  const CfSplit sp = cf_split(n_valid, a.V, (int)gridDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(e_full + b, 1); umma::mbar_init(e_free + b, 3);
      umma::mbar_init(s_full + b, 1); umma::mbar_init(s_empty + b, 8);
      umma::mbar_init(dl_full + b, 8); umma::mbar_init(dl_empty + b, 2);
    }
    umma::mbar_init(acc_done, 2); umma::mbar_init(drained, 16);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(&tmT);
    umma::prefetch_tensormap(&tmE);
  }
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHR) reinterpret_cast<uint32_t*>(smem + OFF_ONES)[i] = 0x3F803F80u;   // bf16 1.0
  umma::fence_proxy_async();
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      Walk w;
      w.init(sp, blockIdx.x, gridDim.x);
      for (int it = 0; w.valid; w.next(), ++it) {
        const int st = it & 1;
        const int i0 = w.c() * sp.MR, mrc = min(sp.MR, sp.mt - i0), j0 = w.jr * VR, vrc = min(VR, sp.vt - j0);
        if (it >= 2) umma::mbar_wait(e_free + st, ((it - 2) >> 1) & 1);                 // stage free: MMAs of item it-2 done
        if (w.first && it > 0) umma::mbar_wait(e_free + ((it - 1) & 1), ((it - 1) >> 1) & 1);   // row tiles free: previous super-item done
        umma::mbar_expect_tx(e_full + st, (uint32_t)(vrc + (w.first ? mrc : 0)) * TILE);
        for (int j = 0; j < vrc; ++j) umma::tma_load_2d(smem + OFF_E + st * E_STAGE + j * TILE, &tmE, 0, (j0 + j) * FT, e_full + st);
        if (w.first)
          for (int i = 0; i < mrc; ++i) umma::tma_load_2d(smem + OFF_T + i * TILE, &tmT, 0, (i0 + i) * FT, e_full + st);
      }
    }
  } else if (warp <= 3) {
    // ===================================================================== MMA issuers: three single threads, one per accumulator
    // chain -- warp 1: S = T_i . E_j^T, warp 2: dT += dl . E_j, warp 3: dE += dl^T . [T_i | 1].  A single thread issuing all 20
    // MMAs of a pair was the bottleneck of the kernel (ncu: tensor pipe 18 % active, epilogue warps waiting on barriers).
    if (lane == 0) {
      constexpr uint32_t id1 = idesc(FT, FT, 0, 0);          // S   = T_i (K-major) . E_j (K-major)
      constexpr uint32_t id2a = idesc(FT, FH, 0, 1);         // dT += dl (K-major) . E_j (MN-major)
      constexpr uint32_t id2b = idesc(FT, DE_COLS, 1, 1);    // dE += dl^T (MN-major) . [T_i | ones] (MN-major)
      const uint32_t t_base = umma::smem_addr(smem + OFF_T), e_base0 = umma::smem_addr(smem + OFF_E);
      const uint32_t dl_base = umma::smem_addr(smem + OFF_DL), ones = umma::smem_addr(smem + OFF_ONES);
      uint32_t g = 0;
      Walk w;
      w.init(sp, blockIdx.x, gridDim.x);
      for (int it = 0; w.valid; w.next(), ++it) {
        const int st = it & 1;
        const int i0 = w.c() * sp.MR, mrc = min(sp.MR, sp.mt - i0), j0 = w.jr * VR, vrc = min(VR, sp.vt - j0);
        const int np = mrc * vrc;
        const uint32_t e_base = e_base0 + st * E_STAGE;
        umma::mbar_wait(e_full + st, (it >> 1) & 1);
        if (warp != 1 && it > 0) umma::mbar_wait(drained, (it - 1) & 1);   // all 16 epilogue warps have read the previous item's dT / dE
        umma::fence_after_sync();
        if (warp == 1) {
          // S[buf] is refilled as soon as the epilogue of pair p-2 has READ it (s_empty, early in that epilogue)
          for (int p = 0; p < np; ++p) {
            const uint32_t gp = g + p, buf = gp & 1;
            const int i = p / vrc, j = p - i * vrc;
            umma::mbar_wait(s_empty + buf, ((gp >> 1) & 1) ^ 1);
            umma::fence_after_sync();
            CF_STAMP(3, gp, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma::mma_bf16_ss(tmem + C_S + buf * FT, umma::make_desc_k_sw128(t_base + i * TILE + k * 32),
                                umma::make_desc_k_sw128(e_base + j * TILE + k * 32), id1, k ? 1u : 0u);
            umma::mma_commit(s_full + buf);
            CF_STAMP(3, gp, 1);
          }
        } else {
          for (int p = 0; p < np; ++p) {
            const uint32_t gp = g + p, buf = gp & 1;
            const int i = p / vrc, j = p - i * vrc;
            umma::mbar_wait(dl_full + buf, (gp >> 1) & 1);   // (the epilogue drains dT of row tile i-1 before it arrives here for j = 0)
            umma::fence_after_sync();
            const uint32_t dl = dl_base + buf * DL_BYTES, tt = t_base + i * TILE, ee = e_base + j * TILE;
            if (warp == 2) CF_STAMP(2, gp, 0);
            if (warp == 2) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma::mma_bf16_ss(tmem + C_DT, umma::make_desc_k_sw128(dl + (kk >> 2) * TILE + (kk & 3) * 32), desc_mn(ee + kk * 2048, TILE),
                                  id2a, (j | kk) ? 1u : 0u);
            } else {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma::mma_bf16_ss(tmem + C_DE + j * DE_COLS, desc_mn(dl + kk * 2048, TILE),
                                  desc_mn(tt + kk * 2048, ones - (tt + kk * 2048)), id2b, (i | kk) ? 1u : 0u);
            }
            if (warp == 2) CF_STAMP(2, gp, 1);
            umma::mma_commit(dl_empty + buf);   // two arrivals: the dT and the dE chain
            if (warp == 2) CF_STAMP(2, gp, 2);
          }
        }
        g += np;
        umma::mma_commit(e_free + st);          // three arrivals
        if (warp != 1) umma::mma_commit(acc_done);   // two arrivals
      }
    }
  } else {
    // ===================================================================== epilogue (warps 4..19)
    // Two groups of 8 warps take alternate pairs (group = S / dl buffer index), so that one group's TMEM-load, barrier and
    // fence latencies overlap the other's arithmetic; inside a group a thread owns one row and 64 columns (`half`).
    // The item tail (accumulator drains) uses all 16 warps: quadrant x column quarter `cq`.
    const int quad = warp & 3, cq = (warp - 4) >> 2;   // TMEM lane quadrant, column quarter
    const int grp = (warp - 4) >> 3, half = cq & 1;
    const int row_in_tile = quad * 32 + lane;
    const int e = threadIdx.x - 128;                    // 0..511
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    uint32_t g = 0;
    Walk w;
    w.init(sp, blockIdx.x, gridDim.x);
    for (int it = 0; w.valid; w.next(), ++it) {
      const int c = w.c(), slot = w.q();
      const bool first = w.first;   // first vocabulary range of the super-item: the dT slot is written, afterwards added to
      const int i0 = c * sp.MR, mrc = min(sp.MR, sp.mt - i0), j0 = w.jr * VR, vrc = min(VR, sp.vt - j0);
      const int np = mrc * vrc;
      auto drain_dt = [&](int i) {
        uint32_t r[16];
        umma::tmem_ld16(tmem + C_DT + lane_off + cq * 16, r);
        umma::tmem_ld_wait();
        const int m = (i0 + i) * FT + row_in_tile;
        if (m < n_valid) {
          float* dst = a.dt_part + ((size_t)slot * a.M_cap + m) * FH + cq * 16;
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            put4(dst + q, first, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        }
      };
      // bias of this item: written into buffer it & 1 by the previous item's tail (prefetched a whole item ahead), directly for
      // the first item; row vectors once per super-item
      if (it == 0)
        for (int x = e; x < vrc * FT; x += 512) {
          const int v = j0 * FT + x;
          sBias[x] = v < a.V ? a.vbias[v] * LOG2E : -INFINITY;
        }
      asm volatile("bar.sync 1, 512;\n" ::: "memory");   // bias visible; the previous super-item's row vectors are no longer read
      if (first) {
        for (int x = e; x < mrc * FT; x += 512) {
          const int m = i0 * FT + x;
          const bool on = m < n_valid && a.row_w[m] > 0.f;
          sLse[x] = on ? a.lse[m] * LOG2E : INFINITY;
          sLab[x] = on ? a.labels[m] : -1;
        }
        asm volatile("bar.sync 1, 512;\n" ::: "memory");
      }
      float next_bias = 0.f;   // the next item's bias value of column e (threads 0..255), in flight during this item's pairs
      {
        Walk wn = w;
        wn.next();
        if (wn.valid && e < VR * FT) {
          const int v = wn.jr * VR * FT + e;
          next_bias = v < a.V ? a.vbias[v] * LOG2E : -INFINITY;
        }
      }
      const float* bias_it = sBias + (it & 1) * VR * FT;
      if (warp == 4) CF_STAMP(3, (uint32_t)it + 16, 3);
      for (int p = 0; p < np; ++p) {
        const uint32_t gp = g + p, buf = gp & 1;
        if ((int)buf != grp) continue;
        const int i = p / vrc, j = p - i * vrc;
        const int lrow = i * FT + row_in_tile;
        const uint64_t nra2 = pack2(-sLse[lrow], -sLse[lrow]), l2 = pack2(LOG2E, LOG2E);
        const int rel_label = sLab[lrow] - ((j0 + j) * FT + half * 64);   // label column relative to this thread's 64 columns
        const float* vec = bias_it + j * FT + half * 64;
        unsigned char* rowp = smem + OFF_DL + buf * DL_BYTES + half * TILE + row_in_tile * 128;
        umma::mbar_wait(s_full + buf, (gp >> 1) & 1);
        umma::fence_after_sync();
        const bool stamp = ((warp - 4) & 7) == 0;
        if (stamp) CF_STAMP(grp, gp, 0);
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t r[32];
          umma::tmem_ld32(tmem + C_S + buf * FT + lane_off + half * 64 + c2 * 32, r);
          umma::tmem_ld_wait();
          if (c2 == 1) {
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(s_empty + buf);   // S buffer free: MMA1 of pair gp+2 may overwrite it
          }
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float2 b2 = *reinterpret_cast<const float2*>(vec + c2 * 32 + q);
            float x0, x1;
            fma2_bias(x0, x1, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), l2, b2.x, b2.y, nra2);
            pk[q >> 1] = pack_bf162(ex2_approx(x0), ex2_approx(x1));
          }
          if (c2 == 0) {
            if (stamp) CF_STAMP(grp, gp, 1);
            umma::mbar_wait(dl_empty + buf, ((gp >> 1) & 1) ^ 1);   // MMA2 of pair gp-2 has consumed this dl buffer
            if (stamp) CF_STAMP(grp, gp, 2);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c2 * 4 + q) ^ (row_in_tile & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        if (rel_label >= 0 && rel_label < 64) {
          // the one-hot term: the label column lives in this thread's row segment (written just above): p -> p - 1 in place
          const int chunk = (rel_label >> 3) ^ (row_in_tile & 7);
          __nv_bfloat16* el = reinterpret_cast<__nv_bfloat16*>(rowp + chunk * 16) + (rel_label & 7);
          *el = __float2bfloat16_rn(__bfloat162float(*el) - 1.f);
        }
        if (j == 0 && i > 0) {
          // dT of the previous row tile is final once the MMAs of pair gp-1 (the other group's) have completed; drain it before
          // MMA2a of this pair (accumulate = 0) may overwrite it -- that MMA waits for this group's dl_full arrivals below
          umma::mbar_wait(dl_empty + ((gp - 1) & 1), ((gp - 1) >> 1) & 1);
          umma::fence_after_sync();
          uint32_t r[32];
          umma::tmem_ld32(tmem + C_DT + lane_off + half * 32, r);
          umma::tmem_ld_wait();
          const int m = (i0 + i - 1) * FT + row_in_tile;
          if (m < n_valid) {
            float* dst = a.dt_part + ((size_t)slot * a.M_cap + m) * FH + half * 32;
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              put4(dst + q, first, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
          }
        }
        umma::fence_before_sync();
        umma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(dl_full + buf);
        if (stamp) CF_STAMP(grp, gp, 3);
      }
      g += np;
      // ---- item tail: last dT tile, the dE accumulators and the bias column
      if (warp == 4) CF_STAMP(3, (uint32_t)it, 2);
      umma::mbar_wait(acc_done, it & 1);
      umma::fence_after_sync();
      if (warp == 4) CF_STAMP(3, (uint32_t)it, 3);
      drain_dt(mrc - 1);
      if (warp == 4) CF_STAMP(2, (uint32_t)it, 3);
      for (int j = 0; j < vrc; ++j) {
        const int v = (j0 + j) * FT + row_in_tile;
        uint32_t r[16];
        umma::tmem_ld16(tmem + C_DE + j * DE_COLS + lane_off + cq * 16, r);
        umma::tmem_ld_wait();
        if (v < a.V) {
          float* dst = a.dE_part + ((size_t)c * a.V + v) * FH + cq * 16;
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            *reinterpret_cast<float4*>(dst + q) =
                make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        }
        if (cq == 0) {
          umma::tmem_ld16(tmem + C_DE + j * DE_COLS + lane_off + FH, r);
          umma::tmem_ld_wait();
          if (v < a.V) a.db_part[(size_t)c * a.V + v] = __uint_as_float(r[0]);
        }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(drained);
      if (e < VR * FT) sBias[((it + 1) & 1) * VR * FT + e] = next_bias;   // (buffer (it+1)&1 was last read in item it-1)
      if (warp == 4) CF_STAMP(3, (uint32_t)it + 16, 2);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<512>(tmem);
  }
}

// table / output-bias gradient = sum of the row-chunk partials (chunk count from the device-side row count)
__global__ void __launch_bounds__(256) ce_bwd_fused_reduce_kernel(const float* __restrict__ dE_part, const float* __restrict__ db_part,
                                                                  const int* __restrict__ d_counts, int M_cap, int V,
                                                                  float* __restrict__ g_table, float* __restrict__ g_bias) {
  pdl_grid_sync();
  const int nch = cf_split(min(M_cap, d_counts[0]), V, 1).nch;
  const long long n4 = (long long)V * FH / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < nch; k0 += 8) {      // eight chunk partials in flight per thread (the chunk count is a run-time value)
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + u < nch) v[u] = __ldcs(reinterpret_cast<const float4*>(dE_part + (size_t)(k0 + u) * V * FH) + idx);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    reinterpret_cast<float4*>(g_table)[idx] = acc;
  }
  if (idx < V) {
    float s = 0.f;
    for (int k = 0; k < nch; ++k) s += db_part[(size_t)k * V + idx];
    g_bias[idx] = s;
  }
}

}  // namespace

bool ce_bwd_fused_supported(int H) { return H == FH && getenv("B4R_DISABLE_CE_FUSED") == nullptr; }
int ce_bwd_fused_max_chunks(int M_cap) { return cf_split(M_cap, 1, 1).nch; }

cudaError_t launch_ce_bwd_fused(const CeUmmaMaps& maps, const CeBwdFusedArgs& a, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(ce_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    done = true;
  }
  Dev d;
  d.vbias = a.vbias; d.lse = a.lse; d.row_w = a.row_w; d.labels = a.labels; d.d_counts = a.d_counts; d.M_cap = a.M_cap; d.V = a.V;
  d.dt_part = a.dt_part; d.dE_part = a.dE_part; d.db_part = a.db_part; d.dbg = a.dbg;
  const CUtensorMap& tmT = *reinterpret_cast<const CUtensorMap*>(maps.a);
  const CUtensorMap& tmE = *reinterpret_cast<const CUtensorMap*>(maps.b);
  return launch_pdl(ce_bwd_fused_kernel, dim3(a.ctas > 0 ? a.ctas : 148), dim3(NTHR), (size_t)SMEM, st, tmT, tmE, d);
}

cudaError_t launch_ce_bwd_fused_reduce(const CeBwdFusedArgs& a, float* g_table, float* g_bias, cudaStream_t st) {
  const long long n4 = (long long)a.V * FH / 4;
  return launch_pdl(ce_bwd_fused_reduce_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), (size_t)0, st, a.dE_part, a.db_part, a.d_counts,
                    a.M_cap, a.V, g_table, g_bias);
}

}  // namespace b4r
