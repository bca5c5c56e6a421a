// Generation 2 of the fused tied-projection x online-softmax cross-entropy forward:
// tcgen05.mma (bf16 -> fp32 accumulators in TMEM) fed by TMA (128-byte-swizzled K-major tiles), warp-specialised:
//   warp 0    : TMA producer (t tile once, then the stream of E tiles through a STAGES-deep smem ring)
//   warp 1    : TMEM allocator + single-thread MMA issuer (double-buffered 128x128 fp32 accumulators)
//   warps 2-17: epilogue, thread = (accumulator row, 32-column quarter of the tile): one tcgen05.ld of 32 columns, + bias, online
//               (max, sum-exp, label logit, first arg-max) entirely thread-local -- no shuffles, logits never leave chip.
// Same partial format / finalize kernel as generation 1 (k_ce.cu); reference ops replaced: tfm MaskedLM projection +
// SparseSoftmaxCrossEntropyWithLogits + argmax metrics (bert4rec_model.py:143, trainer_utils.py:12-23,49-60).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace b4r {

constexpr int UM_BM = 128, UM_BN = 128;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct CeUmmaDev {
  const float* vbias; const int* labels; const int* d_counts;
  int M_cap, v_begin, v_end, target_ctas, max_splits, debug;
  float* part;
};

template <int H>
struct CeUmmaCfg {
  static constexpr int KB = H / 64;                         // 64-column (128-byte) k-blocks
  static constexpr int STAGES = H == 64 ? 4 : (H == 128 ? 3 : 2);
  static constexpr int A_BYTES = KB * UM_BM * 128;
  static constexpr int B_STAGE_BYTES = KB * UM_BN * 128;
  // epilogue column parts per tile: hidden 64 runs two CTAs per SM (2 x 8 epilogue warps); one CTA per SM gets 16 warps itself
  static constexpr int NP = H == 64 ? 2 : 4;
  static constexpr int THREADS = 64 + 128 * NP;
  static constexpr int SMEM = A_BYTES + STAGES * B_STAGE_BYTES + 2 * UM_BN * 4 + 3 * 128 * 5 * 4 + 8 + 256 + 1024;  // + bias x2, merge, barriers, align slack
};

// Work decomposition is decided ON THE DEVICE from the actual row count (the grid is sized for the capacity):
//   mtiles = ceil(n_rows/128), vsplits = clamp(ceil(target_ctas / mtiles), 1, min(max_splits, ntiles)),
//   CTA id -> (m-tile = id / vsplits, split = id % vsplits).  ce_finalize uses the same formula.
__host__ __device__ inline int ce_umma_dyn_splits(int n_rows, int ntiles, int target_ctas, int max_splits) {
  static_assert(UM_BM == 128, "ce_dyn_splits128 assumes 128-row tiles");
  return ce_dyn_splits128(n_rows, ntiles, target_ctas, max_splits);
}

template <int H>
__global__ void __launch_bounds__(CeUmmaCfg<H>::THREADS, H == 64 ? 2 : 1) ce_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, CeUmmaDev a) {
  pdl_grid_sync();
  using Cfg = CeUmmaCfg<H>;
  constexpr int KB = Cfg::KB, STAGES = Cfg::STAGES, NP = Cfg::NP, CW = UM_BN / NP;   // CW = columns of a tile per epilogue thread
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + Cfg::A_BYTES;
  float* sBias = reinterpret_cast<float*>(sB + STAGES * Cfg::B_STAGE_BYTES);  // [2][UM_BN]
  float* sMerge = sBias + 2 * UM_BN;                                          // [3][128][5] results of column quarters 1..3
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMerge + 3 * 128 * 5 + 2);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* afull = tempty + 2;          // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(afull + 1);

  const int n_rows = min(a.M_cap, a.d_counts[1]);
  const int ntiles = (a.v_end - a.v_begin + UM_BN - 1) / UM_BN;
  const int vsplits = ce_umma_dyn_splits(n_rows, ntiles, a.target_ctas, a.max_splits);
  const int mtile = blockIdx.x / vsplits, split = blockIdx.x % vsplits;
  const int m0 = mtile * UM_BM;
  if (m0 >= n_rows) return;  // uniform: before any barrier / TMEM allocation
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tps = (ntiles + vsplits - 1) / vsplits;
  const int tile_lo = split * tps, tile_hi = min(ntiles, tile_lo + tps);
  const int my_tiles = max(0, tile_hi - tile_lo);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { umma::mbar_init(full + i, 1); umma::mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(tfull + i, 1); umma::mbar_init(tempty + i, 4 * NP); }
    umma::mbar_init(afull, 1);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(&tmA);
    umma::prefetch_tensormap(&tmB);
  }
  if (warp == 1) umma::tmem_alloc<256>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  // debug timestamps (B4R_CE_DEBUG & 8): CTA 0 only, first 16 tiles, into the unused tail of the partial buffer
  long long* dbg = reinterpret_cast<long long*>(a.part + (size_t)(a.max_splits - 1) * a.M_cap * 6);
  const bool dbg_on = (a.debug & 8) && blockIdx.x == 0;
#define DBG_T(role, i, slot) do { if (dbg_on && (i) < 16) dbg[((role) * 16 + (i)) * 8 + (slot)] = clock64(); } while (0)

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0 && my_tiles > 0) {
      umma::mbar_expect_tx(afull, Cfg::A_BYTES);
      for (int kb = 0; kb < KB; ++kb) umma::tma_load_2d(sA + kb * UM_BM * 128, &tmA, kb * 64, m0, afull);
      for (int i = 0; i < my_tiles; ++i) {
        const int stage = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        umma::mbar_wait(empty + stage, ph ^ 1);
        DBG_T(0, i, 0);
        umma::mbar_expect_tx(full + stage, Cfg::B_STAGE_BYTES);
        const int v0 = a.v_begin + (tile_lo + i) * UM_BN;
        for (int kb = 0; kb < KB; ++kb)
          umma::tma_load_2d(sB + stage * Cfg::B_STAGE_BYTES + kb * UM_BN * 128, &tmB, kb * 64, v0, full + stage);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc = umma::make_idesc_bf16(UM_BM, UM_BN);
      umma::mbar_wait(afull, 0);
      for (int i = 0; i < my_tiles; ++i) {
        const int stage = i % STAGES, acc = i & 1;
        umma::mbar_wait(tempty + acc, ((i >> 1) & 1) ^ 1);
        DBG_T(1, i, 0);
        umma::mbar_wait(full + stage, (i / STAGES) & 1);
        DBG_T(1, i, 1);
        umma::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * UM_BN;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a_addr = umma::smem_addr(sA + kb * UM_BM * 128);
          const uint32_t b_addr = umma::smem_addr(sB + stage * Cfg::B_STAGE_BYTES + kb * UM_BN * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (!(a.debug & 4)) umma::mma_bf16_ss(d_tmem, umma::make_desc_k_sw128(a_addr + k * 32), umma::make_desc_k_sw128(b_addr + k * 32), idesc,
                              (kb | k) ? 1u : 0u);
        }
        umma::mma_commit(empty + stage);  // smem stage reusable once these MMAs have read it
        umma::mma_commit(tfull + acc);    // accumulator ready for the epilogue
        DBG_T(1, i, 2);
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..17)
    // NP warps per TMEM lane quadrant; member `part` owns columns [part*CW, part*CW+CW) of every tile.  With one CTA per SM sixteen
    // warps (four per scheduler) hide the tcgen05.ld / MUFU latencies that eight left exposed.
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int row = m0 + row_in_tile;
    const int et = threadIdx.x - 64;           // index among the 128 * NP epilogue threads
    const int label = row < n_rows ? a.labels[row] : -1;
    float run_max = -INFINITY, run_sum = 0.f, lab_logit = -INFINITY, best_v = -INFINITY;
    int best_i = 0x7fffffff;
    constexpr float LOG2E = 1.4426950408889634f;
    // The bias of the thread's columns comes straight from global memory (L1: every row of the tile reads the same 128 values),
    // issued before the wait for the accumulator.  No shared staging and therefore no CTA-wide barrier per tile: the epilogue warps
    // drift apart instead of hitting the tcgen05.ld and the MUFU in the same cycles.
    const bool bias_vec = (reinterpret_cast<uintptr_t>(a.vbias) & 15) == 0 && (a.v_begin & 3) == 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      const int v0 = a.v_begin + (tile_lo + i) * UM_BN;
#pragma unroll 1
      for (int c = 0; c < CW / 32; ++c) {
        const int col0 = v0 + part * CW + c * 32;
        float bias[32];
        if (bias_vec && col0 + 32 <= a.v_end) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.vbias + col0 + j));
            bias[j] = b4.x; bias[j + 1] = b4.y; bias[j + 2] = b4.z; bias[j + 3] = b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) bias[j] = col0 + j < a.v_end ? __ldg(a.vbias + col0 + j) : -INFINITY;   // -inf masks the out-of-range columns
        }
        if (c == 0) {
          if (et == 0) DBG_T(2, i, 1);
          umma::mbar_wait(tfull + acc, (i >> 1) & 1);
          if (et == 0) DBG_T(2, i, 2);
          umma::fence_after_sync();
        }
        uint32_t r[32];
        if (!(a.debug & 2)) {
          umma::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * UM_BN + part * CW + c * 32, r);
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint((float)(j + c));
        }
        float v[32];
        float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          v[j] = __uint_as_float(r[j]) + bias[j]; v[j + 1] = __uint_as_float(r[j + 1]) + bias[j + 1];
          v[j + 2] = __uint_as_float(r[j + 2]) + bias[j + 2]; v[j + 3] = __uint_as_float(r[j + 3]) + bias[j + 3];
          cm[0] = fmaxf(cm[0], v[j]); cm[1] = fmaxf(cm[1], v[j + 1]); cm[2] = fmaxf(cm[2], v[j + 2]); cm[3] = fmaxf(cm[3], v[j + 3]);
        }
        const float cmax = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
        if (cmax > best_v) {  // first maximum wins (columns are visited in increasing order)
          best_v = cmax;
#pragma unroll
          for (int j = 31; j >= 0; --j)
            if (v[j] == cmax) best_i = col0 + j;
        }
        if (label >= col0 && label < col0 + 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j == label) lab_logit = v[j];
        }
        if (cmax > run_max) {
          run_sum *= ex2_approx((run_max - cmax) * LOG2E);  // ex2(-inf) = 0 on the first chunk
          run_max = cmax;
        }
        if (run_max != -INFINITY && !(a.debug & 1)) {
          const float ms = run_max * LOG2E;
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += ex2_approx(fmaf(v[j], LOG2E, -ms));
          run_sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        }
      }
      if (et == 0) DBG_T(2, i, 3);
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(tempty + acc);
      if (et == 0) DBG_T(2, i, 4);
    }
    // merge the column parts of each row (parts 1.. -> smem -> part 0, fixed order)
    if (part > 0) {
      float* d = sMerge + ((part - 1) * 128 + row_in_tile) * 5;
      d[0] = run_max; d[1] = run_sum; d[2] = lab_logit; d[3] = best_v; d[4] = __int_as_float(best_i);
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(128 * NP) : "memory");
    if (part == 0 && row < n_rows) {
      float mn = run_max, l = run_sum;
#pragma unroll
      for (int q = 0; q < NP - 1; ++q) {
        const float* o = sMerge + (q * 128 + row_in_tile) * 5;
        const float m2 = fmaxf(mn, o[0]);
        if (m2 != -INFINITY) l = l * ex2_approx((mn - m2) * LOG2E) + o[1] * ex2_approx((o[0] - m2) * LOG2E);
        mn = m2;
        const int bi1 = __float_as_int(o[4]);
        if (o[3] > best_v || (o[3] == best_v && bi1 < best_i)) { best_v = o[3]; best_i = bi1; }
        lab_logit = fmaxf(lab_logit, o[2]);
      }
      float* out = a.part + ((size_t)split * a.M_cap + row) * 6;
      out[0] = mn; out[1] = l; out[2] = lab_logit; out[3] = best_v; out[4] = __int_as_float(best_i); out[5] = 0.f;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<256>(tmem_base);
  }
}

bool ce_umma_make_maps(CeUmmaMaps* maps, const bf16* t, int M_cap, const bf16* E, int V, int H) {
  static_assert(sizeof(CUtensorMap) == sizeof(maps->a), "CUtensorMap size");
  bool ok = make_tmap_bf16_sw128(reinterpret_cast<CUtensorMap*>(maps->a), t, (uint64_t)M_cap, (uint64_t)H, (uint64_t)H, UM_BM);
  ok = ok && make_tmap_bf16_sw128(reinterpret_cast<CUtensorMap*>(maps->b), E, (uint64_t)V, (uint64_t)H, (uint64_t)H, UM_BN);
  ok = ok && make_tmap_bf16_sw128(reinterpret_cast<CUtensorMap*>(maps->a64), t, (uint64_t)M_cap, (uint64_t)H, (uint64_t)H, 64);
  ok = ok && make_tmap_bf16_sw128(reinterpret_cast<CUtensorMap*>(maps->b64), E, (uint64_t)V, (uint64_t)H, (uint64_t)H, 64);
  return ok;
}

int ce_umma_block_m() { return UM_BM; }
int ce_umma_splits(int n_rows, int ntiles, int target_ctas, int max_splits) { return ce_umma_dyn_splits(n_rows, ntiles, target_ctas, max_splits); }

cudaError_t launch_ce_fwd_umma(const CeUmmaMaps& maps, const CeArgs& a, cudaStream_t st) {
  CeUmmaDev d;
  d.vbias = a.vbias; d.labels = a.labels; d.d_counts = a.d_counts; d.M_cap = a.M_cap; d.v_begin = a.v_begin; d.v_end = a.v_end;
  d.target_ctas = a.target_ctas; d.max_splits = a.max_splits; d.part = a.part;
  { const char* e = getenv("B4R_CE_DEBUG"); d.debug = e ? atoi(e) : 0; }
  // capacity grid: enough CTAs for every (m-tile, split) pair any row count can produce
  dim3 grid(ce_dyn_grid(a.target_ctas, (a.M_cap + UM_BM - 1) / UM_BM, (a.v_end - a.v_begin + UM_BN - 1) / UM_BN, a.max_splits));
  const CUtensorMap& tmA = *reinterpret_cast<const CUtensorMap*>(maps.a);
  const CUtensorMap& tmB = *reinterpret_cast<const CUtensorMap*>(maps.b);
#define B4R_UM(HH)                                                                                                  \
  case HH: {                                                                                                        \
    static bool done_##HH = false;                                                                                  \
    if (!done_##HH) {                                                                                               \
      cudaFuncSetAttribute(ce_fwd_umma_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, CeUmmaCfg<HH>::SMEM); \
      done_##HH = true;                                                                                             \
    }                                                                                                               \
    launch_pdl(ce_fwd_umma_kernel<HH>, grid, dim3(CeUmmaCfg<HH>::THREADS), (size_t)CeUmmaCfg<HH>::SMEM, st, tmA, tmB, d);              \
    break;                                                                                                          \
  }
  switch (a.H) {
    B4R_UM(64)
    B4R_UM(128)
    B4R_UM(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_UM
  return cudaGetLastError();
}

}  // namespace b4r
