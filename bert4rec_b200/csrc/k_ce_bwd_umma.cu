// Generation 2 of the tied-projection / softmax-CE BACKWARD: two tcgen05 passes that recompute the logits tile in
// TMEM and never write the [M, V] gradient-of-logits matrix to memory (generation 1 materialises it in bf16 and
// re-reads it twice).  Both passes are the same kernel with the operand roles swapped:
//
//   R  = resident operand tile (128 rows x H, loaded once by TMA),  X = streamed operand tiles (128 rows x H each)
//   MMA1:  S[128 R-rows x 128 X-rows] = R . X^T              (A = R K-major, B = X K-major, fp32 in TMEM)
//   epilogue (one thread per R row / TMEM lane): dl = (exp(S + bias_v - lse_m) - [v == label_m]) * w_m  -> bf16 tile in
//            shared memory, written row-contiguous in the 128-byte-swizzled K-major layout
//   MMA2:  ACC[128 R-rows x H] += dl . X                     (A = dl K-major, B = the SAME X tile read MN-major)
//
//   ROW_IS_M = true : R = t rows (masked slots), X = E tiles  -> ACC = dT rows   (split over the vocabulary)
//   ROW_IS_M = false: R = E rows (vocabulary),   X = t tiles  -> ACC = dE rows, thread-local row sums = d(output bias)
//
// Gradient of the SUM loss (the 1/n_valid normaliser is folded into the optimizer).  Replaces the backward of
// tfm MaskedLM's projection + SparseSoftmaxCrossEntropyWithLogits (bert4rec_model.py:166-167).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace b4r {

constexpr int CB_T = 128;  // rows of the resident tile R (and of X for hidden <= 128; hidden 256 streams 64-row X tiles)

__device__ __forceinline__ float ex2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major operand tile stored as KB sub-tiles of [128 k-rows][64 mn-elements] (128-byte rows, SWIZZLE_128B):
// start = first k-row of the 16-row slice, LBO = distance between 64-element mn blocks, SBO = 8 k-rows = 1024 B.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t byte_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((byte_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16: D fp32, A bf16 K-major, B bf16 MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct CeBwdDev {
  const float* vbias; const float* lse; const float* row_w; const int* labels; const int* d_counts;
  int M_cap, V;
  int target_ctas, max_splits;   // ROW_IS_M: dynamic vocabulary splits (same formula as the forward)
  int msplits;                   // !ROW_IS_M: static number of row splits
  float* out;                    // ROW_IS_M: dT partials [split][M_cap][H] ; else dE partials [msplit][V][H]
  float* dbias_out;              // !ROW_IS_M: [msplit][V]
};

template <int H>
struct CeBwdCfg {
  static constexpr int KB = H / 64;
  static constexpr int XT = H <= 128 ? 128 : 64;           // rows of a streamed X tile (= columns of the S / dl tile)
  // Hidden 256 runs ONE CTA per SM (TMEM: 320 of 512 columns, 200 KB of shared memory), so nothing overlaps a tile's epilogue
  // with the next tile's first MMA unless the CTA does it itself: the S accumulator and the dl tile are double-buffered (SB = 2)
  // and MMA1 of tile i+1 is issued BEFORE the issuer waits for the dl tile of tile i.  Hidden <= 128 runs two CTAs per SM that
  // overlap each other; they keep single buffers (256 TMEM columns each).
  static constexpr int SB = H == 256 ? 2 : 1;
  static constexpr int XSTAGES = H == 128 ? 2 : 3;
  static constexpr int R_BYTES = KB * CB_T * 128;
  static constexpr int X_BYTES = KB * XT * 128;
  static constexpr int DL_BYTES = (XT / 64) * CB_T * 128;  // [128][XT] bf16 as [128][64] sub-tiles
  static constexpr int VEC_BYTES = 2 * 3 * CB_T * 4;       // double-buffered per-column vectors (3 x 128 x 4 B)
  static constexpr int SMEM = R_BYTES + XSTAGES * X_BYTES + SB * DL_BYTES + VEC_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = SB * XT + H <= 256 ? 256 : 512;   // S (SB x XT) + ACC (H)
  static constexpr int CTAS_PER_SM = TMEM_COLS == 256 ? 2 : 1;
};

__host__ __device__ inline int ce_bwd_dyn_splits(int n_rows, int ntiles, int target_ctas, int max_splits) {
  static_assert(CB_T == 128, "ce_dyn_splits128 assumes 128-row tiles");
  return ce_dyn_splits128(n_rows, ntiles, target_ctas, max_splits);
}

template <int H, bool ROW_IS_M>
__global__ void __launch_bounds__(320, CeBwdCfg<H>::CTAS_PER_SM) ce_bwd_umma_kernel(const __grid_constant__ CUtensorMap tmT,
                                                             const __grid_constant__ CUtensorMap tmE, CeBwdDev a) {
  pdl_grid_wait();
  using Cfg = CeBwdCfg<H>;
  constexpr int KB = Cfg::KB, XS = Cfg::XSTAGES, XT = Cfg::XT, SB = Cfg::SB;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sR = smem;
  unsigned char* sX = sR + Cfg::R_BYTES;
  unsigned char* sDl = sX + XS * Cfg::X_BYTES;
  float* sVec = reinterpret_cast<float*>(sDl + SB * Cfg::DL_BYTES);  // [2][3][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sVec + 2 * 3 * CB_T);
  uint64_t* xfull = bars;             // [XS]
  uint64_t* xempty = bars + XS;       // [XS]
  uint64_t* rfull = bars + 2 * XS;    // R tile landed
  uint64_t* s_full = rfull + 1;       // [2] MMA1 done -> epilogue may read S (per S buffer)
  uint64_t* s_empty = s_full + 2;     // [2] epilogue done reading S
  uint64_t* dl_full = s_empty + 2;    // [2] epilogue wrote the dl tile (per dl buffer)
  uint64_t* dl_empty = dl_full + 2;   // [2] MMA2 done reading the dl tile
  uint64_t* acc_full = dl_empty + 2;  // all MMA2 done
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

  // rows with a gradient = the valid masked slots; the zero-weight aux rows behind them (metric parity only) are skipped
  const int n_rows = min(a.M_cap, a.d_counts[0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- work assignment
  int r0, x_lo, x_hi, split;  // R tile start row, streamed tile range [x_lo, x_hi), partial slot
  if (ROW_IS_M) {
    const int ntiles = (a.V + XT - 1) / XT;
    const int vs = ce_bwd_dyn_splits(n_rows, ntiles, a.target_ctas, a.max_splits);
    const int mtile = blockIdx.x / vs;
    split = blockIdx.x % vs;
    r0 = mtile * CB_T;
    if (r0 >= n_rows) return;
    const int tps = (ntiles + vs - 1) / vs;
    x_lo = split * tps; x_hi = min(ntiles, x_lo + tps);
    if (x_lo >= x_hi) {  // trailing split without tiles: its partial is summed by the consumer, so define it
      for (int i = threadIdx.x; i < CB_T * H; i += blockDim.x) {
        const int r = r0 + i / H;
        if (r < n_rows) a.out[((size_t)split * a.M_cap + r) * H + (i % H)] = 0.f;
      }
      return;
    }
  } else {
    const int vtile = blockIdx.x / a.msplits;
    split = blockIdx.x % a.msplits;
    r0 = vtile * CB_T;
    const int mt = (n_rows + XT - 1) / XT;
    const int per = (mt + a.msplits - 1) / a.msplits;
    x_lo = split * per; x_hi = min(mt, x_lo + per);
    if (x_lo >= x_hi) {  // empty row range: this partial slot must still be defined
      for (int i = threadIdx.x; i < CB_T * H; i += blockDim.x) {
        const int r = r0 + i / H;
        if (r < a.V) a.out[((size_t)split * a.V + r) * H + (i % H)] = 0.f;
      }
      for (int i = threadIdx.x; i < CB_T; i += blockDim.x)
        if (r0 + i < a.V) a.dbias_out[(size_t)split * a.V + r0 + i] = 0.f;
      return;
    }
  }
  const int my_tiles = max(0, x_hi - x_lo);
  const CUtensorMap* mapR = ROW_IS_M ? &tmT : &tmE;
  const CUtensorMap* mapX = ROW_IS_M ? &tmE : &tmT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < XS; ++i) { umma::mbar_init(xfull + i, 1); umma::mbar_init(xempty + i, 1); }
    umma::mbar_init(rfull, 1);
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(s_full + i, 1);
      umma::mbar_init(s_empty + i, 8);
      umma::mbar_init(dl_full + i, 8);
      umma::mbar_init(dl_empty + i, 1);
    }
    umma::mbar_init(acc_full, 1);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(mapR);
    umma::prefetch_tensormap(mapX);
  }
  if (warp == 1) umma::tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_s = tmem_base, tmem_acc = tmem_base + SB * XT;
  // buffer / phase of tile i: one S and dl buffer -> (0, i & 1); two -> (i & 1, (i >> 1) & 1)
  auto bsel = [](int i) { return SB == 2 ? (i & 1) : 0; };
  auto bpar = [](int i) { return (uint32_t)(SB == 2 ? ((i >> 1) & 1) : (i & 1)); };

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0 && my_tiles > 0) {
      umma::mbar_expect_tx(rfull, Cfg::R_BYTES);
      for (int kb = 0; kb < KB; ++kb) umma::tma_load_2d(sR + kb * CB_T * 128, mapR, kb * 64, r0, rfull);
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % XS;
        umma::mbar_wait(xempty + st, ((i / XS) & 1) ^ 1);
        umma::mbar_expect_tx(xfull + st, Cfg::X_BYTES);
        const int xr0 = (x_lo + i) * XT;
        for (int kb = 0; kb < KB; ++kb) umma::tma_load_2d(sX + st * Cfg::X_BYTES + kb * XT * 128, mapX, kb * 64, xr0, xfull + st);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc1 = umma::make_idesc_bf16(CB_T, XT);
      constexpr uint32_t idesc2 = make_idesc_bf16_bmn(CB_T, H);
      umma::mbar_wait(rfull, 0);
      auto issue_mma1 = [&](int i) {      // S[buffer of i] = R . X_i^T
        const int st = i % XS;
        const uint32_t x_addr = umma::smem_addr(sX + st * Cfg::X_BYTES);
        umma::mbar_wait(xfull + st, (i / XS) & 1);
        umma::mbar_wait(s_empty + bsel(i), bpar(i) ^ 1);     // the epilogue that last used this S buffer has read it
        umma::fence_after_sync();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t r_addr = umma::smem_addr(sR + kb * CB_T * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16_ss(tmem_s + bsel(i) * XT, umma::make_desc_k_sw128(r_addr + k * 32),
                              umma::make_desc_k_sw128(x_addr + kb * XT * 128 + k * 32), idesc1, (kb | k) ? 1u : 0u);
        }
        umma::mma_commit(s_full + bsel(i));
      };
      if (SB == 2) issue_mma1(0);
      for (int i = 0; i < my_tiles; ++i) {
        const int st = i % XS;
        const uint32_t x_addr = umma::smem_addr(sX + st * Cfg::X_BYTES);
        if (SB == 2) { if (i + 1 < my_tiles) issue_mma1(i + 1); }   // runs on the tensor pipe while the epilogue works on tile i
        else issue_mma1(i);
        // second MMA: ACC += dl (K-major, K = the 128 streamed rows) . X (MN-major: N = H feature columns)
        umma::mbar_wait(dl_full + bsel(i), bpar(i));
        umma::fence_after_sync();
        const uint32_t dl_addr = umma::smem_addr(sDl + bsel(i) * Cfg::DL_BYTES);
#pragma unroll
        for (int kk = 0; kk < XT / 16; ++kk)
          umma::mma_bf16_ss(tmem_acc, umma::make_desc_k_sw128(dl_addr + (kk >> 2) * CB_T * 128 + (kk & 3) * 32),
                            make_desc_mn_sw128(x_addr + kk * 16 * 128, XT * 128), idesc2, (i | kk) ? 1u : 0u);
        umma::mma_commit(xempty + st);
        umma::mma_commit(dl_empty + bsel(i));
      }
      umma::mma_commit(acc_full);
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int row = r0 + row_in_tile;
    const int et = threadIdx.x - 64;  // 0..255
    constexpr float LOG2E = 1.4426950408889634f;
    // per-row scalars
    float row_a = 0.f;   // ROW_IS_M: lse*log2e ; else bias*log2e
    bool row_on = false; // row contributes
    int row_label = -1;  // ROW_IS_M only
    if (ROW_IS_M) {
      if (row < n_rows) { row_a = a.lse[row] * LOG2E; row_on = a.row_w[row] > 0.f; row_label = a.labels[row]; }
    } else {
      if (row < a.V) { row_a = a.vbias[row] * LOG2E; row_on = true; }
    }
    float rowsum = 0.f;
    // per-column vectors of the streamed tile, double-buffered in smem, prefetched one tile ahead into registers
    //   ROW_IS_M: v0[c] = bias[v]*log2e (-inf beyond V)
    //   else    : v0[c] = lse[m]*log2e, v1[c] = w[m] (0 beyond n_rows), v2[c] = label[m] (as int bits)
    float pre0 = 0.f, pre1 = 0.f, pre2 = 0.f;
    auto fetch = [&](int tile_idx) {
      const int c = (x_lo + tile_idx) * XT + et;
      if (ROW_IS_M) {
        pre0 = c < a.V ? a.vbias[c] * LOG2E : -INFINITY;
      } else {
        const bool ok = c < n_rows;
        pre0 = ok ? a.lse[c] * LOG2E : 0.f;
        pre1 = ok ? a.row_w[c] : 0.f;
        pre2 = __int_as_float(ok ? a.labels[c] : -1);
      }
    };
    auto stash = [&](int buf) {
      float* d = sVec + buf * 3 * CB_T;
      d[et] = pre0;
      if (!ROW_IS_M) { d[CB_T + et] = pre1; d[2 * CB_T + et] = pre2; }
    };
    if (et < XT && my_tiles > 0) { fetch(0); stash(0); }
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      if (et < XT && i + 1 < my_tiles) fetch(i + 1);
      asm volatile("bar.sync 1, 256;\n" ::: "memory");  // vectors of tile i visible; buffer buf^1 free
      umma::mbar_wait(s_full + bsel(i), bpar(i));
      umma::fence_after_sync();
      umma::mbar_wait(dl_empty + bsel(i), bpar(i) ^ 1);   // the MMA2 that last read this dl buffer is done
      const float* vec = sVec + buf * 3 * CB_T + half * (XT / 2);
      const int x0 = (x_lo + i) * XT + half * (XT / 2);   // first streamed row (= S column) of this thread's half
#pragma unroll 1
      for (int c = 0; c < XT / 64; ++c) {
        uint32_t r[32];
        umma::tmem_ld32(tmem_s + ((uint32_t)(quad * 32) << 16) + bsel(i) * XT + half * (XT / 2) + c * 32, r);
        umma::tmem_ld_wait();
        uint32_t pk[16];
        if (ROW_IS_M) {
          const int col0 = x0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 b2 = *reinterpret_cast<const float2*>(vec + c * 32 + j);
            float p0 = ex2f_approx(fmaf(__uint_as_float(r[j]), LOG2E, b2.x) - row_a);
            float p1 = ex2f_approx(fmaf(__uint_as_float(r[j + 1]), LOG2E, b2.y) - row_a);
            if (col0 + j == row_label) p0 -= 1.f;
            if (col0 + j + 1 == row_label) p1 -= 1.f;
            pk[j >> 1] = row_on ? pack_bf162(p0, p1) : 0u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 l2 = *reinterpret_cast<const float2*>(vec + c * 32 + j);
            const float2 w2 = *reinterpret_cast<const float2*>(vec + CB_T + c * 32 + j);
            const float2 lb = *reinterpret_cast<const float2*>(vec + 2 * CB_T + c * 32 + j);
            float p0 = ex2f_approx(fmaf(__uint_as_float(r[j]), LOG2E, row_a) - l2.x);
            float p1 = ex2f_approx(fmaf(__uint_as_float(r[j + 1]), LOG2E, row_a) - l2.y);
            if (__float_as_int(lb.x) == row) p0 -= 1.f;
            if (__float_as_int(lb.y) == row) p1 -= 1.f;
            p0 = (w2.x > 0.f && row_on) ? p0 : 0.f;
            p1 = (w2.y > 0.f && row_on) ? p1 : 0.f;
            const uint32_t u = pack_bf162(p0, p1);
            pk[j >> 1] = u;
            const float2 q = unpack_bf162(u);   // the bias gradient sums what the MMA will see
            rowsum += q.x + q.y;
          }
        }
        // row-contiguous store into the 128B-swizzled K-major sub-tile `half`: 16-byte chunk q -> q ^ (row & 7)
        const int dcol = half * (XT / 2) + c * 32;   // first dl column of this chunk: sub-tile dcol / 64, 16-byte chunk (dcol % 64) / 8
        unsigned char* rowp = sDl + bsel(i) * Cfg::DL_BYTES + (dcol >> 6) * (CB_T * 128) + row_in_tile * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (((dcol & 63) >> 3) + q) ^ (row_in_tile & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      umma::fence_before_sync();
      umma::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) { umma::mbar_arrive(s_empty + bsel(i)); umma::mbar_arrive(dl_full + bsel(i)); }
      if (et < XT && i + 1 < my_tiles) stash(buf ^ 1);
    }
    // ---- accumulator -> global partial
    umma::mbar_wait(acc_full, 0);
    umma::fence_after_sync();
    const bool store_row = ROW_IS_M ? (row < n_rows) : (row < a.V);
    const size_t rows_total = ROW_IS_M ? (size_t)a.M_cap : (size_t)a.V;
    constexpr int HC = H / 2;  // columns handled by this half
    float* dst = a.out + ((size_t)split * rows_total + row) * H + half * HC;
#pragma unroll 1
    for (int c = 0; c < HC / 32; ++c) {
      uint32_t r[32];
      umma::tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + half * HC + c * 32, r);
      umma::tmem_ld_wait();
      if (store_row) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c * 32 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
    if (!ROW_IS_M) {
      // combine the two column halves' row sums through smem (sVec is free now)
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      if (half == 1) sVec[row_in_tile] = rowsum;
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      if (half == 0 && row < a.V) a.dbias_out[(size_t)split * a.V + row] = rowsum + sVec[row_in_tile];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

bool ce_bwd_umma_supported(int H) { return H == 64 || H == 128 || H == 256; }
int ce_bwd_umma_xtile(int H) { return H <= 128 ? 128 : 64; }
int ce_bwd_umma_dt_splits(int n_rows, int V, int target_ctas, int max_splits) {
  return ce_bwd_dyn_splits(n_rows, (V + CB_T - 1) / CB_T, target_ctas, max_splits);   // (callers pass the X-tile count themselves)
}

template <int H, bool ROW_IS_M>
static cudaError_t launch_ce_bwd_t(const CUtensorMap& tmT, const CUtensorMap& tmE, const CeBwdDev& d, int grid, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(ce_bwd_umma_kernel<H, ROW_IS_M>, cudaFuncAttributeMaxDynamicSharedMemorySize, CeBwdCfg<H>::SMEM);
    done = true;
  }
  launch_pdl(ce_bwd_umma_kernel<H, ROW_IS_M>, dim3(grid), dim3(320), (size_t)(CeBwdCfg<H>::SMEM), st, tmT, tmE, d);
  return cudaGetLastError();
}

cudaError_t launch_ce_bwd_umma(const CeUmmaMaps& maps, const CeBwdArgs& a, bool row_is_m, cudaStream_t st) {
  CeBwdDev d;
  d.vbias = a.vbias; d.lse = a.lse; d.row_w = a.row_w; d.labels = a.labels; d.d_counts = a.d_counts;
  d.M_cap = a.M_cap; d.V = a.V; d.target_ctas = a.target_ctas; d.max_splits = a.max_splits; d.msplits = a.msplits;
  d.out = a.out; d.dbias_out = a.dbias_out;
  // resident operand: 128-row boxes; streamed operand: XT-row boxes (64 for hidden 256)
  const bool x64 = ce_bwd_umma_xtile(a.H) == 64;
  const CUtensorMap& tmT = *reinterpret_cast<const CUtensorMap*>(row_is_m || !x64 ? maps.a : maps.a64);
  const CUtensorMap& tmE = *reinterpret_cast<const CUtensorMap*>(!row_is_m || !x64 ? maps.b : maps.b64);
  const int mtiles_cap = (a.M_cap + CB_T - 1) / CB_T, vtiles = (a.V + CB_T - 1) / CB_T;
  const int xtiles = (a.V + ce_bwd_umma_xtile(a.H) - 1) / ce_bwd_umma_xtile(a.H);
  const int grid = row_is_m ? ce_dyn_grid(a.target_ctas, mtiles_cap, xtiles, a.max_splits) : vtiles * a.msplits;
  if (a.H == 64) return row_is_m ? launch_ce_bwd_t<64, true>(tmT, tmE, d, grid, st) : launch_ce_bwd_t<64, false>(tmT, tmE, d, grid, st);
  if (a.H == 128) return row_is_m ? launch_ce_bwd_t<128, true>(tmT, tmE, d, grid, st) : launch_ce_bwd_t<128, false>(tmT, tmE, d, grid, st);
  if (a.H == 256) return row_is_m ? launch_ce_bwd_t<256, true>(tmT, tmE, d, grid, st) : launch_ce_bwd_t<256, false>(tmT, tmE, d, grid, st);
  return cudaErrorInvalidValue;
}

}  // namespace b4r
