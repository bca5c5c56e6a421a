// Generation 2 of the generic GEMM (k_gemm.cu gemm_kernel): C[M,N] = A[M,K] . B  with the same fused epilogues, on
// tcgen05.mma + TMEM + TMA.  Persistent, warp-specialised:
//   warp 0    : TMA producer -- A tile [128 rows][64 k] and B tile through a 4-stage shared-memory ring (SWIZZLE_128B)
//   warp 1    : MMA issuer (one elected lane): 4 x (M128 N128 K16) per k-block into one of TWO 128-column fp32
//               accumulators in TMEM, so the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-17: epilogue, thread = (accumulator row, 32-column part): one tcgen05.ld of 32 columns -> bias / GELU /
//               GELU' (+ per-tile column sums = bias gradient) / fp32 residual -> global.  bf16 outputs leave through a
//               swizzled shared-memory staging tile and TMA stores (cp.async.bulk.tensor): one thread writing the 64 bytes
//               of its own row made every store instruction touch 32 half-used sectors -- the stores, not the MMAs or the
//               math, were 55-70 % of these kernels (FFN1 + GELU at C4: 371 us with, 112 us without its stores)
// B is consumed in place in either layout: [K,N] row-major (forward: TF kernels are [in,out]) as an MN-major operand,
// [N,K] row-major (data gradients read the SAME weights transposed) as a K-major operand -- no transposed copies.
// Used for hidden sizes whose GEMMs are real GEMMs (K, N multiples of 64; e.g. C4: M = 204 800 tokens, K/N in
// {256, 768, 1024}); reference ops: Keras EinsumDense projections of tfm TransformerEncoderBlock
// (bert4rec_encoder.py:136-147) and their gradients.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "enc_fused.cuh"

namespace b4r {
using namespace encf;

namespace {
constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 64, TG_STAGES = 4;   // (4 ring stages allocated for the barrier arrays; see tg_stages)
constexpr int TG_EPI_WARPS = 16, TG_THREADS = 64 + 32 * TG_EPI_WARPS;   // producer + issuer + 16 epilogue warps
constexpr int TG_A_BYTES = TG_BM * 128, TG_B_BYTES = TG_BN * 128, TG_STAGE = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_OUT_BYTES = TG_BM * TG_BN * 2;            // bf16 staging of one output tile: two [128][64] swizzled halves
__host__ __device__ constexpr int tg_nout(int epi) { return epi == EPI_BIAS_GELU ? 2 : (epi == EPI_F32_RES ? 0 : 1); }   // bf16 output tiles staged
__host__ __device__ constexpr int tg_nf32(int epi) { return epi == EPI_F32_RES ? 2 : 0; }   // fp32 residual -> output tile, in place: 4 x [128][32] fp32 = 64 KB
// GELU' reads its second operand (the saved pre-activation tile) through TMA as well: two 32 KB landing buffers, paid for with one ring stage
__host__ __device__ constexpr int tg_naux(int epi) { return epi == EPI_GELU_GRAD ? 2 : 0; }
__host__ __device__ constexpr int tg_stages(int epi) { return epi == EPI_GELU_GRAD ? 3 : 4; }
__host__ __device__ constexpr int tg_smem(int epi) { return tg_stages(epi) * TG_STAGE + (tg_nout(epi) + tg_naux(epi) + tg_nf32(epi)) * TG_OUT_BYTES + 2 * 4 * TG_BN * 4 + 256 + 1024; }   // ring + staging + column sums + barriers + align

struct TGemmDev {
  int M, N, K;
  const float* bias;
  bf16* out_bf16; int ld_out;
  bf16* out2_bf16;
  const bf16* aux_bf16; int ld_aux;
  float* out_f32; int ld_f32;
  const float* res_f32;
  float* colsum_part;
};
}  // namespace

template <int EPI, bool B_MN>
__global__ void __launch_bounds__(TG_THREADS, 1) tgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                                                       TGemmDev a) {
  pdl_grid_wait_single_wave();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NST = tg_stages(EPI);
  unsigned char* sOut = smem + NST * TG_STAGE;                            // [NOUT][2 halves][128][64] bf16, 128-byte swizzle
  unsigned char* sAux = sOut + tg_nout(EPI) * TG_OUT_BYTES;               // [2 acc][2 halves][128][64] bf16 (GELU' only)
  unsigned char* sF32 = sAux + tg_naux(EPI) * TG_OUT_BYTES;               // [4 column blocks][128][32] fp32 (F32_RES only)
  float* sCol = reinterpret_cast<float*>(sF32 + tg_nf32(EPI) * TG_OUT_BYTES);   // [2 acc][4 quads][128] column-sum partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + 2 * 4 * TG_BN);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + TG_STAGES;          // [STAGES]
  uint64_t* tfull = bars + 2 * TG_STAGES;      // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint64_t* auxfull = tempty + 2;              // [2]
  uint64_t* resfull = auxfull + 2;             // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(resfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (a.M + TG_BM - 1) / TG_BM, tiles_n = (a.N + TG_BN - 1) / TG_BN;
  const int n_tiles = tiles_m * tiles_n, kblocks = a.K / TG_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TG_STAGES; ++i) { umma::mbar_init(full + i, 1); umma::mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(tfull + i, 1); umma::mbar_init(tempty + i, TG_EPI_WARPS); umma::mbar_init(auxfull + i, 1); }
    umma::mbar_init(resfull, 1);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(&tmA);
    umma::prefetch_tensormap(&tmB);
    if (tg_nout(EPI) > 0 || tg_nf32(EPI) > 0) umma::prefetch_tensormap(&tmO);
    if (tg_nout(EPI) > 1 || tg_naux(EPI) > 0 || tg_nf32(EPI) > 0) umma::prefetch_tensormap(&tmO2);
  }
  if (warp == 1) umma::tmem_alloc<256>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * TG_BM, n0 = (t % tiles_n) * TG_BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int st = it % NST;
          umma::mbar_wait(empty + st, ((it / NST) & 1) ^ 1);
          umma::mbar_expect_tx(full + st, TG_STAGE);
          unsigned char* sA = smem + st * TG_STAGE;
          unsigned char* sB = sA + TG_A_BYTES;
          umma::tma_load_2d(sA, &tmA, kb * TG_BK, m0, full + st);
          if (B_MN) {   // B [K,N]: two [64 k-rows][64 n] blocks
            umma::tma_load_2d(sB, &tmB, n0, kb * TG_BK, full + st);
            umma::tma_load_2d(sB + 8192, &tmB, n0 + 64, kb * TG_BK, full + st);
          } else {      // B [N,K]: [128 n-rows][64 k]
            umma::tma_load_2d(sB, &tmB, kb * TG_BK, n0, full + st);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
      const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), 8192);
      constexpr uint32_t idesc = idesc_gen(TG_BM, TG_BN, 0, B_MN ? 1 : 0);
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
        const int acc = ti & 1;
        umma::mbar_wait(tempty + acc, ((ti >> 1) & 1) ^ 1);
        umma::fence_after_sync();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int st = it % NST;
          umma::mbar_wait(full + st, (it / NST) & 1);
          umma::fence_after_sync();
          const uint32_t offA = st * TG_STAGE, offB = offA + TG_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_at(DK0, offA + k * 32);
            const uint64_t db = B_MN ? desc_at(DMN0, offB + k * 2048) : desc_at(DK0, offB + k * 32);
            umma::mma_bf16_ss(tmem + acc * TG_BN, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma::mma_commit(empty + st);
        }
        umma::mma_commit(tfull + acc);
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 2..17): (lane quadrant, 32-column part)
    const int quad = warp & 3, cpart = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const bool store_leader = threadIdx.x == 64;     // first epilogue thread: issues and tracks the bulk stores
    int ti = 0;
    // staging protocol per tile: [leader: earlier bulk stores have read the staging tiles] -> bar -> st_tile -> proxy fence -> bar ->
    // [leader: TMA stores + commit]
    auto stage_begin = [&]() {
      if (store_leader) umma::tma_store_wait_read<0>();
      asm volatile("bar.sync 2, 512;\n" ::: "memory");
    };
    auto stage_put = [&](int which, const uint32_t (&pk)[16]) {
      st_tile<4>(sOut + which * TG_OUT_BYTES + (cpart >> 1) * TILE_B, row_in_tile, (cpart & 1) * 4, pk);
    };
    auto stage_commit = [&](int m0, int n0) {
      umma::fence_proxy_async();
      asm volatile("bar.sync 2, 512;\n" ::: "memory");
      if (store_leader) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          if (n0 + hh * 64 < a.N) {
            umma::tma_store_2d(&tmO, sOut + hh * TILE_B, n0 + hh * 64, m0);
            if (tg_nout(EPI) > 1) umma::tma_store_2d(&tmO2, sOut + TG_OUT_BYTES + hh * TILE_B, n0 + hh * 64, m0);
          }
        umma::tma_store_commit();
      }
    };
    auto aux_load = [&](int tt, int tti) {       // the pre-activation tile of output tile tt -> landing buffer tti & 1
      if (EPI == EPI_GELU_GRAD && store_leader && tt < n_tiles) {
        const int am0 = (tt / tiles_n) * TG_BM, an0 = (tt % tiles_n) * TG_BN;
        unsigned char* dst = sAux + (tti & 1) * TG_OUT_BYTES;
        umma::mbar_expect_tx(auxfull + (tti & 1), TG_OUT_BYTES);
        umma::tma_load_2d(dst, &tmO2, an0, am0, auxfull + (tti & 1));
        umma::tma_load_2d(dst + TILE_B, &tmO2, an0 + 64, am0, auxfull + (tti & 1));
      }
    };
    auto res_load = [&](int tt) {                // fp32 residual tile of output tile tt -> the in-place buffer (tmO2 = residual map)
      if (EPI == EPI_F32_RES && store_leader && tt < n_tiles) {
        const int am0 = (tt / tiles_n) * TG_BM, an0 = (tt % tiles_n) * TG_BN;
        umma::mbar_expect_tx(resfull, 2 * TG_OUT_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::tma_load_2d(sF32 + j * TILE_B, &tmO2, an0 + j * 32, am0, resfull);
      }
    };
    aux_load(blockIdx.x, 0);
    res_load(blockIdx.x);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const int acc = ti & 1;
      const int mt = t / tiles_n, m0 = mt * TG_BM, n0 = (t % tiles_n) * TG_BN;
      aux_load(t + gridDim.x, ti + 1);           // (its buffer was last read two tiles ago: every thread has passed that tile's staging barrier)
      const int m = m0 + row_in_tile;
      const bool mok = m < a.M;
      umma::mbar_wait(tfull + acc, (ti >> 1) & 1);
      umma::fence_after_sync();
      {
        const int n = n0 + cpart * 32;                   // first of this thread's 32 columns
        float v[32];
        tmem_ld_f32(tmem + ((uint32_t)(quad * 32) << 16) + acc * TG_BN + cpart * 32, v);
        const bool nok = n < a.N;                         // N is a multiple of 32 on this path
        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU) {
          if (nok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n + i));
              v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
          }
        }
        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BF16) {
          uint32_t pk[16];
          pack_n<32>(v, pk);
          stage_begin();
          stage_put(0, pk);
          stage_commit(m0, n0);
        } else if (EPI == EPI_BIAS_GELU) {
          uint32_t pk[16], pk2[16];
          round_n<32>(v, pk);   // GELU of the bf16-rounded pre-activation backward re-reads
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          pack_n<32>(v, pk2);
          stage_begin();
          stage_put(0, pk);
          stage_put(1, pk2);
          stage_commit(m0, n0);
        } else if (EPI == EPI_GELU_GRAD) {
          uint32_t pk[16];
          umma::mbar_wait(auxfull + acc, (ti >> 1) & 1);
          if (mok && nok) {
            float x[32];
            ld_tile<4>(sAux + acc * TG_OUT_BYTES + (cpart >> 1) * TILE_B, row_in_tile, (cpart & 1) * 4, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(x[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          round_n<32>(v, pk);   // the bias gradient sums what the weight-gradient GEMM will see
          stage_begin();
          stage_put(0, pk);
          stage_commit(m0, n0);
          if (a.colsum_part) {
            // column sums over the 32 rows of this warp (recursive halving: 31 shuffles for 32 columns), then over the 4
            // lane quadrants through smem
            float a16[16], a8[8], a4[4], a2[2];
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float keep = h16 ? v[i + 16] : v[i], send = h16 ? v[i] : v[i + 16];
              a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float keep = h8 ? a16[i + 8] : a16[i], send = h8 ? a16[i] : a16[i + 8];
              a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float keep = h4 ? a8[i + 4] : a8[i], send = h4 ? a8[i] : a8[i + 4];
              a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float keep = h2 ? a4[i + 2] : a4[i], send = h2 ? a4[i] : a4[i + 2];
              a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            const float keep = h1 ? a2[1] : a2[0], send = h1 ? a2[0] : a2[1];
            const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            const int col = (h16 ? 16 : 0) + (h8 ? 8 : 0) + (h4 ? 4 : 0) + (h2 ? 2 : 0) + (h1 ? 1 : 0);   // == lane
            sCol[(acc * 4 + quad) * TG_BN + cpart * 32 + col] = tot;
          }
        } else if (EPI == EPI_F32_RES) {
          // out = acc + residual, fp32: the residual tile came in by TMA; add in place, TMA store, then fetch the next tile's residual
          umma::mbar_wait(resfull, ti & 1);
          unsigned char* bt = sF32 + cpart * TILE_B;       // this thread's 32 columns = one 128-byte row of column block cpart
          {
            const unsigned char* rp = bt + row_in_tile * 128;
            uint32_t o[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + ((q ^ (row_in_tile & 7)) << 4));
              o[4 * q] = __float_as_uint(v[4 * q] + r4.x); o[4 * q + 1] = __float_as_uint(v[4 * q + 1] + r4.y);
              o[4 * q + 2] = __float_as_uint(v[4 * q + 2] + r4.z); o[4 * q + 3] = __float_as_uint(v[4 * q + 3] + r4.w);
            }
            st_tile<8>(bt, row_in_tile, 0, o);
          }
          umma::fence_proxy_async();
          asm volatile("bar.sync 2, 512;\n" ::: "memory");
          if (store_leader) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n0 + j * 32 < a.N) umma::tma_store_2d(&tmO, sF32 + j * TILE_B, n0 + j * 32, m0);
            umma::tma_store_commit();
            umma::tma_store_wait_read<0>();
            res_load(t + gridDim.x);
          }
        }
      }
      umma::fence_before_sync();
      if (EPI == EPI_GELU_GRAD && a.colsum_part) {
        asm volatile("bar.sync 1, 512;\n" ::: "memory");   // the 16 epilogue warps: column partials of this tile complete
        const int et = threadIdx.x - 64;
        if (et < TG_BN && n0 + et < a.N) {
          const float* sc = sCol + acc * 4 * TG_BN + et;
          a.colsum_part[(size_t)mt * a.N + n0 + et] = (sc[0] + sc[TG_BN]) + (sc[2 * TG_BN] + sc[3 * TG_BN]);
        }
      }
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(tempty + acc);
    }
    if ((tg_nout(EPI) > 0 || tg_nf32(EPI) > 0) && store_leader) umma::tma_store_wait<0>();   // every bulk store has completed before the CTA retires
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<256>(tmem);
  }
}

bool tgemm_supported(int epi, const GemmArgs& a) {
  if (getenv("B4R_DISABLE_TGEMM")) return false;
  if (!(epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU || epi == EPI_GELU_GRAD || epi == EPI_BF16 || epi == EPI_F32_RES)) return false;
  if (a.a_trans || a.a_rows || a.d_M || a.splits > 1) return false;
  // one 64-wide k-block is enough for the TMA ring (hidden-64 FFN1: 36 -> 30 us at C1); B4R_TGEMM_MINK raises the threshold
  static const int min_k = getenv("B4R_TGEMM_MINK") ? atoi(getenv("B4R_TGEMM_MINK")) : 64;
  if (a.K % TG_BK || a.N % 64 || a.K < min_k || a.M < 256) return false;   // small / narrow problems stay on the portable kernel
  if ((a.a_kmax && a.a_kmax != a.K) || (a.b_kmax && a.b_kmax != a.K)) return false;
  if (a.lda % 8 || a.ldb % 8 || ((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15)) return false;
  if (tg_nf32(epi) > 0 && (a.ld_f32 % 4 || ((uintptr_t)a.out_f32 & 15) || ((uintptr_t)a.res_f32 & 15) || a.N % 32)) return false;
  if (tg_naux(epi) > 0 && (a.ld_aux % 8 || ((uintptr_t)a.aux_bf16 & 15))) return false;
  if (tg_nout(epi) > 0 && (a.ld_out % 8 || ((uintptr_t)a.out_bf16 & 15) || (tg_nout(epi) > 1 && ((uintptr_t)a.out2_bf16 & 15)))) return false;   // TMA stores
  return true;
}

template <int EPI, bool B_MN>
static cudaError_t launch_tgemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                                  const TGemmDev& d, int grid, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(tgemm_kernel<EPI, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem(EPI));
    if (e != cudaSuccess) return e;
    done = true;
  }
  launch_pdl(tgemm_kernel<EPI, B_MN>, dim3(grid), dim3(TG_THREADS), (size_t)(tg_smem(EPI)), st, tmA, tmB, tmO, tmO2, d);
  return cudaGetLastError();
}

cudaError_t launch_tgemm(int epi, const GemmArgs& a, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  if (!make_tmap_bf16_sw128(&tmA, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, TG_BM)) return cudaErrorInvalidValue;
  const bool bmn = a.b_trans;   // b_trans: B memory is [K][N]
  if (bmn ? !make_tmap_bf16_sw128(&tmB, a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, 64)
          : !make_tmap_bf16_sw128(&tmB, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, TG_BN))
    return cudaErrorInvalidValue;
  CUtensorMap tmO = tmA, tmO2 = tmA;   // (placeholders when the epilogue has no bf16 output)
  if (tg_nout(epi) > 0 && !make_tmap_bf16_sw128(&tmO, a.out_bf16, (uint64_t)a.M, (uint64_t)a.N, (uint64_t)a.ld_out, TG_BM)) return cudaErrorInvalidValue;
  if (tg_nout(epi) > 1 && !make_tmap_bf16_sw128(&tmO2, a.out2_bf16, (uint64_t)a.M, (uint64_t)a.N, (uint64_t)a.ld_out, TG_BM)) return cudaErrorInvalidValue;
  if (tg_nf32(epi) > 0 && (!make_tmap_f32_sw128(&tmO, a.out_f32, (uint64_t)a.M, (uint64_t)a.N, (uint64_t)a.ld_f32, TG_BM) ||
                           !make_tmap_f32_sw128(&tmO2, a.res_f32, (uint64_t)a.M, (uint64_t)a.N, (uint64_t)a.ld_f32, TG_BM)))
    return cudaErrorInvalidValue;
  if (tg_naux(epi) > 0 && !make_tmap_bf16_sw128(&tmO2, a.aux_bf16, (uint64_t)a.M, (uint64_t)a.N, (uint64_t)a.ld_aux, TG_BM)) return cudaErrorInvalidValue;
  TGemmDev d;
  d.M = a.M; d.N = a.N; d.K = a.K; d.bias = a.bias; d.out_bf16 = a.out_bf16; d.ld_out = a.ld_out; d.out2_bf16 = a.out2_bf16;
  d.aux_bf16 = a.aux_bf16; d.ld_aux = a.ld_aux; d.out_f32 = a.out_f32; d.ld_f32 = a.ld_f32; d.res_f32 = a.res_f32;
  d.colsum_part = a.colsum_part;
  const int tiles = ((a.M + TG_BM - 1) / TG_BM) * ((a.N + TG_BN - 1) / TG_BN);
  const int grid = tiles < 148 ? tiles : 148;
#define B4R_TG(E)                                                                                              \
  case E:                                                                                                      \
    return bmn ? launch_tgemm_t<E, true>(tmA, tmB, tmO, tmO2, d, grid, st) : launch_tgemm_t<E, false>(tmA, tmB, tmO, tmO2, d, grid, st);
  switch (epi) {
    B4R_TG(EPI_BIAS_BF16)
    B4R_TG(EPI_BIAS_GELU)
    B4R_TG(EPI_GELU_GRAD)
    B4R_TG(EPI_BF16)
    B4R_TG(EPI_F32_RES)
  }
#undef B4R_TG
  return cudaErrorInvalidValue;
}

}  // namespace b4r

// ======================================================================================= weight gradient (generation 2)
// dW[M,N] = X^T dY, contraction over the T token rows.  Both operands are read in place as MN-major tiles ([64 token rows][64
// features] TMA boxes); one CTA owns a 256 x 256 output tile (two M = 128 accumulators of 256 TMEM columns each) and a contiguous
// range of token rows (split over T, deterministic fp32 partials summed by grad_reduce_kernel).  256 x 256 tiles read X N/256 times
// and dY M/256 times: HBM-bound by design.  M and N need not fill the tile (hidden 64 / 128 models: 64 x 192, 64 x 256, 256 x 64 ...):
// feature boxes that lie completely outside the matrix are not loaded, the second accumulator is skipped when M <= 128 and the
// MMA's N shrinks to the valid columns; a partly valid box is zero-filled by TMA.  Rows / columns past M / N are never stored.
namespace b4r {
using namespace encf;
namespace {
constexpr int TW_STAGES = 3, TW_STAGE = 2 * 64 * 256 * 2;   // 32 KB of X + 32 KB of dY per 64-token k-block
constexpr int TW_SMEM = TW_STAGES * TW_STAGE + 256 + 1024;
struct TWgradDev { int M, N, T, t_per_split, tiles_n; float* out; size_t split_stride; int ld_out; float* colsum; };
}  // namespace

__global__ void __launch_bounds__(320, 1) twgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                        TWgradDev a) {
  pdl_grid_wait_single_wave();
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TW_STAGES * TW_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + TW_STAGES;
  uint64_t* tfull = bars + 2 * TW_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, z = blockIdx.y;
  const int m0 = (tile / a.tiles_n) * 256, n0 = (tile % a.tiles_n) * 256;
  const int t_begin = z * a.t_per_split, t_end = min(a.T, t_begin + a.t_per_split);
  const int kblocks = t_end > t_begin ? (t_end - t_begin + 63) / 64 : 0;
  const int m_ext = min(256, a.M - m0), n_ext = min(256, a.N - n0);      // valid part of this CTA's 256 x 256 tile
  const int nbA = (m_ext + 63) / 64, nbB = (n_ext + 63) / 64;            // 64-feature boxes to load per k-block
  const int nh = m_ext > 128 ? 2 : 1;                                    // M = 128 accumulators in use
  // column sums of dY (= the bias gradient of the layer that produced it) ride along: the eight epilogue warps are idle during
  // the main loop and read every dY stage from shared memory before it is released (one more arrival per warp on `empty`)
  const bool do_colsum = a.colsum != nullptr && m0 == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TW_STAGES; ++i) { umma::mbar_init(full + i, 1); umma::mbar_init(empty + i, do_colsum ? 9 : 1); }
    umma::mbar_init(tfull, 1);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(&tmX);
    umma::prefetch_tensormap(&tmY);
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % TW_STAGES;
        umma::mbar_wait(empty + st, ((kb / TW_STAGES) & 1) ^ 1);
        umma::mbar_expect_tx(full + st, (uint32_t)((nbA + nbB) * 8192));
        unsigned char* sA = smem + st * TW_STAGE;
        unsigned char* sB = sA + TW_STAGE / 2;
        const int t0 = t_begin + kb * 64;
        for (int j = 0; j < nbA; ++j) umma::tma_load_2d(sA + j * 8192, &tmX, m0 + j * 64, t0, full + st);
        for (int j = 0; j < nbB; ++j) umma::tma_load_2d(sB + j * 8192, &tmY, n0 + j * 64, t0, full + st);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), 8192);
      // (rows of an accumulator that come from boxes that were not loaded hold whatever the ring held: the MMA keeps rows
      // independent and those rows are not stored)
      const uint32_t idesc = idesc_gen(128, (n_ext + 15) & ~15, 1, 1);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % TW_STAGES;
        umma::mbar_wait(full + st, (kb / TW_STAGES) & 1);
        umma::fence_after_sync();
        const uint32_t offA = st * TW_STAGE, offB = offA + TW_STAGE / 2;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          for (int h = 0; h < nh; ++h)
            umma::mma_bf16_ss(tmem + h * 256, desc_at(DMN0, offA + h * 16384 + k * 2048), desc_at(DMN0, offB + k * 2048), idesc,
                              (kb | k) ? 1u : 0u);
        umma::mma_commit(empty + st);
      }
      umma::mma_commit(tfull);
    }
    __syncwarp();
  } else {
    // epilogue: warps 2..9 -> (lane quadrant, m-half); 256 fp32 columns per thread row
    const int quad = warp & 3, h = (warp - 2) >> 2;
    const int m = m0 + h * 128 + quad * 32 + lane;
    float* dst = a.out + (size_t)z * a.split_stride + (size_t)m * a.ld_out + n0;
    if (do_colsum) {
      // thread = (feature pair p of the 256 columns, half of the 64 token rows of a stage); tiles are [64 rows][64 features] boxes,
      // 128-byte rows, 16-byte chunks XOR-swizzled with the row
      const int et = threadIdx.x - 64, p = et & 127, rh = et >> 7;
      const uint32_t box = (uint32_t)(p >> 5) * 8192u, chunk = (uint32_t)(p & 31) >> 2, inner = (uint32_t)(p & 3) * 4u;
      float c0 = 0.f, c1 = 0.f;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % TW_STAGES;
        umma::mbar_wait(full + st, (kb / TW_STAGES) & 1);
        const uint32_t sB = umma::smem_addr(smem + st * TW_STAGE + TW_STAGE / 2) + box + inner;
#pragma unroll 8
        for (int r = rh * 32; r < rh * 32 + 32; ++r) {
          uint32_t w;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(sB + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
          const float2 f = unpack_bf162(w);
          c0 += f.x; c1 += f.y;
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(empty + st);
      }
      // the two row halves meet in the (now idle) first stage buffer; fixed order: rows 0-31 of every stage first
      if (kblocks > 0) umma::mbar_wait(tfull, 0);       // every MMA has read its operands: the ring is free
      float2* sx = reinterpret_cast<float2*>(smem);
      if (rh == 1) sx[p] = make_float2(c0, c1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (rh == 0) {
        const float2 o = sx[p];
        if (2 * p < n_ext)
          *reinterpret_cast<float2*>(a.colsum + (size_t)z * a.N + n0 + 2 * p) = make_float2(kblocks > 0 ? c0 + o.x : 0.f, kblocks > 0 ? c1 + o.y : 0.f);
      }
    }
    if (kblocks > 0) {
      umma::mbar_wait(tfull, 0);
      umma::fence_after_sync();
    }
#pragma unroll 1
    for (int c = 0; c < 8 && c * 32 < n_ext && h < nh; ++c) {
      float v[32];
      if (kblocks > 0) tmem_ld_f32(tmem + ((uint32_t)(quad * 32) << 16) + h * 256 + c * 32, v);   // (warp-uniform: every lane takes part)
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (m < a.M) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (c * 32 + i < n_ext) *reinterpret_cast<float4*>(dst + c * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
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

bool twgrad_shape_ok(int M, int N, int T) { return !getenv("B4R_DISABLE_TGEMM") && M % 8 == 0 && N % 8 == 0 && M >= 64 && N >= 64 && T >= 4096; }
int twgrad_splits(int M, int N, int T) {
  const int tiles = ((M + 255) / 256) * ((N + 255) / 256);
  int s = 148 / tiles;
  const int cap = T / 512;   // at least 8 k-blocks per CTA
  if (s > cap) s = cap;
  return s < 1 ? 1 : s;
}
bool twgrad_supported(const WgradArgs& a) {
  if (!twgrad_shape_ok(a.M, a.N, a.T)) return false;
  if (a.x_rows || a.d_T || a.accumulate || a.splits < 1) return false;
  if (a.ldx % 8 || a.ldy % 8 || ((uintptr_t)a.X & 15) || ((uintptr_t)a.dY & 15) || (a.ld_out % 4)) return false;
  return true;
}
cudaError_t launch_twgrad(const WgradArgs& a, cudaStream_t st) {
  CUtensorMap tmX, tmY;
  if (!make_tmap_bf16_sw128(&tmX, a.X, (uint64_t)a.T, (uint64_t)a.M, (uint64_t)a.ldx, 64)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_sw128(&tmY, a.dY, (uint64_t)a.T, (uint64_t)a.N, (uint64_t)a.ldy, 64)) return cudaErrorInvalidValue;
  TWgradDev d;
  d.M = a.M; d.N = a.N; d.T = a.T; d.tiles_n = (a.N + 255) / 256; d.out = a.out; d.split_stride = a.split_stride; d.ld_out = a.ld_out; d.colsum = a.colsum_part;
  d.t_per_split = ((a.T + a.splits - 1) / a.splits + 63) / 64 * 64;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(twgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM);
    if (e != cudaSuccess) return e;
    done = true;
  }
  dim3 grid(((a.M + 255) / 256) * ((a.N + 255) / 256), a.splits);
  launch_pdl(twgrad_kernel, dim3(grid), dim3(320), (size_t)(TW_SMEM), st, tmX, tmY, d);
  return cudaGetLastError();
}
}  // namespace b4r

// ======================================================================================= GEMM + full-row epilogue (generation 2)
// y = LayerNorm(drop(A W + bias) + residual) for hidden 256 and 64 (ROW_RES_DROP_LN of k_gemm.cu): one CTA tile = 128 whole
// rows x H columns (an H-column TMEM accumulator, double-buffered: all 512 columns at hidden 256), so the row statistics never
// leave the CTA.  Epilogue thread = (row, H/4-column quarter): sixteen epilogue warps (four per scheduler) -- with eight the two
// passes over the tile were paced by the latency of each warp's own tcgen05.ld / Philox / shared-memory chain.  The tile's residual rows arrive by TMA in a 64 KB swizzled buffer;
// pass 1 turns them IN PLACE into the bf16 pre-LN value (accumulating sum and sum of squares, exchanged between the two halves
// through shared memory) which leaves by TMA store; pass 2 normalises in place and the LN output leaves by TMA store.  No
// per-thread global row accesses remain (they touched 32 half-used sectors per instruction and cost ~17 us per tile).
// Same persistent producer / issuer / epilogue split as tgemm_kernel.
namespace b4r {
using namespace encf;
namespace {
constexpr int TR_STAGES = 3;
__host__ __device__ constexpr int tr_stage(int H) { return 128 * 128 + 64 * H * 2; }   // 16 KB of A + 64 x H of W per 64-wide k-block
__host__ __device__ constexpr int tr_buf(int H) { return (H / 64) * 128 * 128; }        // residual -> pre-LN -> LN output tile: H/64 x [128][64] bf16
constexpr int TR_PARTS = 4, TR_THREADS = 64 + 128 * TR_PARTS;
__host__ __device__ constexpr int tr_smem(int H) { return TR_STAGES * tr_stage(H) + tr_buf(H) + 2 * TR_PARTS * 128 * 2 * 4 + 256 + 1024; }
struct TRowDev {
  int M, K;
  const float* bias; const float* gamma; const float* beta;
  const bf16* residual; bf16* pre; bf16* y; float* mean; float* rstd;
  uint32_t thr16; float inv_keep; unsigned long long seed; uint32_t site; uint32_t step; const long long* d_step;
};
}  // namespace

template <int H>
__global__ void __launch_bounds__(TR_THREADS, 1) trowln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                        const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmP,
                                                        const __grid_constant__ CUtensorMap tmY, TRowDev a) {
  pdl_grid_wait_single_wave();
  constexpr int TR_STAGE = tr_stage(H), TR_BUF = tr_buf(H), NB = H / 64, HC = H / TR_PARTS, NCH = HC / 16;   // column blocks, columns / 16-column chunks per part
  constexpr int TCOLS = 2 * H < 32 ? 32 : 2 * H;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sBuf = smem + TR_STAGES * TR_STAGE;
  float* sStat = reinterpret_cast<float*>(sBuf + TR_BUF);   // [2 acc][TR_PARTS][128 rows][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * TR_PARTS * 128 * 2);
  uint64_t* full = bars;
  uint64_t* empty = bars + TR_STAGES;
  uint64_t* tfull = bars + 2 * TR_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resfull = tempty + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(resfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.M + 127) / 128, kblocks = a.K / 64;
  const uint32_t step = a.step + (a.d_step ? (uint32_t)(*a.d_step) : 0u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TR_STAGES; ++i) { umma::mbar_init(full + i, 1); umma::mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(tfull + i, 1); umma::mbar_init(tempty + i, 4 * TR_PARTS); }
    umma::mbar_init(resfull, 1);
    umma::fence_barrier_init();
    umma::prefetch_tensormap(&tmA);
    umma::prefetch_tensormap(&tmW);
    umma::prefetch_tensormap(&tmR); umma::prefetch_tensormap(&tmP); umma::prefetch_tensormap(&tmY);
  }
  if (warp == 1) umma::tmem_alloc<TCOLS>(tmem_holder);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int st = it % TR_STAGES;
          umma::mbar_wait(empty + st, ((it / TR_STAGES) & 1) ^ 1);
          umma::mbar_expect_tx(full + st, TR_STAGE);
          unsigned char* sA = smem + st * TR_STAGE;
          unsigned char* sB = sA + 128 * 128;
          umma::tma_load_2d(sA, &tmA, kb * 64, t * 128, full + st);
#pragma unroll
          for (int j = 0; j < NB; ++j) umma::tma_load_2d(sB + j * 8192, &tmW, j * 64, kb * 64, full + st);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint64_t DK0 = umma::make_desc_k_sw128(umma::smem_addr(smem));
      const uint64_t DMN0 = desc_mn_sw128(umma::smem_addr(smem), 8192);
      constexpr uint32_t idesc = idesc_gen(128, H, 0, 1);
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
        const int acc = ti & 1;
        umma::mbar_wait(tempty + acc, ((ti >> 1) & 1) ^ 1);
        umma::fence_after_sync();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int st = it % TR_STAGES;
          umma::mbar_wait(full + st, (it / TR_STAGES) & 1);
          umma::fence_after_sync();
          const uint32_t offA = st * TR_STAGE, offB = offA + 128 * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16_ss(tmem + acc * H, desc_at(DK0, offA + k * 32), desc_at(DMN0, offB + k * 2048), idesc, (kb | k) ? 1u : 0u);
          umma::mma_commit(empty + st);
        }
        umma::mma_commit(tfull + acc);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const Philox ph(a.seed);
    const bool leader = threadIdx.x == 64;
    auto load_residual = [&](int tt) {
      umma::mbar_expect_tx(resfull, TR_BUF);
#pragma unroll
      for (int j = 0; j < NB; ++j) umma::tma_load_2d(sBuf + j * TILE_B, &tmR, j * 64, tt * 128, resfull);
    };
    if (leader && (int)blockIdx.x < n_tiles) load_residual(blockIdx.x);
    int ti = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
      const int acc = ti & 1;
      const int m = t * 128 + row_in_tile;
      const bool mok = m < a.M;
      umma::mbar_wait(tfull + acc, (ti >> 1) & 1);
      umma::fence_after_sync();
      umma::mbar_wait(resfull, ti & 1);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int n = part * HC + c * 16;
        unsigned char* bt = sBuf + (n >> 6) * TILE_B;
        const int ch = (n & 63) >> 3;        // first 16-byte chunk of these 16 columns inside the [128][64] block
        float v[16];
        tmem_ld_f16(tmem + ((uint32_t)(quad * 32) << 16) + acc * H + n, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n + i));
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
        if (a.thr16 > 0 && mok) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t bits = keep_bits8(ph, (uint32_t)m, (uint32_t)(n / 8 + q), a.site, step, a.thr16);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * q + i] = ((bits >> i) & 1u) ? v[8 * q + i] * a.inv_keep : 0.f;
          }
        }
        float res[16];
        ld_tile<2>(bt, row_in_tile, ch, res);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += res[i];
        uint32_t pk[8];
        round_n<16>(v, pk);   // LN statistics are taken on the bf16-rounded value that backward will re-read
        st_tile<2>(bt, row_in_tile, ch, pk);
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1 += v[i]; s2 += v[i] * v[i]; }
      }
      // the accumulator has been read completely: hand it back to the MMA warp before the second pass
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(tempty + acc);
      float* st = sStat + ((acc * TR_PARTS + part) * 128 + row_in_tile) * 2;
      st[0] = s1; st[1] = s2;
      umma::fence_proxy_async();
      asm volatile("bar.sync 1, %0;\n" ::"n"(128 * TR_PARTS) : "memory");
      if (leader) {
#pragma unroll
        for (int j = 0; j < NB; ++j) umma::tma_store_2d(&tmP, sBuf + j * TILE_B, j * 64, t * 128);
        umma::tma_store_commit();
      }
      float t1 = 0.f, t2 = 0.f;       // the four parts of a row in part order: every thread of the row gets the same statistics
#pragma unroll
      for (int q = 0; q < TR_PARTS; ++q) {
        const float* so = sStat + ((acc * TR_PARTS + q) * 128 + row_in_tile) * 2;
        t1 += so[0]; t2 += so[1];
      }
      const float mean = t1 * (1.0f / H);
      const float var = fmaxf(t2 * (1.0f / H) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + kLnEps);
      if (leader) umma::tma_store_wait_read<0>();
      asm volatile("bar.sync 1, %0;\n" ::"n"(128 * TR_PARTS) : "memory");     // the pre-LN store has read the buffer: normalise in place
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int n = part * HC + c * 16;
        unsigned char* bt = sBuf + (n >> 6) * TILE_B;
        const int ch = (n & 63) >> 3;
        float v[16];
        ld_tile<2>(bt, row_in_tile, ch, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma + n + i));
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta + n + i));
          v[i] = (v[i] - mean) * rstd * g.x + b.x; v[i + 1] = (v[i + 1] - mean) * rstd * g.y + b.y;
          v[i + 2] = (v[i + 2] - mean) * rstd * g.z + b.z; v[i + 3] = (v[i + 3] - mean) * rstd * g.w + b.w;
        }
        uint32_t pk[8];
        pack_n<16>(v, pk);
        st_tile<2>(bt, row_in_tile, ch, pk);
      }
      if (mok && part == 0) { a.mean[m] = mean; a.rstd[m] = rstd; }
      umma::fence_proxy_async();
      asm volatile("bar.sync 1, %0;\n" ::"n"(128 * TR_PARTS) : "memory");
      if (leader) {
#pragma unroll
        for (int j = 0; j < NB; ++j) umma::tma_store_2d(&tmY, sBuf + j * TILE_B, j * 64, t * 128);
        umma::tma_store_commit();
        umma::tma_store_wait_read<0>();
        if (t + (int)gridDim.x < n_tiles) load_residual(t + gridDim.x);
      }
    }
    if (leader) umma::tma_store_wait<0>();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    umma::fence_after_sync();
    umma::tmem_dealloc<TCOLS>(tmem);
  }
}

bool trowln_supported(int mode, const RowLnArgs& a) {
  if (getenv("B4R_DISABLE_TGEMM")) return false;
  if (mode != ROW_RES_DROP_LN || (a.H != 256 && a.H != 64) || a.a_rows || a.d_M) return false;
  if (a.K % 64 || a.K < 64 || a.M < 256 || a.lda % 8 || ((uintptr_t)a.A & 15) || ((uintptr_t)a.W & 15)) return false;
  if (((uintptr_t)a.residual & 15) || ((uintptr_t)a.pre & 15) || ((uintptr_t)a.y & 15)) return false;   // TMA tiles
  return true;
}
cudaError_t launch_trowln(const RowLnArgs& a, cudaStream_t st) {
  CUtensorMap tmA, tmW;
  if (!make_tmap_bf16_sw128(&tmA, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, 128)) return cudaErrorInvalidValue;
  const uint64_t HH = (uint64_t)a.H;
  if (!make_tmap_bf16_sw128(&tmW, a.W, (uint64_t)a.K, HH, HH, 64)) return cudaErrorInvalidValue;
  CUtensorMap tmR, tmP, tmY;
  if (!make_tmap_bf16_sw128(&tmR, a.residual, (uint64_t)a.M, HH, HH, 128) || !make_tmap_bf16_sw128(&tmP, a.pre, (uint64_t)a.M, HH, HH, 128) ||
      !make_tmap_bf16_sw128(&tmY, a.y, (uint64_t)a.M, HH, HH, 128))
    return cudaErrorInvalidValue;
  TRowDev d;
  d.M = a.M; d.K = a.K; d.bias = a.bias; d.gamma = a.gamma; d.beta = a.beta; d.residual = a.residual; d.pre = a.pre; d.y = a.y;
  d.mean = a.mean; d.rstd = a.rstd;
  d.thr16 = drop_threshold16(a.drop_rate);
  d.inv_keep = 1.0f / (1.0f - (float)d.thr16 / 65536.0f);
  d.seed = a.seed; d.site = a.site; d.step = a.step; d.d_step = a.d_step;
  const int tiles = (a.M + 127) / 128;
  if (a.H == 256) {
    static bool done = false;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(trowln_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem(256));
      if (e != cudaSuccess) return e;
      done = true;
    }
    launch_pdl(trowln_kernel<256>, dim3(tiles < 148 ? tiles : 148), dim3(TR_THREADS), (size_t)tr_smem(256), st, tmA, tmW, tmR, tmP, tmY, d);
  } else {
    static bool done = false;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(trowln_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem(64));
      if (e != cudaSuccess) return e;
      done = true;
    }
    launch_pdl(trowln_kernel<64>, dim3(tiles < 148 ? tiles : 148), dim3(TR_THREADS), (size_t)tr_smem(64), st, tmA, tmW, tmR, tmP, tmY, d);
  }
  return cudaGetLastError();
}
}  // namespace b4r
