// CTA-level bf16 GEMM main loop (cp.async multi-stage pipeline + ldmatrix + mma.sync m16n8k16),
// shared by every GEMM-shaped kernel of the first (portable-ISA) generation of this path.
// The tcgen05/TMEM generation lives in umma_*.cuh and replaces the hot instances.
#pragma once
#include "common.cuh"

namespace b4r {

struct GemmOperands {
  const bf16* A;      // !A_TRANS: memory [m][k] ; A_TRANS: memory [k][m]
  int lda;            // elements between consecutive memory rows of A
  const int* a_rows;  // optional gather applied to A's *memory row* index (m or k)
  const bf16* B;      // !B_TRANS: memory [n][k] ; B_TRANS: memory [k][n]
  int ldb;
  int a_mmax;         // valid extent of A along m (rows if !A_TRANS, else columns; multiple of 8 if columns)
  int a_kmax;         // valid extent of A along k
  int b_nmax;         // valid extent of B along n
  int b_kmax;         // valid extent of B along k
  int k_begin, k_end; // contraction range handled by this CTA
};

template <int BM_, int BN_, int BK_, int WARPS_M_, int WARPS_N_, bool A_TRANS_, bool B_TRANS_, int STAGES_>
struct GemmTile {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, STAGES = STAGES_;
  static constexpr bool A_TRANS = A_TRANS_, B_TRANS = B_TRANS_;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
  static constexpr int MI = WTM / 16, NI = WTN / 8;
  static constexpr int A_ROWS = A_TRANS ? BK : BM, A_COLS = A_TRANS ? BM : BK, A_LD = A_COLS + 8;
  static constexpr int B_ROWS = B_TRANS ? BK : BN, B_COLS = B_TRANS ? BN : BK, B_LD = B_COLS + 8;
  static constexpr int A_ELEMS = A_ROWS * A_LD, B_ELEMS = B_ROWS * B_LD;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t PIPE_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(bf16);
  static_assert(WTM % 16 == 0 && WTN % 16 == 0, "warp tile must be a multiple of 16x16");
  static_assert(BK % 16 == 0, "BK must be a multiple of 16");
};

template <class T>
__device__ __forceinline__ void gemm_load_stage(const GemmOperands& op, int m0, int n0, int k0, bf16* sA, bf16* sB) {
  const int tid = threadIdx.x;
  // ---- A
  constexpr int A_CH = T::A_COLS / 8;
  for (int c = tid; c < T::A_ROWS * A_CH; c += T::THREADS) {
    int r = c / A_CH, cc = (c % A_CH) * 8;
    int mem_row, mem_col;
    bool ok;
    if (!T::A_TRANS) { mem_row = m0 + r; mem_col = k0 + cc; ok = mem_row < op.a_mmax && mem_col < op.a_kmax && mem_col < op.k_end; }
    else             { mem_row = k0 + r; mem_col = m0 + cc; ok = mem_row < op.a_kmax && mem_row < op.k_end && mem_col < op.a_mmax; }
    const bf16* src = op.A;
    if (ok) {
      int pr = op.a_rows ? op.a_rows[mem_row] : mem_row;
      src = op.A + (size_t)pr * op.lda + mem_col;
    }
    cp_async16(sA + r * T::A_LD + cc, src, ok);
  }
  // ---- B
  constexpr int B_CH = T::B_COLS / 8;
  for (int c = tid; c < T::B_ROWS * B_CH; c += T::THREADS) {
    int r = c / B_CH, cc = (c % B_CH) * 8;
    int mem_row, mem_col;
    bool ok;
    if (!T::B_TRANS) { mem_row = n0 + r; mem_col = k0 + cc; ok = mem_row < op.b_nmax && mem_col < op.b_kmax && mem_col < op.k_end; }
    else             { mem_row = k0 + r; mem_col = n0 + cc; ok = mem_row < op.b_kmax && mem_row < op.k_end && mem_col < op.b_nmax; }
    const bf16* src = ok ? op.B + (size_t)mem_row * op.ldb + mem_col : op.B;
    cp_async16(sB + r * T::B_LD + cc, src, ok);
  }
}

template <class T>
__device__ __forceinline__ void gemm_compute_stage(const bf16* sA, const bf16* sB, int warp_m, int warp_n, int lane,
                                                   float (&acc)[T::MI][T::NI][4]) {
#pragma unroll
  for (int kk = 0; kk < T::BK; kk += 16) {
    uint32_t a[T::MI][4];
#pragma unroll
    for (int mi = 0; mi < T::MI; ++mi)
      load_a_frag<T::A_TRANS>(a[mi], sA, T::A_LD, warp_m * T::WTM + mi * 16, kk, lane);
#pragma unroll
    for (int np = 0; np < T::NI / 2; ++np) {
      uint32_t b[4];
      load_b_frag<T::B_TRANS>(b, sB, T::B_LD, warp_n * T::WTN + np * 16, kk, lane);
#pragma unroll
      for (int mi = 0; mi < T::MI; ++mi) {
        mma_bf16(acc[mi][2 * np], a[mi], b[0], b[1]);
        mma_bf16(acc[mi][2 * np + 1], a[mi], b[2], b[3]);
      }
    }
  }
}

// acc[mi][ni][e] covers C(m0 + warp_m*WTM + mi*16 + frag_row(lane,e), n0 + warp_n*WTN + ni*8 + frag_col(lane,e)).
// On return all cp.async groups are drained and the CTA is synchronised (pipeline smem reusable).
template <class T>
__device__ __forceinline__ void gemm_mainloop(const GemmOperands& op, int m0, int n0, bf16* smem,
                                              float (&acc)[T::MI][T::NI][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_m = warp % T::WARPS_M, warp_n = warp / T::WARPS_M;
#pragma unroll
  for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][ni][e] = 0.f;

  const int nk = (op.k_end - op.k_begin + T::BK - 1) / T::BK;
#pragma unroll
  for (int s = 0; s < T::STAGES - 1; ++s) {
    if (s < nk) gemm_load_stage<T>(op, m0, n0, op.k_begin + s * T::BK, smem + s * T::STAGE_ELEMS,
                                   smem + s * T::STAGE_ELEMS + T::A_ELEMS);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<T::STAGES - 2>();
    __syncthreads();
    int nxt = kt + T::STAGES - 1;
    if (nxt < nk) {
      int s = nxt % T::STAGES;
      gemm_load_stage<T>(op, m0, n0, op.k_begin + nxt * T::BK, smem + s * T::STAGE_ELEMS,
                         smem + s * T::STAGE_ELEMS + T::A_ELEMS);
    }
    cp_async_commit();
    int s = kt % T::STAGES;
    gemm_compute_stage<T>(smem + s * T::STAGE_ELEMS, smem + s * T::STAGE_ELEMS + T::A_ELEMS, warp_m, warp_n, lane, acc);
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace b4r
