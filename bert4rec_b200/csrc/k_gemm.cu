// GEMM-shaped kernels, generation 1: mma.sync tiles with fused epilogues.
//   gemm_kernel      - Y = op(A) op(B) with bias / GELU / GELU' / residual / scatter / split-K epilogues
//   gemm_rowln_kernel- Y = A W with a full-row epilogue: bias (+dropout +residual | +GELU) + LayerNorm(eps 1e-12)
//   wgrad_kernel     - dW = X^T dY (split over the token dimension, deterministic partials)
// Reference ops replaced: Keras EinsumDense / MultiHeadAttention projections and LayerNormalization inside
// tfm.nlp.layers.TransformerEncoderBlock (bert4rec_encoder.py:136-147) and tfm MaskedLM dense+LN
// (bert4rec_model.py:76-81); see SURVEY.md 2b rows K3-K5, K7.
#include "gemm.cuh"
#include "common.cuh"
#include "kernels.h"

namespace b4r {

// ======================================================================================= generic GEMM
typedef GemmTile<128, 64, 32, 4, 2, false, false, 3> TileNT;  // A [m][k], B [n][k]
typedef GemmTile<128, 64, 32, 4, 2, false, true, 3> TileNN;   // A [m][k], B [k][n]

int gemm_block_m() { return 128; }

struct EpiDev {
  int M, N;
  const float* bias;
  bf16* out_bf16; int ld_out;
  bf16* out2_bf16;
  const bf16* aux_bf16; int ld_aux;
  float* out_f32; int ld_f32;
  const float* res_f32;
  const int* scatter_rows;
  float* colsum_part;
  size_t split_stride;
};

template <class T, int EPI>
__global__ void __launch_bounds__(T::THREADS) gemm_kernel(GemmOperands op, EpiDev ep, const int* d_M, int d_M_off, int k_per_split) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* smem = reinterpret_cast<bf16*>(smem_raw);
  int M = ep.M;
  if (d_M) M = min(M, *d_M - d_M_off);
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  if (m0 >= M) {
    if (EPI == EPI_GELU_GRAD && ep.colsum_part) {  // keep partials defined for skipped tiles
      for (int c = threadIdx.x; c < T::BN; c += T::THREADS)
        if (n0 + c < ep.N) ep.colsum_part[(size_t)blockIdx.y * ep.N + n0 + c] = 0.f;
    }
    return;
  }
  op.a_mmax = min(op.a_mmax, M);
  op.k_begin = blockIdx.z * k_per_split;
  op.k_end = min(op.k_end, op.k_begin + k_per_split);

  float acc[T::MI][T::NI][4];
  gemm_mainloop<T>(op, m0, n0, smem, acc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_m = warp % T::WARPS_M, warp_n = warp / T::WARPS_M;
  float* s_col = reinterpret_cast<float*>(smem_raw);  // [WARPS_M][BN] column partial sums (EPI_GELU_GRAD)
  float colsum[T::NI][2];
#pragma unroll
  for (int ni = 0; ni < T::NI; ++ni) colsum[ni][0] = colsum[ni][1] = 0.f;

#pragma unroll
  for (int mi = 0; mi < T::MI; ++mi) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + warp_m * T::WTM + mi * 16 + (lane >> 2) + half * 8;
      const bool mok = m < M;
#pragma unroll
      for (int ni = 0; ni < T::NI; ++ni) {
        const int n = n0 + warp_n * T::WTN + ni * 8 + ((lane & 3) << 1);
        if (!mok || n >= ep.N) continue;  // N is even everywhere on this path
        float v0 = acc[mi][ni][half * 2], v1 = acc[mi][ni][half * 2 + 1];
        const bool has1 = (EPI != EPI_BIAS_F32) || (n + 1 < ep.N);  // only the logits GEMM can have odd N
        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_F32) {
          v0 += ep.bias[n];
          if (has1) v1 += ep.bias[n + 1];
        }
        if (EPI == EPI_BIAS_BF16 || EPI == EPI_BF16) {
          *reinterpret_cast<uint32_t*>(ep.out_bf16 + (size_t)m * ep.ld_out + n) = pack_bf162(v0, v1);
        } else if (EPI == EPI_BIAS_GELU) {
          *reinterpret_cast<uint32_t*>(ep.out_bf16 + (size_t)m * ep.ld_out + n) = pack_bf162(v0, v1);
          // GELU is applied to the bf16-rounded pre-activation so that backward (which re-reads it) is consistent
          float2 r = unpack_bf162(pack_bf162(v0, v1));
          *reinterpret_cast<uint32_t*>(ep.out2_bf16 + (size_t)m * ep.ld_out + n) = pack_bf162(gelu_erf(r.x), gelu_erf(r.y));
        } else if (EPI == EPI_GELU_GRAD) {
          float2 pre = unpack_bf162(*reinterpret_cast<const uint32_t*>(ep.aux_bf16 + (size_t)m * ep.ld_aux + n));
          v0 *= gelu_erf_grad(pre.x); v1 *= gelu_erf_grad(pre.y);
          uint32_t pk = pack_bf162(v0, v1);
          *reinterpret_cast<uint32_t*>(ep.out_bf16 + (size_t)m * ep.ld_out + n) = pk;
          float2 r = unpack_bf162(pk);  // bias gradient sums what the weight-gradient GEMM will see
          colsum[ni][0] += r.x; colsum[ni][1] += r.y;
        } else if (EPI == EPI_F32_RES) {
          const float2 r = *reinterpret_cast<const float2*>(ep.res_f32 + (size_t)m * ep.ld_f32 + n);
          *reinterpret_cast<float2*>(ep.out_f32 + (size_t)m * ep.ld_f32 + n) = make_float2(v0 + r.x, v1 + r.y);
        } else if (EPI == EPI_SCATTER_F32) {
          const int dst = ep.scatter_rows[m];
          *reinterpret_cast<float2*>(ep.out_f32 + (size_t)dst * ep.ld_f32 + n) = make_float2(v0, v1);
        } else if (EPI == EPI_F32_PARTIAL) {
          *reinterpret_cast<float2*>(ep.out_f32 + blockIdx.z * ep.split_stride + (size_t)m * ep.ld_f32 + n) = make_float2(v0, v1);
        } else if (EPI == EPI_BIAS_F32) {
          ep.out_f32[(size_t)m * ep.ld_f32 + n] = v0;
          if (has1) ep.out_f32[(size_t)m * ep.ld_f32 + n + 1] = v1;
        }
      }
    }
  }
  if (EPI == EPI_GELU_GRAD && ep.colsum_part) {
    // reduce over the 8 row-groups of the warp (lanes sharing lane&3), then over WARPS_M warps through smem
#pragma unroll
    for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = colsum[ni][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        colsum[ni][e] = v;
      }
    if ((lane >> 2) == 0) {
#pragma unroll
      for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          s_col[warp_m * T::BN + warp_n * T::WTN + ni * 8 + ((lane & 3) << 1) + e] = colsum[ni][e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < T::BN; c += T::THREADS) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < T::WARPS_M; ++w) v += s_col[w * T::BN + c];
      if (n0 + c < ep.N) ep.colsum_part[(size_t)blockIdx.y * ep.N + n0 + c] = v;
    }
  }
}

template <class T, int EPI>
static cudaError_t launch_gemm_t(const GemmArgs& a, cudaStream_t st) {
  GemmOperands op;
  op.A = a.A; op.lda = a.lda; op.a_rows = a.a_rows; op.B = a.B; op.ldb = a.ldb;
  op.a_mmax = a.M; op.a_kmax = a.a_kmax ? a.a_kmax : a.K;
  op.b_nmax = a.N; op.b_kmax = a.b_kmax ? a.b_kmax : a.K;
  op.k_begin = 0; op.k_end = a.K;
  EpiDev ep;
  ep.M = a.M; ep.N = a.N; ep.bias = a.bias; ep.out_bf16 = a.out_bf16; ep.ld_out = a.ld_out; ep.out2_bf16 = a.out2_bf16;
  ep.aux_bf16 = a.aux_bf16; ep.ld_aux = a.ld_aux; ep.out_f32 = a.out_f32; ep.ld_f32 = a.ld_f32; ep.res_f32 = a.res_f32;
  ep.scatter_rows = a.scatter_rows; ep.colsum_part = a.colsum_part; ep.split_stride = a.split_stride;
  int splits = a.splits > 0 ? a.splits : 1;
  int kps = ((a.K + splits - 1) / splits + T::BK - 1) / T::BK * T::BK;
  dim3 grid((a.N + T::BN - 1) / T::BN, (a.M + T::BM - 1) / T::BM, splits);
  size_t smem = T::PIPE_BYTES;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_kernel<T, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  launch_pdl(gemm_kernel<T, EPI>, dim3(grid), dim3(T::THREADS), (size_t)(smem), st, op, ep, a.d_M, a.d_M_off, kps);
  return cudaGetLastError();
}

cudaError_t launch_gemm(int epi, const GemmArgs& a, cudaStream_t st) {
  if (a.a_trans) return cudaErrorInvalidValue;
  if (tgemm_supported(epi, a)) return launch_tgemm(epi, a, st);   // generation 2: tcgen05 + TMEM + TMA
#define B4R_CASE(E)                                                          \
  case E:                                                                    \
    return a.b_trans ? launch_gemm_t<TileNN, E>(a, st) : launch_gemm_t<TileNT, E>(a, st);
  switch (epi) {
    B4R_CASE(EPI_BIAS_BF16)
    B4R_CASE(EPI_BIAS_GELU)
    B4R_CASE(EPI_GELU_GRAD)
    B4R_CASE(EPI_BF16)
    B4R_CASE(EPI_F32_RES)
    B4R_CASE(EPI_SCATTER_F32)
    B4R_CASE(EPI_F32_PARTIAL)
    B4R_CASE(EPI_BIAS_F32)
  }
#undef B4R_CASE
  return cudaErrorInvalidValue;
}

// ======================================================================================= GEMM + full-row epilogue
template <int H>
struct RowTile {
  // BM = 64 rows, BN = H columns (the whole row), 8 warps
  typedef GemmTile<64, H, 32, (H >= 256 ? 2 : 4), (H >= 256 ? 4 : 2), false, true, 3> T;
  static constexpr int LDS = H + 4;  // fp32 staging row stride
  static constexpr size_t STAGE_BYTES = (size_t)64 * LDS * sizeof(float);
  static constexpr size_t SMEM = T::PIPE_BYTES > STAGE_BYTES ? T::PIPE_BYTES : STAGE_BYTES;
};

struct RowLnDev {
  int M, H;
  const float* bias; const float* gamma; const float* beta;
  const bf16* residual; bf16* pre; bf16* act; bf16* y; float* mean; float* rstd;
  uint32_t thr16; float inv_keep; uint64_t seed; uint32_t site; uint32_t step; const long long* d_step;
};

template <int H, int MODE>
__global__ void __launch_bounds__(256) gemm_rowln_kernel(GemmOperands op, RowLnDev ep, const int* d_M) {
  pdl_grid_sync();
  typedef typename RowTile<H>::T T;
  constexpr int LDS = RowTile<H>::LDS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* smem = reinterpret_cast<bf16*>(smem_raw);
  int M = ep.M;
  if (d_M) M = min(M, *d_M);
  const int m0 = blockIdx.x * T::BM;
  if (m0 >= M) return;
  if (ep.d_step) ep.step += (uint32_t)(*ep.d_step);
  op.a_mmax = min(op.a_mmax, M);

  float acc[T::MI][T::NI][4];
  gemm_mainloop<T>(op, m0, 0, smem, acc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_m = warp % T::WARPS_M, warp_n = warp / T::WARPS_M;
  float* stage = reinterpret_cast<float*>(smem_raw);
#pragma unroll
  for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < T::NI; ++ni)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int r = warp_m * T::WTM + mi * 16 + (lane >> 2) + half * 8;
        int c = warp_n * T::WTN + ni * 8 + ((lane & 3) << 1);
        *reinterpret_cast<float2*>(stage + r * LDS + c) = make_float2(acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1]);
      }
  __syncthreads();

  // row pass: LPR lanes per row, 8 columns per lane
  constexpr int LPR = H / 8;
  constexpr int RPW = 32 / LPR;  // rows per warp pass
  const int sub = lane / LPR, l = lane % LPR;
  const Philox ph(ep.seed);
  for (int r = warp * RPW + sub; r < T::BM; r += 8 * RPW) {
    const int m = m0 + r;
    const bool ok = m < M;  // uniform within each LPR group; all lanes run the shuffles
    const int c0 = l * 8;
    float v[8];
    {
      float4 a = *reinterpret_cast<const float4*>(stage + r * LDS + c0);
      float4 b = *reinterpret_cast<const float4*>(stage + r * LDS + c0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += ep.bias[c0 + i];
    if (MODE == ROW_RES_DROP_LN) {
      if (ep.thr16 > 0 && ok) {
        uint32_t bits = keep_bits8(ph, (uint32_t)m, (uint32_t)l, ep.site, ep.step, ep.thr16);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * ep.inv_keep : 0.f;
      }
      if (ok) {
        uint4 rr = *reinterpret_cast<const uint4*>(ep.residual + (size_t)m * H + c0);
        uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(w[i]); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
        uint4 o;
        o.x = pack_bf162(v[0], v[1]); o.y = pack_bf162(v[2], v[3]); o.z = pack_bf162(v[4], v[5]); o.w = pack_bf162(v[6], v[7]);
        *reinterpret_cast<uint4*>(ep.pre + (size_t)m * H + c0) = o;
        // LN statistics are taken on the bf16-rounded value that backward will re-read
        uint32_t ww[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(ww[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      }
    } else {
      if (ok) {
        uint4 o;
        o.x = pack_bf162(v[0], v[1]); o.y = pack_bf162(v[2], v[3]); o.z = pack_bf162(v[4], v[5]); o.w = pack_bf162(v[6], v[7]);
        *reinterpret_cast<uint4*>(ep.pre + (size_t)m * H + c0) = o;
        uint32_t ww[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = unpack_bf162(ww[i]);
          v[2 * i] = gelu_erf(f.x); v[2 * i + 1] = gelu_erf(f.y);
        }
        o.x = pack_bf162(v[0], v[1]); o.y = pack_bf162(v[2], v[3]); o.z = pack_bf162(v[4], v[5]); o.w = pack_bf162(v[6], v[7]);
        *reinterpret_cast<uint4*>(ep.act + (size_t)m * H + c0) = o;
        uint32_t w2[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(w2[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    s = group_sum<LPR>(s);
    const float mean = s * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float d = v[i] - mean; q += d * d; }
    q = group_sum<LPR>(q);
    const float rstd = rsqrtf(q * (1.0f / H) + kLnEps);
    if (ok) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * ep.gamma[c0 + i] + ep.beta[c0 + i];
      uint4 ov;
      ov.x = pack_bf162(o[0], o[1]); ov.y = pack_bf162(o[2], o[3]); ov.z = pack_bf162(o[4], o[5]); ov.w = pack_bf162(o[6], o[7]);
      *reinterpret_cast<uint4*>(ep.y + (size_t)m * H + c0) = ov;
      if (l == 0) { ep.mean[m] = mean; ep.rstd[m] = rstd; }
    }
  }
}

template <int H, int MODE>
static cudaError_t launch_rowln_t(const RowLnArgs& a, cudaStream_t st) {
  typedef typename RowTile<H>::T T;
  GemmOperands op;
  op.A = a.A; op.lda = a.lda; op.a_rows = a.a_rows; op.B = a.W; op.ldb = H;
  op.a_mmax = a.M; op.a_kmax = a.K; op.b_nmax = H; op.b_kmax = a.K; op.k_begin = 0; op.k_end = a.K;
  RowLnDev ep;
  ep.M = a.M; ep.H = H; ep.bias = a.bias; ep.gamma = a.gamma; ep.beta = a.beta; ep.residual = a.residual;
  ep.pre = a.pre; ep.act = a.act; ep.y = a.y; ep.mean = a.mean; ep.rstd = a.rstd;
  ep.thr16 = drop_threshold16(a.drop_rate);
  ep.inv_keep = 1.0f / (1.0f - (float)ep.thr16 / 65536.0f);
  ep.seed = a.seed; ep.site = a.site; ep.step = a.step; ep.d_step = a.d_step;
  size_t smem = RowTile<H>::SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_rowln_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  launch_pdl(gemm_rowln_kernel<H, MODE>, dim3((a.M + T::BM - 1) / T::BM), dim3(256), (size_t)smem, st, op, ep, a.d_M);
  return cudaGetLastError();
}

cudaError_t launch_gemm_rowln(int mode, const RowLnArgs& a, cudaStream_t st) {
  if (trowln_supported(mode, a)) return launch_trowln(a, st);   // generation 2: tcgen05 + TMEM + TMA
#define B4R_ROW(HH)                                                                       \
  case HH:                                                                                \
    return mode == ROW_RES_DROP_LN ? launch_rowln_t<HH, ROW_RES_DROP_LN>(a, st) : launch_rowln_t<HH, ROW_GELU_LN>(a, st);
  switch (a.H) {
    B4R_ROW(64)
    B4R_ROW(128)
    B4R_ROW(256)
  }
#undef B4R_ROW
  return cudaErrorInvalidValue;
}

// ======================================================================================= weight gradient
typedef GemmTile<64, 64, 32, 2, 2, true, true, 4> TileWG;  // A = X^T ([t][m]), B = dY ([t][n])

__global__ void __launch_bounds__(TileWG::THREADS) wgrad_kernel(GemmOperands op, float* out, size_t split_stride,
                                                                int ld_out, int M, int N, int accumulate,
                                                                const int* d_T, int d_T_off, int t_per_split) {
  pdl_grid_wait();
  typedef TileWG T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* smem = reinterpret_cast<bf16*>(smem_raw);
  int Tn = op.k_end;
  if (d_T) Tn = max(0, min(Tn, *d_T - d_T_off));
  op.a_kmax = min(op.a_kmax, Tn);
  op.b_kmax = min(op.b_kmax, Tn);
  op.k_begin = blockIdx.z * t_per_split;
  op.k_end = min(Tn, op.k_begin + t_per_split);
  if (op.k_end < op.k_begin) op.k_end = op.k_begin;
  const int m0 = blockIdx.y * T::BM, n0 = blockIdx.x * T::BN;
  float acc[T::MI][T::NI][4];
  gemm_mainloop<T>(op, m0, n0, smem, acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_m = warp % T::WARPS_M, warp_n = warp / T::WARPS_M;
  float* dst = out + blockIdx.z * split_stride;
#pragma unroll
  for (int mi = 0; mi < T::MI; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + warp_m * T::WTM + mi * 16 + (lane >> 2) + half * 8;
      if (m >= M) continue;
#pragma unroll
      for (int ni = 0; ni < T::NI; ++ni) {
        const int n = n0 + warp_n * T::WTN + ni * 8 + ((lane & 3) << 1);
        if (n >= N) continue;
        float2* p = reinterpret_cast<float2*>(dst + (size_t)m * ld_out + n);
        float2 v = make_float2(acc[mi][ni][half * 2], acc[mi][ni][half * 2 + 1]);
        if (accumulate) { float2 o = *p; v.x += o.x; v.y += o.y; }
        *p = v;
      }
    }
}

cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  if (twgrad_supported(a)) return launch_twgrad(a, st);   // generation 2: tcgen05 + TMEM + TMA
  typedef TileWG T;
  GemmOperands op;
  op.A = a.X; op.lda = a.ldx; op.a_rows = a.x_rows; op.B = a.dY; op.ldb = a.ldy;
  op.a_mmax = a.x_mmax ? a.x_mmax : (a.M + 7) / 8 * 8;
  op.a_kmax = a.T; op.b_nmax = a.N; op.b_kmax = a.T; op.k_begin = 0; op.k_end = a.T;
  int splits = a.splits > 0 ? a.splits : 1;
  int tps = ((a.T + splits - 1) / splits + T::BK - 1) / T::BK * T::BK;
  dim3 grid((a.N + T::BN - 1) / T::BN, (a.M + T::BM - 1) / T::BM, splits);
  size_t smem = T::PIPE_BYTES;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  launch_pdl(wgrad_kernel, dim3(grid), dim3(T::THREADS), (size_t)(smem), st, op, a.out, a.split_stride, a.ld_out, a.M, a.N, a.accumulate, a.d_T, a.d_T_off, tps);
  return cudaGetLastError();
}

}  // namespace b4r
