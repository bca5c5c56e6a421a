// MLM transform backward, hidden 64, ONE launch (replaces head_bwd_rows + wgrad:head_wt + gemm:head_dx_scatter):
//   dt      = sum of the split partials of the tied-projection backward (dT pass)
//   d_act   = LayerNorm backward (gamma/beta gradients)          [t = LN(gelu(t_pre)), t_pre = x Wt + bt]
//   d_tpre  = d_act * gelu'(t_pre), rounded to bf16              (bias gradient = column sums of the rounded value)
//   dx      = d_tpre Wt^T  -> fp32 rows scattered to the token rows of the masked slots (d_out of the last layer)
//   dWt    += x^T d_tpre   -> register accumulators, one [64,64] partial per CTA
// The problem is ~2 k rows x 64: one warp per row, lane l owns columns 2l, 2l+1; no tensor cores needed.
// Reference op: backward of tfm.nlp.layers.MaskedLM's dense + LayerNorm (bert4rec_model.py:76-81,166-167).
#include "common.cuh"
#include "kernels.h"

namespace b4r {

constexpr int HF_H = 64;
constexpr int HF_LD = HF_H + 1;   // odd row stride: the transposing read of the final sum is conflict-free
constexpr int HF_SMEM = (8 * HF_H * HF_LD + 8 * 3 * HF_H) * 4;

__global__ void __launch_bounds__(256) head_bwd_fused_kernel(const float* __restrict__ dt_part, int nsplit, size_t split_stride,
                                                             const bf16* __restrict__ t_pre, const bf16* __restrict__ t_act,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const bf16* __restrict__ wt,
                                                             const bf16* __restrict__ x, const int* __restrict__ rows,
                                                             const int* __restrict__ d_counts, int M_cap, float* __restrict__ dx_out,
                                                             float* __restrict__ p_ln, float* __restrict__ p_wt, int dyn_vtiles,
                                                             int dyn_target, int dyn_max) {
  pdl_grid_sync();
  __shared__ float sW[HF_H][HF_H + 1];      // Wt[in][out] as fp32
  extern __shared__ float sAccW[];          // [8 warps][out][HF_LD]: every warp's dWt accumulator, TRANSPOSED ([out][in]); then
                                            // [8][3][HF_H]: its LayerNorm / bias sums -- summed over the warps in fixed order
  float* sLnW = sAccW + 8 * HF_H * HF_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = min(M_cap, d_counts[0]);    // valid masked slots only: aux rows carry no gradient
  if (dyn_max < 0) {                        // one-pass CE backward: slots = vocabulary groups (dyn_vtiles = V, dyn_target = its CTAs)
    nsplit = cf_split(M, dyn_vtiles, dyn_target).b;
  } else if (dyn_max > 0) {                 // split count chosen on the device by the generation-2 dT pass
    nsplit = ce_dyn_splits128(M, dyn_vtiles, dyn_target, dyn_max);
  }
  for (int i = tid; i < HF_H * HF_H / 8; i += 256) {   // 8 bf16 per 16-byte load
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wt) + i);
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
    const int r = (i * 8) / HF_H, c = (i * 8) % HF_H;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf162(w4[k]); sW[r][c + 2 * k] = f.x; sW[r][c + 2 * k + 1] = f.y; }
  }
  __syncthreads();
  const int c0 = 2 * lane;
  const float g0 = gamma[c0], g1 = gamma[c0 + 1];
  float acc0[HF_H], acc1[HF_H];             // dWt rows c0, c0+1
#pragma unroll
  for (int o = 0; o < HF_H; ++o) { acc0[o] = 0.f; acc1[o] = 0.f; }
  float a_g0 = 0.f, a_g1 = 0.f, a_b0 = 0.f, a_b1 = 0.f, a_c0 = 0.f, a_c1 = 0.f;
  const int nw = gridDim.x * 8;
  for (int m = blockIdx.x * 8 + warp; m < M; m += nw) {
    float d0 = 0.f, d1 = 0.f;
    // every load of the row is issued before the first use (independent L2 round trips, not a latency chain)
    const int rm = rows[m];
    const uint32_t act_u = *reinterpret_cast<const uint32_t*>(t_act + (size_t)m * HF_H + c0);
    const uint32_t tp_u = *reinterpret_cast<const uint32_t*>(t_pre + (size_t)m * HF_H + c0);
    const float mu = mean[m], rs = rstd[m];
    for (int s0 = 0; s0 < nsplit; s0 += 16) {   // 16 independent loads in flight per lane
      float2 v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        v[k] = s0 + k < nsplit ? *reinterpret_cast<const float2*>(dt_part + (s0 + k) * split_stride + (size_t)m * HF_H + c0) : make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 16; ++k) { d0 += v[k].x; d1 += v[k].y; }
    }
    const uint32_t x_u = *reinterpret_cast<const uint32_t*>(x + (size_t)rm * HF_H + c0);
    const float2 act = unpack_bf162(act_u), tp = unpack_bf162(tp_u), xv = unpack_bf162(x_u);
    const float xh0 = (act.x - mu) * rs, xh1 = (act.y - mu) * rs;
    a_g0 += d0 * xh0; a_g1 += d1 * xh1; a_b0 += d0; a_b1 += d1;
    const float dxh0 = d0 * g0, dxh1 = d1 * g1;
    const float s1 = warp_sum(dxh0 + dxh1) * (1.0f / HF_H);
    const float s2 = warp_sum(dxh0 * xh0 + dxh1 * xh1) * (1.0f / HF_H);
    float e0 = rs * (dxh0 - s1 - xh0 * s2) * gelu_erf_grad(tp.x);
    float e1 = rs * (dxh1 - s1 - xh1 * s2) * gelu_erf_grad(tp.y);
    const float2 er = unpack_bf162(pack_bf162(e0, e1));   // the bf16 value the layered path stores and re-reads
    e0 = er.x; e1 = er.y;
    a_c0 += e0; a_c1 += e1;
    float dx0 = 0.f, dx1 = 0.f;
#pragma unroll
    for (int o = 0; o < HF_H; o += 2) {
      const float eo0 = __shfl_sync(0xffffffffu, e0, o >> 1), eo1 = __shfl_sync(0xffffffffu, e1, o >> 1);
      dx0 += eo0 * sW[c0][o] + eo1 * sW[c0][o + 1];
      dx1 += eo0 * sW[c0 + 1][o] + eo1 * sW[c0 + 1][o + 1];
      acc0[o] += xv.x * eo0; acc0[o + 1] += xv.x * eo1;
      acc1[o] += xv.y * eo0; acc1[o + 1] += xv.y * eo1;
    }
    *reinterpret_cast<float2*>(dx_out + (size_t)rm * HF_H + c0) = make_float2(dx0, dx1);
  }
  // CTA reduction: every warp parks its accumulators in its own shared-memory slab, then all threads sum the slabs of the warps
  // that had rows, in warp order (fixed order -> deterministic)
  const bool had_rows = blockIdx.x * 8 + warp < M;
  if (had_rows) {
    float* slab = sAccW + warp * HF_H * HF_LD;
#pragma unroll
    for (int o = 0; o < HF_H; ++o) { slab[o * HF_LD + c0] = acc0[o]; slab[o * HF_LD + c0 + 1] = acc1[o]; }
    float* ln = sLnW + warp * 3 * HF_H;
    ln[c0] = a_g0; ln[c0 + 1] = a_g1; ln[HF_H + c0] = a_b0; ln[HF_H + c0 + 1] = a_b1; ln[2 * HF_H + c0] = a_c0; ln[2 * HF_H + c0 + 1] = a_c1;
  }
  __syncthreads();
  int nwarps = M - (int)blockIdx.x * 8;
  nwarps = nwarps < 0 ? 0 : (nwarps > 8 ? 8 : nwarps);
  for (int i = tid; i < HF_H * HF_H; i += 256) {   // i = in * 64 + out; slab index [out][in]
    const int in = i / HF_H, out = i % HF_H;
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += sAccW[w * HF_H * HF_LD + out * HF_LD + in];
    p_wt[(size_t)blockIdx.x * HF_H * HF_H + i] = v;
  }
  if (tid < 3 * HF_H) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += sLnW[w * 3 * HF_H + tid];
    p_ln[(size_t)blockIdx.x * 3 * HF_H + tid] = v;
  }
}

int head_bwd_fused_ctas() { return 148; }
bool head_bwd_fused_supported(int H) { return H == HF_H; }

cudaError_t launch_head_bwd_fused(const float* dt_part, int nsplit, size_t split_stride, const bf16* t_pre, const bf16* t_act,
                                  const float* mean, const float* rstd, const float* gamma, const bf16* wt, const bf16* x,
                                  const int* rows, const int* d_counts, int M_cap, float* dx_out, float* p_ln, float* p_wt,
                                  cudaStream_t st, int dyn_vtiles, int dyn_target, int dyn_max) {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(head_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM);
    done = true;
  }
  return launch_pdl(head_bwd_fused_kernel, dim3(head_bwd_fused_ctas()), dim3(256), (size_t)HF_SMEM, st, dt_part, nsplit, split_stride, t_pre, t_act,
                    mean, rstd, gamma, wt, x, rows, d_counts, M_cap, dx_out, p_ln, p_wt, dyn_vtiles, dyn_target, dyn_max);
}

}  // namespace b4r
