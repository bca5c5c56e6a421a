// Candidate scoring + ranking for the evaluator: fused gather-dot over the candidate ids, warp-level stable
// descending rank (lower candidate index first on ties = tf.argsort(direction="DESCENDING") / TopKV2 semantics),
// 1-based rank of the ground truth, rank histogram -> HR@k / NDCG@k / MAP.
// Replaces BERT4RecModel.rank_items' per-position tf.gather + tf.argsort + tf.gather (bert4rec_model.py:224-239)
// and the rank lookup + metric updates of BERT4RecEvaluator.evaluate_batch (bert4rec_evaluator.py:112-120,
// evaluation_metrics.py:47-112); SURVEY.md 2b row K13.
#include "common.cuh"
#include "kernels.h"

namespace b4r {

// one warp per slot; dynamic smem per warp: C floats (scores) + C int64 (ids)
template <int H>
__global__ void __launch_bounds__(128) rank_candidates_kernel(const bf16* __restrict__ t, int ldt, const bf16* __restrict__ E,
                                                              const float* __restrict__ vbias, const int64_t* __restrict__ cand,
                                                              const int64_t* __restrict__ gt, int M, int C, int V,
                                                              const int* __restrict__ d_counts,
                                                              int64_t* __restrict__ ranking, float* __restrict__ scores,
                                                              int* __restrict__ rank, unsigned long long* __restrict__ hist) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  if (d_counts && m >= d_counts[0]) {   // no selected slot behind this row (caller passed more candidate lists than slots): rank 0
    if (lane == 0 && rank) rank[m] = 0;
    return;
  }
  long long* s_id = reinterpret_cast<long long*>(smem_raw) + (size_t)warp * C;
  float* s_sc = reinterpret_cast<float*>(reinterpret_cast<long long*>(smem_raw) + (size_t)4 * C) + (size_t)warp * C;
  // the slot's hidden row, fp32, in shared memory (read as a broadcast by every lane)
  float* s_t = reinterpret_cast<float*>(reinterpret_cast<long long*>(smem_raw) + (size_t)4 * C) + (size_t)4 * C + (size_t)warp * H;
  for (int k = lane * 2; k < H; k += 64) {
    const float2 f = unpack_bf162(*reinterpret_cast<const uint32_t*>(t + (size_t)m * ldt + k));
    s_t[k] = f.x; s_t[k + 1] = f.y;
  }
  __syncwarp();
  // one candidate per lane: its whole table row is fetched with independent 16-byte loads (no cross-lane reduction, all
  // gathers of the slot in flight at once)
  for (int c = lane; c < C; c += 32) {
    const long long id = cand[(size_t)m * C + c];
    if (id < 0 || id >= V) {   // not an item of this model's catalogue (e.g. a sampler built over a larger vocabulary): ranked last
      s_sc[c] = -INFINITY;
      s_id[c] = id;
      continue;
    }
    const uint4* row = reinterpret_cast<const uint4*>(E + (size_t)id * H);
    float a4[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains; summed in a fixed order
#pragma unroll
    for (int j = 0; j < H / 8; ++j) {
      const uint4 v = __ldg(row + j);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf162(w[i]);
        a4[i] = fmaf(s_t[j * 8 + 2 * i], f.x, a4[i]);
        a4[i] = fmaf(s_t[j * 8 + 2 * i + 1], f.y, a4[i]);
      }
    }
    s_sc[c] = ((a4[0] + a4[1]) + (a4[2] + a4[3])) + vbias[id];
    s_id[c] = id;
  }
  __syncwarp();
  const long long g = gt ? gt[m] : -1;
  int gt_rank = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float sc = s_sc[c];
    int r = 0;
    for (int k = 0; k < C; ++k) {
      const float sk = s_sc[k];
      r += (sk > sc) || (sk == sc && k < c);
    }
    if (ranking) ranking[(size_t)m * C + r] = s_id[c];
    if (scores) scores[(size_t)m * C + c] = sc;
    if (s_id[c] == g) gt_rank = min(gt_rank, r + 1);  // first occurrence in the ranking (np.where(...)[0][0])
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gt_rank = min(gt_rank, __shfl_xor_sync(0xffffffffu, gt_rank, o));
  if (lane == 0) {
    const int rr = gt_rank == 0x7fffffff ? 0 : gt_rank;  // 0 = ground truth not among the candidates
    if (rank) rank[m] = rr;
    if (hist && rr > 0) atomicAdd(hist + rr, 1ull);
  }
}

cudaError_t launch_rank_candidates(const bf16* t, int ldt, const bf16* E, const float* vbias, const int64_t* cand,
                                   const int64_t* gt, int M, int C, int H, int V, const int* d_counts, int64_t* ranking,
                                   float* scores, int* rank, unsigned long long* hist, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  size_t smem = (size_t)4 * C * (sizeof(long long) + sizeof(float)) + (size_t)4 * H * sizeof(float);
  int grid = (M + 3) / 4;
#define B4R_RK(HH)                                                                                            \
  case HH:                                                                                                    \
    { static size_t cap_##HH = 0; if (smem > cap_##HH) { cudaFuncSetAttribute(rank_candidates_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cap_##HH = smem; } } \
    launch_pdl(rank_candidates_kernel<HH>, dim3(grid), dim3(128), smem, st, t, ldt, E, vbias, cand, gt, M, C, V, d_counts, ranking, scores, rank, hist); \
    break;
  switch (H) {
    B4R_RK(64)
    B4R_RK(128)
    B4R_RK(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_RK
  return cudaGetLastError();
}

// out = {n, NDCG@k[0..nk), HR@k[0..nk), MAP} from hist[1..max_rank]; fp64, fixed summation order.
__global__ void __launch_bounds__(1024) metrics_from_hist_kernel(const unsigned long long* __restrict__ hist, int max_rank,
                                                                 const int* __restrict__ ks, int nk, double* __restrict__ out) {
  __shared__ double s[1024];
  const int tid = threadIdx.x;
  const int nout = 2 + 2 * nk;
  for (int o = 0; o < nout; ++o) {
    double acc = 0.0;
    for (int r = 1 + tid; r <= max_rank; r += 1024) {
      const double h = (double)hist[r];
      if (h == 0.0) continue;
      if (o == 0) acc += h;
      else if (o <= nk) { if (r <= ks[o - 1]) acc += h * (r == 1 ? 1.0 : 1.0 / log2((double)r + 1.0)); }
      else if (o <= 2 * nk) { if (r <= ks[o - 1 - nk]) acc += h; }
      else acc += h / (double)r;
    }
    s[tid] = acc;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
      if (tid < w) s[tid] += s[tid + w];
      __syncthreads();
    }
    if (tid == 0) out[o] = s[0];
    __syncthreads();
  }
  if (tid == 0) {
    const double n = out[0] > 0 ? out[0] : 1.0;
    for (int o = 1; o < nout; ++o) out[o] /= n;
  }
}

cudaError_t launch_metrics_from_hist(const unsigned long long* hist, int max_rank, const int* ks, int nk, double* out,
                                     cudaStream_t st) {
  metrics_from_hist_kernel<<<1, 1024, 0, st>>>(hist, max_rank, ks, nk, out);
  return cudaGetLastError();
}

}  // namespace b4r
