// Vocabulary-sharded tied projection (large catalogues): the small row-side kernels around the tcgen05 CE passes.
//
// Every rank owns the catalogue rows [v_begin, v_end) of the tied table for the projection.  The masked-slot rows of ALL ranks
// are all-gathered ([n][M_cap] padded layout), packed valid-rows-first (the CE kernels work on [0, n_valid) / [0, n_rows)),
// scored against the local shard, the per-row partials (max, sum-exp, label logit, best value, best index) are exchanged and
// merged, the backward produces the shard's table gradient (complete: it saw every row) and a PARTIAL dT for every row that a
// reduce-scatter sums back to the owning rank.
#include "kernels.h"

namespace b4r {

// prefix of the compact layout: valid rows of rank 0, 1, ... then the zero-weight aux rows of rank 0, 1, ...
__device__ __forceinline__ int shard_dst_row(const int* __restrict__ counts, int n, int r, int i) {
  int tot_valid = 0, pv = 0, pa = 0;
  for (int q = 0; q < n; ++q) {
    const int nv = counts[q * 2], nr = counts[q * 2 + 1];
    if (q < r) { pv += nv; pa += nr - nv; }
    tot_valid += nv;
  }
  const int nv = counts[r * 2], nr = counts[r * 2 + 1];
  if (i < nv) return pv + i;
  if (i < nr) return tot_valid + pa + (i - nv);
  return -1;
}

// one warp per gathered row
__global__ void __launch_bounds__(256) shard_pack_kernel(const bf16* __restrict__ rows_in, const int* __restrict__ labels_in,
                                                         const float* __restrict__ w_in, const int* __restrict__ mult_in,
                                                         const int* __restrict__ counts_in, int n, int M_cap, int H, int v_begin,
                                                         bf16* __restrict__ rows, int* __restrict__ lab_local, int* __restrict__ lab_global,
                                                         float* __restrict__ w, int* __restrict__ mult, int* __restrict__ counts) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g == 0 && lane == 0) {
    int tv = 0, tr = 0;
    for (int q = 0; q < n; ++q) { tv += counts_in[q * 2]; tr += counts_in[q * 2 + 1]; }
    counts[0] = tv; counts[1] = tr;
  }
  if (g >= n * M_cap) return;
  const int r = g / M_cap, i = g % M_cap;
  const int dst = shard_dst_row(counts_in, n, r, i);
  if (dst < 0) return;
  const uint4* src4 = reinterpret_cast<const uint4*>(rows_in + (size_t)g * H);
  uint4* dst4 = reinterpret_cast<uint4*>(rows + (size_t)dst * H);
  for (int c = lane; c < H / 8; c += 32) dst4[c] = src4[c];
  if (lane == 0) {
    const int l = labels_in[g];
    lab_global[dst] = l; lab_local[dst] = l - v_begin;   // a label outside the shard never matches a local column
    w[dst] = w_in[g]; mult[dst] = mult_in[g];
  }
}

// merge the device-chosen number of vocabulary splits of the forward into ONE partial per row; the best index becomes global
__global__ void __launch_bounds__(256) shard_part_merge_kernel(const float* __restrict__ part, const int* __restrict__ counts, int M_cap,
                                                               int ntiles, int target_ctas, int max_splits, int v_begin,
                                                               float* __restrict__ out) {
  const int n_rows = min(M_cap, counts[1]);
  const int vs = ce_dyn_splits128(n_rows, ntiles, target_ctas, max_splits);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M_cap) return;
  float mn = -INFINITY, l = 0.f, lab = -INFINITY, bv = -INFINITY;
  int bi = 0x7fffffff;
  if (r < n_rows) {
    for (int s = 0; s < vs; ++s) {
      const float2* p = reinterpret_cast<const float2*>(part + ((size_t)s * M_cap + r) * 6);
      const float2 p01 = p[0], p23 = p[1], p45 = p[2];
      const float m3 = fmaxf(mn, p01.x);
      if (m3 != -INFINITY) l = l * __expf(mn - m3) + p01.y * __expf(p01.x - m3);
      mn = m3;
      lab = fmaxf(lab, p23.x);
      const int bi2 = __float_as_int(p45.x);
      if (p23.y > bv || (p23.y == bv && bi2 < bi)) { bv = p23.y; bi = bi2; }
    }
    if (bi != 0x7fffffff) bi += v_begin;
  }
  float2* o = reinterpret_cast<float2*>(out + (size_t)r * 6);
  o[0] = make_float2(mn, l); o[1] = make_float2(lab, bv); o[2] = make_float2(__int_as_float(bi), 0.f);
}

// dT partials of the shard ([vs][cap][H], vs chosen on the device by the dT pass) -> summed, in the padded [n][M_cap][H] layout
// the reduce-scatter wants (rows without a gradient are zero).  One thread per 4 columns.
__global__ void __launch_bounds__(256) shard_dt_unpack_kernel(const float* __restrict__ dt_part, const int* __restrict__ counts,
                                                              const int* __restrict__ counts_in, int n, int M_cap, int cap, int H,
                                                              int xtiles, int target_ctas, int max_splits, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int hq = H / 4;
  if (idx >= (long long)n * M_cap * hq) return;
  const int g = (int)(idx / hq), c = (int)(idx % hq) * 4;
  const int r = g / M_cap, i = g % M_cap;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < counts_in[r * 2]) {
    const int src = shard_dst_row(counts_in, n, r, i);
    const int n_valid = min(cap, counts[0]);
    const int vs = ce_dyn_splits128(n_valid, xtiles, target_ctas, max_splits);
    for (int s = 0; s < vs; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(dt_part + ((size_t)s * cap + src) * H + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  *reinterpret_cast<float4*>(out + (size_t)g * H + c) = acc;
}

cudaError_t launch_shard_pack(const bf16* rows_in, const int* labels_in, const float* w_in, const int* mult_in, const int* counts_in,
                              int n, int M_cap, int H, int v_begin, bf16* rows, int* lab_local, int* lab_global, float* w, int* mult,
                              int* counts, cudaStream_t st) {
  const long long threads = (long long)n * M_cap * 32;
  shard_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rows_in, labels_in, w_in, mult_in, counts_in, n, M_cap, H, v_begin,
                                                                     rows, lab_local, lab_global, w, mult, counts);
  return cudaGetLastError();
}

cudaError_t launch_shard_part_merge(const float* part, const int* counts, int M_cap, int ntiles, int target_ctas, int max_splits,
                                    int v_begin, float* out, cudaStream_t st) {
  shard_part_merge_kernel<<<(M_cap + 255) / 256, 256, 0, st>>>(part, counts, M_cap, ntiles, target_ctas, max_splits, v_begin, out);
  return cudaGetLastError();
}

cudaError_t launch_shard_dt_unpack(const float* dt_part, const int* counts, const int* counts_in, int n, int M_cap, int cap, int H,
                                   int xtiles, int target_ctas, int max_splits, float* out, cudaStream_t st) {
  const long long threads = (long long)n * M_cap * (H / 4);
  shard_dt_unpack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dt_part, counts, counts_in, n, M_cap, cap, H, xtiles,
                                                                          target_ctas, max_splits, out);
  return cudaGetLastError();
}

}  // namespace b4r
