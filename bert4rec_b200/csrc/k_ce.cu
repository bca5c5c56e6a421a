// MLM head: valid-slot compaction and the tied item-vocabulary projection fused with an online-max softmax
// cross-entropy (+ running arg-max for the accuracy metrics).  The [B,P,V] logits tensor of the reference
// (tfm MaskedLM, bert4rec_model.py:76-81,143; MaskedSparseCategoricalCrossentropy / masked_accuracy,
// trainer_utils.py:12-23,49-60) is never materialised in forward.
// Generation 1: mma.sync tiles, A tile (64 rows of t) resident in smem, E streamed in 128-row tiles.
#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace b4r {

// ------------------------------------------------------------------------------------------------ compaction
// Single CTA, 4 consecutive slots per thread per pass.  Order of the compacted rows = row-major (b, p) order of the
// valid slots (= tf.boolean_mask order).
__global__ void __launch_bounds__(1024) mlm_select_kernel(const int64_t* __restrict__ positions, const int64_t* __restrict__ ids,
                                                          const int64_t* __restrict__ weights, int use_weights, int B, int S,
                                                          int P, int want_aux, int* __restrict__ rows, int* __restrict__ labels,
                                                          float* __restrict__ row_w, int* __restrict__ row_mult,
                                                          int* __restrict__ counts) {
  pdl_grid_sync();
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  const int n = B * P;
  // block exclusive scan of one int per thread; returns the exclusive prefix, *total = block total
  auto block_scan = [&](int x, int* total) -> int {
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int excl = (warp > 0 ? s_warp[warp - 1] : 0) + incl - x;
    *total = s_warp[31];
    __syncthreads();
    return excl;
  };
  for (int base = 0; base < n; base += 4096) {
    const int i0 = base + tid * 4;
    long long idv[4];
    int val[4], cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      const bool ok = i < n;
      idv[k] = ok ? ids[i] : 0;
      val[k] = !ok ? 0 : (use_weights == 2 ? 1 : (use_weights ? (weights[i] != 0) : (idv[k] != 0)));
      cnt += val[k];
    }
    int total;
    const int carry = s_carry;   // read before the scan's barriers; thread 0 updates it after them
    int excl = carry + block_scan(cnt, &total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (val[k]) {
        const int i = i0 + k;
        long long pos = positions[i];
        pos = pos < 0 ? 0 : (pos >= S ? S - 1 : pos);
        rows[excl] = (i / P) * S + (int)pos;
        labels[excl] = (int)idv[k];
        row_w[excl] = idv[k] != 0 ? 1.f : 0.f;
        row_mult[excl] = 1;
        ++excl;
      }
    }
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  const int n_valid = s_carry;
  if (want_aux) {
    // one aux row per sequence with padded slots: logits of (b, position 0) against label 0, for
    // SparseCategoricalAccuracy over ALL slots (bert4rec_trainer.py:30).  One thread per sequence, independent loads.
    for (int base = 0; base < B; base += 1024) {
      int npad = 0;
      if (base + tid < B) {
        const int64_t* src = (use_weights ? weights : ids) + (size_t)(base + tid) * P;
        int nv = 0;
        if (use_weights == 2) nv = P;
        else {
#pragma unroll 8
          for (int p = 0; p < P; ++p) nv += src[p] != 0;
        }
        npad = P - nv;
      }
      const int b = base + tid;
      const int has = (b < B && npad > 0) ? 1 : 0;
      int total;
      const int carry = s_carry;
      const int excl = carry + block_scan(has, &total);
      if (has) {
        rows[excl] = b * S;
        labels[excl] = 0;
        row_w[excl] = 0.f;
        row_mult[excl] = npad;
      }
      if (tid == 0) s_carry += total;
      __syncthreads();
    }
  }
  if (tid == 0) { counts[0] = n_valid; counts[1] = s_carry; }
}

cudaError_t launch_mlm_select(const int64_t* positions, const int64_t* ids, const int64_t* weights, int use_weights,
                              int B, int S, int P, int want_aux, int* rows, int* labels, float* row_w, int* row_mult,
                              int* counts, cudaStream_t st) {
  launch_pdl(mlm_select_kernel, dim3(1), dim3(1024), (size_t)0, st, positions, ids, weights, use_weights, B, S, P, want_aux, rows, labels, row_w,
             row_mult, counts);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fused linear + CE
constexpr int CE_BM = 64, CE_BN = 128;
int ce_block_m() { return CE_BM; }

struct CeDev {
  const bf16* t; int ldt; const bf16* E; const float* vbias;
  const int* labels; const float* row_w; const int* row_mult; const int* d_counts;
  int M_cap, V, v_begin, v_end, vsplits, batch, target_ctas, max_splits;
  float* part; float* lse; float* lab_out; float* stats; float* step_stats;
  bf16* dlogits; int ld_dl; int row_begin, row_count;
};

template <int H>
__device__ __forceinline__ void ce_load_b_tile(const CeDev& a, int v0, bf16* sB, float* sBias) {
  constexpr int LD = H + 8, CH = H / 8;
  for (int c = threadIdx.x; c < CE_BN * CH; c += 256) {
    const int r = c / CH, cc = (c % CH) * 8;
    const bool ok = v0 + r < a.v_end;
    cp_async16(sB + r * LD + cc, a.E + (size_t)(ok ? v0 + r : 0) * H + cc, ok);
  }
  if (threadIdx.x < CE_BN) {
    const int v = v0 + threadIdx.x;
    sBias[threadIdx.x] = v < a.v_end ? a.vbias[v] : 0.f;
  }
}

template <int H>
__device__ __forceinline__ void ce_tile_mma(const bf16* sA, const bf16* sB, int warp_m, int warp_n, int lane, float (&acc)[8][4]) {
  constexpr int LD = H + 8;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < H; kk += 16) {
    uint32_t af[4];
    load_a_frag<false>(af, sA, LD, warp_m * 16, kk, lane);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      load_b_frag<false>(bb, sB, LD, warp_n * 64 + np * 16, kk, lane);
      mma_bf16(acc[2 * np], af, bb[0], bb[1]);
      mma_bf16(acc[2 * np + 1], af, bb[2], bb[3]);
    }
  }
}

// grid = (vsplits, ceil(M_cap/64)); 256 threads = 4 (M) x 2 (N) warps, warp tile 16 x 64.
template <int H>
__global__ void __launch_bounds__(256) ce_fwd_kernel(CeDev a) {
  pdl_grid_wait();
  constexpr int LD = H + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB0 = sA + CE_BM * LD;
  bf16* sB1 = sB0 + CE_BN * LD;
  float* sBias0 = reinterpret_cast<float*>(sB1 + CE_BN * LD);
  float* sBias1 = sBias0 + CE_BN;
  float* sMerge = sBias1 + CE_BN;  // [2][64][5]

  const int n_rows = min(a.M_cap, a.d_counts[1]);
  const int m0 = blockIdx.y * CE_BM;
  if (m0 >= n_rows) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  const int ntiles = (a.v_end - a.v_begin + CE_BN - 1) / CE_BN;
  const int tps = (ntiles + a.vsplits - 1) / a.vsplits;
  const int tile_lo = blockIdx.x * tps, tile_hi = min(ntiles, tile_lo + tps);

  constexpr int CH = H / 8;
  for (int c = tid; c < CE_BM * CH; c += 256) {
    const int r = c / CH, cc = (c % CH) * 8;
    const bool ok = m0 + r < n_rows;
    cp_async16(sA + r * LD + cc, a.t + (size_t)(ok ? m0 + r : 0) * a.ldt + cc, ok);
  }
  if (tile_lo < tile_hi) ce_load_b_tile<H>(a, a.v_begin + tile_lo * CE_BN, sB0, sBias0);
  cp_async_commit();

  int label[2];
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, lab_logit[2] = {-INFINITY, -INFINITY};
  float best_v[2] = {-INFINITY, -INFINITY};
  int best_i[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = m0 + warp_m * 16 + g + h * 8;
    label[h] = r < n_rows ? a.labels[r] : -1;
  }

  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int buf = (tile - tile_lo) & 1;
    bf16* sB = buf ? sB1 : sB0;
    float* sBias = buf ? sBias1 : sBias0;
    cp_async_wait<0>();
    __syncthreads();  // tile landed (and sBias stores visible); previous tile's buffer is free
    if (tile + 1 < tile_hi) ce_load_b_tile<H>(a, a.v_begin + (tile + 1) * CE_BN, buf ? sB0 : sB1, buf ? sBias0 : sBias1);
    cp_async_commit();
    float acc[8][4];
    ce_tile_mma<H>(sA, sB, warp_m, warp_n, lane, acc);
    const int v0 = a.v_begin + tile * CE_BN + warp_n * 64;
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = warp_n * 64 + nt * 8 + 2 * t4 + (e & 1);
        const int col = v0 + nt * 8 + 2 * t4 + (e & 1);
        const int h = e >> 1;
        float v = col < a.v_end ? acc[nt][e] + sBias[cl] : -INFINITY;
        acc[nt][e] = v;
        tmax[h] = fmaxf(tmax[h], v);
        if (v > best_v[h]) { best_v[h] = v; best_i[h] = col; }
        if (col == label[h]) lab_logit[h] = v;
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float m_new = fmaxf(m_run[h], tmax[h]);
      if (m_new != -INFINITY) {
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) sum += __expf(acc[nt][h * 2] - m_new) + __expf(acc[nt][h * 2 + 1] - m_new);
        l_run[h] = l_run[h] * __expf(m_run[h] - m_new) + sum;  // exp(-inf - finite) = 0
        m_run[h] = m_new;
      }
    }
  }
  cp_async_wait<0>();
  // merge the 4 lanes of a quad (same rows, different columns)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m_run[h], o), l2 = __shfl_xor_sync(0xffffffffu, l_run[h], o);
      const float bv2 = __shfl_xor_sync(0xffffffffu, best_v[h], o);
      const int bi2 = __shfl_xor_sync(0xffffffffu, best_i[h], o);
      const float lb2 = __shfl_xor_sync(0xffffffffu, lab_logit[h], o);
      const float mn = fmaxf(m_run[h], m2);
      if (mn != -INFINITY) l_run[h] = l_run[h] * __expf(m_run[h] - mn) + l2 * __expf(m2 - mn);
      m_run[h] = mn;
      if (bv2 > best_v[h] || (bv2 == best_v[h] && bi2 < best_i[h])) { best_v[h] = bv2; best_i[h] = bi2; }
      lab_logit[h] = fmaxf(lab_logit[h], lb2);
    }
  }
  __syncthreads();
  if (t4 == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* d = sMerge + ((warp_n * 64) + warp_m * 16 + g + h * 8) * 5;
      d[0] = m_run[h]; d[1] = l_run[h]; d[2] = lab_logit[h]; d[3] = best_v[h]; d[4] = __int_as_float(best_i[h]);
    }
  }
  __syncthreads();
  if (tid < CE_BM && m0 + tid < n_rows) {
    const float* p0 = sMerge + tid * 5;
    const float* p1 = sMerge + (64 + tid) * 5;
    const float mn = fmaxf(p0[0], p1[0]);
    float l = 0.f;
    if (mn != -INFINITY) l = p0[1] * __expf(p0[0] - mn) + p1[1] * __expf(p1[0] - mn);
    float bv = p0[3]; int bi = __float_as_int(p0[4]);
    const int bi1 = __float_as_int(p1[4]);
    if (p1[3] > bv || (p1[3] == bv && bi1 < bi)) { bv = p1[3]; bi = bi1; }
    float* out = a.part + ((size_t)blockIdx.x * a.M_cap + m0 + tid) * 6;
    out[0] = mn; out[1] = l; out[2] = fmaxf(p0[2], p1[2]); out[3] = bv; out[4] = __int_as_float(bi); out[5] = 0.f;
  }
}

// Multi-CTA: merge the vsplits partials per row, emit lse; per-CTA partial statistics; the LAST CTA to finish (atomic
// ticket) sums them in CTA-index order (deterministic) and accumulates the step / running statistics.
__global__ void __launch_bounds__(256) ce_finalize_kernel(CeDev a, float* __restrict__ fin_part, int* __restrict__ ticket) {
  pdl_grid_sync();
  __shared__ float s_red[8][5];
  __shared__ int s_last;
  const int n_rows = min(a.M_cap, a.d_counts[1]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.vsplits < 0) {  // generation-2 forward: same device-side split formula as ce_fwd_umma_kernel (128-row tiles)
    const int ntiles = (a.v_end - a.v_begin + 127) / 128;
    a.vsplits = ce_dyn_splits128(n_rows, ntiles, a.target_ctas, a.max_splits);
  }
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // loss_sum, n_valid, correct_masked, correct_all, n_all
  // 8 lanes per row: lane q merges splits q, q+8, ...; the 8 lanes are then merged by an xor butterfly (fixed order)
  const int q = lane & 7, rsub = lane >> 3;
  for (int r0 = blockIdx.x * 32; r0 < n_rows; r0 += gridDim.x * 32) {
    const int r = r0 + warp * 4 + rsub;
    float mn = -INFINITY, l = 0.f, lab = -INFINITY, bv = -INFINITY;
    int bi = 0x7fffffff;
    // the row's weight / multiplicity / label do not depend on the merge: requested up front, one round trip instead of two
    float w = 0.f; int mult = 0, label_r = -1;
    if (q == 0 && r < n_rows) { w = a.row_w[r]; mult = a.row_mult[r]; label_r = a.labels[r]; }
    if (r < n_rows) {
      for (int s = q; s < a.vsplits; s += 8) {
        const float2* p = reinterpret_cast<const float2*>(a.part + ((size_t)s * a.M_cap + r) * 6);
        const float2 p01 = p[0], p23 = p[1], p45 = p[2];
        const float m3 = fmaxf(mn, p01.x);
        if (m3 != -INFINITY) l = l * __expf(mn - m3) + p01.y * __expf(p01.x - m3);
        mn = m3;
        lab = fmaxf(lab, p23.x);
        const int bi2 = __float_as_int(p45.x);
        if (p23.y > bv || (p23.y == bv && bi2 < bi)) { bv = p23.y; bi = bi2; }
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, mn, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
      const float lab2 = __shfl_xor_sync(0xffffffffu, lab, o), bv2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int bi2 = __shfl_xor_sync(0xffffffffu, bi, o);
      const float m3 = fmaxf(mn, m2);
      if (m3 != -INFINITY) l = l * __expf(mn - m3) + l2 * __expf(m2 - m3);
      mn = m3;
      lab = fmaxf(lab, lab2);
      if (bv2 > bv || (bv2 == bv && bi2 < bi)) { bv = bv2; bi = bi2; }
    }
    if (q == 0 && r < n_rows) {
      const float lse = mn + logf(l);
      a.lse[r] = lse;
      if (a.lab_out) a.lab_out[r] = lab;
      const int correct = (bi == label_r);
      if (w > 0.f) acc[0] += (lse - lab);
      acc[1] += w;
      acc[2] += (w > 0.f && correct) ? 1.f : 0.f;
      acc[3] += correct ? (float)mult : 0.f;
      acc[4] += (float)mult;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float v = warp_sum(acc[k]);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  __shared__ float s_tot[5];
  if (tid < 5) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_red[w][tid];
    fin_part[blockIdx.x * 5 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    // fixed-order sum of the per-CTA statistics: thread t takes CTAs t, t + 256, ...; the 256 partial sums are added in index order
    __shared__ float s_fin[256][5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float v = 0.f;
      for (unsigned b = tid; b < gridDim.x; b += 256) v += __ldcg(fin_part + b * 5 + k);
      s_fin[tid][k] = v;
    }
    __syncthreads();
    if (tid < 5) {
      float v = 0.f;
      const int n = gridDim.x < 256 ? gridDim.x : 256;
      for (int t = 0; t < n; ++t) v += s_fin[t][tid];
      s_tot[tid] = v;
      a.step_stats[tid] = v;
      if (a.stats) a.stats[tid] += v;
    }
  }
  if (tid == 0) *ticket = 0;
  __syncthreads();
  if (tid == 0 && a.stats) {
    // Keras running means: loss = Mean(batch loss, weight = batch size); masked_accuracy = Mean over batches
    const float nv = fmaxf(s_tot[1], 1.f);
    a.stats[5] += (s_tot[0] / nv) * (float)a.batch;
    a.stats[6] += (float)a.batch;
    a.stats[7] += s_tot[2] / nv;
    a.stats[8] += 1.f;
  }
}

// dlogits tile: dl = (softmax - onehot) * w  (gradient of the SUM loss; the 1/n_valid normaliser is folded into the
// optimizer's gradient scale).  grid = (ceil(Vshard/128), ceil(row_count/64)).
template <int H>
__global__ void __launch_bounds__(256) ce_dlogits_kernel(CeDev a) {
  pdl_grid_wait();
  constexpr int LD = H + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB = sA + CE_BM * LD;
  float* sBias = reinterpret_cast<float*>(sB + CE_BN * LD);
  const int n_rows = min(a.M_cap, a.d_counts[1]);
  const int m0 = a.row_begin + blockIdx.y * CE_BM;
  const int row_end = min(n_rows, a.row_begin + a.row_count);
  if (m0 >= row_end) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;
  const int v0 = a.v_begin + blockIdx.x * CE_BN;
  constexpr int CH = H / 8;
  for (int c = tid; c < CE_BM * CH; c += 256) {
    const int r = c / CH, cc = (c % CH) * 8;
    const bool ok = m0 + r < row_end;
    cp_async16(sA + r * LD + cc, a.t + (size_t)(ok ? m0 + r : 0) * a.ldt + cc, ok);
  }
  ce_load_b_tile<H>(a, v0, sB, sBias);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  float acc[8][4];
  ce_tile_mma<H>(sA, sB, warp_m, warp_n, lane, acc);
  float lse[2], w[2];
  int label[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = m0 + warp_m * 16 + g + h * 8;
    const bool ok = r < row_end;
    lse[h] = ok ? a.lse[r] : 0.f;
    w[h] = ok ? a.row_w[r] : 0.f;
    label[h] = ok ? a.labels[r] : -1;
  }
  __syncthreads();  // all warps done reading sB -> reuse it as the per-warp staging area
  bf16* stage = sB + warp * (16 * 72);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float d[2];
#pragma unroll
      for (int e1 = 0; e1 < 2; ++e1) {
        const int cl = warp_n * 64 + nt * 8 + 2 * t4 + e1;
        const int col = v0 + cl;
        float v = 0.f;
        if (col < a.v_end && w[h] > 0.f) {
          v = __expf(acc[nt][h * 2 + e1] + sBias[cl] - lse[h]);
          if (col == label[h]) v -= 1.f;
          v *= w[h];
        }
        d[e1] = v;
      }
      *reinterpret_cast<uint32_t*>(stage + (g + h * 8) * 72 + nt * 8 + 2 * t4) = pack_bf162(d[0], d[1]);
    }
  __syncwarp();
  // coalesced write-out: 8 lanes x 16 B per row, 4 rows per pass
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = pass * 4 + (lane >> 3), c = (lane & 7) * 8;
    const int row = m0 + warp_m * 16 + r;
    if (row < row_end) {
      const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 72 + c);
      const size_t col = (size_t)(blockIdx.x * CE_BN + warp_n * 64 + c);
      *reinterpret_cast<uint4*>(a.dlogits + (size_t)(row - a.row_begin) * a.ld_dl + col) = v;
    }
  }
}

// Full-catalogue rank of the ground truth (rank_items(items=None) + evaluator rank lookup):
//   beat[m] += #{v in shard : s_v > s_gt[m]  or  (s_v == s_gt[m] and v < gt[m])} ;  rank = 1 + sum over shards.
// s_gt comes from the label-logit of ce_fwd (same MMA order -> identical bits for the gt column itself).
template <int H>
__global__ void __launch_bounds__(256) ce_count_kernel(CeDev a, const float* __restrict__ s_gt, int* __restrict__ beat) {
  pdl_grid_wait();
  constexpr int LD = H + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB0 = sA + CE_BM * LD;
  bf16* sB1 = sB0 + CE_BN * LD;
  float* sBias0 = reinterpret_cast<float*>(sB1 + CE_BN * LD);
  float* sBias1 = sBias0 + CE_BN;
  const int n_rows = min(a.M_cap, a.d_counts[1]);
  const int m0 = blockIdx.y * CE_BM;
  if (m0 >= n_rows) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;
  const int ntiles = (a.v_end - a.v_begin + CE_BN - 1) / CE_BN;
  const int tps = (ntiles + a.vsplits - 1) / a.vsplits;
  const int tile_lo = blockIdx.x * tps, tile_hi = min(ntiles, tile_lo + tps);
  constexpr int CH = H / 8;
  for (int c = tid; c < CE_BM * CH; c += 256) {
    const int r = c / CH, cc = (c % CH) * 8;
    const bool ok = m0 + r < n_rows;
    cp_async16(sA + r * LD + cc, a.t + (size_t)(ok ? m0 + r : 0) * a.ldt + cc, ok);
  }
  if (tile_lo < tile_hi) ce_load_b_tile<H>(a, a.v_begin + tile_lo * CE_BN, sB0, sBias0);
  cp_async_commit();
  int label[2], cnt[2] = {0, 0};
  float sg[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = m0 + warp_m * 16 + g + h * 8;
    label[h] = r < n_rows ? a.labels[r] : -1;
    sg[h] = r < n_rows ? s_gt[r] : INFINITY;
  }
  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int buf = (tile - tile_lo) & 1;
    bf16* sB = buf ? sB1 : sB0;
    float* sBias = buf ? sBias1 : sBias0;
    cp_async_wait<0>();
    __syncthreads();
    if (tile + 1 < tile_hi) ce_load_b_tile<H>(a, a.v_begin + (tile + 1) * CE_BN, buf ? sB0 : sB1, buf ? sBias0 : sBias1);
    cp_async_commit();
    float acc[8][4];
    ce_tile_mma<H>(sA, sB, warp_m, warp_n, lane, acc);
    const int v0 = a.v_begin + tile * CE_BN + warp_n * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = warp_n * 64 + nt * 8 + 2 * t4 + (e & 1);
        const int col = v0 + nt * 8 + 2 * t4 + (e & 1);
        const int h = e >> 1;
        if (col < a.v_end) {
          const float v = acc[nt][e] + sBias[cl];
          // the label column never competes with itself (its score here may differ from s_gt in the last bit)
          cnt[h] += col != label[h] && ((v > sg[h]) || (v == sg[h] && col < label[h]));
        }
      }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    cnt[h] += __shfl_xor_sync(0xffffffffu, cnt[h], 1);
    cnt[h] += __shfl_xor_sync(0xffffffffu, cnt[h], 2);
    const int r = m0 + warp_m * 16 + g + h * 8;
    if (t4 == 0 && r < n_rows && cnt[h]) atomicAdd(beat + r, cnt[h]);
  }
}

// ------------------------------------------------------------------------------------------------ full-catalogue top-k
// rank_items(items=None) for serving / full-catalogue evaluation (bert4rec_model.py:235-236: tf.argsort over the vocabulary,
// DESCENDING, stable = lower item id first among equal logits; apps/recommender.py:14-63): the K best items of every selected row
// over the vocabulary shard [v_begin, v_end), without materialising any logits.  A (score, id) pair is ONE unsigned 64-bit key,
//   key = order_preserving_bits(score) << 32 | (0xFFFFFFFF - id),
// so "larger key" == "ranks earlier" (higher score, then lower id) and every selection / merge below is a plain integer maximum:
// exact, order independent and therefore deterministic.
//   topk_partial_kernel: grid (vocabulary splits, 64-row tiles); the score tiles of ce_fwd_kernel (mma.sync, E streamed through a
//     double-buffered tile); each CTA keeps, per row, the K largest keys of its vocabulary range in shared memory (replace-the-minimum
//     under a per-row lock; after the first few tiles almost every score fails the racy threshold pre-check and costs one compare).
//   topk_merge_kernel: one CTA per row, bitonic sort of the nlists x K partial keys, first K -> sorted keys.  The same kernel merges
//     the per-rank lists after the NCCL all-gather of a vocabulary-sharded catalogue.
__host__ __device__ inline unsigned long long topk_key(float score, int id) {
#ifdef __CUDA_ARCH__
  unsigned u = __float_as_uint(score);
#else
  unsigned u; memcpy(&u, &score, 4);
#endif
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)id);
}
constexpr int TOPK_MAX = 128;

template <int H>
__global__ void __launch_bounds__(256) topk_partial_kernel(CeDev a, int K, int n_rows_static, unsigned long long* __restrict__ part) {
  pdl_grid_wait();
  constexpr int LD = H + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB0 = sA + CE_BM * LD;
  bf16* sB1 = sB0 + CE_BN * LD;
  float* sBias0 = reinterpret_cast<float*>(sB1 + CE_BN * LD);
  float* sBias1 = sBias0 + CE_BN;
  unsigned long long* sList = reinterpret_cast<unsigned long long*>(sBias1 + CE_BN);   // [64][K]
  unsigned long long* sThr = sList + (size_t)CE_BM * K;                                  // [64] current minimum of a full list (0 until full)
  int* sCnt = reinterpret_cast<int*>(sThr + CE_BM);                                      // [64]
  int* sMinPos = sCnt + CE_BM;                                                           // [64]
  int* sLock = sMinPos + CE_BM;                                                          // [64]
  const int n_rows = a.d_counts ? min(n_rows_static, a.d_counts[0]) : n_rows_static;
  const int m0 = blockIdx.y * CE_BM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = warp & 3, warp_n = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;
  const int ntiles = (a.v_end - a.v_begin + CE_BN - 1) / CE_BN;
  const int tps = (ntiles + a.vsplits - 1) / a.vsplits;
  const int tile_lo = blockIdx.x * tps, tile_hi = min(ntiles, tile_lo + tps);
  unsigned long long* out = part + ((size_t)blockIdx.x * n_rows_static + m0) * K;        // [split][row][K]
  if (m0 >= n_rows) return;
  for (int i = tid; i < CE_BM; i += 256) { sThr[i] = 0ull; sCnt[i] = 0; sMinPos[i] = 0; sLock[i] = 0; }
  constexpr int CH = H / 8;
  for (int c = tid; c < CE_BM * CH; c += 256) {
    const int r = c / CH, cc = (c % CH) * 8;
    const bool ok = m0 + r < n_rows;
    cp_async16(sA + r * LD + cc, a.t + (size_t)(ok ? m0 + r : 0) * a.ldt + cc, ok);
  }
  if (tile_lo < tile_hi) ce_load_b_tile<H>(a, a.v_begin + tile_lo * CE_BN, sB0, sBias0);
  cp_async_commit();
  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int buf = (tile - tile_lo) & 1;
    bf16* sB = buf ? sB1 : sB0;
    float* sBias = buf ? sBias1 : sBias0;
    cp_async_wait<0>();
    __syncthreads();
    if (tile + 1 < tile_hi) ce_load_b_tile<H>(a, a.v_begin + (tile + 1) * CE_BN, buf ? sB0 : sB1, buf ? sBias0 : sBias1);
    cp_async_commit();
    float acc[8][4];
    ce_tile_mma<H>(sA, sB, warp_m, warp_n, lane, acc);
    const int v0 = a.v_begin + tile * CE_BN + warp_n * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = warp_n * 64 + nt * 8 + 2 * t4 + (e & 1);
        const int col = v0 + nt * 8 + 2 * t4 + (e & 1);
        const int rl = warp_m * 16 + g + (e >> 1) * 8;
        if (col < a.v_end && m0 + rl < n_rows) {
          const unsigned long long key = topk_key(acc[nt][e] + sBias[cl], col);
          if (key > *reinterpret_cast<volatile unsigned long long*>(sThr + rl)) {
            bool done = false;
            while (!done) {                       // (lock taken and released inside one iteration: lanes of a warp may share a row)
              if (atomicCAS(sLock + rl, 0, 1) == 0) {
                __threadfence_block();
                volatile unsigned long long* lst = sList + (size_t)rl * K;
                const int cnt = *reinterpret_cast<volatile int*>(sCnt + rl);
                bool rescan = false;
                if (cnt < K) {
                  lst[cnt] = key;
                  *reinterpret_cast<volatile int*>(sCnt + rl) = cnt + 1;
                  rescan = cnt + 1 == K;
                } else if (key > *reinterpret_cast<volatile unsigned long long*>(sThr + rl)) {
                  lst[*reinterpret_cast<volatile int*>(sMinPos + rl)] = key;
                  rescan = true;
                }
                if (rescan) {
                  unsigned long long mn = lst[0]; int mp = 0;
                  for (int i = 1; i < K; ++i) { const unsigned long long x = lst[i]; if (x < mn) { mn = x; mp = i; } }
                  *reinterpret_cast<volatile int*>(sMinPos + rl) = mp;
                  *reinterpret_cast<volatile unsigned long long*>(sThr + rl) = mn;
                }
                __threadfence_block();
                atomicExch(sLock + rl, 0);
                done = true;
              }
            }
          }
        }
      }
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int i = tid; i < CE_BM * K; i += 256) {
    const int r = i / K, j = i % K;
    if (m0 + r < n_rows) out[(size_t)r * K + j] = j < sCnt[r] ? sList[(size_t)r * K + j] : 0ull;
  }
}

// keys_in: [nlists][n_rows][K] (0 = empty slot); one CTA per row: bitonic sort (descending) of the nlists*K keys padded to a power of
// two, the first K -> keys_out [n_rows][K] and, if requested, ids_out int64 / scores_out fp32 (-1 / -inf for empty slots)
__global__ void __launch_bounds__(256) topk_merge_kernel(const unsigned long long* __restrict__ keys_in, int nlists, int n_rows, int K,
                                                         const int* __restrict__ d_counts, unsigned long long* __restrict__ keys_out,
                                                         long long* __restrict__ ids_out, float* __restrict__ scores_out, int npow2) {
  pdl_grid_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);
  const int row = blockIdx.x, tid = threadIdx.x;
  const bool live = !d_counts || row < d_counts[0];
  const int total = nlists * K;
  for (int i = tid; i < npow2; i += 256) {
    unsigned long long k = 0ull;
    if (live && i < total) k = keys_in[((size_t)(i / K) * n_rows + row) * K + (i % K)];
    s[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npow2 / 2; i += 256) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long x = s[lo], y = s[hi];
        if (desc ? x < y : x > y) { s[lo] = y; s[hi] = x; }
      }
      __syncthreads();
    }
  for (int j = tid; j < K; j += 256) {
    const unsigned long long k = s[j];
    if (keys_out) keys_out[(size_t)row * K + j] = k;
    if (ids_out) ids_out[(size_t)row * K + j] = k ? (long long)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)) : -1ll;
    if (scores_out) {
      unsigned u = (unsigned)(k >> 32);
      u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
      scores_out[(size_t)row * K + j] = k ? __uint_as_float(u) : -INFINITY;
    }
  }
}

static CeDev to_dev(const CeArgs& a) {
  CeDev d;
  d.t = a.t; d.ldt = a.ldt; d.E = a.E; d.vbias = a.vbias; d.labels = a.labels; d.row_w = a.row_w; d.row_mult = a.row_mult;
  d.d_counts = a.d_counts; d.M_cap = a.M_cap; d.V = a.V; d.v_begin = a.v_begin; d.v_end = a.v_end;
  d.vsplits = a.vsplits != 0 ? a.vsplits : 1; d.batch = a.batch; d.target_ctas = a.target_ctas; d.max_splits = a.max_splits;
  d.part = a.part; d.lse = a.lse; d.lab_out = a.lab_out; d.stats = a.stats; d.step_stats = a.step_stats; d.dlogits = a.dlogits; d.ld_dl = a.ld_dl;
  d.row_begin = a.row_begin; d.row_count = a.row_count;
  return d;
}

template <int H>
static size_t ce_fwd_smem() {
  return (size_t)(CE_BM + 2 * CE_BN) * (H + 8) * sizeof(bf16) + 2 * CE_BN * sizeof(float) + 2 * 64 * 5 * sizeof(float);
}
template <int H>
static size_t ce_dl_smem() {
  size_t b = (size_t)CE_BN * (H + 8) * sizeof(bf16);
  size_t stage = (size_t)8 * 16 * 72 * sizeof(bf16);
  return (size_t)CE_BM * (H + 8) * sizeof(bf16) + (b > stage ? b : stage) + CE_BN * sizeof(float);
}

cudaError_t launch_ce_fwd(const CeArgs& a, cudaStream_t st) {
  CeDev d = to_dev(a);
  dim3 grid(d.vsplits, (a.M_cap + CE_BM - 1) / CE_BM);
#define B4R_CE(HH)                                                                                         \
  case HH: {                                                                                               \
    size_t smem = ce_fwd_smem<HH>();                                                                       \
    static bool done_##HH = false;                                                                         \
    if (!done_##HH) { cudaFuncSetAttribute(ce_fwd_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done_##HH = true; }        \
    launch_pdl(ce_fwd_kernel<HH>, dim3(grid), dim3(256), (size_t)(smem), st, d);                                                         \
    break;                                                                                                 \
  }
  switch (a.H) {
    B4R_CE(64)
    B4R_CE(128)
    B4R_CE(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_CE
  return cudaGetLastError();
}

cudaError_t launch_ce_count(const CeArgs& a, const float* s_gt, int* beat, cudaStream_t st) {
  CeDev d = to_dev(a);
  dim3 grid(d.vsplits, (a.M_cap + CE_BM - 1) / CE_BM);
#define B4R_CC(HH)                                                                                         \
  case HH: {                                                                                               \
    size_t smem = ce_fwd_smem<HH>();                                                                       \
    static bool done_##HH = false;                                                                         \
    if (!done_##HH) { cudaFuncSetAttribute(ce_count_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done_##HH = true; }      \
    launch_pdl(ce_count_kernel<HH>, dim3(grid), dim3(256), (size_t)(smem), st, d, s_gt, beat);                                           \
    break;                                                                                                 \
  }
  switch (a.H) {
    B4R_CC(64)
    B4R_CC(128)
    B4R_CC(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_CC
  return cudaGetLastError();
}


int topk_splits(int n_rows, int v_len, int K) {
  const int mt = (n_rows + CE_BM - 1) / CE_BM, ntiles = (v_len + CE_BN - 1) / CE_BN;
  int vs = (2 * 148 + mt - 1) / mt;                 // about two waves of CTAs
  if (vs > ntiles) vs = ntiles;
  while (vs > 1 && (long long)vs * K > 4096) --vs;  // the merge sorts vs * K keys per row in shared memory
  return vs < 1 ? 1 : vs;
}
size_t topk_scratch_bytes(int n_rows, int v_len, int K) { return (size_t)topk_splits(n_rows, v_len, K) * n_rows * K * sizeof(unsigned long long); }

cudaError_t launch_topk_merge(const unsigned long long* keys_in, int nlists, int n_rows, int K, const int* d_counts,
                              unsigned long long* keys_out, long long* ids_out, float* scores_out, cudaStream_t st) {
  if (n_rows <= 0) return cudaSuccess;
  int np2 = 2;
  while (np2 < nlists * K) np2 <<= 1;
  if (np2 > 8192) return cudaErrorInvalidValue;
  static bool done = false;
  if (!done) { cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8); done = true; }
  launch_pdl(topk_merge_kernel, dim3(n_rows), dim3(256), (size_t)np2 * 8, st, keys_in, nlists, n_rows, K, d_counts, keys_out, ids_out, scores_out, np2);
  return cudaGetLastError();
}

// keys_out [n_rows][K] sorted (best first) over [a.v_begin, a.v_end); scratch: topk_scratch_bytes(n_rows, v_len, K)
cudaError_t launch_topk_full(const CeArgs& a, int n_rows, int K, unsigned long long* scratch, unsigned long long* keys_out,
                             long long* ids_out, float* scores_out, cudaStream_t st) {
  if (K < 1 || K > TOPK_MAX || n_rows < 1) return cudaErrorInvalidValue;
  CeDev d = to_dev(a);
  d.vsplits = topk_splits(n_rows, a.v_end - a.v_begin, K);
  dim3 grid(d.vsplits, (n_rows + CE_BM - 1) / CE_BM);
#define B4R_TK(HH)                                                                                         \
  case HH: {                                                                                               \
    size_t smem = (size_t)(CE_BM + 2 * CE_BN) * (HH + 8) * sizeof(bf16) + 2 * CE_BN * sizeof(float) + (size_t)CE_BM * K * 8 + CE_BM * (8 + 12); \
    static size_t cap_##HH = 0;                                                                            \
    if (smem > cap_##HH) { cudaError_t e = cudaFuncSetAttribute(topk_partial_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; cap_##HH = smem; } \
    launch_pdl(topk_partial_kernel<HH>, dim3(grid), dim3(256), smem, st, d, K, n_rows, scratch);          \
    break;                                                                                                 \
  }
  switch (a.H) {
    B4R_TK(64)
    B4R_TK(128)
    B4R_TK(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_TK
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_topk_merge(scratch, d.vsplits, n_rows, K, a.d_counts, keys_out, ids_out, scores_out, st);
}

cudaError_t launch_ce_finalize(const CeArgs& a, cudaStream_t st) {
  int blocks = (a.M_cap + 31) / 32;
  if (blocks > kCeFinalizeMaxBlocks) blocks = kCeFinalizeMaxBlocks;   // one 32-row pass per CTA up to four CTAs per SM
  launch_pdl(ce_finalize_kernel, dim3(blocks), dim3(256), (size_t)0, st, to_dev(a), a.fin_part, a.ticket);
  return cudaGetLastError();
}

cudaError_t launch_ce_dlogits(const CeArgs& a, cudaStream_t st) {
  CeDev d = to_dev(a);
  dim3 grid((a.v_end - a.v_begin + CE_BN - 1) / CE_BN, (a.row_count + CE_BM - 1) / CE_BM);
#define B4R_DL(HH)                                                                                         \
  case HH: {                                                                                               \
    size_t smem = ce_dl_smem<HH>();                                                                        \
    static bool done_##HH = false;                                                                         \
    if (!done_##HH) { cudaFuncSetAttribute(ce_dlogits_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done_##HH = true; }    \
    launch_pdl(ce_dlogits_kernel<HH>, dim3(grid), dim3(256), (size_t)(smem), st, d);                                                     \
    break;                                                                                                 \
  }
  switch (a.H) {
    B4R_DL(64)
    B4R_DL(128)
    B4R_DL(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_DL
  return cudaGetLastError();
}

}  // namespace b4r
