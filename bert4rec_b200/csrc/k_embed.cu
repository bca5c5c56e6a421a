// Embedding stage: item-table gather + absolute position embedding + LayerNorm(eps 1e-12) + dropout, and its
// backward (LN backward, position-gradient reduction, scatter-add into the item-table gradient).
// Reference ops replaced: tfm OnDeviceEmbedding / PositionEmbedding / keras LayerNormalization / Dropout at
// bert4rec_encoder.py:198-211.  HBM-bound: 128-bit vectorised row accesses, H/8 lanes per row.
#include "common.cuh"
#include "kernels.h"

namespace b4r {

constexpr int kEmbedRowsPerGroup = 4;

template <int H>
__global__ void __launch_bounds__(256) embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const bf16* __restrict__ table,
                                                           const bf16* __restrict__ pos, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ out, int T,
                                                           int S, int V, uint32_t thr16, float inv_keep, uint64_t seed,
                                                           uint32_t step, const long long* __restrict__ d_step) {
  if (d_step) step += (uint32_t)(*d_step);
  constexpr int LPR = H / 8;       // lanes per row (16 B each)
  constexpr int RPW = 32 / LPR;    // rows per warp and pass
  constexpr int R = kEmbedRowsPerGroup;   // passes per CTA: the ids and the gathered rows of all passes are requested before any is used
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int c0 = l * 8;
  const int t0 = blockIdx.x * (8 * RPW * R) + warp * RPW + sub;
  long long id[R];
  uint4 e[R], p[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = t0 + r * 8 * RPW;
    id[r] = t < T ? ids[t] : 0;
    id[r] = id[r] < 0 ? 0 : (id[r] >= V ? V - 1 : id[r]);  // clamp like a GPU gather; the reference would raise on CPU
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = t0 + r * 8 * RPW;
    e[r] = __ldg(reinterpret_cast<const uint4*>(table + (size_t)id[r] * H + c0));
    p[r] = __ldg(reinterpret_cast<const uint4*>(pos + (size_t)((t < T ? t : 0) % S) * H + c0));
  }
  float gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { gm[i] = gamma[c0 + i]; bt[i] = beta[c0 + i]; }
  const Philox ph(seed);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = t0 + r * 8 * RPW;
    const bool ok = t < T;
    float v[8];
    const uint32_t ew[4] = {e[r].x, e[r].y, e[r].z, e[r].w}, pw[4] = {p[r].x, p[r].y, p[r].z, p[r].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 a = unpack_bf162(ew[i]), b = unpack_bf162(pw[i]);
      v[2 * i] = a.x + b.x; v[2 * i + 1] = a.y + b.y;
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
      for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * gm[i] + bt[i];
      if (thr16 > 0) {
        uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)l, site_id(SITE_EMB, 0), step, thr16);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((bits >> i) & 1u) ? o[i] * inv_keep : 0.f;
      }
      uint4 ov;
      ov.x = pack_bf162(o[0], o[1]); ov.y = pack_bf162(o[2], o[3]); ov.z = pack_bf162(o[4], o[5]); ov.w = pack_bf162(o[6], o[7]);
      *reinterpret_cast<uint4*>(out + (size_t)t * H + c0) = ov;
    }
  }
}

cudaError_t launch_embed_ln_fwd(const int64_t* ids, const bf16* table, const bf16* pos, const float* gamma,
                                const float* beta, bf16* out, int B, int S, int H, int V, float drop_rate,
                                uint64_t seed, uint32_t step, const long long* d_step, cudaStream_t st) {
  const int T = B * S;
  uint32_t thr = drop_threshold16(drop_rate);
  float inv_keep = 1.0f / (1.0f - (float)thr / 65536.0f);
#define B4R_E(HH)                                                                                              \
  case HH: {                                                                                                   \
    int rows_per_cta = 8 * (32 / (HH / 8)) * kEmbedRowsPerGroup;                                               \
    embed_ln_fwd_kernel<HH><<<(T + rows_per_cta - 1) / rows_per_cta, 256, 0, st>>>(ids, table, pos, gamma, beta, out, T, S, \
                                                                                   V, thr, inv_keep, seed, step, d_step);  \
    break;                                                                                                     \
  }
  switch (H) {
    B4R_E(64)
    B4R_E(128)
    B4R_E(256)
    default: return cudaErrorInvalidValue;
  }
#undef B4R_E
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ backward
// grid = (S, bsplits).  CTA (s, z) walks batch rows b = z, z+bsplits, ... for position s: recomputes the LN input
// (gather again - cheaper than saving it), applies dropout' and LN backward, writes the row's dx (dx_rows may alias d_out:
// in place) and accumulates the position gradient / LN parameter gradients locally.  The item-table gradient is NOT
// scattered here: k_tablegrad.cu sums the dx rows per item in a fixed order (bit-reproducible, no hot-row atomics).
int embed_bwd_bsplits(int B) { return B >= 64 ? 4 : 1; }

template <int H>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t* __restrict__ ids, const bf16* __restrict__ table,
                                                        const bf16* __restrict__ pos, const float* __restrict__ gamma,
                                                        const float* d_out, const bf16* __restrict__ d_branch, float* dx_rows,
                                                        float* __restrict__ dpos_part, float* __restrict__ dln_part, int B,
                                                        int S, int V, uint32_t thr16, float inv_keep, uint64_t seed,
                                                        uint32_t step, const long long* __restrict__ d_step) {
  pdl_grid_wait();
  if (d_step) step += (uint32_t)(*d_step);
  constexpr int LPR = H / 8, RPW = 32 / LPR, RPC = 8 * RPW;  // rows per CTA pass
  __shared__ float s_red[3][RPC][H + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR, c0 = l * 8;
  const int s = blockIdx.x, z = blockIdx.y, nz = gridDim.y;
  const int slot = warp * RPW + sub;
  float a_pos[8], a_g[8], a_b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a_pos[i] = a_g[i] = a_b[i] = 0.f;
  float pv[8], gm[8];
  {
    const uint4 p = __ldg(reinterpret_cast<const uint4*>(pos + (size_t)s * H + c0));
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(pw[i]); pv[2 * i] = f.x; pv[2 * i + 1] = f.y; }
#pragma unroll
    for (int i = 0; i < 8; ++i) gm[i] = gamma[c0 + i];
  }
  const Philox ph(seed);
  const int nb = (B - z + nz - 1) / nz;  // batch rows owned by this z
  for (int it = 0; it * RPC < nb; ++it) {
    const int bi = it * RPC + slot;
    const bool ok = bi < nb;
    const int b = z + bi * nz;
    const int t = b * S + s;
    float x[8], dy[8];
    long long id = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = 0.f; dy[i] = 0.f; }
    if (ok) {
      id = ids[t];
      id = id < 0 ? 0 : (id >= V ? V - 1 : id);
      const uint4 e = __ldg(reinterpret_cast<const uint4*>(table + (size_t)id * H + c0));
      const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(ew[i]); x[2 * i] = f.x + pv[2 * i]; x[2 * i + 1] = f.y + pv[2 * i + 1]; }
      const float4 d0 = *reinterpret_cast<const float4*>(d_out + (size_t)t * H + c0);
      const float4 d1 = *reinterpret_cast<const float4*>(d_out + (size_t)t * H + c0 + 4);
      dy[0] = d0.x; dy[1] = d0.y; dy[2] = d0.z; dy[3] = d0.w; dy[4] = d1.x; dy[5] = d1.y; dy[6] = d1.z; dy[7] = d1.w;
      if (d_branch) {   // + the bf16 output of the first layer's QKV data-gradient GEMM
        const uint4 bv = *reinterpret_cast<const uint4*>(d_branch + (size_t)t * H + c0);
        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(bw[i]); dy[2 * i] += f.x; dy[2 * i + 1] += f.y; }
      }
      if (thr16 > 0) {
        uint32_t bits = keep_bits8(ph, (uint32_t)t, (uint32_t)l, site_id(SITE_EMB, 0), step, thr16);
#pragma unroll
        for (int i = 0; i < 8; ++i) dy[i] = ((bits >> i) & 1u) ? dy[i] * inv_keep : 0.f;
      }
    }
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sm += x[i];
    sm = group_sum<LPR>(sm);
    const float mean = sm * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float d = x[i] - mean; q += d * d; }
    q = group_sum<LPR>(q);
    const float rstd = rsqrtf(q * (1.0f / H) + kLnEps);
    float xh[8], dxh[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[i] = (x[i] - mean) * rstd;
      dxh[i] = dy[i] * gm[i];
      s1 += dxh[i]; s2 += dxh[i] * xh[i];
    }
    s1 = group_sum<LPR>(s1) * (1.0f / H);
    s2 = group_sum<LPR>(s2) * (1.0f / H);
    float dx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dx[i] = 0.f;
    if (ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dx[i] = rstd * (dxh[i] - s1 - xh[i] * s2);
        a_pos[i] += dx[i];
        a_g[i] += dy[i] * xh[i];
        a_b[i] += dy[i];
      }
    }
    // dx of the gathered row replaces d_out in place (same thread, same addresses); the item-table gradient is the
    // fixed-order segmented sum of these rows over the id-sorted token list (k_tablegrad.cu)
    if (ok) {
      *reinterpret_cast<float4*>(dx_rows + (size_t)t * H + c0) = make_float4(dx[0], dx[1], dx[2], dx[3]);
      *reinterpret_cast<float4*>(dx_rows + (size_t)t * H + c0 + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
    }
  }
  // CTA reduction over the RPC row slots (fixed order -> deterministic for dpos / dgamma / dbeta)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_red[0][slot][c0 + i] = a_pos[i];
    s_red[1][slot][c0 + i] = a_g[i];
    s_red[2][slot][c0 + i] = a_b[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * H; c += 256) {
    const int k = c / H, col = c % H;
    float v = 0.f;
    for (int r = 0; r < RPC; ++r) v += s_red[k][r][col];
    if (k == 0) dpos_part[((size_t)z * S + s) * H + col] = v;
    else dln_part[((size_t)(z * S + s)) * 2 * H + (k - 1) * H + col] = v;
  }
}

cudaError_t launch_embed_bwd(const int64_t* ids, const bf16* table, const bf16* pos, const float* gamma,
                             const float* d_out, const bf16* d_branch, float* dx_rows, float* dpos_part, float* dln_part, int B, int S,
                             int H, int V, float drop_rate, uint64_t seed, uint32_t step, const long long* d_step,
                             int bsplits, cudaStream_t st) {
  uint32_t thr = drop_threshold16(drop_rate);
  float inv_keep = 1.0f / (1.0f - (float)thr / 65536.0f);
  dim3 grid(S, bsplits);
  switch (H) {
    case 64: launch_pdl(embed_bwd_kernel<64>, dim3(grid), dim3(256), (size_t)(0), st, ids, table, pos, gamma, d_out, d_branch, dx_rows, dpos_part, dln_part, B, S, V, thr, inv_keep, seed, step, d_step); break;
    case 128: launch_pdl(embed_bwd_kernel<128>, dim3(grid), dim3(256), (size_t)(0), st, ids, table, pos, gamma, d_out, d_branch, dx_rows, dpos_part, dln_part, B, S, V, thr, inv_keep, seed, step, d_step); break;
    case 256: launch_pdl(embed_bwd_kernel<256>, dim3(grid), dim3(256), (size_t)(0), st, ids, table, pos, gamma, d_out, d_branch, dx_rows, dpos_part, dln_part, B, S, V, thr, inv_keep, seed, step, d_step); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace b4r
