// Row-wise / elementwise kernels: LayerNorm backward (+dropout'), deterministic partial reduction of all
// parameter gradients, global-norm clip + AdamW with warm-up / linear decay (one fused multi-tensor launch
// over the flat parameter buffer), casts, column sums.
// Reference ops replaced: tape.gradient pieces of keras LayerNormalization/Dropout (bert4rec_model.py:166-167),
// tf.clip_by_global_norm + AdamWeightDecay._resource_apply_dense + WarmUp/PolynomialDecay
// (adam_w_optimizer.py:22-36,91-136; optimizers/__init__.py:38-46) -- SURVEY.md 2b rows K10-K12.
#include "common.cuh"
#include "kernels.h"

namespace b4r {

// ------------------------------------------------------------------------------------------------ LN backward
int ln_bwd_parts(int M) { return (M + 63) / 64; }

// HEAD=false: residual-branch LayerNorm (d_pre written, dropout' applied to the branch gradient)
// HEAD=true : MLM transform LayerNorm: d_out = sum of nsplit split-K partials, the LN input is gelu(t_pre) so the
//             branch gradient is additionally multiplied by gelu'(t_pre); no d_pre output.
template <int H, bool HEAD>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* d_out, const bf16* __restrict__ branch, const bf16* __restrict__ pre,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, float* d_pre,
                                                     bf16* __restrict__ d_branch, float* __restrict__ partials, int M,
                                                     uint32_t thr16, float inv_keep, uint64_t seed, uint32_t site,
                                                     uint32_t step, int nsplit, size_t split_stride,
                                                     const bf16* __restrict__ t_pre, const int* __restrict__ d_M,
                                                     const long long* __restrict__ d_step, int dyn_vtiles = 0,
                                                     int dyn_target = 0, int dyn_max = 0) {
  pdl_grid_wait();
  if (d_M) M = min(M, *d_M);
  if (HEAD && dyn_max > 0) {  // split count chosen on the device by the generation-2 dT pass (same formula)
    nsplit = ce_dyn_splits128(M, dyn_vtiles, dyn_target, dyn_max);
  }
  if (d_step) step += (uint32_t)(*d_step);
  constexpr int LPR = H / 8, RPW = 32 / LPR, RPC = 8 * RPW;
  __shared__ float s_red[3][RPC][H + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR, c0 = l * 8;
  const int slot = warp * RPW + sub;
  float a_g[8], a_b[8], a_c[8], gm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a_g[i] = a_b[i] = a_c[i] = 0.f; gm[i] = gamma[c0 + i]; }
  const Philox ph(seed);
  for (int it = 0; it < 64 / RPC; ++it) {
    const int m = blockIdx.x * 64 + it * RPC + slot;
    const bool ok = m < M;
    float dy[8], xh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dy[i] = 0.f; xh[i] = 0.f; }
    float rs = 0.f;
    if (ok) {
#pragma unroll 4
      for (int sp = 0; sp < (HEAD ? nsplit : 1); ++sp) {
        const float* src = d_out + sp * split_stride + (size_t)m * H + c0;
        const float4 d0 = *reinterpret_cast<const float4*>(src);
        const float4 d1 = *reinterpret_cast<const float4*>(src + 4);
        dy[0] += d0.x; dy[1] += d0.y; dy[2] += d0.z; dy[3] += d0.w; dy[4] += d1.x; dy[5] += d1.y; dy[6] += d1.z; dy[7] += d1.w;
      }
      if (!HEAD && branch) {   // gradient that arrived through the sublayer above (bf16 output of its data-gradient GEMM)
        const uint4 bv = *reinterpret_cast<const uint4*>(branch + (size_t)m * H + c0);
        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(bw[i]); dy[2 * i] += f.x; dy[2 * i + 1] += f.y; }
      }
      const uint4 pv = *reinterpret_cast<const uint4*>(pre + (size_t)m * H + c0);
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
      const float mu = mean[m];
      rs = rstd[m];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = unpack_bf162(pw[i]);
        xh[2 * i] = (f.x - mu) * rs; xh[2 * i + 1] = (f.y - mu) * rs;
      }
    }
    float dxh[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dxh[i] = dy[i] * gm[i]; s1 += dxh[i]; s2 += dxh[i] * xh[i]; }
    s1 = group_sum<LPR>(s1) * (1.0f / H);
    s2 = group_sum<LPR>(s2) * (1.0f / H);
    if (ok) {
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dx[i] = rs * (dxh[i] - s1 - xh[i] * s2);
        a_g[i] += dy[i] * xh[i];
        a_b[i] += dy[i];
      }
      if (!HEAD) {
        *reinterpret_cast<float4*>(d_pre + (size_t)m * H + c0) = make_float4(dx[0], dx[1], dx[2], dx[3]);
        *reinterpret_cast<float4*>(d_pre + (size_t)m * H + c0 + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
      } else {
        const uint4 tv = *reinterpret_cast<const uint4*>(t_pre + (size_t)m * H + c0);
        const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = unpack_bf162(tw[i]);
          dx[2 * i] *= gelu_erf_grad(f.x); dx[2 * i + 1] *= gelu_erf_grad(f.y);
        }
      }
      if (!HEAD && thr16 > 0) {
        uint32_t bits = keep_bits8(ph, (uint32_t)m, (uint32_t)l, site, step, thr16);
#pragma unroll
        for (int i = 0; i < 8; ++i) dx[i] = ((bits >> i) & 1u) ? dx[i] * inv_keep : 0.f;
      }
      uint4 ov;
      ov.x = pack_bf162(dx[0], dx[1]); ov.y = pack_bf162(dx[2], dx[3]); ov.z = pack_bf162(dx[4], dx[5]); ov.w = pack_bf162(dx[6], dx[7]);
      *reinterpret_cast<uint4*>(d_branch + (size_t)m * H + c0) = ov;
      const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = unpack_bf162(ow[i]); a_c[2 * i] += f.x; a_c[2 * i + 1] += f.y; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_red[0][slot][c0 + i] = a_g[i];
    s_red[1][slot][c0 + i] = a_b[i];
    s_red[2][slot][c0 + i] = a_c[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * H; c += 256) {
    const int k = c / H, col = c % H;
    float v = 0.f;
    for (int r = 0; r < RPC; ++r) v += s_red[k][r][col];
    partials[(size_t)blockIdx.x * 3 * H + c] = v;
  }
}

cudaError_t launch_ln_bwd(const float* d_out, const bf16* branch, const bf16* pre, const float* mean, const float* rstd,
                          const float* gamma, float* d_pre, bf16* d_branch, float* partials, int M, int H,
                          float drop_rate, uint64_t seed, uint32_t site, uint32_t step, const long long* d_step,
                          cudaStream_t st) {
  uint32_t thr = drop_threshold16(drop_rate);
  float inv_keep = 1.0f / (1.0f - (float)thr / 65536.0f);
  int grid = ln_bwd_parts(M);
  switch (H) {
    case 64: launch_pdl(ln_bwd_kernel<64, false>, dim3(grid), dim3(256), (size_t)(0), st, d_out, branch, pre, mean, rstd, gamma, d_pre, d_branch, partials, M, thr, inv_keep, seed, site, step, 1, 0, nullptr, nullptr, d_step, 0, 0, 0); break;
    case 128: launch_pdl(ln_bwd_kernel<128, false>, dim3(grid), dim3(256), (size_t)(0), st, d_out, branch, pre, mean, rstd, gamma, d_pre, d_branch, partials, M, thr, inv_keep, seed, site, step, 1, 0, nullptr, nullptr, d_step, 0, 0, 0); break;
    case 256: launch_pdl(ln_bwd_kernel<256, false>, dim3(grid), dim3(256), (size_t)(0), st, d_out, branch, pre, mean, rstd, gamma, d_pre, d_branch, partials, M, thr, inv_keep, seed, site, step, 1, 0, nullptr, nullptr, d_step, 0, 0, 0); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_head_bwd_rows(const float* dt_part, int nsplit, size_t split_stride, const bf16* t_pre,
                                 const bf16* act, const float* mean, const float* rstd, const float* gamma,
                                 bf16* d_tpre, float* partials, int M_cap, const int* d_counts, int H,
                                 cudaStream_t st, int dyn_vtiles, int dyn_target, int dyn_max) {
  int grid = ln_bwd_parts(M_cap);
  const int* d_M = d_counts;   // n_valid: aux rows carry no gradient
  switch (H) {
    case 64: launch_pdl(ln_bwd_kernel<64, true>, dim3(grid), dim3(256), (size_t)(0), st, dt_part, nullptr, act, mean, rstd, gamma, nullptr, d_tpre, partials, M_cap, 0, 1.f, 0, 0, 0, nsplit, split_stride, t_pre, d_M, nullptr, dyn_vtiles, dyn_target, dyn_max); break;
    case 128: launch_pdl(ln_bwd_kernel<128, true>, dim3(grid), dim3(256), (size_t)(0), st, dt_part, nullptr, act, mean, rstd, gamma, nullptr, d_tpre, partials, M_cap, 0, 1.f, 0, 0, 0, nsplit, split_stride, t_pre, d_M, nullptr, dyn_vtiles, dyn_target, dyn_max); break;
    case 256: launch_pdl(ln_bwd_kernel<256, true>, dim3(grid), dim3(256), (size_t)(0), st, dt_part, nullptr, act, mean, rstd, gamma, nullptr, d_tpre, partials, M_cap, 0, 1.f, 0, 0, 0, nsplit, split_stride, t_pre, d_M, nullptr, dyn_vtiles, dyn_target, dyn_max); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ partial reduction
// One launch reduces every parameter-gradient partial buffer of the step (fixed summation order -> deterministic).
// block = 32 columns x 8 part-lanes; grid = (ceil(max_len/32), njobs).
__global__ void __launch_bounds__(256) grad_reduce_kernel(const ReduceJob* __restrict__ jobs) {
  pdl_grid_sync();
  __shared__ float4 s[8][33];
  const ReduceJob job = jobs[blockIdx.y];
  // grid.x is capped (launch_grad_reduce): every block strides over the job's column blocks.  A grid sized for the LONGEST job
  // of the launch made the short jobs (biases, LayerNorm vectors) launch thousands of blocks that exited at once -- at C4
  // (70 jobs x 8192 column blocks) the empty blocks, not the 0.6 GB of partials, were most of this kernel's 555 us.
  if (job.nparts <= 8) {
    // few, long partials (split-K weight gradients, vocabulary-sized buffers): 1024 columns per block, one thread per
    // column, coalesced
    for (int c0 = blockIdx.x * 1024; c0 < job.len; c0 += gridDim.x * 1024) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = c0 + k * 256 + threadIdx.x;
        if (col < job.len) {
          float x[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) x[p] = p < job.nparts ? job.src[(size_t)p * job.part_stride + col] : 0.f;   // all in flight
          float v = 0.f;
#pragma unroll
          for (int p = 0; p < 8; ++p) v += x[p];
          if (job.accumulate) v += job.dst[col];
          job.dst[col] = v;
        }
      }
    }
    return;
  }
  // many partials (per-CTA bias / LayerNorm sums, token-split weight gradients): 32 column groups x 8 part-lanes per block; a
  // column group is 4 consecutive columns (one 16-byte load) when the job's layout allows it, else 1 column
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool wide = (job.len & 3) == 0 && (job.part_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(job.src) | reinterpret_cast<uintptr_t>(job.dst)) & 15) == 0;
  const int cw = wide ? 4 : 1;
  for (int cb = blockIdx.x * 32 * cw; cb < job.len; cb += gridDim.x * 32 * cw) {
    const int col = cb + tx * cw;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < job.len) {
      const float* p = job.src + col;
      if (wide) {
        // eight (predicated) loads in flight per thread: a lane's share of up to 64 partials costs one memory round trip
        for (int k = ty; k < job.nparts; k += 64) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k + 8 * u < job.nparts) v[u] = *reinterpret_cast<const float4*>(p + (size_t)(k + 8 * u) * job.part_stride);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
      } else {
        int k = ty;
        for (; k + 24 < job.nparts; k += 32) {
          float a = p[(size_t)k * job.part_stride], b = p[(size_t)(k + 8) * job.part_stride];
          float c = p[(size_t)(k + 16) * job.part_stride], d = p[(size_t)(k + 24) * job.part_stride];
          acc.x += a; acc.x += b; acc.x += c; acc.x += d;
        }
        for (; k < job.nparts; k += 8) acc.x += p[(size_t)k * job.part_stride];
      }
    }
    __syncthreads();            // (the previous round's readers are done with s)
    s[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < job.len) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 8; ++r) { v.x += s[r][tx].x; v.y += s[r][tx].y; v.z += s[r][tx].z; v.w += s[r][tx].w; }
      if (wide) {
        float4* d = reinterpret_cast<float4*>(job.dst + col);
        if (job.accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *d = v;
      } else {
        if (job.accumulate) v.x += job.dst[col];
        job.dst[col] = v.x;
      }
    }
  }
}

int grad_reduce_blocks(int nparts, int len) { return nparts <= 8 ? (len + 1023) / 1024 : (len + 31) / 32; }

cudaError_t launch_grad_reduce(const ReduceJob* d_jobs, int njobs, int max_blocks, cudaStream_t st) {
  if (njobs <= 0) return cudaSuccess;
  const int cap = njobs > 8 ? 96 : 1184;                 // many jobs of very different lengths: few blocks per job, striding
  dim3 grid(max_blocks < cap ? max_blocks : cap, njobs);
  return launch_pdl(grad_reduce_kernel, grid, dim3(256), (size_t)0, st, d_jobs);
}

// ------------------------------------------------------------------------------------------------ column sums
// partials[z][n] = sum over rows r = z*8+ty, step 8*gridDim.y of src[r][n]
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ src, int ld, int M, int N,
                                                          const int* __restrict__ d_M, int d_M_off, float* __restrict__ part) {
  pdl_grid_wait();
  __shared__ float s[8][65];
  if (d_M) M = min(M, *d_M - d_M_off);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 64 + tx * 2;
  float a0 = 0.f, a1 = 0.f;
  if (col < N) {
    for (int r = blockIdx.y * 8 + ty; r < M; r += 8 * gridDim.y) {
      float2 f = unpack_bf162(*reinterpret_cast<const uint32_t*>(src + (size_t)r * ld + col));
      a0 += f.x; a1 += f.y;
    }
  }
  s[ty][tx * 2] = a0; s[ty][tx * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) v += s[r][threadIdx.x];
    if (c < N) part[(size_t)blockIdx.y * N + c] = v;
  }
}
// N <= 2048, N % 8 == 0: one thread per 8 consecutive columns (16-byte loads), RL row lanes per block
__global__ void __launch_bounds__(256) colsum8_bf16_kernel(const bf16* __restrict__ src, int ld, int M, int N,
                                                           const int* __restrict__ d_M, int d_M_off, float* __restrict__ part) {
  pdl_grid_wait();
  extern __shared__ float s_cs[];   // [RL][N]
  if (d_M) M = min(M, *d_M - d_M_off);
  const int groups = N >> 3, RL = 256 / groups;
  const int gi = threadIdx.x % groups, rl = threadIdx.x / groups;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < RL) {
    const int rows_per_block = (M + gridDim.y - 1) / gridDim.y;
    const int r_begin = blockIdx.y * rows_per_block, r_end = min(M, r_begin + rows_per_block);
#pragma unroll 4
    for (int r = r_begin + rl; r < r_end; r += RL) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + gi * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf162(w[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_cs[rl * N + gi * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    float v = 0.f;
    for (int r = 0; r < RL; ++r) v += s_cs[r * N + c];
    part[(size_t)blockIdx.y * N + c] = v;
  }
}
cudaError_t launch_colsum_bf16(const bf16* src, int ld, int M, int N, const int* d_M, int d_M_off, float* part,
                               int splits, cudaStream_t st) {
  if (N % 8 == 0 && N <= 2048 && ld % 8 == 0) {
    const int groups = N / 8, RL = 256 / groups;
    launch_pdl(colsum8_bf16_kernel, dim3(dim3(1, splits)), dim3(256), (size_t)((size_t)RL * N * sizeof(float)), st, src, ld, M, N, d_M, d_M_off, part);
    return cudaGetLastError();
  }
  dim3 grid((N + 63) / 64, splits);
  launch_pdl(colsum_bf16_kernel, dim3(grid), dim3(256), (size_t)(0), st, src, ld, M, N, d_M, d_M_off, part);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ norm + AdamW
struct AdamWDev {
  float* p; bf16* shadow; const float* g; float* m; float* v;
  long long n_decay, n;
  float* sq_part; int n_sq_part;   // [n_sq_part] partial squared norms, then coef[4] = {gscale*clip, lr, alpha, -} and the ticket
  const float* d_count; float grad_scale;
  long long* d_step;
  float init_lr, end_lr; float num_train_steps, num_warmup_steps;
  float wd, beta1, beta2, eps, clip;
  float* d_lr_out;
};

// Stage 1: squared-norm partials of the gradient.  The LAST block to finish (atomic ticket) sums them in a fixed order,
// advances optimizer.iterations and evaluates the step coefficients ONCE (global-norm clip scale, warm-up / decayed
// learning rate, Adam bias correction in double precision) -- not per block of the update kernel.
__global__ void __launch_bounds__(256) sqnorm_kernel(AdamWDev a, int with_coef) {
  pdl_grid_sync();
  __shared__ float s[8];
  __shared__ int s_last;
  float acc = 0.f;
  const long long n4 = a.n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.g);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) { float v = a.g[(n4 << 2) + threadIdx.x]; acc += v * v; }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += s[i];
    a.sq_part[blockIdx.x] = v;
  }
  if (!with_coef) return;
  int* ticket = reinterpret_cast<int*>(a.sq_part + a.n_sq_part + 4);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float tot = 0.f;
    for (int i = threadIdx.x; i < a.n_sq_part; i += 32) tot += __ldcg(a.sq_part + i);
    tot = warp_sum(tot);
    if (threadIdx.x == 0) {
      *ticket = 0;
      const long long t = *a.d_step + 1;   // optimizer.iterations -> t = iterations + 1
      *a.d_step = t;
      float gscale = a.grad_scale;
      if (a.d_count) gscale /= fmaxf(*a.d_count, 1.0f);
      const float gn = sqrtf(tot) * gscale;  // global norm of the (normalised) gradient
      float clip_scale = 1.0f;
      if (a.clip > 0.f) clip_scale = a.clip * fminf(1.0f / gn, 1.0f / a.clip);  // tf.clip_by_global_norm
      const float it = (float)(t - 1);
      float lr;
      if (a.num_warmup_steps > 0.f && it < a.num_warmup_steps) {
        lr = a.init_lr * (it / a.num_warmup_steps);                        // WarmUp, power 1
      } else {
        const float st = fminf(it, a.num_train_steps);
        lr = (a.init_lr - a.end_lr) * (1.0f - st / a.num_train_steps) + a.end_lr;  // PolynomialDecay, power 1
      }
      const double b1p = pow((double)a.beta1, (double)t), b2p = pow((double)a.beta2, (double)t);
      const float alpha = (float)((double)lr * sqrt(1.0 - b2p) / (1.0 - b1p));
      float* coef = a.sq_part + a.n_sq_part;
      coef[0] = gscale * clip_scale; coef[1] = lr; coef[2] = alpha;
      if (a.d_lr_out) { a.d_lr_out[0] = lr; a.d_lr_out[1] = gn; }
    }
  }
}

cudaError_t launch_sqnorm(const float* g, long long n, float* out_part, int nblocks, cudaStream_t st) {
  AdamWDev d{};
  d.g = g; d.n = n; d.sq_part = out_part; d.n_sq_part = nblocks;
  sqnorm_kernel<<<nblocks, 256, 0, st>>>(d, 0);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamWDev a) {
  pdl_grid_sync();
  const float* coef = a.sq_part + a.n_sq_part;
  const float gs = coef[0], lr = coef[1], alpha = coef[2];
  const float ob1 = 1.0f - a.beta1, ob2 = 1.0f - a.beta2;
  const long long n4 = a.n >> 2;  // segments are padded to multiples of 8 elements, so n % 4 == 0
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    const bool decay = (i << 2) < a.n_decay;  // n_decay is a multiple of 8
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x * gs, g.y * gs, g.z * gs, g.w * gs};
    float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (decay) pp[k] -= lr * pp[k] * a.wd;
      mm[k] += (gg[k] - mm[k]) * ob1;
      vv[k] += (gg[k] * gg[k] - vv[k]) * ob2;
      pp[k] -= alpha * mm[k] / (sqrtf(vv[k]) + a.eps);
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    uint2 sh;
    sh.x = pack_bf162(pp[0], pp[1]); sh.y = pack_bf162(pp[2], pp[3]);
    reinterpret_cast<uint2*>(a.shadow)[i] = sh;
  }
}

cudaError_t launch_adamw(const AdamWArgs& a, cudaStream_t st) {
  AdamWDev d;
  d.p = a.p; d.shadow = a.shadow; d.g = a.g; d.m = a.m; d.v = a.v; d.n_decay = a.n_decay; d.n = a.n;
  d.sq_part = const_cast<float*>(a.sq_part); d.n_sq_part = a.n_sq_part; d.d_count = a.d_count; d.grad_scale = a.grad_scale;
  d.d_step = a.d_step_out; d.init_lr = a.init_lr; d.end_lr = a.end_lr;
  d.num_train_steps = (float)a.num_train_steps; d.num_warmup_steps = (float)a.num_warmup_steps;
  d.wd = a.wd; d.beta1 = a.beta1; d.beta2 = a.beta2; d.eps = a.eps; d.clip = a.clip; d.d_lr_out = a.d_lr_out;
  // stage 1: squared-norm partials + (last block) step counter and coefficients ; stage 2: update
  launch_pdl(sqnorm_kernel, dim3(a.n_sq_part), dim3(256), (size_t)0, st, d, 1);
  long long n4 = a.n >> 2;
  int blocks = (int)((n4 + 1023) / 1024);   // 4 float4 per thread
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return launch_pdl(adamw_kernel, dim3(blocks), dim3(256), (size_t)0, st, d);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}
cudaError_t launch_cast_bf16(const float* src, bf16* dst, long long n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<<<blocks, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ test helper
__global__ void dropout_mask_dump_kernel(uint8_t* out, int rows, int cols, uint32_t thr16, uint64_t seed, uint32_t site,
                                         uint32_t step) {
  const Philox ph(seed);
  const int c8n = (cols + 7) / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * c8n;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / c8n), c8 = (int)(i % c8n);
    uint32_t bits = thr16 > 0 ? keep_bits8(ph, (uint32_t)r, (uint32_t)c8, site, step, thr16) : 0xFFu;
    for (int k = 0; k < 8; ++k)
      if (c8 * 8 + k < cols) out[(size_t)r * cols + c8 * 8 + k] = (bits >> k) & 1u;
  }
}
cudaError_t launch_dropout_mask_dump(uint8_t* out, int rows, int cols, float rate, uint64_t seed, uint32_t site,
                                     uint32_t step, cudaStream_t st) {
  dropout_mask_dump_kernel<<<148 * 4, 256, 0, st>>>(out, rows, cols, drop_threshold16(rate), seed, site, step);
  return cudaGetLastError();
}

}  // namespace b4r
