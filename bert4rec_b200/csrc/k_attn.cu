// Masked bidirectional self-attention (forward + backward), one (sequence, head) per CTA column.
// Replaces Keras MultiHeadAttention inside tfm TransformerEncoderBlock (bert4rec_encoder.py:136-147,216-222;
// SURVEY.md 2b rows K2/K3): the [B,S,S] mask and the [B,N,S,S] score/prob tensors are never materialised; the
// key-padding mask is an additive -1e9 applied in registers (Keras Softmax(mask) semantics: a fully masked row
// degenerates to uniform attention, exactly as in the reference), attention-prob dropout keep bits are written
// bit-packed (1 bit / prob) in forward and re-read in backward.
// Generation 1: mma.sync m16n8k16 with flash-style online softmax.  S <= 256, D in {32, 64}.
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace b4r {

int attn_mask_words(int S) { return (S + 63) / 64; }

template <int D>
struct AttnSmem {
  static constexpr int LD = D + 8;
};

// ------------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, const int64_t* __restrict__ mask,
                                                       bf16* __restrict__ ctx, float* __restrict__ lse,
                                                       uint64_t* __restrict__ keep_bits, int S, int H, int N,
                                                       uint32_t thr16, float inv_keep, uint64_t seed, uint32_t site,
                                                       uint32_t step, const long long* __restrict__ d_step) {
  pdl_grid_wait();
  if (d_step) step += (uint32_t)(*d_step);
  constexpr int LD = D + 8;
  constexpr int KD = D / 16;  // k-steps over the head dim
  constexpr int ND = D / 8;   // output n-tiles
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S16 = (S + 15) & ~15;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + 64 * LD;
  bf16* sV = sK + S16 * LD;
  float* sMask = reinterpret_cast<float*>(sV + S16 * LD);

  const int bn = blockIdx.y, b = bn / N, n = bn % N;
  const int q0 = blockIdx.x * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bf16* base = qkv + (size_t)b * S * 3 * H + n * D;
  constexpr int CH = D / 8;
  for (int c = tid; c < 64 * CH; c += 128) {
    int r = c / CH, cc = (c % CH) * 8;
    bool ok = q0 + r < S;
    cp_async16(sQ + r * LD + cc, base + (size_t)(ok ? q0 + r : 0) * 3 * H + cc, ok);
  }
  for (int c = tid; c < S16 * CH; c += 128) {
    int r = c / CH, cc = (c % CH) * 8;
    bool ok = r < S;
    const bf16* src = base + (size_t)(ok ? r : 0) * 3 * H + cc;
    cp_async16(sK + r * LD + cc, src + H, ok);
    cp_async16(sV + r * LD + cc, src + 2 * H, ok);
  }
  cp_async_commit();
  const int S64 = (S + 63) & ~63;
  for (int j = tid; j < S64; j += 128)
    sMask[j] = j < S ? (mask[(size_t)b * S + j] != 0 ? 0.f : -1e9f) : -INFINITY;
  cp_async_wait<0>();
  __syncthreads();

  const float scale = rsqrtf((float)D);
  uint32_t aq[KD][4];
#pragma unroll
  for (int kk = 0; kk < KD; ++kk) load_a_frag<false>(aq[kk], sQ, LD, warp * 16, kk * 16, lane);

  float o[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const Philox ph(seed);
  const int W = (S + 63) / 64;
  const int row_g[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

  for (int kb = 0; kb * 64 < S16; ++kb) {
    const int np_max = min(4, (S16 - kb * 64) / 16);
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KD; ++kk)
#pragma unroll
      for (int np = 0; np < 4; ++np)
        if (np < np_max) {
          uint32_t bb[4];
          load_b_frag<false>(bb, sK, LD, kb * 64 + np * 16, kk * 16, lane);
          mma_bf16(s[2 * np], aq[kk], bb[0], bb[1]);
          mma_bf16(s[2 * np + 1], aq[kk], bb[2], bb[3]);
        }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = kb * 64 + nt * 8 + 2 * t4 + (e & 1);
        float v = (nt < 2 * np_max) ? s[nt][e] * scale + sMask[j] : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    float corr[2], m_use[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m_run[h], mx[h]);
      m_use[h] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[h] = (m_run[h] == -INFINITY) ? 0.f : __expf(m_run[h] - m_use[h]);
      m_run[h] = m_new;
    }
    float psum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float p = __expf(s[nt][e] - m_use[e >> 1]);
        s[nt][e] = p;
        psum[e >> 1] += p;
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + psum[h];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[i][e] *= corr[e >> 1];

    if (thr16 > 0) {
      // keep bits: Philox counter (global query row, kb*8 + t4*2 + call, site, step); 16-bit lane (nt%4)*2 + (e&1)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t grow = (uint32_t)(bn * S + row_g[h]);
        unsigned long long word = 0ull;
#pragma unroll
        for (int call = 0; call < 2; ++call) {
          uint4 r = ph(grow, (uint32_t)(kb * 8 + t4 * 2 + call), site, step);
          const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t u = (q & 1) ? (w[q >> 1] >> 16) : (w[q >> 1] & 0xFFFFu);
            const int nt = call * 4 + (q >> 1), e = h * 2 + (q & 1);
            const bool keep = u >= thr16;
            s[nt][e] = keep ? s[nt][e] * inv_keep : 0.f;
            word |= (unsigned long long)(keep ? 1u : 0u) << (nt * 8 + 2 * t4 + (q & 1));
          }
        }
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        if (t4 == 0 && row_g[h] < S) keep_bits[((size_t)bn * S + row_g[h]) * W + kb] = word;
      }
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      if (ks < np_max) {
        uint32_t ap[4];
        ap[0] = pack_bf162(s[2 * ks][0], s[2 * ks][1]);
        ap[1] = pack_bf162(s[2 * ks][2], s[2 * ks][3]);
        ap[2] = pack_bf162(s[2 * ks + 1][0], s[2 * ks + 1][1]);
        ap[3] = pack_bf162(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
        for (int dp = 0; dp < ND / 2; ++dp) {
          uint32_t bb[4];
          load_b_frag<true>(bb, sV, LD, dp * 16, kb * 64 + ks * 16, lane);
          mma_bf16(o[2 * dp], ap, bb[0], bb[1]);
          mma_bf16(o[2 * dp + 1], ap, bb[2], bb[3]);
        }
      }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float l = l_run[h];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int i = row_g[h];
    if (i < S) {
      const float inv = 1.0f / l;
      bf16* dst = ctx + ((size_t)b * S + i) * H + n * D;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
        *reinterpret_cast<uint32_t*>(dst + nd * 8 + 2 * t4) = pack_bf162(o[nd][h * 2] * inv, o[nd][h * 2 + 1] * inv);
      if (t4 == 0) lse[(size_t)bn * S + i] = m_run[h] + __logf(l);
    }
  }
}

static size_t attn_fwd_smem(int S, int D) {
  int S16 = (S + 15) & ~15;
  return (size_t)(64 + 2 * S16) * (D + 8) * sizeof(bf16) + (size_t)((S + 63) & ~63) * sizeof(float);
}

cudaError_t launch_attn_fwd(const AttnArgs& a, cudaStream_t st) {
  if (!a.no_tcgen05 && fattn_supported(a) && !getenv("B4R_FATTN_NO_FWD")) return launch_fattn_fwd(a, st);   // generation 2: tcgen05 + TMEM + TMA
  if (tattn_fwd_supported(a)) return launch_tattn_fwd(a, st);                 // first tcgen05 forward (opt-in)
  const int D = a.H / a.N;
  if (a.S > 256 || (D != 32 && D != 64)) return cudaErrorInvalidValue;
  uint32_t thr = drop_threshold16(a.drop_rate);
  float inv_keep = 1.0f / (1.0f - (float)thr / 65536.0f);
  dim3 grid((a.S + 63) / 64, a.B * a.N);
  size_t smem = attn_fwd_smem(a.S, D);
  static size_t cap32 = 0, cap64 = 0;
  if (D == 32) {
    if (smem > cap32) { cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cap32 = smem; }
    launch_pdl(attn_fwd_kernel<32>, dim3(grid), dim3(128), (size_t)(smem), st, a.qkv, a.mask, a.ctx, a.lse, a.keep_bits, a.S, a.H, a.N, thr, inv_keep, a.seed, a.site, a.step, a.d_step);
  } else {
    if (smem > cap64) { cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cap64 = smem; }
    launch_pdl(attn_fwd_kernel<64>, dim3(grid), dim3(128), (size_t)(smem), st, a.qkv, a.mask, a.ctx, a.lse, a.keep_bits, a.S, a.H, a.N, thr, inv_keep, a.seed, a.site, a.step, a.d_step);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (sequence, head); 4 warps; key blocks of 64 (16 keys per warp), query blocks of 64.
// Works in the transposed arrangement (rows = keys, cols = queries) so that P^T and dS^T feed the dV / dK MMAs
// straight from registers; dS^T goes through shared memory once to produce dQ.
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const bf16* __restrict__ qkv, const int64_t* __restrict__ mask,
                                                       const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                                                       const float* __restrict__ lse, const uint64_t* __restrict__ keep_bits,
                                                       bf16* __restrict__ dqkv, int S, int H, int N, uint32_t thr16,
                                                       float inv_keep) {
  pdl_grid_wait();
  constexpr int LD = D + 8;
  constexpr int KD = D / 16, ND = D / 8;
  constexpr int LDS = 72;  // dS^T tile row stride (64 queries + 8)
  constexpr int LDQ = D + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S16 = (S + 15) & ~15;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + S16 * LD;
  bf16* sV = sK + S16 * LD;
  bf16* sdO = sV + S16 * LD;
  bf16* sDS = sdO + S16 * LD;
  float* sdQ = reinterpret_cast<float*>(sDS + 64 * LDS);
  float* sLse = sdQ + S16 * LDQ;
  float* sDelta = sLse + S16;
  float* sMask = sDelta + S16;
  unsigned long long* sBits = reinterpret_cast<unsigned long long*>(sMask + S16);  // [64]

  const int bn = blockIdx.x, b = bn / N, n = bn % N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bf16* base = qkv + (size_t)b * S * 3 * H + n * D;
  const bf16* dobase = dctx + (size_t)b * S * H + n * D;
  const bf16* obase = ctx + (size_t)b * S * H + n * D;
  constexpr int CH = D / 8;
  for (int c = tid; c < S16 * CH; c += 128) {
    int r = c / CH, cc = (c % CH) * 8;
    bool ok = r < S;
    const bf16* src = base + (size_t)(ok ? r : 0) * 3 * H + cc;
    cp_async16(sQ + r * LD + cc, src, ok);
    cp_async16(sK + r * LD + cc, src + H, ok);
    cp_async16(sV + r * LD + cc, src + 2 * H, ok);
    cp_async16(sdO + r * LD + cc, dobase + (size_t)(ok ? r : 0) * H + cc, ok);
  }
  cp_async_commit();
  for (int i = tid; i < S16; i += 128) {
    float dl = 0.f;
    if (i < S) {
      const uint4* po = reinterpret_cast<const uint4*>(obase + (size_t)i * H);
      const uint4* pd = reinterpret_cast<const uint4*>(dobase + (size_t)i * H);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        uint4 ov = po[c], dv = pd[c];
        const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 a = unpack_bf162(ow[k]), d2 = unpack_bf162(dw[k]);
          dl += a.x * d2.x + a.y * d2.y;
        }
      }
    }
    sDelta[i] = dl;
    sLse[i] = i < S ? lse[(size_t)bn * S + i] : INFINITY;
    sMask[i] = i < S ? (mask[(size_t)b * S + i] != 0 ? 0.f : -1e9f) : -INFINITY;
  }
  for (int i = tid; i < S16 * LDQ; i += 128) sdQ[i] = 0.f;
  cp_async_wait<0>();
  __syncthreads();

  const float scale = rsqrtf((float)D);
  const int W = (S + 63) / 64;

  for (int kb = 0; kb * 64 < S16; ++kb) {
    const int key0 = kb * 64 + warp * 16;
    const bool warp_active = key0 < S16;
    const int nk_pairs = min(4, (S16 - kb * 64) / 16);
    uint32_t ak[KD][4], av[KD][4];
    float dk[ND][4], dv[ND][4];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
    if (warp_active) {
#pragma unroll
      for (int kk = 0; kk < KD; ++kk) {
        load_a_frag<false>(ak[kk], sK, LD, key0, kk * 16, lane);
        load_a_frag<false>(av[kk], sV, LD, key0, kk * 16, lane);
      }
    }
    const float madd[2] = {warp_active ? sMask[key0 + g] : 0.f, warp_active ? sMask[key0 + g + 8] : 0.f};

    for (int qb = 0; qb * 64 < S16; ++qb) {
      const int nq_pairs = min(4, (S16 - qb * 64) / 16);
      if (thr16 > 0) {
        if (tid < 64) {
          const int i = qb * 64 + tid;
          sBits[tid] = i < S ? keep_bits[((size_t)bn * S + i) * W + kb] : 0ull;
        }
        __syncthreads();
      }
      if (warp_active) {
        float st[8][4], dpt[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) st[i][e] = dpt[i][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KD; ++kk)
#pragma unroll
          for (int np = 0; np < 4; ++np)
            if (np < nq_pairs) {
              uint32_t bb[4];
              load_b_frag<false>(bb, sQ, LD, qb * 64 + np * 16, kk * 16, lane);
              mma_bf16(st[2 * np], ak[kk], bb[0], bb[1]);
              mma_bf16(st[2 * np + 1], ak[kk], bb[2], bb[3]);
              load_b_frag<false>(bb, sdO, LD, qb * 64 + np * 16, kk * 16, lane);
              mma_bf16(dpt[2 * np], av[kk], bb[0], bb[1]);
              mma_bf16(dpt[2 * np + 1], av[kk], bb[2], bb[3]);
            }
        // st -> P_d^T (kept in st), dpt -> dS^T (kept in dpt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (nt < 2 * nq_pairs) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int il = nt * 8 + 2 * t4 + (e & 1);  // query inside the block
              const int i = qb * 64 + il;
              const int h = e >> 1;
              const float val = st[nt][e] * scale + madd[h];
              const float p = __expf(val - sLse[i]);
              float pd = p, dp = dpt[nt][e];
              if (thr16 > 0) {
                const bool keep = (sBits[il] >> (warp * 16 + g + h * 8)) & 1ull;
                pd = keep ? p * inv_keep : 0.f;
                dp = keep ? dp * inv_keep : 0.f;
              }
              st[nt][e] = pd;
              dpt[nt][e] = p * (dp - sDelta[i]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) st[nt][e] = dpt[nt][e] = 0.f;
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < nq_pairs) {
            uint32_t ap[4], ad[4];
            ap[0] = pack_bf162(st[2 * ks][0], st[2 * ks][1]);
            ap[1] = pack_bf162(st[2 * ks][2], st[2 * ks][3]);
            ap[2] = pack_bf162(st[2 * ks + 1][0], st[2 * ks + 1][1]);
            ap[3] = pack_bf162(st[2 * ks + 1][2], st[2 * ks + 1][3]);
            ad[0] = pack_bf162(dpt[2 * ks][0], dpt[2 * ks][1]);
            ad[1] = pack_bf162(dpt[2 * ks][2], dpt[2 * ks][3]);
            ad[2] = pack_bf162(dpt[2 * ks + 1][0], dpt[2 * ks + 1][1]);
            ad[3] = pack_bf162(dpt[2 * ks + 1][2], dpt[2 * ks + 1][3]);
#pragma unroll
            for (int dp2 = 0; dp2 < ND / 2; ++dp2) {
              uint32_t bb[4];
              load_b_frag<true>(bb, sdO, LD, dp2 * 16, qb * 64 + ks * 16, lane);
              mma_bf16(dv[2 * dp2], ap, bb[0], bb[1]);
              mma_bf16(dv[2 * dp2 + 1], ap, bb[2], bb[3]);
              load_b_frag<true>(bb, sQ, LD, dp2 * 16, qb * 64 + ks * 16, lane);
              mma_bf16(dk[2 * dp2], ad, bb[0], bb[1]);
              mma_bf16(dk[2 * dp2 + 1], ad, bb[2], bb[3]);
            }
          }
        // dS^T tile -> smem [key_local][query_local]
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            *reinterpret_cast<uint32_t*>(sDS + (warp * 16 + g + h * 8) * LDS + nt * 8 + 2 * t4) =
                pack_bf162(dpt[nt][h * 2], dpt[nt][h * 2 + 1]);
      }
      __syncthreads();
      // dQ[q-block rows owned by this warp] += dS[queries x keys(kb)] * K[keys x D]
      if (qb * 64 + warp * 16 < S16) {
        float dq[ND][4];
#pragma unroll
        for (int i = 0; i < ND; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) dq[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks < nk_pairs) {
            uint32_t a[4];
            load_a_frag<true>(a, sDS, LDS, warp * 16, ks * 16, lane);
#pragma unroll
            for (int dp2 = 0; dp2 < ND / 2; ++dp2) {
              uint32_t bb[4];
              load_b_frag<true>(bb, sK, LD, dp2 * 16, kb * 64 + ks * 16, lane);
              mma_bf16(dq[2 * dp2], a, bb[0], bb[1]);
              mma_bf16(dq[2 * dp2 + 1], a, bb[2], bb[3]);
            }
          }
#pragma unroll
        for (int nd = 0; nd < ND; ++nd)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = qb * 64 + warp * 16 + g + (e >> 1) * 8;
            sdQ[i * LDQ + nd * 8 + 2 * t4 + (e & 1)] += dq[nd][e];
          }
      }
      __syncthreads();
    }
    if (warp_active) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = key0 + g + h * 8;
        if (j < S) {
          bf16* dst = dqkv + ((size_t)b * S + j) * 3 * H + n * D;
#pragma unroll
          for (int nd = 0; nd < ND; ++nd) {
            *reinterpret_cast<uint32_t*>(dst + H + nd * 8 + 2 * t4) = pack_bf162(dk[nd][h * 2] * scale, dk[nd][h * 2 + 1] * scale);
            *reinterpret_cast<uint32_t*>(dst + 2 * H + nd * 8 + 2 * t4) = pack_bf162(dv[nd][h * 2], dv[nd][h * 2 + 1]);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < S * (D / 2); c += 128) {
    const int i = c / (D / 2), d2 = (c % (D / 2)) * 2;
    *reinterpret_cast<uint32_t*>(dqkv + ((size_t)b * S + i) * 3 * H + n * D + d2) =
        pack_bf162(sdQ[i * LDQ + d2] * scale, sdQ[i * LDQ + d2 + 1] * scale);
  }
}

static size_t attn_bwd_smem(int S, int D) {
  int S16 = (S + 15) & ~15;
  return (size_t)4 * S16 * (D + 8) * sizeof(bf16) + 64 * 72 * sizeof(bf16) + (size_t)S16 * (D + 1) * sizeof(float) +
         3 * (size_t)S16 * sizeof(float) + 64 * sizeof(unsigned long long) + 16;
}

cudaError_t launch_attn_bwd(const AttnArgs& a, cudaStream_t st) {
  if (!a.no_tcgen05 && fattn_supported(a) && !getenv("B4R_FATTN_NO_BWD")) return launch_fattn_bwd(a, st);   // generation 2: tcgen05 + TMEM + TMA
  const int D = a.H / a.N;
  if (a.S > 256 || (D != 32 && D != 64)) return cudaErrorInvalidValue;
  uint32_t thr = drop_threshold16(a.drop_rate);
  float inv_keep = 1.0f / (1.0f - (float)thr / 65536.0f);
  size_t smem = attn_bwd_smem(a.S, D);
  static size_t cap32 = 0, cap64 = 0;
  if (D == 32) {
    if (smem > cap32) { cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cap32 = smem; }
    launch_pdl(attn_bwd_kernel<32>, dim3(a.B * a.N), dim3(128), (size_t)(smem), st, a.qkv, a.mask, a.ctx, a.dctx, a.lse, a.keep_bits, a.dqkv, a.S, a.H, a.N, thr, inv_keep);
  } else {
    if (smem > cap64) { cudaFuncSetAttribute(attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cap64 = smem; }
    launch_pdl(attn_bwd_kernel<64>, dim3(a.B * a.N), dim3(128), (size_t)(smem), st, a.qkv, a.mask, a.ctx, a.dctx, a.lse, a.keep_bits, a.dqkv, a.S, a.H, a.N, thr, inv_keep);
  }
  return cudaGetLastError();
}

}  // namespace b4r
