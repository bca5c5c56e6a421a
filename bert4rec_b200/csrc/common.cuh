// Shared device helpers for the bert4rec_b200 kernels (sm_100a only).
#pragma once
#include <cstdlib>
#include <utility>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH__
#define B4R_HOST 1
#endif

namespace b4r {

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

constexpr int kWarp = 32;
constexpr float kLnEps = 1e-12f;  // reference: bert4rec_encoder.py:116-117

// ---------------------------------------------------------------- dropout site ids (Philox c2)
enum DropSite : uint32_t { SITE_EMB = 1, SITE_ATTN_OUT = 2, SITE_FFN_OUT = 3, SITE_ATTN_PROBS = 4 };
__host__ __device__ inline uint32_t site_id(uint32_t site, uint32_t layer) { return site | (layer << 8); }

// ---------------------------------------------------------------- Philox4x32-7 (the Crush-resistant round count of the Random123 paper; 10 is its
// safety-margin default): the dropout masks cost 13 of the ~23 instructions per attention probability at 10 rounds
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// Dropout threshold on 16-bit lanes: keep iff u16 >= thr16  (P[drop] = thr16 / 65536).
__host__ __device__ inline uint32_t drop_threshold16(float rate) {
  float t = rate * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)t);
}
// 8 keep flags (bit i = keep element col0+i) for the elementwise [rows, cols] dropout sites.
// Canonical mapping: Philox counter = (row, col/8, site, step), 16-bit lane = col % 8.
__device__ __forceinline__ uint32_t keep_bits8(const Philox& ph, uint32_t row, uint32_t col8, uint32_t site,
                                              uint32_t step, uint32_t thr16) {
  uint4 r = ph(row, col8, site, step);
  uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((w[i] & 0xFFFFu) >= thr16 ? 1u : 0u) << (2 * i);
    bits |= ((w[i] >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

// ---------------------------------------------------------------- math
// Exact-erf GELU (keras.activations.gelu(approximate=False), bert4rec_encoder.py:96) and its derivative, branch-free:
// Phi(x) = 1/2 erfc(-x / sqrt 2) with erfc from Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7, i.e. fp32 round-off level),
// whose exponential exp(-x^2 / 2) is also the density term of the derivative.  ~15 instructions instead of erff's ~35: the
// GELU epilogues of the FFN GEMMs were instruction-bound on erff (profiles/r01_ncu_tgemm_c4.md).
__device__ __forceinline__ void gelu_terms(float x, float& Phi, float& dens) {
  const float t = __fdividef(1.0f, fmaf(0.23164189f, fabsf(x), 1.0f));     // 1 / (1 + p |x| / sqrt 2), p = 0.3275911
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));   // exp(-x^2 / 2)
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float hc = 0.5f * q * t * e;                                         // 1/2 erfc(|x| / sqrt 2)
  Phi = x >= 0.f ? 1.0f - hc : hc;
  dens = e * 0.3989422804014327f;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float Phi, dens;
  gelu_terms(x, Phi, dens);
  return x * Phi;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float Phi, dens;
  gelu_terms(x, Phi, dens);
  return fmaf(x, dens, Phi);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {  // sum over aligned groups of W lanes
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  bf162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf162(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}

// ---------------------------------------------------------------- async copy / ldmatrix / mma
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte cp.async; when !valid the destination is zero-filled (src-size 0) and src is not read.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  uint32_t d = smem_u32(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 rows x 16 k) at (row0, k0).
//  TRANS=false: smem holds A[m][k] (k contiguous), ld = elements per row.
//  TRANS=true : smem holds A^T, i.e. [k][m] (m contiguous).
template <bool TRANS>
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const bf16* s, int ld, int row0, int k0, int lane) {
  if (!TRANS) {
    const bf16* p = s + (size_t)(row0 + (lane & 15)) * ld + k0 + ((lane >> 4) << 3);
    ldsm_x4(a, smem_u32(p));
  } else {
    int j = lane >> 3, r = lane & 7;
    const bf16* p = s + (size_t)(k0 + ((j >> 1) << 3) + r) * ld + row0 + ((j & 1) << 3);
    ldsm_x4_t(a, smem_u32(p));
  }
}
// B fragments for two adjacent n-tiles (16 n x 16 k) at (n0, k0): b[0],b[1] -> n-tile n0; b[2],b[3] -> n-tile n0+8.
//  TRANS=false: smem holds B as [n][k] (k contiguous)   ("NT": C = A * B^T with B row-major [N,K])
//  TRANS=true : smem holds B as [k][n] (n contiguous)   ("NN": C = A * B   with B row-major [K,N])
template <bool TRANS>
__device__ __forceinline__ void load_b_frag(uint32_t (&b)[4], const bf16* s, int ld, int n0, int k0, int lane) {
  int j = lane >> 3, r = lane & 7;
  if (!TRANS) {
    const bf16* p = s + (size_t)(n0 + ((j >> 1) << 3) + r) * ld + k0 + ((j & 1) << 3);
    ldsm_x4(b, smem_u32(p));
  } else {
    const bf16* p = s + (size_t)(k0 + ((j & 1) << 3) + r) * ld + n0 + ((j >> 1) << 3);
    ldsm_x4_t(b, smem_u32(p));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched through launch_pdl() may be SCHEDULED while its predecessor in
// the stream is still running -- its grid launch, CTA placement and parameter fetch overlap the predecessor's tail -- and
// blocks at pdl_grid_sync() until the predecessor has completed and flushed its memory.  Every kernel on the step's path
// starts with pdl_grid_sync() (wait for the predecessor, then allow the successor to be scheduled), so the dependency
// chain is unchanged; only the ~2-3 us of launch latency between dependent kernels of the captured step is hidden.
// Without the launch attribute both instructions are no-ops.  B4R_DISABLE_PDL=1 launches without the attribute.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// multi-wave kernels: wait only -- the successor is released when this grid's CTAs exit (an explicit early trigger lets the
// successor's CTAs take SM slots from this grid's later waves: measured slower)
__device__ __forceinline__ void pdl_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// persistent kernels whose whole grid is resident at once (at most one CTA per SM): there are no later waves the successor's CTAs
// could displace, so the successor is released at once and its launch latency hides behind this kernel; a larger grid only waits
constexpr int kNumSMs = 148;
__device__ __forceinline__ void pdl_grid_wait_single_wave() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (gridDim.x * gridDim.y * gridDim.z <= (unsigned)kNumSMs) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
inline bool pdl_enabled() {
  static const bool on = getenv("B4R_DISABLE_PDL") == nullptr;
  return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// C-fragment coordinates: element e of acc[4] -> (row, col) inside a 16x8 tile.
__device__ __forceinline__ int frag_row(int lane, int e) { return (lane >> 2) + ((e >> 1) << 3); }
__device__ __forceinline__ int frag_col(int lane, int e) { return ((lane & 3) << 1) + (e & 1); }

}  // namespace b4r
