// Gradient all-reduce of the data-parallel step by this library's own kernel over NVLink peer mappings (no NCCL launch on the
// step): every rank's flat fp32 gradient lives in symmetric memory, i.e. every rank holds a device pointer to every other rank's
// buffer (and to a small flag array) that is valid for plain loads and stores through NVLink / NVSwitch.
//
// Two-shot, deterministic, in place.  The buffer is cut into `world` slices.
//   barrier A   every rank's gradient is final (its stream reached this kernel)
//   phase 1     rank r reads slice r of EVERY rank (peer loads, 16 bytes per lane, several in flight) and adds them in rank order
//               0, 1, 2, ... -- the same order whoever reduces, so the sum does not depend on the rank count's timing -- and writes
//               the sum into slice r of its own buffer
//   barrier B   all slices are reduced
//   phase 2     rank r copies slice q != r from rank q's buffer (the one place where it was reduced) into its own
//   barrier C   nobody still reads this rank's buffer: the next step may overwrite it
// Each GPU moves 2 (world - 1) / world of the buffer over NVLink; at 7.4 MB (C3) and 8 ranks that is 13 MB per GPU.  A barrier is a
// flag per (barrier, source rank) in the target's flag array holding the call number (monotonic: no reset, no ABA); waiting is
// polling LOCAL memory.  Inside a GPU the CTAs of the (co-resident, <= one per SM) grid meet at an atomic counter before CTA 0
// signals.  Every spin is bounded: a rank that never arrives makes the others give up after about a minute with an error flag
// instead of hanging the device.  Peer data is read with ld.global.cg (peer lines must not be served from a stale L1 line of the previous step).
// Reference op: the data-parallel gradient mean of the trainer (bert4rec/trainers/trainer_utils.py:22 semantics).
#include "common.cuh"
#include "kernels.h"

namespace b4r {
namespace {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PThreads = 512;
constexpr unsigned kSpinLimit = 1u << 26;   // ~1 us per poll: a rank may lag by tens of seconds (start-up, another rank's eager first step)

__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

struct P2PDev {
  float* const* bufs;        // [world] this buffer on every rank
  uint32_t* const* flags;    // [world] flag array on every rank: [3 barriers][kP2PMaxWorld sources]
  size_t off, n;             // element range of the buffer
  int rank, world;
  uint32_t* state;           // local: [0] completed calls, [1] CTA arrivals (monotonic), [7] error
};

// all CTAs of the grid have arrived `target` times in total; then CTA 0 tells every rank, then every CTA waits for every rank
__device__ void grid_then_world_barrier(const P2PDev& a, int bar, uint32_t epoch, uint32_t target, bool local_first) {
  __syncthreads();
  if (local_first) {
    if (threadIdx.x == 0) {
      __threadfence_system();
      atomicAdd(a.state + 1, 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned it = 0;
      while ((int)(atomicAdd(a.state + 1, 0u) - target) < 0 && ++it < kSpinLimit) {}
      if (it >= kSpinLimit) a.state[7] = 100u + bar;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_flag(a.flags[threadIdx.x] + bar * kP2PMaxWorld + a.rank, epoch);
  }
  if ((int)threadIdx.x < a.world) {
    const uint32_t* f = a.flags[a.rank] + bar * kP2PMaxWorld + threadIdx.x;
    unsigned it = 0;
    while ((int)(ld_flag(f) - epoch) < 0 && ++it < kSpinLimit) {}
    if (it >= kSpinLimit) a.state[7] = 200u + bar;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kP2PThreads, 1) p2p_allreduce_kernel(P2PDev a) {
  pdl_grid_wait();
  const uint32_t epoch = a.state[0] + 1u;      // (CTA 0 publishes it at the very end, behind two grid-wide meetings)
  const int R = a.rank, W = a.world;
  const size_t n4 = a.n / 4, per4 = (n4 + W - 1) / W;
  const size_t gtid = (size_t)blockIdx.x * kP2PThreads + threadIdx.x, gstride = (size_t)gridDim.x * kP2PThreads;
  float* own = a.bufs[R] + a.off;

  grid_then_world_barrier(a, 0, epoch, 0u, false);
  // ---- phase 1: reduce my slice in rank order
  {
    const size_t lo = min(n4, (size_t)R * per4), hi = min(n4, lo + per4);
    if (W <= 4) {
      // few ranks: four elements per lane and trip, so that 4 x world peer loads are in flight per lane (an NVLink round trip is
      // ~1 us; with one element per trip the slice costs its trip count in microseconds)
      for (size_t i = lo + gtid; i < hi; i += 4 * gstride) {
        float4 v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < W) v[u][q] = i + u * gstride < hi ? ld_peer(reinterpret_cast<const float4*>(a.bufs[q] + a.off) + i + u * gstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float4 s = v[u][0];
#pragma unroll
          for (int q = 1; q < 4; ++q)
            if (q < W) { s.x += v[u][q].x; s.y += v[u][q].y; s.z += v[u][q].z; s.w += v[u][q].w; }
          if (i + u * gstride < hi) reinterpret_cast<float4*>(own)[i + u * gstride] = s;
        }
      }
    } else {
      for (size_t i = lo + gtid; i < hi; i += gstride) {
        float4 v[kP2PMaxWorld];                  // all ranks' values requested before the first add (world loads in flight per lane)
#pragma unroll
        for (int q = 0; q < kP2PMaxWorld; ++q)
          if (q < W) v[q] = ld_peer(reinterpret_cast<const float4*>(a.bufs[q] + a.off) + i);
        float4 s = v[0];
#pragma unroll
        for (int q = 1; q < kP2PMaxWorld; ++q)
          if (q < W) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
        reinterpret_cast<float4*>(own)[i] = s;
      }
    }
    if (R == W - 1 && gtid < (a.n & 3)) {       // the last (n mod 4) elements belong to the last rank
      const size_t i = n4 * 4 + gtid;
      float s = 0.f;
      for (int q = 0; q < W; ++q) s += ld_peer1(a.bufs[q] + a.off + i);
      own[i] = s;
    }
  }
  grid_then_world_barrier(a, 1, epoch, (2u * epoch - 1u) * gridDim.x, true);
  // ---- phase 2: gather the other slices from where they were reduced
  {
    // one flat index space over the (world - 1) foreign slices, so that a lane's loads go to several peers at once instead of one
    // round trip per peer; every rank starts with a different peer
    const size_t total = (size_t)(W - 1) * per4;
    for (size_t t0 = gtid; t0 < total; t0 += 4 * gstride) {
      float4 v[4];
      size_t idx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t t = t0 + u * gstride;
        idx[u] = (size_t)-1;
        if (t < total) {
          const int q = (R + 1 + (int)(t / per4)) % W;
          const size_t i = (size_t)q * per4 + t % per4;
          if (i < n4) { idx[u] = i; v[u] = ld_peer(reinterpret_cast<const float4*>(a.bufs[q] + a.off) + i); }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (idx[u] != (size_t)-1) reinterpret_cast<float4*>(own)[idx[u]] = v[u];
    }
    if (R != W - 1 && gtid < (a.n & 3)) own[n4 * 4 + gtid] = ld_peer1(a.bufs[W - 1] + a.off + n4 * 4 + gtid);
  }
  grid_then_world_barrier(a, 2, epoch, 2u * epoch * gridDim.x, true);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.state[0] = epoch;
}

// ---- variant for NVSwitch systems with multicast (NVLS): `mc` is the multicast mapping of the same buffer.  One pass: a lane reads
// the element of ALL ranks with one multimem.ld_reduce (the switch adds the ranks' values) and writes the sum to ALL ranks with one
// multimem.st -- each GPU moves 1/world of the buffer in and 1/world out, and there is no gather phase: two barriers instead of three.
// The same lane reads and then overwrites an element, and nothing else touches the slice, so the operation stays in place.  The sum
// is taken once per element (every rank receives the same bits); its association order is the switch's, not rank order.
__device__ __forceinline__ float4 mm_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float mm_ld_reduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kP2PThreads, 1) p2p_allreduce_nvls_kernel(P2PDev a, float* mc) {
  pdl_grid_wait();
  const uint32_t epoch = a.state[0] + 1u;
  const int R = a.rank, W = a.world;
  const size_t n4 = a.n / 4, per4 = (n4 + W - 1) / W;
  const size_t gtid = (size_t)blockIdx.x * kP2PThreads + threadIdx.x, gstride = (size_t)gridDim.x * kP2PThreads;
  float4* mc4 = reinterpret_cast<float4*>(mc + a.off);
  grid_then_world_barrier(a, 0, epoch, 0u, false);
  const size_t lo = min(n4, (size_t)R * per4), hi = min(n4, lo + per4);
  for (size_t i = lo + gtid; i < hi; i += 4 * gstride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * gstride < hi) v[u] = mm_ld_reduce(mc4 + i + u * gstride);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * gstride < hi) mm_st(mc4 + i + u * gstride, v[u]);
  }
  if (R == W - 1 && gtid < (a.n & 3)) {
    float* pe = mc + a.off + n4 * 4 + gtid;
    mm_st1(pe, mm_ld_reduce1(pe));
  }
  grid_then_world_barrier(a, 1, epoch, epoch * gridDim.x, true);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.state[0] = epoch;
}

}  // namespace

int p2p_allreduce_max_world() { return kP2PMaxWorld; }

cudaError_t launch_p2p_allreduce(float* const* bufs, uint32_t* const* flags, float* multicast, size_t off, size_t n, int rank, int world,
                                 uint32_t* state, cudaStream_t st) {
  if (world < 2 || world > kP2PMaxWorld || rank < 0 || rank >= world) return cudaErrorInvalidValue;
  P2PDev a;
  a.bufs = bufs; a.flags = flags; a.off = off; a.n = n; a.rank = rank; a.world = world; a.state = state;
  // 96 of the 148 SMs (measured at 7.4 MB, 2 ranks: 16 CTAs 59 us, 32: 44, 64: 39, 96: 41, 128: 39): enough lanes to keep the NVLink
  // loads in flight, and every CTA becomes resident whatever little else the device runs (the CTAs meet at a counter)
  static const int ctas = getenv("B4R_P2P_CTAS") ? atoi(getenv("B4R_P2P_CTAS")) : 96;   // (development aid; must be the same on every call)
  if (multicast) return launch_pdl(p2p_allreduce_nvls_kernel, dim3(ctas), dim3(kP2PThreads), (size_t)0, st, a, multicast);
  return launch_pdl(p2p_allreduce_kernel, dim3(ctas), dim3(kP2PThreads), (size_t)0, st, a);
}

}  // namespace b4r
