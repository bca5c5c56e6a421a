// Deterministic item-table gradient of the embedding gather (reference: the scatter-add TensorFlow performs for
// tf.gather's gradient, bert4rec/models/components/networks/bert4rec_encoder.py:103-108 -> OnDeviceEmbedding):
//   grad_table[id] += sum_{t : ids[t] == id} dx[t]
// as a fixed-order segmented sum instead of floating-point atomics.
//   1. token sort (integer, side branch of the step): stable LSD radix sort of (id, t), 8-bit digits, three small kernels per
//      pass (block histograms, one-block exclusive scan, stable scatter).  Stability makes the order inside a run (= the
//      tokens of one item) the token order, independent of scheduling.
//   2. long-run list (side branch): runs that span more than kMaxParts chunks, found from samples + binary search.
//   3. chunk sums: the sorted list is cut into chunks of 64 tokens; a group of H/4 threads walks one chunk, keeps the running
//      sum of the current run in registers and, at a run end, either adds it to the table (run completely inside the chunk:
//      exactly one writer per item) or stores it as the chunk's head / tail carry.
//   4. boundary runs: the carries of a run are added in chunk order by one group (short runs) or one CTA (long runs).
// HBM-bound: T rows of H fp32 are gathered once (4H bytes per token) + the touched table rows read and written once.
#include "common.cuh"
#include "kernels.h"

namespace b4r {

namespace {
constexpr int kSortThreads = 256, kSortWarps = 8, kSortRounds = 8;
constexpr int kSortTile = kSortThreads * kSortRounds;   // 2048 keys per CTA; warp w owns 256 contiguous keys
// tokens per chunk of the segmented sum: 64 for large batches; 16 for small ones, where the walk over a chunk is a chain of
// dependent memory round trips that only more (shorter) chunks can hide
constexpr int kMaxParts = 16;                           // runs spanning more chunks than this are summed by a whole CTA
constexpr int kUnroll = 16;
__host__ __device__ constexpr int sample_stride(int chunk) { return 8 * chunk; }   // every long run contains a multiple of this

__device__ __forceinline__ uint32_t load_key(const long long* ids, const uint32_t* keys_in, int idx, int V) {
  if (ids) {
    long long id = ids[idx];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    return (uint32_t)id;
  }
  return keys_in[idx];
}

// per-warp digit counts of this CTA's tile; keys stay in registers for the scatter
__device__ __forceinline__ void count_tile(const long long* ids, const uint32_t* keys_in, int T, int V, int shift,
                                           uint32_t (*cnt)[256], uint32_t (&key)[kSortRounds]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile + warp * 32 * kSortRounds;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int idx = base + r * 32 + lane;
    key[r] = idx < T ? load_key(ids, keys_in, idx, V) : 0u;
  }
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int idx = base + r * 32 + lane;
    const bool ok = idx < T;
    const uint32_t d = ok ? ((key[r] >> shift) & 255u) : (256u + lane);   // tail lanes: singleton groups, not counted
    const unsigned m = __match_any_sync(0xffffffffu, d);
    if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads) tg_hist_kernel(const long long* __restrict__ ids, const uint32_t* __restrict__ keys_in,
                                                               int T, int V, int shift, uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t cnt[kSortWarps][256];
  uint32_t key[kSortRounds];
  count_tile(ids, keys_in, T, V, shift, cnt, key);
  uint32_t tot = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) tot += cnt[w][threadIdx.x];
  hist[(size_t)threadIdx.x * nblk + blockIdx.x] = tot;
}

// exclusive scan of n counters in place, one CTA (n = 256 digits x sort CTAs: 25 600 at C4; any n works, a thread owns n / 1024 of them)
__global__ void __launch_bounds__(1024) tg_scan_kernel(uint32_t* __restrict__ h, int n) {
  __shared__ uint32_t wsum[32];
  const int per = (n + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, n);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += h[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  uint32_t run = wsum[warp] + inc - s;
  for (int i = lo; i < hi; ++i) { const uint32_t v = h[i]; h[i] = run; run += v; }
}

__global__ void __launch_bounds__(kSortThreads) tg_scatter_kernel(const long long* __restrict__ ids, const uint32_t* __restrict__ keys_in,
                                                                  const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                                  uint32_t* __restrict__ vals_out, int T, int V, int shift,
                                                                  const uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t cnt[kSortWarps][256];
  uint32_t key[kSortRounds];
  count_tile(ids, keys_in, T, V, shift, cnt, key);
  {
    uint32_t run = hist[(size_t)threadIdx.x * nblk + blockIdx.x];   // first output slot of (digit, this CTA)
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) { const uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = blockIdx.x * kSortTile + warp * 32 * kSortRounds;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const int idx = base + r * 32 + lane;
    const bool ok = idx < T;
    const uint32_t d = ok ? ((key[r] >> shift) & 255u) : (256u + lane);
    const unsigned m = __match_any_sync(0xffffffffu, d);
    uint32_t pos = 0;
    if (ok) pos = cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
    __syncwarp();
    if (ok) {
      keys_out[pos] = key[r];
      vals_out[pos] = ids ? (uint32_t)idx : vals_in[idx];
    }
  }
}

// long runs: {lo, hi, id}.  One CTA; every run spanning more than kMaxParts chunk boundaries contains a sample position.
__global__ void __launch_bounds__(1024) tg_long_runs_kernel(const uint32_t* __restrict__ sid, int T, int* __restrict__ out, int cap,
                                                            int kChunk) {
  const int kSample = sample_stride(kChunk);
  __shared__ int n;
  if (threadIdx.x == 0) n = 0;
  __syncthreads();
  for (int p = threadIdx.x * kSample; p < T; p += 1024 * kSample) {
    const uint32_t id = sid[p];
    int lo = 0, hi = p;                       // lower bound of id in [0, p]
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sid[mid] < id) lo = mid + 1; else hi = mid; }
    const int first = lo;
    if (p - kSample >= first) continue;       // an earlier sample of the same run reports it
    lo = p + 1; hi = T;                       // upper bound of id in (p, T]
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sid[mid] <= id) lo = mid + 1; else hi = mid; }
    const int end = lo;
    if ((end - 1) / kChunk - first / kChunk > kMaxParts) {
      const int k = atomicAdd(&n, 1);
      if (k < cap) { out[1 + 4 * k] = first; out[2 + 4 * k] = end; out[3 + 4 * k] = (int)id; out[4 + 4 * k] = 0; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = n < cap ? n : cap;
}

// Small batches (T <= kSmallSortTokens): the whole sort and the long-run list in ONE launch of one CTA -- a chain of seven small
// launches on the side branch costs the step more in scheduling gaps than the sort itself.  Warp w owns a contiguous slice of the
// keys; per pass: warp-private digit counts, offsets by (digit, warp), stable scatter.  The ping-pong buffers live in global memory
// (L2); they are read with ld.cg after a CTA barrier.
constexpr int kSmallSortTokens = 16384;
__global__ void __launch_bounds__(1024) tg_sort_small_kernel(const long long* __restrict__ ids, int T, int V, int passes, uint32_t* k0,
                                                             uint32_t* v0, uint32_t* k1, uint32_t* v1, int* __restrict__ long_runs,
                                                             int cap, int kChunk) {
  __shared__ uint32_t cnt[32][256];
  __shared__ uint32_t dsum[256];
  __shared__ int n_long;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = ((T + 31) / 32 + 31) / 32 * 32;
  const int base = warp * per_warp;
  for (int p = 0; p < passes; ++p) {
    const uint32_t* kin = (p & 1) ? k1 : k0;
    const uint32_t* vin = (p & 1) ? v1 : v0;
    uint32_t* kout = (p & 1) ? k0 : k1;
    uint32_t* vout = (p & 1) ? v0 : v1;
    const int shift = 8 * p;
    for (int i = threadIdx.x; i < 32 * 256; i += 1024) (&cnt[0][0])[i] = 0;
    __syncthreads();
    for (int r = 0; r < per_warp; r += 32) {
      const int idx = base + r + lane;
      const bool ok = idx < T;
      uint32_t key = 0;
      if (ok) {
        if (p == 0) { long long id = ids[idx]; id = id < 0 ? 0 : (id >= V ? V - 1 : id); key = (uint32_t)id; }
        else key = __ldcg(kin + idx);
      }
      const uint32_t d = ok ? ((key >> shift) & 255u) : (256u + lane);
      const unsigned m = __match_any_sync(0xffffffffu, d);
      if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t run = 0;
      for (int w = 0; w < 32; ++w) { const uint32_t c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += c; }
      dsum[threadIdx.x] = run;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t pre = 0;
      for (int d = 0; d < (int)threadIdx.x; ++d) pre += dsum[d];
      for (int w = 0; w < 32; ++w) cnt[w][threadIdx.x] += pre;
    }
    __syncthreads();
    for (int r = 0; r < per_warp; r += 32) {
      const int idx = base + r + lane;
      const bool ok = idx < T;
      uint32_t key = 0, val = (uint32_t)idx;
      if (ok) {
        if (p == 0) { long long id = ids[idx]; id = id < 0 ? 0 : (id >= V ? V - 1 : id); key = (uint32_t)id; }
        else { key = __ldcg(kin + idx); val = __ldcg(vin + idx); }
      }
      const uint32_t d = ok ? ((key >> shift) & 255u) : (256u + lane);
      const unsigned m = __match_any_sync(0xffffffffu, d);
      uint32_t pos = 0;
      if (ok) pos = cnt[warp][d] + __popc(m & ((1u << lane) - 1u));
      __syncwarp();
      if (ok && lane == __ffs(m) - 1) cnt[warp][d] += __popc(m);
      __syncwarp();
      if (ok) { __stcg(kout + pos, key); __stcg(vout + pos, val); }
    }
    __syncthreads();
  }
  const uint32_t* sid = (passes & 1) ? k1 : k0;
  const int kSample = sample_stride(kChunk);
  if (threadIdx.x == 0) n_long = 0;
  __syncthreads();
  for (int p = threadIdx.x * kSample; p < T; p += 1024 * kSample) {
    const uint32_t id = __ldcg(sid + p);
    int lo = 0, hi = p;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldcg(sid + mid) < id) lo = mid + 1; else hi = mid; }
    const int first = lo;
    if (p - kSample >= first) continue;
    lo = p + 1; hi = T;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldcg(sid + mid) <= id) lo = mid + 1; else hi = mid; }
    const int end = lo;
    if ((end - 1) / kChunk - first / kChunk > kMaxParts) {
      const int k = atomicAdd(&n_long, 1);
      if (k < cap) { long_runs[1 + 4 * k] = first; long_runs[2 + 4 * k] = end; long_runs[3 + 4 * k] = (int)id; long_runs[4 + 4 * k] = 0; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) long_runs[0] = n_long < cap ? n_long : cap;
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// grid = ceil(chunks / groups per CTA), 256 threads; a group of H/4 threads owns one chunk, a thread 4 columns of it
template <int H, int kChunk>
__global__ void __launch_bounds__(256) tg_chunk_sum_kernel(const uint32_t* __restrict__ sid, const uint32_t* __restrict__ perm,
                                                           const float* __restrict__ dx, float* table, float* __restrict__ carry, int T) {
  pdl_grid_wait();
  constexpr int LPR = H / 4, GPB = 256 / LPR;
  const int g = threadIdx.x / LPR, col = (threadIdx.x % LPR) * 4;
  const int c = blockIdx.x * GPB + g;
  const int start = c * kChunk;
  if (start >= T) return;
  const int end = min(start + kChunk, T);
  const bool head_cont = start > 0 && sid[start - 1] == sid[start];
  const bool tail_cont = end < T && sid[end] == sid[end - 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool first_run = true;
  constexpr int U = kChunk == 16 ? 16 : 8;
  for (int i = start; i < end; i += U) {
    // one round trip for the ids / token indices of U tokens, one for their dx rows and for the table rows of the runs that end
    // among them (fetched ahead so that finishing a run is an add and a store, not a dependent read-modify-write)
    uint32_t id[U + 1], tk[U];
    float4 r[U], tb[U];
#pragma unroll
    for (int u = 0; u <= U; ++u) id[u] = sid[min(i + u, T - 1)];
#pragma unroll
    for (int u = 0; u < U; ++u) tk[u] = perm[min(i + u, end - 1)];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = *reinterpret_cast<const float4*>(dx + (size_t)tk[u] * H + col);
      const bool ends = i + u < end && (i + u == end - 1 || id[u + 1] != id[u]);
      tb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ends) tb[u] = *reinterpret_cast<const float4*>(table + (size_t)id[u] * H + col);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u < end) {
        add4(acc, r[u]);
        const bool last = i + u == end - 1;
        if (last || id[u + 1] != id[u]) {               // the run of item id[u] ends here
          float* dst;
          if (first_run && head_cont) dst = carry + ((size_t)c * 2 + 0) * H + col;
          else if (last && tail_cont) dst = carry + ((size_t)c * 2 + 1) * H + col;
          else {
            dst = table + (size_t)id[u] * H + col;
            add4(acc, tb[u]);
          }
          *reinterpret_cast<float4*>(dst) = acc;
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
          first_run = false;
        }
      }
    }
  }
}

// CTAs [0, nA): a group of H/4 threads per chunk finishes the run that STARTS in its chunk and continues into the next ones
// (tail carry of the chunk + head carries of the following chunks, at most kMaxParts of them).  CTAs [nA, ...): one long run each.
template <int H, int kChunk>
__global__ void __launch_bounds__(1024) tg_boundary_kernel(const uint32_t* __restrict__ sid, const float* __restrict__ carry,
                                                           const int* __restrict__ long_runs, float* table, int T, int nA) {
  pdl_grid_wait();
  constexpr int LPR = H / 4, GPB = 1024 / LPR;
  __shared__ float4 part[GPB][LPR];
  const int g = threadIdx.x / LPR, lane = threadIdx.x % LPR, col = lane * 4;
  const int nchunks = (T + kChunk - 1) / kChunk;
  if ((int)blockIdx.x < nA) {
    const int c = blockIdx.x * GPB + g;
    if (c >= nchunks - 1) return;                               // the last chunk has no successor
    const int start = c * kChunk, end = start + kChunk;
    const uint32_t id = sid[end - 1], nxt = sid[end], fst = sid[start], prv = start > 0 ? sid[start - 1] : ~0u;
    if (nxt != id) return;                                      // tail run does not continue
    if (fst == id && prv == id) return;                         // pass-through chunk: the run started earlier
    // the table row and the first two parts are needed in any case: fetch them while the extent of the run is being found
    float* dst = table + (size_t)id * H + col;
    const float4 tb = *reinterpret_cast<const float4*>(dst);
    float4 acc = *reinterpret_cast<const float4*>(carry + ((size_t)c * 2 + 1) * H + col);
    const float4 h1 = *reinterpret_cast<const float4*>(carry + ((size_t)(c + 1) * 2 + 0) * H + col);
    int m = 0;                                                  // head carries to add: chunks c+1 .. c+m
    for (;;) {
      ++m;
      const int k = c + m, ks = k * kChunk, ke = min(ks + kChunk, T);
      const bool through = ke < T && sid[ke - 1] == id && sid[ke] == id;
      if (!through) break;
      if (m > kMaxParts) return;                                // long run: a whole CTA sums it
    }
    if (m > kMaxParts) return;
    add4(acc, h1);
    for (int i = 2; i <= m; ++i) add4(acc, *reinterpret_cast<const float4*>(carry + ((size_t)(c + i) * 2 + 0) * H + col));
    add4(acc, tb);
    *reinterpret_cast<float4*>(dst) = acc;
    return;
  }
  const int r = blockIdx.x - nA;
  if (r >= long_runs[0]) return;
  const int first = long_runs[1 + 4 * r], endt = long_runs[2 + 4 * r], id = long_runs[3 + 4 * r];
  const int c0 = first / kChunk, c1 = (endt - 1) / kChunk;
  const int n = c1 - c0 + 1;                                    // part 0 = tail carry of c0, part i = head carry of c0 + i
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kU = 8;
  for (int i = g; i < n; i += GPB * kU) {
    float4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int p = i + u * GPB;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < n) v[u] = *reinterpret_cast<const float4*>(carry + ((size_t)(c0 + p) * 2 + (p == 0 ? 1 : 0)) * H + col);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) add4(acc, v[u]);
  }
  part[g][lane] = acc;
  __syncthreads();
  if (g == 0) {
    float4 tot = part[0][lane];
    for (int k = 1; k < GPB; ++k) add4(tot, part[k][lane]);
    float* dst = table + (size_t)id * H + col;
    add4(tot, *reinterpret_cast<const float4*>(dst));
    *reinterpret_cast<float4*>(dst) = tot;
  }
}

int sort_passes(int V) {
  int bits = 1;
  while (bits < 32 && (1ll << bits) < (long long)V) ++bits;
  return (bits + 7) / 8;
}
}  // namespace

int table_grad_sort_blocks(int T) { return (T + kSortTile - 1) / kSortTile; }
int table_grad_chunk_tokens(int T) { return T < 65536 ? 16 : 64; }
int table_grad_chunks(int T) { const int c = table_grad_chunk_tokens(T); return (T + c - 1) / c; }
int table_grad_max_long_runs(int T) { return T / (kMaxParts * table_grad_chunk_tokens(T)) + 1; }
int table_grad_sorted_buf(int V) { return sort_passes(V) & 1; }
int table_grad_sort_launches(int T, int V) { return T <= kSmallSortTokens ? 1 : 3 * sort_passes(V) + 1; }

cudaError_t launch_token_sort(const TableGradArgs& a, cudaStream_t st) {
  const int nblk = table_grad_sort_blocks(a.T);
  const int passes = sort_passes(a.V);
  if (a.T <= kSmallSortTokens) {
    tg_sort_small_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const long long*>(a.ids), a.T, a.V, passes, a.keys[0], a.vals[0], a.keys[1],
                                             a.vals[1], a.long_runs, table_grad_max_long_runs(a.T), table_grad_chunk_tokens(a.T));
    return cudaGetLastError();
  }
  for (int p = 0; p < passes; ++p) {
    const long long* ids = p == 0 ? reinterpret_cast<const long long*>(a.ids) : nullptr;
    const uint32_t* kin = a.keys[p & 1];
    const uint32_t* vin = a.vals[p & 1];
    tg_hist_kernel<<<nblk, kSortThreads, 0, st>>>(ids, kin, a.T, a.V, 8 * p, a.hist, nblk);
    tg_scan_kernel<<<1, 1024, 0, st>>>(a.hist, 256 * nblk);
    tg_scatter_kernel<<<nblk, kSortThreads, 0, st>>>(ids, kin, vin, a.keys[(p + 1) & 1], a.vals[(p + 1) & 1], a.T, a.V, 8 * p, a.hist, nblk);
  }
  tg_long_runs_kernel<<<1, 1024, 0, st>>>(a.keys[passes & 1], a.T, a.long_runs, table_grad_max_long_runs(a.T), table_grad_chunk_tokens(a.T));
  return cudaGetLastError();
}

template <int H, int CH>
static cudaError_t table_grad_h(const TableGradArgs& a, cudaStream_t st) {
  const int fin = table_grad_sorted_buf(a.V);
  const int chunks = table_grad_chunks(a.T);
  constexpr int LPR = H / 4;
  const int g1 = (chunks + 256 / LPR - 1) / (256 / LPR);
  cudaError_t e = launch_pdl(tg_chunk_sum_kernel<H, CH>, dim3(g1), dim3(256), (size_t)0, st, (const uint32_t*)a.keys[fin],
                             (const uint32_t*)a.vals[fin], a.dx, a.grad_table, a.carry, a.T);
  if (e != cudaSuccess) return e;
  const int nA = (chunks + 1024 / LPR - 1) / (1024 / LPR);
  return launch_pdl(tg_boundary_kernel<H, CH>, dim3(nA + table_grad_max_long_runs(a.T)), dim3(1024), (size_t)0, st,
                    (const uint32_t*)a.keys[fin], (const float*)a.carry, (const int*)a.long_runs, a.grad_table, a.T, nA);
}

cudaError_t launch_table_grad(const TableGradArgs& a, cudaStream_t st) {
  const bool small = table_grad_chunk_tokens(a.T) == 16;
  switch (a.H) {
    case 64: return small ? table_grad_h<64, 16>(a, st) : table_grad_h<64, 64>(a, st);
    case 128: return small ? table_grad_h<128, 16>(a, st) : table_grad_h<128, 64>(a, st);
    case 256: return small ? table_grad_h<256, 16>(a, st) : table_grad_h<256, 64>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace b4r
