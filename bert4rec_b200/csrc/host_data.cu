// Host data path of the hot loop: Cloze masking and negative sampling for a whole batch, in C++ threads,
// BIT-EXACT with the reference's Python:
//   * apply_dynamic_masking_task + process_element layout   (dataloader_utils.py:186-261, bert4rec_preprocessor.py:47-116)
//     -- CPython's `random` module: MT19937 seeded by init_by_array(seed words), random() = 53-bit from two draws,
//        shuffle / choice through _randbelow (rejection on getrandbits(bit_length(n)));
//   * RandomSampler.sample / PopularSampler.sample / PopularRandomSampler.sample
//     (random_sampler.py:63-79, popular_sampler.py:53-71, popular_random_sampler.py:77-117)
//     -- numpy's LEGACY RandomState: np.random.seed(int) = init_genrand, choice(replace=False) = permutation(n)[:size] with the
//        masked-rejection `random_interval`, choice(replace=True) = masked-rejection randint, choice(p=...) = cumsum / searchsorted.
// No GPU involved; nothing here touches CUDA.  Errors mirror the ValueErrors of the Python code (b4r_last_error()).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <unistd.h>
#include <map>
#include <mutex>
#include <tuple>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/b4r.h"

namespace b4r { void set_last_error(const char* msg); }

namespace {

struct MT19937 {
  uint32_t mt[624];
  int pos;
  void init_genrand(uint32_t s) {   // numpy mt19937_seed / reference MT init
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    pos = 624;
  }
  void init_by_array(const uint32_t* key, int len) {   // CPython random_seed for an int
    static const MT19937 base = [] { MT19937 b; b.init_genrand(19650218u); return b; }();   // constant first stage: computed once
    memcpy(mt, base.mt, sizeof(mt));
    int i = 1, j = 0;
    for (int k = 624 > len ? 624 : len; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
      ++i; ++j;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
      if (j >= len) j = 0;
    }
    for (int k = 623; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
      ++i;
      if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    pos = 624;
  }
  // The state is twisted LAZILY, one word per draw, in place and in index order: word k of the next generation depends on the old
  // words k, k+1 and on word k+397 (old for k < 227, already new for k >= 227), exactly the values the block refill of the
  // reference implementation sees when it reaches k.  A masked sequence draws ~100 numbers, not 624: most of the twist is never done.
  uint32_t next32() {
    static const uint32_t mag[2] = {0u, 0x9908b0dfu};
    if (pos >= 624) pos = 0;
    const int k = pos, k1 = k == 623 ? 0 : k + 1, km = k < 227 ? k + 397 : k - 227;
    const uint32_t t = (mt[k] & 0x80000000u) | (mt[k1] & 0x7fffffffu);
    uint32_t y = mt[k] = mt[km] ^ (t >> 1) ^ mag[t & 1u];
    ++pos;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double next_double() {   // CPython random() and numpy random_sample(): identical construction
    const uint32_t a = next32() >> 5, b = next32() >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
};

// ---- CPython `random`
inline void py_seed(MT19937& g, uint64_t seed) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  g.init_by_array(key, key[1] ? 2 : 1);
}
// Four generators seeded in lockstep: init_by_array is a chain of ~1250 dependent multiply-xor steps (the whole cost of masking a
// short sequence is this latency chain); four independent chains interleaved in one loop keep the multiplier busy.
inline void py_seed4(MT19937* g, const uint64_t* seed, int count) {
  static const MT19937 base = [] { MT19937 b; b.init_genrand(19650218u); return b; }();
  uint32_t key[4][2];
  int len[4], j[4];
  for (int q = 0; q < 4; ++q) {
    const uint64_t sd = seed[q < count ? q : 0];
    key[q][0] = (uint32_t)sd; key[q][1] = (uint32_t)(sd >> 32);
    len[q] = key[q][1] ? 2 : 1; j[q] = 0;
    memcpy(g[q].mt, base.mt, sizeof(base.mt));
  }
  uint32_t* __restrict__ m0 = g[0].mt; uint32_t* __restrict__ m1 = g[1].mt;
  uint32_t* __restrict__ m2 = g[2].mt; uint32_t* __restrict__ m3 = g[3].mt;
  uint32_t p0 = m0[0], p1 = m1[0], p2 = m2[0], p3 = m3[0];   // mt[i-1] of every chain, carried in registers
  int i = 1;
  for (int k = 624; k; --k) {
    p0 = m0[i] = (m0[i] ^ ((p0 ^ (p0 >> 30)) * 1664525u)) + key[0][j[0]] + (uint32_t)j[0];
    p1 = m1[i] = (m1[i] ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + key[1][j[1]] + (uint32_t)j[1];
    p2 = m2[i] = (m2[i] ^ ((p2 ^ (p2 >> 30)) * 1664525u)) + key[2][j[2]] + (uint32_t)j[2];
    p3 = m3[i] = (m3[i] ^ ((p3 ^ (p3 >> 30)) * 1664525u)) + key[3][j[3]] + (uint32_t)j[3];
    for (int q = 0; q < 4; ++q)
      if (++j[q] >= len[q]) j[q] = 0;
    if (++i >= 624) { m0[0] = p0; m1[0] = p1; m2[0] = p2; m3[0] = p3; i = 1; }
  }
  for (int k = 623; k; --k) {
    p0 = m0[i] = (m0[i] ^ ((p0 ^ (p0 >> 30)) * 1566083941u)) - (uint32_t)i;
    p1 = m1[i] = (m1[i] ^ ((p1 ^ (p1 >> 30)) * 1566083941u)) - (uint32_t)i;
    p2 = m2[i] = (m2[i] ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - (uint32_t)i;
    p3 = m3[i] = (m3[i] ^ ((p3 ^ (p3 >> 30)) * 1566083941u)) - (uint32_t)i;
    if (++i >= 624) { m0[0] = p0; m1[0] = p1; m2[0] = p2; m3[0] = p3; i = 1; }
  }
  for (int q = 0; q < 4; ++q) { g[q].mt[0] = 0x80000000u; g[q].pos = 624; }
}
inline int bit_length(uint64_t n) { return n ? 64 - __builtin_clzll(n) : 0; }
inline uint64_t py_randbelow(MT19937& g, uint64_t n) {   // n >= 1, n < 2^32 here
  const int k = bit_length(n);
  uint64_t r;
  if (k <= 32) {
    do { r = g.next32() >> (32 - k); } while (r >= n);
  } else {   // getrandbits(k) for 32 < k <= 64: low word first
    do {
      const uint64_t lo = g.next32();
      const uint64_t hi = g.next32() >> (64 - k);
      r = lo | (hi << 32);
    } while (r >= n);
  }
  return r;
}

// ---- numpy legacy RandomState
inline uint64_t np_interval(MT19937& g, uint64_t max) {   // random_interval: uniform on [0, max]
  if (max == 0) return 0;
  uint64_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
  uint64_t v;
  if (max <= 0xffffffffull) {
    while ((v = (g.next32() & mask)) > max) {}
  } else {
    while ((v = ((((uint64_t)g.next32()) << 32 | g.next32()) & mask)) > max) {}
  }
  return v;
}

// Persistent worker pool: a batch call is a few hundred microseconds of work, a std::thread start costs tens of microseconds, so
// the workers are created once (lazily, re-created in a forked child) and woken per job.  One job at a time (callers serialise on
// `run_mu`); the calling thread works too.  The pool is never destroyed (detached workers die with the process).
class Pool {
 public:
  static Pool& get() {
    static std::mutex mu;
    static Pool* pool = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (!pool || owner != getpid()) {   // first use, or a fork()ed child (threads do not survive fork)
      pool = new Pool();
      owner = getpid();
    }
    return *pool;
  }
  // runs job() on up to `threads` threads (including the caller) and returns when all of them are done
  void run(int threads, const std::function<void()>& job) {
    std::lock_guard<std::mutex> serial(run_mu_);
    const int helpers = std::min(threads - 1, (int)workers_);
    if (helpers <= 0) { job(); return; }
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &job; wanted_ = helpers; running_ = helpers; ++gen_;
    }
    cv_.notify_all();
    job();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return running_ == 0; });
    job_ = nullptr;
  }
  int size() const { return (int)workers_ + 1; }

 private:
  Pool() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (hw > 64) hw = 64;
    workers_ = hw - 1;
    for (unsigned i = 0; i < workers_; ++i) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void()>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (wanted_ > 0) { --wanted_; job = job_; }   // only `helpers` workers take part in this job
      }
      if (!job) continue;
      (*job)();
      std::lock_guard<std::mutex> lk(mu_);
      if (--running_ == 0) done_.notify_all();
    }
  }
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void()>* job_ = nullptr;
  unsigned long gen_ = 0;
  int wanted_ = 0, running_ = 0;
  unsigned workers_ = 0;
};

template <class F>
int run_parallel(int n, int n_threads, F&& body, int grain = 1) {   // body(i) -> 0 or error; first error wins
  if (n_threads < 1) n_threads = Pool::get().size();
  if (grain > 1 && n_threads > (n + grain - 1) / grain) n_threads = (n + grain - 1) / grain;   // cheap items: a thread per `grain` of them
  if (n_threads > n) n_threads = n > 0 ? n : 1;
  std::atomic<int> next(0), err(0);
  const int chunk = n >= 16 * n_threads ? 16 : (n + n_threads - 1) / n_threads > 0 ? (n + n_threads - 1) / n_threads : 1;
  auto worker = [&]() {
    for (;;) {
      const int i0 = next.fetch_add(chunk);
      if (i0 >= n || err.load()) return;
      const int i1 = i0 + chunk < n ? i0 + chunk : n;
      for (int i = i0; i < i1; ++i) {
        const int e = body(i);
        if (e) { err.store(e); return; }
      }
    }
  };
  if (n_threads <= 1) { worker(); return err.load(); }
  Pool::get().run(n_threads, worker);
  return err.load();
}

// first error of a call (worker threads race for it); copied into the caller's thread-local b4r_last_error() at the end
struct Err {
  std::atomic<bool> set{false};
  char msg[400];
  int fail(const char* fmt, ...) {
    char buf[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    bool expect = false;
    if (set.compare_exchange_strong(expect, true)) memcpy(msg, buf, sizeof(buf));
    return 1;
  }
  int finish(int rc) {
    if (rc) b4r::set_last_error(set.load() ? msg : "host data path failed");
    return rc;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------ Cloze masking
extern "C" int b4r_host_cloze_mask_batch(const int64_t* tokens, const int64_t* offsets, const uint64_t* seeds, int n, int max_seq_len,
                                         int max_pred, int64_t mask_id, int64_t pad_id, const int64_t* special_ids, int n_special,
                                         int64_t vocab_size, double selection_rate, double mask_token_rate, double random_token_rate,
                                         int64_t* labels, int64_t* input_word_ids, int64_t* input_mask, int64_t* masked_lm_ids,
                                         int64_t* masked_lm_positions, int64_t* masked_lm_weights, int n_threads) {
  Err err;
  if (!tokens || !offsets || !seeds || n < 0 || max_seq_len < 1 || max_pred < 1 || n_special < 0 || (n_special && !special_ids))
    return err.finish(err.fail("b4r_host_cloze_mask_batch: bad argument"));
  if (!labels || !input_word_ids || !input_mask || !masked_lm_ids || !masked_lm_positions || !masked_lm_weights)
    return err.finish(err.fail("b4r_host_cloze_mask_batch: null output"));
  // selectable vocab = [i for i in range(vocab_size) if i not in special]: k-th element by skipping the sorted specials
  std::vector<int64_t> sp;
  for (int i = 0; i < n_special; ++i)
    if (special_ids[i] >= 0 && special_ids[i] < vocab_size) sp.push_back(special_ids[i]);
  std::sort(sp.begin(), sp.end());
  sp.erase(std::unique(sp.begin(), sp.end()), sp.end());
  const int64_t n_selectable = vocab_size - (int64_t)sp.size();
  const double threshold = mask_token_rate + random_token_rate;
  const int S = max_seq_len, P = max_pred;
  auto is_special = [&](int64_t v) {
    for (int i = 0; i < n_special; ++i)
      if (special_ids[i] == v) return true;
    return false;
  };
  const int rc = run_parallel((n + 3) / 4, n_threads, [&](int grp) -> int {
   MT19937 gens[4];
   const int b0 = grp * 4, cnt = n - b0 < 4 ? n - b0 : 4;
   py_seed4(gens, seeds + b0, cnt);
   for (int bq = 0; bq < cnt; ++bq) {
    const int b = b0 + bq;
    MT19937& g = gens[bq];
    const int64_t* seq = tokens + offsets[b];
    const int64_t len64 = offsets[b + 1] - offsets[b];
    if (len64 < 0 || len64 > S) return err.fail("sequence %d has %lld tokens, more than max_seq_len %d", b, (long long)len64, S);
    const int len = (int)len64;
    int64_t* lab = labels + (size_t)b * S; int64_t* ids = input_word_ids + (size_t)b * S; int64_t* msk = input_mask + (size_t)b * S;
    int64_t* mid = masked_lm_ids + (size_t)b * P; int64_t* mpos = masked_lm_positions + (size_t)b * P;
    int64_t* mw = masked_lm_weights + (size_t)b * P;
    for (int i = 0; i < S; ++i) {
      const bool in = i < len;
      lab[i] = in ? seq[i] : pad_id; ids[i] = in ? seq[i] : pad_id; msk[i] = in ? 1 : pad_id;
    }
    for (int i = 0; i < P; ++i) { mid[i] = pad_id; mpos[i] = pad_id; mw[i] = pad_id; }
    int n_plain = 0;
    for (int i = 0; i < len; ++i) n_plain += is_special(seq[i]) ? 0 : 1;
    int n_pred = (int)((double)n_plain * selection_rate);   // int(len * rate): truncation of the double product
    if (n_pred < 1) n_pred = 1;
    if (n_pred > P) n_pred = P;
    // random.shuffle(list(range(n_plain)))
    int order_stack[512];
    std::vector<int> order_heap;
    int* order = order_stack;
    if (n_plain > 512) { order_heap.resize(n_plain); order = order_heap.data(); }
    for (int i = 0; i < n_plain; ++i) order[i] = i;
    for (int i = n_plain - 1; i >= 1; --i) {
      const int j = (int)py_randbelow(g, (uint64_t)i + 1);
      const int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    const int n_sel = n_pred < n_plain ? n_pred : n_plain;
    std::sort(order, order + n_sel);
    for (int k = 0; k < n_sel; ++k) {
      const int idx = order[k];
      const double rn = g.next_double();
      int64_t token = seq[idx];
      if (rn < threshold) {
        if (n_selectable < 1) return err.fail("Cannot choose from an empty sequence (no selectable vocabulary)");
        int64_t v = (int64_t)py_randbelow(g, (uint64_t)n_selectable);   // random.choice(selectable_vocab)
        for (size_t s = 0; s < sp.size(); ++s) {
          if (sp[s] <= v) ++v; else break;
        }
        token = v;
      }
      if (rn < mask_token_rate) token = mask_id;
      ids[idx] = token;
      mid[k] = seq[idx]; mpos[k] = idx; mw[k] = 1;
    }
   }
    return 0;
  });
  return err.finish(rc);
}

// ------------------------------------------------------------------------------------------------ negative sampling
namespace {
// np.random.choice(pool, size, replace) without p, after np.random.seed(seed)
int np_choice_uniform(Err& err, MT19937& g, const std::vector<int64_t>& pool, int size, bool replace, int64_t* out,
                      std::vector<int64_t>& perm) {
  const int64_t n = (int64_t)pool.size();
  if (n == 0 && size > 0) return err.fail("'a' cannot be empty unless no samples are taken");
  if (replace) {
    for (int i = 0; i < size; ++i) out[i] = pool[(size_t)np_interval(g, (uint64_t)n - 1)];   // randint: masked rejection
    return 0;
  }
  if (size > n) return err.fail("Cannot take a larger sample than population when 'replace=False'");
  perm.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) perm[(size_t)i] = i;
  if (n - 1 <= 0xffffffffll) {
    // legacy shuffle of arange(n): j = random_interval(i) = masked rejection on 32-bit draws; the mask (smallest 2^k - 1 >= i)
    // only changes when i crosses a power of two, so it is carried instead of recomputed
    uint32_t mask = 0;
    for (uint64_t m = (uint64_t)(n - 1); m; m >>= 1) mask = (mask << 1) | 1u;
    for (int64_t i = n - 1; i >= 1; --i) {
      while ((mask >> 1) >= (uint64_t)i) mask >>= 1;
      uint32_t j;
      while ((j = (g.next32() & mask)) > (uint64_t)i) {}
      const int64_t t = perm[(size_t)i]; perm[(size_t)i] = perm[j]; perm[j] = t;
    }
  } else {
    for (int64_t i = n - 1; i >= 1; --i) {
      const int64_t j = (int64_t)np_interval(g, (uint64_t)i);
      const int64_t t = perm[(size_t)i]; perm[(size_t)i] = perm[(size_t)j]; perm[(size_t)j] = t;
    }
  }
  for (int i = 0; i < size; ++i) out[i] = pool[(size_t)perm[(size_t)i]];
  return 0;
}
}  // namespace

// RandomSampler.sample(without=..., seed=...) for n requests: pool = vocab minus without (vocab order kept), one legacy choice.
// out int64 [n][sample_size].
extern "C" int b4r_host_sample_random_batch(const int64_t* vocab, int64_t n_vocab, const int64_t* without, const int64_t* without_off,
                                            const uint32_t* seeds, int n, int sample_size, int allow_duplicates, int64_t* out,
                                            int n_threads) {
  Err err;
  if (!vocab || n_vocab < 0 || !seeds || n < 0 || sample_size < 0 || !out || (without && !without_off))
    return err.finish(err.fail("b4r_host_sample_random_batch: bad argument"));
  if (!allow_duplicates && sample_size > n_vocab)
    return err.finish(err.fail("When no duplicates are allowed in the final sample then the sample size (given sample size: %d)) can "
                            "not be greater than the length length of the vocab (length of the vocab: %lld)", sample_size,
                            (long long)n_vocab));
  // small non-negative ids (the tokenizer's): exclusion by a per-thread mark array instead of a hash set
  int64_t vmin = 0, vmax = -1;
  for (int64_t i = 0; i < n_vocab; ++i) {
    if (i == 0 || vocab[i] < vmin) vmin = vocab[i];
    if (i == 0 || vocab[i] > vmax) vmax = vocab[i];
  }
  const bool dense = n_vocab > 0 && vmin >= 0 && vmax < (1ll << 26);
  // Fast path (the reproducible-evaluation case: one fixed seed for every call, no duplicates, unique small ids): the legacy
  // permutation depends on (seed, pool size) only, so it is shuffled ONCE per distinct pool size of the batch and its first
  // sample_size entries are mapped through each request's pool by rank-select over the excluded positions -- the same draws as the
  // per-request full shuffle, O(|without| + sample_size * |without|) instead of O(V) per request.
  bool same_seed = n > 0 && !allow_duplicates && dense && sample_size > 0;
  for (int b = 1; same_seed && b < n; ++b) same_seed = seeds[b] == seeds[0];
  if (same_seed) {
    std::vector<int32_t> pos_of((size_t)vmax + 1, -1);
    bool unique = true;
    for (int64_t i = 0; i < n_vocab && unique; ++i) {
      if (pos_of[(size_t)vocab[i]] >= 0) unique = false;
      pos_of[(size_t)vocab[i]] = (int32_t)i;
    }
    if (unique && n_vocab < (1ll << 31)) {
      const int64_t tot_wo = without ? without_off[n] - without_off[0] : 0;
      std::vector<int32_t> ex_pos((size_t)tot_wo);   // per request: sorted unique excluded vocab positions, at without_off[b] - without_off[0]
      std::vector<int32_t> ex_cnt((size_t)n, 0);
      const int kGrain = 32;   // requests per thread for the cheap per-request phases
      int rc = run_parallel(n, n_threads, [&](int b) -> int {
        int32_t* e = ex_pos.data() + (without ? without_off[b] - without_off[0] : 0);
        int m = 0;
        if (without)
          for (int64_t k = without_off[b]; k < without_off[b + 1]; ++k) {
            const int64_t wv = without[k];
            if (wv >= 0 && wv <= vmax && pos_of[(size_t)wv] >= 0) e[m++] = pos_of[(size_t)wv];
          }
        std::sort(e, e + m);
        m = (int)(std::unique(e, e + m) - e);
        ex_cnt[(size_t)b] = m;
        for (int q = 0; q < m; ++q) e[q] -= q;   // thresholds thr[q] = e[q] - q (see the mapping below)
        if (sample_size > n_vocab - m) return err.fail("Cannot take a larger sample than population when 'replace=False'");
        return 0;
      }, kGrain);
      if (rc) return err.finish(rc);
        // distinct pool sizes -> shuffled prefixes
      std::vector<int64_t> sizes;
      for (int b = 0; b < n; ++b) sizes.push_back(n_vocab - ex_cnt[(size_t)b]);
      std::sort(sizes.begin(), sizes.end());
      sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
      // process-wide cache of the shuffled prefixes, keyed by (seed, pool size, sample size): an evaluation loop calls this once
      // per batch with the same sampler, so after the first batches no shuffle is left to do
      static std::mutex cache_mu;
      static std::map<std::tuple<uint32_t, int64_t, int>, std::vector<int64_t>> cache;
      std::vector<std::vector<int64_t>> prefix(sizes.size());
      std::vector<int> todo;
      {
        std::lock_guard<std::mutex> lk(cache_mu);
        for (size_t si = 0; si < sizes.size(); ++si) {
          auto it = cache.find(std::make_tuple(seeds[0], sizes[si], sample_size));
          if (it != cache.end()) prefix[si] = it->second; else todo.push_back((int)si);
        }
      }
        rc = run_parallel((int)todo.size(), n_threads, [&](int ti) -> int {
        thread_local std::vector<int64_t> ident, perm;
        const size_t si = (size_t)todo[(size_t)ti];
        const int64_t m = sizes[si];
        ident.resize((size_t)m);
        for (int64_t i = 0; i < m; ++i) ident[(size_t)i] = i;
        MT19937 g;
        g.init_genrand(seeds[0]);
        prefix[si].resize((size_t)sample_size);
        return np_choice_uniform(err, g, ident, sample_size, false, prefix[si].data(), perm);
      });
      if (rc) return err.finish(rc);
      if (!todo.empty()) {
        std::lock_guard<std::mutex> lk(cache_mu);
        if (cache.size() > 8192) cache.clear();
        for (int si : todo) cache[std::make_tuple(seeds[0], sizes[(size_t)si], sample_size)] = prefix[(size_t)si];
      }
        rc = run_parallel(n, n_threads, [&](int b) -> int {
        const int m = ex_cnt[(size_t)b];
        const int32_t* e = ex_pos.data() + (without ? without_off[b] - without_off[0] : 0);
        const size_t si = (size_t)(std::lower_bound(sizes.begin(), sizes.end(), n_vocab - m) - sizes.begin());
        const int64_t* pf = prefix[si].data();
        int64_t* o = out + (size_t)b * sample_size;
        for (int k = 0; k < sample_size; ++k) {
          // index into the pool -> position in the vocab: the q-th excluded position e[q] precedes pool index i iff e[q] - q <= i
          // (thr[q] = e[q] - q is non-decreasing), so position = i + #{q : thr[q] <= i}
          const int64_t i = pf[k];
          int cnt = 0;
          for (int q = 0; q < m; ++q) cnt += e[q] <= (int32_t)i;   // branch-free (vectorised) count over <= a few dozen thresholds
          o[k] = vocab[(size_t)(i + cnt)];
        }
        return 0;
      }, kGrain);
        return err.finish(rc);
    }
  }
  const int rc = run_parallel(n, n_threads, [&](int b) -> int {
    thread_local std::vector<int64_t> pool, perm;
    thread_local std::unordered_set<int64_t> ex;
    thread_local std::vector<uint8_t> mark;
    pool.clear();
    if (without && without_off[b + 1] > without_off[b] && dense) {
      if ((int64_t)mark.size() < vmax + 1) mark.assign((size_t)vmax + 1, 0);
      for (int64_t k = without_off[b]; k < without_off[b + 1]; ++k)
        if (without[k] >= 0 && without[k] <= vmax) mark[(size_t)without[k]] = 1;
      pool.reserve((size_t)n_vocab);
      for (int64_t i = 0; i < n_vocab; ++i)
        if (!mark[(size_t)vocab[i]]) pool.push_back(vocab[i]);
      for (int64_t k = without_off[b]; k < without_off[b + 1]; ++k)
        if (without[k] >= 0 && without[k] <= vmax) mark[(size_t)without[k]] = 0;
    } else if (without && without_off[b + 1] > without_off[b]) {
      ex.clear();
      ex.insert(without + without_off[b], without + without_off[b + 1]);
      for (int64_t i = 0; i < n_vocab; ++i)
        if (!ex.count(vocab[i])) pool.push_back(vocab[i]);
    } else {
      pool.assign(vocab, vocab + n_vocab);
    }
    MT19937 g;
    g.init_genrand(seeds[b]);
    return np_choice_uniform(err, g, pool, sample_size, allow_duplicates != 0, out + (size_t)b * sample_size, perm);
  });
  return err.finish(rc);
}

// PopularSampler.sample(without=...) with a pre-ranked source (set_source ranks by popularity once): the first sample_size items
// of the ranked list that are not in `without`.  out int64 [n][sample_size], out_len int32 [n] (the list can run short).
extern "C" int b4r_host_sample_popular_batch(const int64_t* ranked, int64_t n_ranked, const int64_t* without, const int64_t* without_off,
                                             int n, int sample_size, int64_t* out, int32_t* out_len, int n_threads) {
  Err err;
  if (!ranked || n_ranked < 0 || n < 0 || sample_size < 0 || !out || !out_len || (without && !without_off))
    return err.finish(err.fail("b4r_host_sample_popular_batch: bad argument"));
  const int rc = run_parallel(n, n_threads, [&](int b) -> int {
    thread_local std::unordered_set<int64_t> ex;
    ex.clear();
    if (without) ex.insert(without + without_off[b], without + without_off[b + 1]);
    int k = 0;
    int64_t* o = out + (size_t)b * sample_size;
    for (int64_t i = 0; i < n_ranked && k < sample_size; ++i)
      if (!ex.count(ranked[i])) o[k++] = ranked[i];
    out_len[b] = k;
    for (int i = k; i < sample_size; ++i) o[i] = 0;
    return 0;
  });
  return err.finish(rc);
}

// PopularRandomSampler.sample(without=..., seed=...): np.random.choice(vocab, sample_size + |set(without)|, replace, p) then the
// items of `without` are removed and the first sample_size kept (popular_random_sampler.py:96-117).  out_len: the list can
// run short when removed items were drawn.  p: the sampler's probability_distribution (float64 [n_vocab]).
extern "C" int b4r_host_sample_pop_random_batch(const int64_t* vocab, const double* p, int64_t n_vocab, const int64_t* without,
                                                const int64_t* without_off, const uint32_t* seeds, int n, int sample_size,
                                                int allow_duplicates, int64_t* out, int32_t* out_len, int n_threads) {
  Err err;
  if (!vocab || !p || n_vocab < 1 || !seeds || n < 0 || sample_size < 0 || !out || !out_len || (without && !without_off))
    return err.finish(err.fail("b4r_host_sample_pop_random_batch: bad argument"));
  // numpy's checks on p (mtrand.pyx choice): non-negative, no NaN, sums to 1 within sqrt(eps) (Kahan sum)
  {
    double sum = 0.0, c = 0.0;
    for (int64_t i = 0; i < n_vocab; ++i) {
      if (!(p[i] >= 0.0)) return err.finish(err.fail(p[i] != p[i] ? "probabilities contain NaN" : "probabilities are not non-negative"));
      const double y = p[i] - c, t = sum + y;
      c = (t - sum) - y; sum = t;
    }
    if (!(sum - 1.0 <= 1.4901161193847656e-08 && 1.0 - sum <= 1.4901161193847656e-08))
      return err.finish(err.fail("probabilities do not sum to 1"));
  }
  int64_t nonzero = 0;
  for (int64_t i = 0; i < n_vocab; ++i) nonzero += p[i] > 0.0;
  // one weighted legacy draw of `size` indices: np.random.choice(n_vocab, size, replace, p) after np.random.seed(seed)
  auto draw = [&](uint32_t seed, int64_t size, std::vector<int64_t>& found) -> int {
    thread_local std::vector<double> pw, cdf, x;
    thread_local std::vector<std::pair<int64_t, int>> uniq;
    thread_local std::vector<std::pair<int, int64_t>> byidx;
    MT19937 g;
    g.init_genrand(seed);
    found.assign((size_t)size, 0);
    cdf.resize((size_t)n_vocab);
    auto make_cdf = [&](const double* q) {
      double acc = 0.0;
      for (int64_t i = 0; i < n_vocab; ++i) { acc += q[i]; cdf[(size_t)i] = acc; }
      const double last = cdf[(size_t)n_vocab - 1];
      for (int64_t i = 0; i < n_vocab; ++i) cdf[(size_t)i] /= last;
    };
    if (allow_duplicates) {
      make_cdf(p);
      for (int64_t i = 0; i < size; ++i) {
        const double u = g.next_double();
        found[(size_t)i] = std::upper_bound(cdf.begin(), cdf.end(), u) - cdf.begin();   // searchsorted(side='right')
      }
    } else {
      if (nonzero < size) return err.fail("Fewer non-zero entries in p than size");
      pw.assign(p, p + n_vocab);
      int64_t n_uniq = 0;
      while (n_uniq < size) {
        const int64_t m = size - n_uniq;
        x.resize((size_t)m);
        for (int64_t i = 0; i < m; ++i) x[(size_t)i] = g.next_double();
        if (n_uniq > 0)
          for (int64_t i = 0; i < n_uniq; ++i) pw[(size_t)found[(size_t)i]] = 0.0;
        make_cdf(pw.data());
        // new = searchsorted ; keep the FIRST occurrence of every value, in draw order (np.unique(return_index) + sort)
        uniq.clear();
        for (int64_t i = 0; i < m; ++i)
          uniq.emplace_back((int64_t)(std::upper_bound(cdf.begin(), cdf.end(), x[(size_t)i]) - cdf.begin()), (int)i);
        std::sort(uniq.begin(), uniq.end());
        byidx.clear();
        for (size_t i = 0; i < uniq.size(); ++i)
          if (i == 0 || uniq[i].first != uniq[i - 1].first) byidx.emplace_back(uniq[i].second, uniq[i].first);
        std::sort(byidx.begin(), byidx.end());
        for (auto& e : byidx) found[(size_t)n_uniq++] = e.second;
      }
    }
    for (int64_t i = 0; i < size; ++i)
      if (found[(size_t)i] >= n_vocab) return err.fail("internal: sampled index out of range");
    return 0;
  };
  // the draw depends on (seed, sample_size + |set(without)|) only -- `without` just filters it afterwards: with one seed for the
  // whole batch (reproducible evaluation) it is computed once per distinct draw size and shared by the requests
  bool same_seed = n > 0;
  for (int b = 1; same_seed && b < n; ++b) same_seed = seeds[b] == seeds[0];
  std::vector<int64_t> draw_sizes;
  std::vector<std::vector<int64_t>> shared;
  if (same_seed) {
    std::vector<int64_t> dsz((size_t)n);
    int rc0 = run_parallel(n, n_threads, [&](int b) -> int {
      thread_local std::vector<int64_t> w;
      w.clear();
      if (without) w.assign(without + without_off[b], without + without_off[b + 1]);
      std::sort(w.begin(), w.end());
      dsz[(size_t)b] = sample_size + (int64_t)(std::unique(w.begin(), w.end()) - w.begin());
      return 0;
    });
    if (rc0) return err.finish(rc0);
    draw_sizes = dsz;
    std::sort(draw_sizes.begin(), draw_sizes.end());
    draw_sizes.erase(std::unique(draw_sizes.begin(), draw_sizes.end()), draw_sizes.end());
    shared.resize(draw_sizes.size());
    rc0 = run_parallel((int)draw_sizes.size(), n_threads, [&](int si) -> int {
      if (!allow_duplicates && draw_sizes[(size_t)si] > n_vocab) return 0;   // reported per request below, with its own message
      return draw(seeds[0], draw_sizes[(size_t)si], shared[(size_t)si]);
    });
    if (rc0) return err.finish(rc0);
  }
  const int rc = run_parallel(n, n_threads, [&](int b) -> int {
    thread_local std::unordered_set<int64_t> ex;
    thread_local std::vector<int64_t> own;
    ex.clear();
    if (without) ex.insert(without + without_off[b], without + without_off[b + 1]);
    const int64_t size = sample_size + (without ? (int64_t)ex.size() : 0);
    if (!allow_duplicates && size > n_vocab)
      return err.fail("The given without list (length: %d reduces the vocab (length: %lld) too much to take a sample of size %d "
                       "(since no duplicates are allowed).", (int)ex.size(), (long long)n_vocab, sample_size);
    const std::vector<int64_t>* found = &own;
    if (same_seed) {
      found = &shared[(size_t)(std::lower_bound(draw_sizes.begin(), draw_sizes.end(), size) - draw_sizes.begin())];
    } else if (draw(seeds[b], size, own)) {
      return 1;
    }
    int k = 0;
    int64_t* o = out + (size_t)b * sample_size;
    for (int64_t i = 0; i < size && k < sample_size; ++i) {
      const int64_t v = vocab[(size_t)(*found)[(size_t)i]];
      if (!ex.count(v)) o[k++] = v;
    }
    out_len[b] = k;
    for (int i = k; i < sample_size; ++i) o[i] = 0;
    return 0;
  });
  return err.finish(rc);
}
