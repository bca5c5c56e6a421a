// C ABI (include/b4r.h): parameter layout, session (workspace carving + the kernel schedule of one
// forward / backward / optimizer step), ranking entry points.  Host-side only: every function enqueues on the
// caller's stream and returns.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b4r.h"
#include "../../include/b4r_debug.h"
#include "common.cuh"
#include "kernels.h"

using namespace b4r;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
namespace b4r { void set_last_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); } }   // host_data.cu
#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Kernel launch through the C ABI: counts the launch and, when profiling is on, brackets it with CUDA events on the
// launching stream.  Inside a stream capture the events become external event-record NODES of the graph, so every
// replay re-times every kernel in place (durations inside the graph, not at eager launch rate).
struct ProfRec { const char* tag; cudaEvent_t a, b; bool in_graph; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static void prof_record(cudaEvent_t e, cudaStream_t st, bool capturing) {
  if (capturing) cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
  else cudaEventRecord(e, st);
}
#define KL_(counter, tag, expr)                                                          \
  do {                                                                                  \
    cudaEvent_t e0__ = nullptr, e1__ = nullptr;                                         \
    bool cap__ = false;                                                                 \
    if (g_prof_on) {                                                                    \
      cudaStreamCaptureStatus cs__ = cudaStreamCaptureStatusNone;                       \
      cudaStreamIsCapturing(st, &cs__);                                                 \
      cap__ = cs__ == cudaStreamCaptureStatusActive;                                    \
      cudaEventCreate(&e0__); cudaEventCreate(&e1__); prof_record(e0__, st, cap__);     \
    }                                                                                   \
    CK(expr);                                                                           \
    counter;                                                                            \
    if (g_prof_on) { prof_record(e1__, st, cap__); g_prof.push_back(ProfRec{tag, e0__, e1__, cap__}); } \
  } while (0)
#define KL(tag, expr) KL_(s->launches++, tag, expr)

extern "C" int b4r_version(void) { return B4R_VERSION; }
extern "C" const char* b4r_last_error(void) { return g_err; }
extern "C" int b4r_device_check(int device) {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return fail("device %d is sm_%d%d; libb4r is built for sm_100a (B200) only", device, p.major, p.minor);
  return 0;
}

// ------------------------------------------------------------------------------------------------ layout
namespace {
struct Seg { std::string name; int64_t off, numel; int rows, cols, group; };
struct Layout {
  std::vector<Seg> segs;
  int64_t n_decay = 0, n_train = 0, n_total = 0;
  int64_t find(const std::string& n) const {
    for (auto& s : segs) if (s.name == n) return s.off;
    return -1;
  }
};
static int64_t pad8(int64_t x) { return (x + 7) / 8 * 8; }

static int check_cfg(const b4r_config* c) {
  if (!c) return fail("null config");
  const int H = c->hidden_size;
  if (H != 64 && H != 128 && H != 256) return fail("hidden_size %d unsupported (64, 128, 256)", H);
  if (c->num_heads <= 0 || H % c->num_heads) return fail("hidden_size %d not divisible by num_heads %d", H, c->num_heads);
  const int D = H / c->num_heads;
  if (D != 32 && D != 64) return fail("head dim %d unsupported (32, 64)", D);
  if (c->inner_dim <= 0 || c->inner_dim % 8) return fail("inner_dim %d must be a positive multiple of 8", c->inner_dim);
  if (c->vocab_size < 3) return fail("vocab_size %d too small", c->vocab_size);
  if (c->max_seq_len < 1 || c->max_seq_len > 256) return fail("max_seq_len %d unsupported (1..256)", c->max_seq_len);
  if (c->num_layers < 1 || c->num_layers > 64) return fail("num_layers %d unsupported", c->num_layers);
  if (c->output_dropout < 0.f || c->output_dropout >= 1.f || c->attention_dropout < 0.f || c->attention_dropout >= 1.f)
    return fail("dropout rates must be in [0,1)");
  return 0;
}

static Layout make_layout(const b4r_config& c) {
  Layout L;
  int64_t off = 0;
  auto add = [&](const std::string& n, int rows, int cols, int group) {
    int64_t numel = (int64_t)rows * (cols ? cols : 1);
    L.segs.push_back({n, off, numel, rows, cols, group});
    off += pad8(numel);
  };
  const int V = c.vocab_size, H = c.hidden_size, I = c.inner_dim, Lyr = c.num_layers;
  add("word_embeddings", V, H, 0);
  add("position_embedding", c.max_seq_len, H, 0);
  for (int l = 0; l < Lyr; ++l) {
    std::string p = "layer_" + std::to_string(l) + "/";
    add(p + "wqkv", H, 3 * H, 0);
    add(p + "wo", H, H, 0);
    add(p + "w1", H, I, 0);
    add(p + "w2", I, H, 0);
  }
  add("head/wt", H, H, 0);
  L.n_decay = off;
  add("emb_ln/gamma", H, 0, 1);
  add("emb_ln/beta", H, 0, 1);
  for (int l = 0; l < Lyr; ++l) {
    std::string p = "layer_" + std::to_string(l) + "/";
    add(p + "bqkv", 3 * H, 0, 1);
    add(p + "bo", H, 0, 1);
    add(p + "ln1/gamma", H, 0, 1);
    add(p + "ln1/beta", H, 0, 1);
    add(p + "b1", I, 0, 1);
    add(p + "b2", H, 0, 1);
    add(p + "ln2/gamma", H, 0, 1);
    add(p + "ln2/beta", H, 0, 1);
  }
  add("head/bt", H, 0, 1);
  add("head/ln/gamma", H, 0, 1);
  add("head/ln/beta", H, 0, 1);
  add("head/output_bias", V, 0, 1);
  L.n_train = off;
  add("pooler/w", H, H, 2);
  add("pooler/b", H, 0, 2);
  L.n_total = off;
  return L;
}
}  // namespace

extern "C" int b4r_param_entries(const b4r_config* cfg, b4r_param_entry* out, int cap) {
  if (check_cfg(cfg)) return -1;
  Layout L = make_layout(*cfg);
  int n = (int)L.segs.size();
  if (out) {
    for (int i = 0; i < n && i < cap; ++i) {
      memset(&out[i], 0, sizeof(out[i]));
      snprintf(out[i].name, sizeof(out[i].name), "%s", L.segs[i].name.c_str());
      out[i].offset = L.segs[i].off; out[i].numel = L.segs[i].numel;
      out[i].rows = L.segs[i].rows; out[i].cols = L.segs[i].cols; out[i].group = L.segs[i].group;
    }
  }
  return n;
}
extern "C" int b4r_param_counts(const b4r_config* cfg, int64_t* n_decay, int64_t* n_trainable, int64_t* n_total) {
  if (check_cfg(cfg)) return 1;
  Layout L = make_layout(*cfg);
  if (n_decay) *n_decay = L.n_decay;
  if (n_trainable) *n_trainable = L.n_train;
  if (n_total) *n_total = L.n_total;
  return 0;
}

// ------------------------------------------------------------------------------------------------ session
namespace {
struct LayerBuf {
  bf16 *qkv, *ctx, *a_pre, *y, *h_pre, *h, *o_pre, *out;
  float *lse, *mean1, *rstd1, *mean2, *rstd2;
  uint64_t* keep;
  // parameter offsets
  int64_t wqkv, wo, w1, w2, bqkv, bo, g1, be1, b1, b2, g2, be2;
  // gradient partial buffers
  float *p_wqkv, *p_wo, *p_w1, *p_w2, *p_ln2, *p_ln1, *p_b1, *p_bqkv;
  int s_wqkv, s_wo, s_w1, s_w2;
};
struct Bump {
  char* base; size_t off, cap; bool dry;
  template <class T> T* take(size_t n) {
    off = (off + 255) / 256 * 256;
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};
}  // namespace

static const int kCeFusedCtas = 148;   // one persistent CTA per SM (k_ce_bwd_fused.cu)

struct b4r_session {
  b4r_config cfg;
  int B, S, P, T, Mcap, H, I, V, N, Vp;
  float* params; bf16* shadow; float* grads;
  Layout lay;
  bf16* x0;
  std::vector<LayerBuf> layers;
  // head
  int *rows, *labels, *row_mult, *counts; float* row_w;
  bf16 *t_pre, *t_act, *t; float *hmean, *hrstd;
  float *ce_part, *lse, *lab, *stats, *step_stats, *fin_part; int* ticket;
  int vsplits, vsplits_umma;
  bool use_umma = false;
  CeUmmaMaps umaps;
  bool use_fattn = true;    // tcgen05 attention of the layered path (k_fattn.cu); false: mma.sync generation (k_attn.cu)
  bool use_fused = false;   // whole-encoder forward in one tcgen05 launch (k_enc_fused.cu)
  bool use_fused_bwd = false, fused_bwd_ok = false;   // ... and the backward (k_enc_fused_bwd.cu)
  float *enc_wpart = nullptr, *enc_bpart = nullptr, *enc_dpos = nullptr, *enc_embln = nullptr;
  ReduceJob* d_jobs_f = nullptr; int n_jobs_f = 0, jobs_f_blocks = 0;   // reduce jobs when the fused backward produced the partials
  void* d_enc_tables = nullptr;
  // internal side stream: kernels that do not depend on each other run as parallel branches (forked from / joined to the
  // caller's stream with events, so the caller still sees ONE ordered stream; inside a capture they become graph branches)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sel = nullptr;
  bool sel_pending = false, overlap_select = false;
  // deterministic item-table gradient (k_tablegrad.cu): the token sort runs on its own branch beside the whole backward
  cudaStream_t side_sort = nullptr;
  cudaEvent_t ev_sort_fork = nullptr, ev_sort = nullptr;
  TableGradArgs tg{};
  bool sort_pending = false;   // the token sort of the current ids was enqueued by b4r_encode(training) and is not joined yet
  ~b4r_session() {
    if (ev_sort_fork) cudaEventDestroy(ev_sort_fork);
    if (ev_sort) cudaEventDestroy(ev_sort);
    if (side_sort) cudaStreamDestroy(side_sort);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_sel) cudaEventDestroy(ev_sel);
    if (side) cudaStreamDestroy(side);
  }
  unsigned long long* dbg_buf = nullptr;
  bf16* dlogits; int dl_rows;
  float* dt_part; int dt_splits, dt_max_splits;
  float *p_dE, *p_dbias2; int me_splits;
  bool ce_fused_ok = false, ce_fused_default = false, use_ce_fused = false; int cf_vranges = 0; float *cf_dt = nullptr, *cf_dE = nullptr, *cf_db = nullptr;   // generation-3 CE backward
  ReduceJob* d_ce_jobs;
  bf16* d_tpre;
  float *p_head_ln, *p_wt, *p_vbias; int s_wt, vb_splits;
  // backward scratch
  float *dxa; bf16 *dg, *d_branch, *dh, *dctx, *dqkv;
  float *p_dpos, *p_embln; int emb_bsplits;
  ReduceJob* d_jobs; int n_jobs, jobs_max_len;
  ReduceJob* d_vb_jobs;  // [2]: first chunk (assign), later chunks (accumulate)
  const int64_t *ids, *mask;
  int select_mode;
  int launches;
};

static const int kColsumSplits = 296;   // CTAs (= partial rows) of the bias-gradient column sums: two per SM
static int wgrad_splits(int M, int N, int T) {
  if (twgrad_shape_ok(M, N, T)) return twgrad_splits(M, N, T);   // the tcgen05 kernel's own split count
  int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  int s = (2 * 148 + tiles - 1) / tiles;
  int cap = T / 256;
  if (cap < 1) cap = 1;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return s;
}

// Row splits of the dE pass of the tcgen05 CE backward: its grid is (vocabulary tiles x splits) CTAs of equal work, `ctas` resident at a
// time, so the pass takes ceil(tiles * splits / ctas) / splits of the unsplit time -- choose the split count (<= 8, partial buffers grow with
// it) that minimises that; e.g. 102 tiles on 148 resident CTAs: 1 split = one 69 %-full wave, 7 splits = 5 waves of 1/7 = 0.71.
static int choose_me_splits(int vtiles, int ctas) {
  int best = 1;
  double best_t = 1e30;
  for (int ms = 1; ms <= 8; ++ms) {
    const double t = (double)((vtiles * ms + ctas - 1) / ctas) / ms;
    if (t < best_t - 1e-9) { best_t = t; best = ms; }
  }
  return best;
}

static size_t carve(b4r_session* s, void* ws, size_t cap, bool dry, std::vector<ReduceJob>* jobs, std::vector<ReduceJob>* jobs_f = nullptr) {
  Bump b{reinterpret_cast<char*>(ws), 0, cap, dry};
  const int T = s->T, H = s->H, I = s->I, V = s->V, N = s->N, B = s->B, S = s->S, Mcap = s->Mcap;
  const int W = attn_mask_words(S);
  auto off = [&](const std::string& n) { return s->lay.find(n); };
  bool in_layers = false;   // jobs pushed inside the layer loop belong to the layered backward only
  auto job = [&](float* src, int64_t dst_off, int nparts, int len, long long stride) {
    if (jobs && s->grads) jobs->push_back(ReduceJob{src, s->grads + dst_off, nparts, len, stride, 0});
    if (jobs_f && s->grads && !in_layers) jobs_f->push_back(ReduceJob{src, s->grads + dst_off, nparts, len, stride, 0});
  };
  s->x0 = b.take<bf16>((size_t)T * H);
  s->layers.resize(s->cfg.num_layers);
  const int ln_parts = ln_bwd_parts(T);
  const int mt128 = (T + gemm_block_m() - 1) / gemm_block_m();
  in_layers = true;
  for (int l = 0; l < s->cfg.num_layers; ++l) {
    LayerBuf& L = s->layers[l];
    std::string p = "layer_" + std::to_string(l) + "/";
    L.wqkv = off(p + "wqkv"); L.wo = off(p + "wo"); L.w1 = off(p + "w1"); L.w2 = off(p + "w2");
    L.bqkv = off(p + "bqkv"); L.bo = off(p + "bo"); L.g1 = off(p + "ln1/gamma"); L.be1 = off(p + "ln1/beta");
    L.b1 = off(p + "b1"); L.b2 = off(p + "b2"); L.g2 = off(p + "ln2/gamma"); L.be2 = off(p + "ln2/beta");
    L.qkv = b.take<bf16>((size_t)T * 3 * H);
    L.ctx = b.take<bf16>((size_t)T * H);
    L.a_pre = b.take<bf16>((size_t)T * H);
    L.y = b.take<bf16>((size_t)T * H);
    L.h_pre = b.take<bf16>((size_t)T * I);
    L.h = b.take<bf16>((size_t)T * I);
    L.o_pre = b.take<bf16>((size_t)T * H);
    L.out = b.take<bf16>((size_t)T * H);
    L.lse = b.take<float>((size_t)B * N * S);
    L.mean1 = b.take<float>(T); L.rstd1 = b.take<float>(T); L.mean2 = b.take<float>(T); L.rstd2 = b.take<float>(T);
    L.keep = b.take<uint64_t>((size_t)B * N * S * W);
    L.s_wqkv = wgrad_splits(H, 3 * H, T); L.s_wo = wgrad_splits(H, H, T);
    L.s_w1 = wgrad_splits(H, I, T); L.s_w2 = wgrad_splits(I, H, T);
    L.p_wqkv = b.take<float>((size_t)L.s_wqkv * H * 3 * H);
    L.p_wo = b.take<float>((size_t)L.s_wo * H * H);
    L.p_w1 = b.take<float>((size_t)L.s_w1 * H * I);
    L.p_w2 = b.take<float>((size_t)L.s_w2 * I * H);
    L.p_ln2 = b.take<float>((size_t)ln_parts * 3 * H);
    L.p_ln1 = b.take<float>((size_t)ln_parts * 3 * H);
    L.p_b1 = b.take<float>((size_t)mt128 * I);
    L.p_bqkv = b.take<float>((size_t)kColsumSplits * 3 * H);
    job(L.p_wqkv, L.wqkv, L.s_wqkv, H * 3 * H, (long long)H * 3 * H);
    job(L.p_wo, L.wo, L.s_wo, H * H, (long long)H * H);
    job(L.p_w1, L.w1, L.s_w1, H * I, (long long)H * I);
    job(L.p_w2, L.w2, L.s_w2, I * H, (long long)I * H);
    job(L.p_ln2, L.g2, ln_parts, H, 3 * H);
    job(L.p_ln2 + H, L.be2, ln_parts, H, 3 * H);
    job(L.p_ln2 + 2 * H, L.b2, ln_parts, H, 3 * H);
    job(L.p_ln1, L.g1, ln_parts, H, 3 * H);
    job(L.p_ln1 + H, L.be1, ln_parts, H, 3 * H);
    job(L.p_ln1 + 2 * H, L.bo, ln_parts, H, 3 * H);
    job(L.p_b1, L.b1, mt128, I, I);
    // hidden 256: the tcgen05 weight-gradient kernel of Wqkv also sums the columns of dQKV (one partial row per token split)
    job(L.p_bqkv, L.bqkv, twgrad_shape_ok(H, 3 * H, T) ? L.s_wqkv : kColsumSplits, 3 * H, 3 * H);
  }
  in_layers = false;
  s->fused_bwd_ok = enc_bwd_fused_supported(H, N, S, I);
  if (s->fused_bwd_ok) {
    const int nc = enc_fused_ctas(B, S), Ln = s->cfg.num_layers;
    const size_t WPn = enc_bwd_wpart_floats(I), PFn = enc_bwd_bpart_floats(I);
    s->enc_wpart = b.take<float>((size_t)Ln * nc * WPn);
    s->enc_bpart = b.take<float>((size_t)Ln * nc * PFn);
    for (int l = 0; l < Ln && jobs_f && s->grads; ++l) {
      // the four kernels of a layer (and its eight bias / LayerNorm vectors) are contiguous in the flat layout
      jobs_f->push_back(ReduceJob{s->enc_wpart + (size_t)l * nc * WPn, s->grads + s->layers[l].wqkv, nc, (int)WPn, (long long)WPn, 0});
      jobs_f->push_back(ReduceJob{s->enc_bpart + (size_t)l * nc * PFn, s->grads + s->layers[l].bqkv, nc, (int)PFn, (long long)PFn, 0});
    }
  }
  // head
  s->rows = b.take<int>(Mcap); s->labels = b.take<int>(Mcap); s->row_mult = b.take<int>(Mcap);
  s->row_w = b.take<float>(Mcap); s->counts = b.take<int>(8);
  s->t_pre = b.take<bf16>((size_t)Mcap * H); s->t_act = b.take<bf16>((size_t)Mcap * H); s->t = b.take<bf16>((size_t)Mcap * H);
  s->hmean = b.take<float>(Mcap); s->hrstd = b.take<float>(Mcap);
  {
    int mtiles = (Mcap + ce_block_m() - 1) / ce_block_m();
    int vtiles = (V + 127) / 128;
    int vs = (2 * 148 + mtiles - 1) / mtiles;
    if (vs > vtiles) vs = vtiles;
    if (vs < 1) vs = 1;
    if (vs > 64) vs = 64;
    s->vsplits = vs;
  }
  {
    int mtiles = (Mcap + ce_umma_block_m() - 1) / ce_umma_block_m();
    int vtiles = (V + 127) / 128;
    (void)mtiles;
    int vs = vtiles < 64 ? vtiles : 64;   // capacity of the partial buffer; the actual split count is chosen on the device
    if (vs < 1) vs = 1;
    s->vsplits_umma = vs;
  }
  s->ce_part = b.take<float>((size_t)(s->vsplits > s->vsplits_umma ? s->vsplits : s->vsplits_umma) * Mcap * 6);
  s->lse = b.take<float>(Mcap); s->lab = b.take<float>(Mcap);
  s->stats = b.take<float>(8); s->step_stats = b.take<float>(8);
  s->fin_part = b.take<float>(kCeFinalizeMaxBlocks * 5); s->ticket = b.take<int>(4);
  {
    // dlogits chunk: at most ~256 MB so that it stays close to the L2 / small in HBM
    size_t max_rows = ((size_t)256 << 20) / ((size_t)s->Vp * 2);
    max_rows = max_rows / 64 * 64;
    if (max_rows < 64) max_rows = 64;
    s->dl_rows = (int)((size_t)((Mcap + 63) / 64 * 64) < max_rows ? (size_t)((Mcap + 63) / 64 * 64) : max_rows);
  }
  s->dlogits = b.take<bf16>((size_t)s->dl_rows * s->Vp);
  {
    int mtiles = (Mcap + gemm_block_m() - 1) / gemm_block_m();
    int sp = (2 * 148 + mtiles - 1) / mtiles;
    int cap2 = s->Vp / 256;
    if (sp > cap2) sp = cap2;
    if (sp < 1) sp = 1;
    if (sp > 32) sp = 32;
    s->dt_splits = sp;
  }
  {
    int vtiles = (V + 127) / 128;
    const int xtiles = (V + ce_bwd_umma_xtile(H) - 1) / ce_bwd_umma_xtile(H);
    s->dt_max_splits = xtiles < 24 ? xtiles : 24;
    s->me_splits = choose_me_splits(vtiles, ce_bwd_umma_xtile(H) == 64 ? 148 : 2 * 148);   // resident CTAs: two per SM, one at hidden 256
  }
  const bool bwd_umma = ce_bwd_umma_supported(H);
  s->dt_part = b.take<float>((size_t)(bwd_umma && s->dt_max_splits > s->dt_splits ? s->dt_max_splits : s->dt_splits) * Mcap * H);
  s->p_dE = b.take<float>(bwd_umma ? (size_t)s->me_splits * V * H : 8);
  s->p_dbias2 = b.take<float>(bwd_umma ? (size_t)s->me_splits * V : 8);
  s->d_ce_jobs = b.take<ReduceJob>(2);
  {
    // generation-3 CE backward (one pass): partial buffers, used when they stay within 1 GB
    const int nvr = ce_bwd_fused_dt_slot_cap(V, kCeFusedCtas), nch = ce_bwd_fused_max_chunks(Mcap);
    const size_t bytes = ((size_t)nvr * Mcap * H + (size_t)nch * V * (H + 1)) * sizeof(float);
    s->ce_fused_ok = bwd_umma && ce_bwd_fused_supported(H) && head_bwd_fused_supported(H) && bytes <= ((size_t)1 << 30);
    s->ce_fused_default = true;
    s->cf_vranges = nvr;   // capacity of the dT partial buffer (slots)
    if (s->ce_fused_ok) {
      s->cf_dt = b.take<float>((size_t)nvr * Mcap * H);
      s->cf_dE = b.take<float>((size_t)nch * V * H);
      s->cf_db = b.take<float>((size_t)nch * V);
    }
  }
  s->d_tpre = b.take<bf16>((size_t)Mcap * H);
  const bool head_fused = head_bwd_fused_supported(H);
  const int head_parts = head_fused ? head_bwd_fused_ctas() : ln_bwd_parts(Mcap);
  s->p_head_ln = b.take<float>((size_t)head_parts * 3 * H);
  s->s_wt = head_fused ? head_bwd_fused_ctas() : wgrad_splits(H, H, Mcap);
  s->p_wt = b.take<float>((size_t)s->s_wt * H * H);
  s->vb_splits = 16;
  s->p_vbias = b.take<float>((size_t)s->vb_splits * V);
  job(s->p_head_ln, off("head/ln/gamma"), head_parts, H, 3 * H);
  job(s->p_head_ln + H, off("head/ln/beta"), head_parts, H, 3 * H);
  job(s->p_head_ln + 2 * H, off("head/bt"), head_parts, H, 3 * H);
  job(s->p_wt, off("head/wt"), s->s_wt, H * H, (long long)H * H);
  // backward scratch
  s->dxa = b.take<float>((size_t)T * H); s->dg = b.take<bf16>((size_t)T * H);
  s->d_branch = b.take<bf16>((size_t)T * H); s->dh = b.take<bf16>((size_t)T * I);
  s->dctx = b.take<bf16>((size_t)T * H); s->dqkv = b.take<bf16>((size_t)T * 3 * H);
  s->emb_bsplits = embed_bwd_bsplits(B);
  s->p_dpos = b.take<float>((size_t)s->emb_bsplits * S * H);
  s->p_embln = b.take<float>((size_t)s->emb_bsplits * S * 2 * H);
  s->tg.T = T; s->tg.V = V; s->tg.H = H;
  for (int i = 0; i < 2; ++i) { s->tg.keys[i] = b.take<uint32_t>((size_t)T); s->tg.vals[i] = b.take<uint32_t>((size_t)T); }
  s->tg.hist = b.take<uint32_t>((size_t)256 * table_grad_sort_blocks(T));
  s->tg.long_runs = b.take<int>((size_t)1 + 4 * table_grad_max_long_runs(T));
  s->tg.carry = b.take<float>((size_t)table_grad_chunks(T) * 2 * H);
  in_layers = true;   // the fused backward reduces its own embedding partials (below)
  job(s->p_dpos, off("position_embedding"), s->emb_bsplits, S * H, (long long)S * H);
  job(s->p_embln, off("emb_ln/gamma"), s->emb_bsplits * S, H, 2 * H);
  job(s->p_embln + H, off("emb_ln/beta"), s->emb_bsplits * S, H, 2 * H);
  in_layers = false;
  if (s->fused_bwd_ok) {
    const int nc = enc_fused_ctas(B, S);
    s->enc_dpos = b.take<float>((size_t)nc * S * H);
    s->enc_embln = b.take<float>((size_t)nc * 128);
    if (jobs_f && s->grads) {
      jobs_f->push_back(ReduceJob{s->enc_dpos, s->grads + off("position_embedding"), nc, S * H, (long long)S * H, 0});
      jobs_f->push_back(ReduceJob{s->enc_embln, s->grads + off("emb_ln/gamma"), nc, 2 * H, (long long)2 * H, 0});
    }
  }
  s->d_jobs = b.take<ReduceJob>(256);
  s->d_jobs_f = b.take<ReduceJob>(256);
  s->d_vb_jobs = b.take<ReduceJob>(2);
  s->d_enc_tables = b.take<unsigned char>(enc_fused_table_bytes(s->cfg.num_layers));
  s->dbg_buf = b.take<unsigned long long>(512);
  return b.off + 256;
}

static b4r_session* make_session_shell(const b4r_config* cfg, int batch, int seq_len, int max_pred) {
  b4r_session* s = new b4r_session();
  s->cfg = *cfg;
  s->B = batch; s->S = seq_len; s->P = max_pred; s->T = batch * seq_len;
  s->Mcap = batch * max_pred + batch;
  s->H = cfg->hidden_size; s->I = cfg->inner_dim; s->V = cfg->vocab_size; s->N = cfg->num_heads;
  s->Vp = (cfg->vocab_size + 127) / 128 * 128;
  s->lay = make_layout(*cfg);
  s->params = nullptr; s->shadow = nullptr; s->grads = nullptr;
  s->launches = 0; s->select_mode = 0; s->ids = nullptr; s->mask = nullptr;
  return s;
}

static int check_dims(const b4r_config* cfg, int batch, int seq_len, int max_pred) {
  if (check_cfg(cfg)) return 1;
  if (batch < 1 || seq_len < 1 || max_pred < 0) return fail("bad session dims batch=%d seq_len=%d max_pred=%d", batch, seq_len, max_pred);
  if (seq_len > cfg->max_seq_len) return fail("seq_len %d exceeds max_sequence_length %d", seq_len, cfg->max_seq_len);
  if ((int64_t)batch * seq_len > (int64_t)1 << 30) return fail("batch*seq_len too large");
  return 0;
}

extern "C" size_t b4r_session_workspace_bytes(const b4r_config* cfg, int batch, int seq_len, int max_pred) {
  if (check_dims(cfg, batch, seq_len, max_pred)) return 0;
  b4r_session* s = make_session_shell(cfg, batch, seq_len, max_pred);
  size_t n = carve(s, nullptr, 0, true, nullptr);
  delete s;
  return n;
}

extern "C" int b4r_session_create(const b4r_config* cfg, int batch, int seq_len, int max_pred, float* params,
                                  void* shadow_bf16, float* grads, void* workspace, size_t workspace_bytes,
                                  b4r_session** out) {
  if (!out) return fail("null out");
  if (check_dims(cfg, batch, seq_len, max_pred)) return 1;
  if (!params || !shadow_bf16 || !workspace) return fail("null buffer");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (b4r_device_check(dev)) return 1;
  if (((uintptr_t)params & 31) || ((uintptr_t)shadow_bf16 & 15) || ((uintptr_t)workspace & 255) || ((uintptr_t)grads & 31))
    return fail("buffers must be aligned (params/grads 32 B, shadow 16 B, workspace 256 B)");
  b4r_session* s = make_session_shell(cfg, batch, seq_len, max_pred);
  s->params = params; s->shadow = reinterpret_cast<bf16*>(shadow_bf16); s->grads = grads;
  std::vector<ReduceJob> jobs, jobs_f;
  size_t need = carve(s, workspace, workspace_bytes, false, &jobs, &jobs_f);
  if (need > workspace_bytes) { delete s; return fail("workspace too small: need %zu bytes, got %zu", need, workspace_bytes); }
  if (jobs.size() > 256) { delete s; return fail("too many reduce jobs"); }
  s->n_jobs = (int)jobs.size();
  s->jobs_max_len = 0;
  for (auto& j : jobs) { int nb = grad_reduce_blocks(j.nparts, j.len); if (nb > s->jobs_max_len) s->jobs_max_len = nb; }
  CK(cudaMemcpy(s->d_jobs, jobs.data(), jobs.size() * sizeof(ReduceJob), cudaMemcpyHostToDevice));
  if (jobs_f.size() > 256) { delete s; return fail("too many reduce jobs"); }
  s->n_jobs_f = (int)jobs_f.size();
  for (auto& j : jobs_f) { int nb = grad_reduce_blocks(j.nparts, j.len); if (nb > s->jobs_f_blocks) s->jobs_f_blocks = nb; }
  if (!jobs_f.empty()) CK(cudaMemcpy(s->d_jobs_f, jobs_f.data(), jobs_f.size() * sizeof(ReduceJob), cudaMemcpyHostToDevice));
  ReduceJob vb[2];
  vb[0] = ReduceJob{s->p_vbias, grads ? grads + s->lay.find("head/output_bias") : nullptr, s->vb_splits, s->V, (long long)s->V, 0};
  vb[1] = vb[0]; vb[1].accumulate = 1;
  CK(cudaMemcpy(s->d_vb_jobs, vb, sizeof(vb), cudaMemcpyHostToDevice));
  ReduceJob cj[2];
  cj[0] = ReduceJob{s->p_dE, grads ? grads + s->lay.find("word_embeddings") : nullptr, s->me_splits, s->V * s->H, (long long)s->V * s->H, 0};
  cj[1] = ReduceJob{s->p_dbias2, grads ? grads + s->lay.find("head/output_bias") : nullptr, s->me_splits, s->V, (long long)s->V, 0};
  CK(cudaMemcpy(s->d_ce_jobs, cj, sizeof(cj), cudaMemcpyHostToDevice));
  CK(cudaMemset(s->stats, 0, 8 * sizeof(float)));
  CK(cudaMemset(s->step_stats, 0, 8 * sizeof(float)));
  CK(cudaMemset(s->ticket, 0, 4 * sizeof(int)));
  s->use_umma = ce_umma_make_maps(&s->umaps, s->t, s->Mcap, s->shadow + s->lay.find("word_embeddings"), s->V, s->H) &&
                getenv("B4R_DISABLE_UMMA") == nullptr;
  s->use_ce_fused = s->use_umma && s->ce_fused_ok && s->ce_fused_default && s->grads != nullptr;
  CK(cudaMemset(s->counts, 0, 8 * sizeof(int)));
  CK(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s->ev_sel, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&s->side_sort, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s->ev_sort_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s->ev_sort, cudaEventDisableTiming));
  if (enc_fused_supported(s->H, s->N, s->S, s->I)) {
    const int Ln = s->cfg.num_layers;
    std::vector<EncFusedLayerHost> lh(Ln);
    std::vector<const bf16*> wp((size_t)Ln * 4);
    for (int l = 0; l < Ln; ++l) {
      LayerBuf& L = s->layers[l];
      const float* P = s->params;
      lh[l] = EncFusedLayerHost{P + L.bqkv, P + L.bo, P + L.g1, P + L.be1, P + L.b1, P + L.b2, P + L.g2, P + L.be2,
                                L.qkv, L.ctx, L.a_pre, L.y, L.h_pre, L.h, L.o_pre, L.out, L.lse, L.mean1, L.rstd1, L.mean2, L.rstd2, L.keep};
      wp[l * 4 + 0] = s->shadow + L.wqkv; wp[l * 4 + 1] = s->shadow + L.wo; wp[l * 4 + 2] = s->shadow + L.w1; wp[l * 4 + 3] = s->shadow + L.w2;
    }
    std::vector<unsigned char> host(enc_fused_table_bytes(Ln));
    if (enc_fused_build_tables(lh.data(), Ln, s->I, wp.data(), host.data(), s->d_enc_tables)) {
      CK(cudaMemcpy(s->d_enc_tables, host.data(), host.size(), cudaMemcpyHostToDevice));
      s->use_fused = true;
      s->use_fused_bwd = s->fused_bwd_ok && s->grads != nullptr;
    }
  }
  *out = s;
  return 0;
}

extern "C" void b4r_session_destroy(b4r_session* s) { delete s; }

extern "C" int b4r_sync_shadow(b4r_session* s, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s) return fail("null session");
  KL("cast_bf16", launch_cast_bf16(s->params, s->shadow, s->lay.n_total, st));
  return 0;
}

// The token sort of the embedding gradient (k_tablegrad.cu) depends on the ids only.  It is enqueued on its own stream branch as early
// as possible -- by the training-mode encode, so that the small sort CTAs run beside the forward (the fused forward leaves 20 SMs
// idle) instead of taking an SM from the persistent 148-CTA kernels at the start of the backward -- and joined right before the
// table gradient needs it.
static int launch_sort_branch(b4r_session* s, cudaStream_t st) {
  s->tg.ids = s->ids;
  CK(cudaEventRecord(s->ev_sort_fork, st));
  CK(cudaStreamWaitEvent(s->side_sort, s->ev_sort_fork, 0));
  {
    cudaStream_t st_main = st; (void)st_main;
    cudaStream_t st = s->side_sort;
    KL("token_sort", launch_token_sort(s->tg, st));
    s->launches += table_grad_sort_launches(s->tg.T, s->tg.V) - 1;
  }
  CK(cudaEventRecord(s->ev_sort, s->side_sort));
  s->sort_pending = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ forward
extern "C" int b4r_encode(b4r_session* s, const int64_t* ids, const int64_t* mask, int training, uint64_t seed,
                          uint32_t step, const int64_t* step_counter, void* stream) {
  const long long* d_step = reinterpret_cast<const long long*>(step_counter);
  if (!s || !ids || !mask) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int T = s->T, H = s->H, I = s->I;
  const float od = training ? s->cfg.output_dropout : 0.f, ad = training ? s->cfg.attention_dropout : 0.f;
  const float* P = s->params;
  const bf16* W = s->shadow;
  s->ids = ids; s->mask = mask;
  s->sort_pending = false;
  if (training && s->grads && launch_sort_branch(s, st)) return 1;
  if (s->use_fused) {
    EncFusedArgs f{};
    f.ids = ids; f.mask = mask; f.table = W + s->lay.find("word_embeddings"); f.pos = W + s->lay.find("position_embedding");
    f.emb_g = P + s->lay.find("emb_ln/gamma"); f.emb_b = P + s->lay.find("emb_ln/beta"); f.x0 = s->x0; f.dev_tables = s->d_enc_tables;
    f.B = s->B; f.S = s->S; f.V = s->V; f.L = s->cfg.num_layers; f.I = I; f.training = training;
    f.out_drop = od; f.attn_drop = ad; f.seed = seed; f.step = step; f.d_step = d_step;
    f.dbg = getenv("B4R_FUSED_DEBUG") ? (void*)s->dbg_buf : nullptr;
    KL("enc_fwd_fused", launch_enc_fwd_fused(f, st));
    return 0;
  }
  KL("embed_ln_fwd", launch_embed_ln_fwd(ids, W + s->lay.find("word_embeddings"), W + s->lay.find("position_embedding"),
                         P + s->lay.find("emb_ln/gamma"), P + s->lay.find("emb_ln/beta"), s->x0, s->B, s->S, H, s->V, od,
                         seed, step, d_step, st));
  const bf16* x = s->x0;
  for (int l = 0; l < s->cfg.num_layers; ++l) {
    LayerBuf& L = s->layers[l];
    GemmArgs g{};
    g.A = x; g.lda = H; g.B = W + L.wqkv; g.ldb = 3 * H; g.b_trans = true; g.M = T; g.N = 3 * H; g.K = H;
    g.bias = P + L.bqkv; g.out_bf16 = L.qkv; g.ld_out = 3 * H;
    KL("gemm:qkv", launch_gemm(EPI_BIAS_BF16, g, st));
    AttnArgs a{};
    a.qkv = L.qkv; a.mask = mask; a.ctx = L.ctx; a.lse = L.lse; a.keep_bits = L.keep; a.B = s->B; a.S = s->S; a.H = H; a.N = s->N;
    a.drop_rate = ad; a.seed = seed; a.site = site_id(SITE_ATTN_PROBS, l); a.step = step; a.d_step = d_step;
    a.no_tcgen05 = !s->use_fattn;
    KL("attn_fwd", launch_attn_fwd(a, st));
    RowLnArgs r{};
    r.A = L.ctx; r.lda = H; r.W = W + L.wo; r.M = T; r.K = H; r.H = H; r.bias = P + L.bo; r.gamma = P + L.g1; r.beta = P + L.be1;
    r.residual = x; r.pre = L.a_pre; r.y = L.y; r.mean = L.mean1; r.rstd = L.rstd1;
    r.drop_rate = od; r.seed = seed; r.site = site_id(SITE_ATTN_OUT, l); r.step = step; r.d_step = d_step;
    KL("rowln:attn_out", launch_gemm_rowln(ROW_RES_DROP_LN, r, st));
    GemmArgs f{};
    f.A = L.y; f.lda = H; f.B = W + L.w1; f.ldb = I; f.b_trans = true; f.M = T; f.N = I; f.K = H;
    f.bias = P + L.b1; f.out_bf16 = L.h_pre; f.out2_bf16 = L.h; f.ld_out = I;
    KL("gemm:ffn1_gelu", launch_gemm(EPI_BIAS_GELU, f, st));
    RowLnArgs r2{};
    r2.A = L.h; r2.lda = I; r2.W = W + L.w2; r2.M = T; r2.K = I; r2.H = H; r2.bias = P + L.b2; r2.gamma = P + L.g2; r2.beta = P + L.be2;
    r2.residual = L.y; r2.pre = L.o_pre; r2.y = L.out; r2.mean = L.mean2; r2.rstd = L.rstd2;
    r2.drop_rate = od; r2.seed = seed; r2.site = site_id(SITE_FFN_OUT, l); r2.step = step; r2.d_step = d_step;
    KL("rowln:ffn2", launch_gemm_rowln(ROW_RES_DROP_LN, r2, st));
    x = L.out;
  }
  return 0;
}

extern "C" int b4r_mlm_select(b4r_session* s, const int64_t* positions, const int64_t* ids, const int64_t* weights,
                              int mode, int want_aux, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !positions) return fail("null argument");
  if (mode == 0 && !ids) return fail("mode 0 needs masked_lm_ids");
  if (mode == 1 && !weights) return fail("mode 1 needs masked_lm_weights");
  if (s->P < 1) return fail("session was created with max_pred = 0");
  // mode 2 (all slots): reuse the weights path with a null test -> handled by passing positions as weights of ones
  s->select_mode = mode;
  const int64_t* id_src = ids ? ids : positions;
  cudaStream_t st_caller = st;
  if (s->overlap_select) {   // the compaction depends on the inputs only: run it as a branch beside what the caller enqueues next
    CK(cudaEventRecord(s->ev_fork, st_caller));
    CK(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
    st = s->side;
  }
  if (mode == 2) {
    // all slots valid: weights := non-null pointer whose values are irrelevant -> use use_weights = 2
    KL("mlm_select", launch_mlm_select(positions, id_src, nullptr, 2, s->B, s->S, s->P, 0, s->rows, s->labels, s->row_w, s->row_mult, s->counts, st));
  } else {
    KL("mlm_select", launch_mlm_select(positions, id_src, weights, mode, s->B, s->S, s->P, want_aux, s->rows, s->labels, s->row_w, s->row_mult, s->counts, st));
  }
  if (s->overlap_select) {
    CK(cudaEventRecord(s->ev_sel, s->side));
    s->sel_pending = true;
  }
  return 0;
}

// joins a pending side-stream selection into the caller's stream (every consumer of rows / labels / counts calls this)
static int join_select(b4r_session* s, cudaStream_t st) {
  if (s->sel_pending) {
    CK(cudaStreamWaitEvent(st, s->ev_sel, 0));
    s->sel_pending = false;
  }
  return 0;
}

extern "C" int b4r_mlm_transform(b4r_session* s, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s) return fail("null session");
  if (join_select(s, st)) return 1;
  const int H = s->H;
  const bf16* x = s->layers.back().out;
  RowLnArgs r{};
  r.A = x; r.lda = H; r.a_rows = s->rows; r.W = s->shadow + s->lay.find("head/wt"); r.M = s->Mcap; r.K = H; r.H = H;
  r.d_M = s->counts + 1;
  r.bias = s->params + s->lay.find("head/bt"); r.gamma = s->params + s->lay.find("head/ln/gamma"); r.beta = s->params + s->lay.find("head/ln/beta");
  r.pre = s->t_pre; r.act = s->t_act; r.y = s->t; r.mean = s->hmean; r.rstd = s->hrstd;
  KL("rowln:mlm_transform", launch_gemm_rowln(ROW_GELU_LN, r, st));
  return 0;
}

static CeArgs ce_args(b4r_session* s) {
  CeArgs c{};
  c.t = s->t; c.ldt = s->H; c.E = s->shadow + s->lay.find("word_embeddings"); c.vbias = s->params + s->lay.find("head/output_bias");
  c.labels = s->labels; c.row_w = s->row_w; c.row_mult = s->row_mult; c.d_counts = s->counts;
  c.M_cap = s->Mcap; c.H = s->H; c.V = s->V; c.v_begin = 0; c.v_end = s->V; c.vsplits = s->vsplits; c.batch = s->B;
  c.part = s->ce_part; c.lse = s->lse; c.lab_out = s->lab; c.stats = nullptr; c.step_stats = s->step_stats;
  c.fin_part = s->fin_part; c.ticket = s->ticket;
  c.dlogits = s->dlogits; c.ld_dl = s->Vp;
  return c;
}

extern "C" int b4r_mlm_loss(b4r_session* s, float* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s) return fail("null session");
  if (join_select(s, st)) return 1;
  CeArgs c = ce_args(s);
  c.stats = stats;
  if (s->use_umma) {
    c.target_ctas = 2 * 148; c.max_splits = s->vsplits_umma;
    KL("ce_fwd_umma", launch_ce_fwd_umma(s->umaps, c, st));
    c.vsplits = -1;  // finalize re-derives the device-side split count
  } else {
    KL("ce_fwd", launch_ce_fwd(c, st));
  }
  KL("ce_finalize", launch_ce_finalize(c, st));
  return 0;
}

extern "C" int b4r_mlm_logits(b4r_session* s, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !out) return fail("null argument");
  if (join_select(s, st)) return 1;
  GemmArgs g{};
  g.A = s->t; g.lda = s->H; g.B = s->shadow + s->lay.find("word_embeddings"); g.ldb = s->H; g.b_trans = false;
  g.M = s->Mcap; g.N = s->V; g.K = s->H; g.d_M = s->counts + 1;
  g.bias = s->params + s->lay.find("head/output_bias"); g.out_f32 = out; g.ld_f32 = s->V;
  KL("gemm:logits", launch_gemm(EPI_BIAS_F32, g, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ backward
// ext_dt != nullptr: the gradient of the transformed rows comes from outside (vocabulary-sharded projection: the reduce-scattered
// sum of every shard's partial dT, fp32 [Mcap][H] in this session's row order); the cross-entropy backward is skipped and the
// table / output-bias gradients are left as the shard wrote them (the embedding backward accumulates on top).
static int backward_impl(b4r_session* s, uint64_t seed, uint32_t step, const int64_t* step_counter, void* stream, const float* ext_dt) {
  const long long* d_step = reinterpret_cast<const long long*>(step_counter);
  if (!s || !s->grads) return fail("session has no gradient buffer");
  if (!s->ids) return fail("b4r_encode must run before b4r_backward");
  cudaStream_t st = (cudaStream_t)stream;
  const int T = s->T, H = s->H, I = s->I, V = s->V, Mcap = s->Mcap;
  const float* P = s->params;
  const bf16* W = s->shadow;
  float* G = s->grads;
  const float od = s->cfg.output_dropout;
  const int64_t oE = s->lay.find("word_embeddings");
  const bool bwd_umma = !ext_dt && s->use_umma && ce_bwd_umma_supported(H);
  const int ce_xt = ce_bwd_umma_xtile(H), ce_xtiles = (V + ce_xt - 1) / ce_xt, ce_ctas = ce_xt == 64 ? 148 : 2 * 148;
  // gradient accumulators that are scatter / accumulate targets
  if (!bwd_umma && !ext_dt) CK(cudaMemsetAsync(G + oE, 0, (size_t)V * H * sizeof(float), st));
  CK(cudaMemsetAsync(s->dxa, 0, (size_t)T * H * sizeof(float), st));
  // token sort of the embedding gradient: already on its branch when the forward ran in training mode (launch_sort_branch)
  s->tg.dx = s->dxa; s->tg.grad_table = G + oE;
  if (!s->sort_pending && launch_sort_branch(s, st)) return 1;
  CeArgs c = ce_args(s);
  bool head_done = false;
  if (join_select(s, st)) return 1;
  if (bwd_umma && s->use_ce_fused) {
    // ---- CE backward, generation 3: ONE tcgen05 pass for dT and dE (hidden 64)
    CeBwdFusedArgs fa{};
    fa.vbias = c.vbias; fa.lse = s->lse; fa.row_w = s->row_w; fa.labels = s->labels; fa.d_counts = s->counts;
    fa.M_cap = Mcap; fa.V = V; fa.ctas = kCeFusedCtas; fa.dt_part = s->cf_dt; fa.dE_part = s->cf_dE; fa.db_part = s->cf_db;
    fa.dbg = getenv("B4R_CF_DEBUG") ? s->dbg_buf : nullptr;
    KL("ce_bwd_fused", launch_ce_bwd_fused(s->umaps, fa, st));
    // the MLM-transform backward needs dT only: it runs as a branch beside the reduction of the dE partials
    CK(cudaEventRecord(s->ev_fork, st));
    CK(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
    {
      cudaStream_t st_main = st; (void)st_main;
      cudaStream_t st = s->side;
      KL("head_bwd_fused", launch_head_bwd_fused(s->cf_dt, s->cf_vranges, (size_t)Mcap * H, s->t_pre, s->t_act, s->hmean, s->hrstd,
                               P + s->lay.find("head/ln/gamma"), W + s->lay.find("head/wt"), s->layers.back().out, s->rows, s->counts,
                               Mcap, s->dxa, s->p_head_ln, s->p_wt, st, V, kCeFusedCtas, -1));
    }
    CK(cudaEventRecord(s->ev_join, s->side));
    head_done = true;
    KL("grad_reduce:ce", launch_ce_bwd_fused_reduce(fa, G + oE, G + s->lay.find("head/output_bias"), st));
    CK(cudaStreamWaitEvent(st, s->ev_join, 0));
  } else if (bwd_umma) {
    // ---- CE backward, generation 2: two tcgen05 passes that recompute the logits tile, nothing [M,V]-sized in memory
    CeBwdArgs ba{};
    ba.vbias = c.vbias; ba.lse = s->lse; ba.row_w = s->row_w; ba.labels = s->labels; ba.d_counts = s->counts;
    ba.M_cap = Mcap; ba.V = V; ba.H = H; ba.target_ctas = ce_ctas; ba.max_splits = s->dt_max_splits; ba.msplits = s->me_splits;
    ba.out = s->dt_part; ba.dbias_out = nullptr;
    KL("ce_bwd_umma:dT", launch_ce_bwd_umma(s->umaps, ba, true, st));
    if (head_bwd_fused_supported(H)) {
      // the MLM-transform backward needs dT only: it runs as a branch beside the dE pass and its reduction
      CK(cudaEventRecord(s->ev_fork, st));
      CK(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
      {
        cudaStream_t st_main = st; (void)st_main;
        cudaStream_t st = s->side;
        KL("head_bwd_fused", launch_head_bwd_fused(s->dt_part, s->dt_splits, (size_t)Mcap * H, s->t_pre, s->t_act, s->hmean, s->hrstd,
                                 P + s->lay.find("head/ln/gamma"), W + s->lay.find("head/wt"), s->layers.back().out, s->rows, s->counts,
                                 Mcap, s->dxa, s->p_head_ln, s->p_wt, st, ce_xtiles, ce_ctas, s->dt_max_splits));
      }
      CK(cudaEventRecord(s->ev_join, s->side));
      head_done = true;
    }
    ba.out = s->p_dE; ba.dbias_out = s->p_dbias2;
    KL("ce_bwd_umma:dE", launch_ce_bwd_umma(s->umaps, ba, false, st));
    KL("grad_reduce:ce", launch_grad_reduce(s->d_ce_jobs, 2, grad_reduce_blocks(s->me_splits, V * H), st));
    if (head_done) CK(cudaStreamWaitEvent(st, s->ev_join, 0));
  }
  // ---- CE backward, generation 1: dlogits materialised in bf16, chunked over rows
  int chunk = 0;
  for (int r0 = 0; r0 < (bwd_umma || ext_dt ? 0 : Mcap); r0 += s->dl_rows, ++chunk) {
    const int rc = (Mcap - r0) < s->dl_rows ? (Mcap - r0) : s->dl_rows;
    c.row_begin = r0; c.row_count = rc;
    KL("ce_dlogits", launch_ce_dlogits(c, st));
    KL("colsum:vbias", launch_colsum_bf16(s->dlogits, s->Vp, rc, V, s->counts + 1, r0, s->p_vbias, s->vb_splits, st));
    KL("grad_reduce:vbias", launch_grad_reduce(s->d_vb_jobs + (chunk > 0 ? 1 : 0), 1, grad_reduce_blocks(s->vb_splits, V), st));
    GemmArgs g{};
    g.A = s->dlogits; g.lda = s->Vp; g.B = W + oE; g.ldb = H; g.b_trans = true; g.M = rc; g.N = H; g.K = s->Vp;
    g.a_kmax = s->Vp; g.b_kmax = V; g.d_M = s->counts + 1; g.d_M_off = r0; g.splits = s->dt_splits;
    g.out_f32 = s->dt_part + (size_t)r0 * H; g.ld_f32 = H; g.split_stride = (size_t)Mcap * H;
    KL("gemm:ce_dT", launch_gemm(EPI_F32_PARTIAL, g, st));
    WgradArgs w{};
    w.X = s->dlogits; w.ldx = s->Vp; w.dY = s->t + (size_t)r0 * H; w.ldy = H; w.M = V; w.N = H; w.T = rc;
    w.d_T = s->counts + 1; w.d_T_off = r0; w.splits = 1; w.out = G + oE; w.ld_out = H; w.accumulate = 1; w.x_mmax = s->Vp;
    KL("wgrad:ce_dE", launch_wgrad(w, st));
  }
  // ---- MLM transform backward
  const bf16* xL = s->layers.back().out;
  const float* dt_src = ext_dt ? ext_dt : s->dt_part;
  const int dt_n = ext_dt ? 1 : s->dt_splits;
  if (head_done) {
    // already issued beside the dE pass
  } else if (head_bwd_fused_supported(H)) {
    KL("head_bwd_fused", launch_head_bwd_fused(dt_src, dt_n, (size_t)Mcap * H, s->t_pre, s->t_act, s->hmean, s->hrstd,
                             P + s->lay.find("head/ln/gamma"), W + s->lay.find("head/wt"), xL, s->rows, s->counts, Mcap, s->dxa,
                             s->p_head_ln, s->p_wt, st, bwd_umma ? ce_xtiles : 0, bwd_umma ? ce_ctas : 0,
                             bwd_umma ? s->dt_max_splits : 0));
  } else {
  KL("head_bwd_rows", launch_head_bwd_rows(dt_src, dt_n, (size_t)Mcap * H, s->t_pre, s->t_act, s->hmean, s->hrstd,
                          P + s->lay.find("head/ln/gamma"), s->d_tpre, s->p_head_ln, Mcap, s->counts, H, st,
                          bwd_umma ? ce_xtiles : 0, bwd_umma ? ce_ctas : 0, bwd_umma ? s->dt_max_splits : 0));
  {
    WgradArgs w{};
    w.X = xL; w.ldx = H; w.x_rows = s->rows; w.dY = s->d_tpre; w.ldy = H; w.M = H; w.N = H; w.T = Mcap;
    w.d_T = s->counts;   // n_valid rows only: aux rows carry no gradient
    w.splits = s->s_wt; w.out = s->p_wt; w.split_stride = (size_t)H * H; w.ld_out = H;
    KL("wgrad:head_wt", launch_wgrad(w, st));
    GemmArgs g{};
    g.A = s->d_tpre; g.lda = H; g.B = W + s->lay.find("head/wt"); g.ldb = H; g.b_trans = false; g.M = Mcap; g.N = H; g.K = H;
    g.d_M = s->counts;  // n_valid only: aux rows carry no gradient and may alias position 0 of a valid slot
    g.out_f32 = s->dxa; g.ld_f32 = H; g.scatter_rows = s->rows;
    KL("gemm:head_dx_scatter", launch_gemm(EPI_SCATTER_F32, g, st));
  }
  }
  // ---- encoder layers, last to first.  d_out lives in dxa at the top of every iteration.
  const bool fbwd = s->use_fused_bwd && s->use_fused;   // the fused backward assumes the fused forward's tables
  if (fbwd) {
    EncBwdArgs f{};
    f.mask = s->mask; f.x0 = s->x0; f.dev_tables = s->d_enc_tables; f.dx = s->dxa; f.wpart = s->enc_wpart; f.bpart = s->enc_bpart;
    f.B = s->B; f.S = s->S; f.L = s->cfg.num_layers; f.I = I;
    f.out_drop = od; f.attn_drop = s->cfg.attention_dropout; f.seed = seed; f.step = step; f.d_step = d_step;
    f.dbg = getenv("B4R_FUSED_DEBUG") ? (void*)s->dbg_buf : nullptr;
    f.ids = s->ids; f.table = W + oE; f.pos = W + s->lay.find("position_embedding"); f.emb_g = P + s->lay.find("emb_ln/gamma");
    f.dx_rows = s->dxa; f.dpos_part = s->enc_dpos; f.embln_part = s->enc_embln; f.V = V;
    KL("enc_bwd_fused", launch_enc_bwd_fused(f, st));
  }
  for (int l = s->cfg.num_layers - 1; l >= 0 && !fbwd; --l) {
    LayerBuf& L = s->layers[l];
    const bf16* x_in = l == 0 ? s->x0 : s->layers[l - 1].out;
    // The gradient of the residual stream stays fp32 in dxa and is updated IN PLACE by the LayerNorm backward kernels; what the two
    // data-gradient GEMMs of a layer add to it travels as bf16 (dg) and is summed by the next consumer -- 12 bytes per element less
    // than an fp32 residual-in / fp32-out GEMM epilogue followed by an fp32 read.
    KL("ln_bwd", launch_ln_bwd(s->dxa, l == s->cfg.num_layers - 1 ? nullptr : s->dg, L.o_pre, L.mean2, L.rstd2, P + L.g2, s->dxa, s->d_branch, L.p_ln2, T, H, od, seed,
                     site_id(SITE_FFN_OUT, l), step, d_step, st));
    {
      WgradArgs w{};
      w.X = L.h; w.ldx = I; w.dY = s->d_branch; w.ldy = H; w.M = I; w.N = H; w.T = T; w.splits = L.s_w2;
      w.out = L.p_w2; w.split_stride = (size_t)I * H; w.ld_out = H;
      KL("wgrad:w2", launch_wgrad(w, st));
    }
    {
      GemmArgs g{};
      g.A = s->d_branch; g.lda = H; g.B = W + L.w2; g.ldb = H; g.b_trans = false; g.M = T; g.N = I; g.K = H;
      g.aux_bf16 = L.h_pre; g.ld_aux = I; g.out_bf16 = s->dh; g.ld_out = I; g.colsum_part = L.p_b1;
      KL("gemm:ffn2_dgrad_gelu", launch_gemm(EPI_GELU_GRAD, g, st));
    }
    {
      WgradArgs w{};
      w.X = L.y; w.ldx = H; w.dY = s->dh; w.ldy = I; w.M = H; w.N = I; w.T = T; w.splits = L.s_w1;
      w.out = L.p_w1; w.split_stride = (size_t)H * I; w.ld_out = I;
      KL("wgrad:w1", launch_wgrad(w, st));
    }
    {
      GemmArgs g{};
      g.A = s->dh; g.lda = I; g.B = W + L.w1; g.ldb = I; g.b_trans = false; g.M = T; g.N = H; g.K = I;
      g.out_bf16 = s->dg; g.ld_out = H;
      KL("gemm:ffn1_dgrad", launch_gemm(EPI_BF16, g, st));
    }
    KL("ln_bwd", launch_ln_bwd(s->dxa, s->dg, L.a_pre, L.mean1, L.rstd1, P + L.g1, s->dxa, s->d_branch, L.p_ln1, T, H, od, seed,
                     site_id(SITE_ATTN_OUT, l), step, d_step, st));
    {
      WgradArgs w{};
      w.X = L.ctx; w.ldx = H; w.dY = s->d_branch; w.ldy = H; w.M = H; w.N = H; w.T = T; w.splits = L.s_wo;
      w.out = L.p_wo; w.split_stride = (size_t)H * H; w.ld_out = H;
      KL("wgrad:wo", launch_wgrad(w, st));
    }
    {
      GemmArgs g{};
      g.A = s->d_branch; g.lda = H; g.B = W + L.wo; g.ldb = H; g.b_trans = false; g.M = T; g.N = H; g.K = H;
      g.out_bf16 = s->dctx; g.ld_out = H;
      KL("gemm:attn_out_dgrad", launch_gemm(EPI_BF16, g, st));
    }
    {
      AttnArgs a{};
      a.qkv = L.qkv; a.mask = s->mask; a.ctx = L.ctx; a.lse = L.lse; a.keep_bits = L.keep; a.B = s->B; a.S = s->S; a.H = H; a.N = s->N;
      a.drop_rate = s->cfg.attention_dropout; a.dctx = s->dctx; a.dqkv = s->dqkv; a.no_tcgen05 = !s->use_fattn;
      KL("attn_bwd", launch_attn_bwd(a, st));
    }
    const bool bias_in_wgrad = twgrad_shape_ok(H, 3 * H, T);
    if (!bias_in_wgrad) KL("colsum:bqkv", launch_colsum_bf16(s->dqkv, 3 * H, T, 3 * H, nullptr, 0, L.p_bqkv, kColsumSplits, st));
    {
      WgradArgs w{};
      w.X = x_in; w.ldx = H; w.dY = s->dqkv; w.ldy = 3 * H; w.M = H; w.N = 3 * H; w.T = T; w.splits = L.s_wqkv;
      w.out = L.p_wqkv; w.split_stride = (size_t)H * 3 * H; w.ld_out = 3 * H;
      if (bias_in_wgrad) {
        w.colsum_part = L.p_bqkv;
        if (!twgrad_supported(w)) return fail("the tcgen05 weight-gradient kernel rejected the Wqkv problem it was planned for");
      }
      KL("wgrad:wqkv", launch_wgrad(w, st));
    }
    {
      GemmArgs g{};
      g.A = s->dqkv; g.lda = 3 * H; g.B = W + L.wqkv; g.ldb = 3 * H; g.b_trans = false; g.M = T; g.N = H; g.K = 3 * H;
      g.out_bf16 = s->dg; g.ld_out = H;
      KL("gemm:qkv_dgrad", launch_gemm(EPI_BF16, g, st));
    }
  }
  if (!fbwd) KL("embed_bwd", launch_embed_bwd(s->ids, W + oE, W + s->lay.find("position_embedding"), P + s->lay.find("emb_ln/gamma"), s->dxa,
                      s->cfg.num_layers > 0 ? s->dg : nullptr, s->dxa, s->p_dpos, s->p_embln, s->B, s->S, H, V, od, seed, step, d_step, s->emb_bsplits, st));
  // item-table gradient of the gather: fixed-order per-item sums of the dx rows on top of the tied-projection part.  It touches the
  // table only, the final reduction everything else: the two run as parallel branches.
  CK(cudaEventRecord(s->ev_sort_fork, st));
  CK(cudaStreamWaitEvent(s->side_sort, s->ev_sort_fork, 0));
  {
    cudaStream_t st_main = st; (void)st_main;
    cudaStream_t st = s->side_sort;
    KL("table_grad", launch_table_grad(s->tg, st));
    s->launches += 1;
  }
  CK(cudaEventRecord(s->ev_sort, s->side_sort));
  s->sort_pending = false;
  if (fbwd) KL("grad_reduce:all", launch_grad_reduce(s->d_jobs_f, s->n_jobs_f, s->jobs_f_blocks, st));
  else KL("grad_reduce:all", launch_grad_reduce(s->d_jobs, s->n_jobs, s->jobs_max_len, st));
  CK(cudaStreamWaitEvent(st, s->ev_sort, 0));
  return 0;
}

extern "C" int b4r_backward(b4r_session* s, uint64_t seed, uint32_t step, const int64_t* step_counter, void* stream) {
  return backward_impl(s, seed, step, step_counter, stream, nullptr);
}
extern "C" int b4r_backward_from_dt(b4r_session* s, const float* dt, uint64_t seed, uint32_t step, const int64_t* step_counter,
                                    void* stream) {
  if (!dt) return fail("null dt");
  return backward_impl(s, seed, step, step_counter, stream, dt);
}
extern "C" const int32_t* b4r_mlm_labels(b4r_session* s) { return s ? s->labels : nullptr; }
extern "C" const float* b4r_mlm_row_weights(b4r_session* s) { return s ? s->row_w : nullptr; }
extern "C" const int32_t* b4r_mlm_row_mult(b4r_session* s) { return s ? s->row_mult : nullptr; }

// ------------------------------------------------------------------------------------------------ vocabulary shard
// The tied output projection of ONE rank's slice [v_begin, v_end) of the catalogue over the masked-slot rows of ALL ranks
// (SURVEY 8e "large catalogue"; the reference computes the full [B,P,V] logits on one device, bert4rec_model.py:139-147).
struct b4r_shard {
  int H, V, v_begin, v_end, Vs, n_ranks, Mcap, cap;
  const bf16* table; const float* vbias; float* g_table; float* g_bias;
  bf16* rows; int *lab_local, *lab_global, *mult, *counts; float* w;
  const int* counts_in;   // [n][2] of the last pack (device, caller-owned: must stay alive until the backward)
  float *ce_part, *lse, *lab, *step_stats, *fin_part; int* ticket;
  int fwd_max_splits, dt_max_splits, me_splits;
  float *dt_part, *p_dE, *p_dbias;
  ReduceJob* d_jobs;
  CeUmmaMaps maps;
};

static size_t shard_carve(b4r_shard* s, void* ws, size_t ws_bytes, bool dry) {
  Bump b{reinterpret_cast<char*>(ws), 0, ws_bytes, dry};
  const int H = s->H, cap = s->cap, Vs = s->Vs;
  const size_t budget = (size_t)1 << 30;   // partial buffers are capped at 1 GB each (large row counts need no splits anyway)
  s->rows = b.take<bf16>((size_t)cap * H);
  s->lab_local = b.take<int>(cap); s->lab_global = b.take<int>(cap); s->mult = b.take<int>(cap); s->w = b.take<float>(cap);
  s->counts = b.take<int>(8);
  const int vtiles = (Vs + 127) / 128;
  {
    size_t ms = budget / ((size_t)cap * 6 * sizeof(float));
    int vs = vtiles < 64 ? vtiles : 64;
    if ((size_t)vs > ms) vs = (int)ms;
    s->fwd_max_splits = vs < 1 ? 1 : vs;
  }
  s->ce_part = b.take<float>((size_t)s->fwd_max_splits * cap * 6);
  s->lse = b.take<float>(cap); s->lab = b.take<float>(cap);
  s->step_stats = b.take<float>(8); s->fin_part = b.take<float>(kCeFinalizeMaxBlocks * 5); s->ticket = b.take<int>(4);
  {
    const int xt = ce_bwd_umma_xtile(H), xtiles = (Vs + xt - 1) / xt;
    size_t ms = budget / ((size_t)cap * H * sizeof(float));
    int vs = xtiles < 24 ? xtiles : 24;
    if ((size_t)vs > ms) vs = (int)ms;
    s->dt_max_splits = vs < 1 ? 1 : vs;
    s->me_splits = choose_me_splits(vtiles, xt == 64 ? 148 : 2 * 148);
  }
  s->dt_part = b.take<float>((size_t)s->dt_max_splits * cap * H);
  s->p_dE = b.take<float>((size_t)s->me_splits * Vs * H);
  s->p_dbias = b.take<float>((size_t)s->me_splits * Vs);
  s->d_jobs = b.take<ReduceJob>(2);
  return b.off;
}

static int shard_check(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end) {
  if (hidden != 64 && hidden != 128 && hidden != 256) return fail("vocabulary shard: hidden %d not in {64, 128, 256}", hidden);
  if (n_ranks < 1 || n_ranks > 64 || rows_per_rank < 1) return fail("vocabulary shard: bad row layout %d x %d", n_ranks, rows_per_rank);
  if ((long long)n_ranks * rows_per_rank > (1ll << 30)) return fail("vocabulary shard: too many rows");
  if (v_begin < 0 || v_end > vocab || v_begin >= v_end) return fail("bad vocabulary shard [%d, %d) of %d", v_begin, v_end, vocab);
  return 0;
}

static b4r_shard* shard_shell(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end) {
  b4r_shard* s = new b4r_shard();
  s->H = hidden; s->V = vocab; s->v_begin = v_begin; s->v_end = v_end; s->Vs = v_end - v_begin;
  s->n_ranks = n_ranks; s->Mcap = rows_per_rank; s->cap = n_ranks * rows_per_rank;
  return s;
}

extern "C" size_t b4r_shard_workspace_bytes(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end) {
  if (shard_check(hidden, vocab, n_ranks, rows_per_rank, v_begin, v_end)) return 0;
  b4r_shard* s = shard_shell(hidden, vocab, n_ranks, rows_per_rank, v_begin, v_end);
  const size_t n = shard_carve(s, nullptr, 0, true);
  delete s;
  return n;
}

extern "C" int b4r_shard_create(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end, const void* table_bf16,
                                const float* output_bias, float* grad_table, float* grad_bias, void* workspace,
                                size_t workspace_bytes, b4r_shard** out) {
  if (!out) return fail("null out");
  if (shard_check(hidden, vocab, n_ranks, rows_per_rank, v_begin, v_end)) return 1;
  if (!table_bf16 || !output_bias || !grad_table || !grad_bias || !workspace) return fail("null buffer");
  if ((uintptr_t)workspace & 255) return fail("workspace must be 256-byte aligned");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (b4r_device_check(dev)) return 1;
  b4r_shard* s = shard_shell(hidden, vocab, n_ranks, rows_per_rank, v_begin, v_end);
  s->table = reinterpret_cast<const bf16*>(table_bf16); s->vbias = output_bias; s->g_table = grad_table; s->g_bias = grad_bias;
  const size_t need = shard_carve(s, workspace, workspace_bytes, false);
  if (need > workspace_bytes) { delete s; return fail("workspace too small: need %zu bytes", need); }
  // zero once: rows behind the packed ones are read (never used) by the tile loads and must stay finite
  if (cudaMemset(workspace, 0, need) != cudaSuccess) { delete s; return fail("cudaMemset failed"); }
  ReduceJob j[2];
  j[0] = ReduceJob{s->p_dE, grad_table + (size_t)v_begin * hidden, s->me_splits, s->Vs * hidden, (long long)s->Vs * hidden, 0};
  j[1] = ReduceJob{s->p_dbias, grad_bias + v_begin, s->me_splits, s->Vs, (long long)s->Vs, 0};
  if (cudaMemcpy(s->d_jobs, j, sizeof(j), cudaMemcpyHostToDevice) != cudaSuccess) { delete s; return fail("cudaMemcpy failed"); }
  if (!ce_umma_make_maps(&s->maps, s->rows, s->cap, s->table + (size_t)v_begin * hidden, s->Vs, hidden)) {
    delete s;
    return fail("vocabulary shard: tensor maps could not be created");
  }
  *out = s;
  return 0;
}

extern "C" void b4r_shard_destroy(b4r_shard* s) { delete s; }

// rows_in bf16 [n][rows_per_rank][hidden], labels_in int32 / weights_in fp32 / mult_in int32 [n][rows_per_rank], counts_in int32 [n][2]
// = every rank's {n_valid, n_rows}: the all-gathered b4r_mlm_hidden / labels / row_weights / row_mult / counts buffers.
extern "C" int b4r_shard_pack(b4r_shard* s, const void* rows_in, const int32_t* labels_in, const float* weights_in,
                              const int32_t* mult_in, const int32_t* counts_in, void* stream) {
  if (!s || !rows_in || !labels_in || !weights_in || !mult_in || !counts_in) return fail("null argument");
  s->counts_in = counts_in;
  CK(launch_shard_pack(reinterpret_cast<const bf16*>(rows_in), labels_in, weights_in, mult_in, counts_in, s->n_ranks, s->Mcap, s->H,
                       s->v_begin, s->rows, s->lab_local, s->lab_global, s->w, s->mult, s->counts, (cudaStream_t)stream));
  return 0;
}

static CeArgs shard_ce_args(b4r_shard* s) {
  CeArgs c{};
  c.t = s->rows; c.ldt = s->H; c.E = s->table + (size_t)s->v_begin * s->H; c.vbias = s->vbias + s->v_begin;
  c.labels = s->lab_local; c.row_w = s->w; c.row_mult = s->mult; c.d_counts = s->counts;
  c.M_cap = s->cap; c.H = s->H; c.V = s->Vs; c.v_begin = 0; c.v_end = s->Vs;
  c.part = s->ce_part; c.lse = s->lse; c.lab_out = s->lab; c.step_stats = s->step_stats; c.fin_part = s->fin_part; c.ticket = s->ticket;
  c.target_ctas = 2 * 148; c.max_splits = s->fwd_max_splits;
  return c;
}

// forward over the shard: part_out fp32 [n*rows_per_rank][6] = per packed row {max, sum exp(x - max), label logit (-inf when
// the label lives in another shard), best logit, best GLOBAL index (int bits), 0}
extern "C" int b4r_shard_ce_partial(b4r_shard* s, float* part_out, void* stream) {
  if (!s || !part_out) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CeArgs c = shard_ce_args(s);
  CK(launch_ce_fwd_umma(s->maps, c, st));
  CK(launch_shard_part_merge(s->ce_part, s->counts, s->cap, (s->Vs + 127) / 128, c.target_ctas, c.max_splits, s->v_begin, part_out, st));
  return 0;
}

// parts fp32 [n_shards][n*rows_per_rank][6] (all-gathered b4r_shard_ce_partial outputs): log-sum-exp per row, loss / accuracy
// sums of the GLOBAL batch into the shard's step statistics and (optional) the running stats (same layout as b4r_mlm_loss)
extern "C" int b4r_shard_ce_merge(b4r_shard* s, const float* parts, int n_shards, int global_batch, float* stats, void* stream) {
  if (!s || !parts || n_shards < 1) return fail("bad argument");
  CeArgs c = shard_ce_args(s);
  c.labels = s->lab_global;
  c.part = const_cast<float*>(parts); c.vsplits = n_shards; c.batch = global_batch; c.stats = stats;
  CK(launch_ce_finalize(c, (cudaStream_t)stream));
  return 0;
}

// backward over the shard.  dt_out fp32 [n][rows_per_rank][hidden]: this shard's partial gradient of every rank's transformed
// rows (reduce-scatter(SUM) over the shards gives each rank its dT for b4r_backward_from_dt).  The shard's slice of the table
// and output-bias gradients is complete and is WRITTEN to grad_table / grad_bias; with zero_all the rest of both is zeroed
// first, so that an all-reduce(SUM) over ranks assembles the whole projection gradient.
extern "C" int b4r_shard_ce_backward(b4r_shard* s, float* dt_out, int zero_all, void* stream) {
  if (!s || !dt_out) return fail("null argument");
  if (!s->counts_in) return fail("b4r_shard_pack must run before b4r_shard_ce_backward");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = s->H, xt = ce_bwd_umma_xtile(H), xtiles = (s->Vs + xt - 1) / xt, ctas = xt == 64 ? 148 : 2 * 148;
  if (zero_all) {
    CK(cudaMemsetAsync(s->g_table, 0, (size_t)s->V * H * sizeof(float), st));
    CK(cudaMemsetAsync(s->g_bias, 0, (size_t)s->V * sizeof(float), st));
  }
  CeBwdArgs ba{};
  ba.vbias = s->vbias + s->v_begin; ba.lse = s->lse; ba.row_w = s->w; ba.labels = s->lab_local; ba.d_counts = s->counts;
  ba.M_cap = s->cap; ba.V = s->Vs; ba.H = H; ba.target_ctas = ctas; ba.max_splits = s->dt_max_splits; ba.msplits = s->me_splits;
  ba.out = s->dt_part; ba.dbias_out = nullptr;
  CK(launch_ce_bwd_umma(s->maps, ba, true, st));
  CK(launch_shard_dt_unpack(s->dt_part, s->counts, s->counts_in, s->n_ranks, s->Mcap, s->cap, H, xtiles, ctas, s->dt_max_splits,
                            dt_out, st));
  ba.out = s->p_dE; ba.dbias_out = s->p_dbias;
  CK(launch_ce_bwd_umma(s->maps, ba, false, st));
  CK(launch_grad_reduce(s->d_jobs, 2, grad_reduce_blocks(s->me_splits, s->Vs * H), st));
  return 0;
}
extern "C" const float* b4r_shard_step_stats(b4r_shard* s) { return s ? s->step_stats : nullptr; }
extern "C" const float* b4r_shard_lse(b4r_shard* s) { return s ? s->lse : nullptr; }
extern "C" const int32_t* b4r_shard_counts(b4r_shard* s) { return s ? s->counts : nullptr; }

// ------------------------------------------------------------------------------------------------ pooler
__global__ void pooler_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ out, int B, int S, int H) {
  const int bi = blockIdx.x;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float acc = b[j];
    for (int k = 0; k < H; ++k) acc += __bfloat162float(x[(size_t)bi * S * H + k]) * w[(size_t)k * H + j];
    out[(size_t)bi * H + j] = tanhf(acc);
  }
}
extern "C" int b4r_pooled_output(b4r_session* s, float* out, void* stream) {
  if (!s || !out) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  pooler_kernel<<<s->B, 128, 0, st>>>(s->layers.back().out, s->params + s->lay.find("pooler/w"),
                                                       s->params + s->lay.find("pooler/b"), out, s->B, s->S, s->H);
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ optimizer
static const int kSqParts = 296;
extern "C" size_t b4r_adamw_scratch_floats(void) { return kSqParts + 16; }   // partials + coefficients + ticket (zero-initialised by the caller)
extern "C" int b4r_adamw_step(float* params, void* shadow_bf16, const float* grads, float* m, float* v, int64_t n_decay,
                              int64_t n_trainable, const b4r_adamw_hparams* hp, const float* count, float grad_scale,
                              int64_t* step_counter, float* scratch, float* lr_out, void* stream) {
  if (!params || !shadow_bf16 || !grads || !m || !v || !hp || !step_counter || !scratch) return fail("null argument");
  if ((n_decay & 7) || (n_trainable & 7)) return fail("segment sizes must be multiples of 8");
  AdamWArgs a{};
  a.p = params; a.shadow = reinterpret_cast<bf16*>(shadow_bf16); a.g = grads; a.m = m; a.v = v;
  a.n_decay = n_decay; a.n = n_trainable; a.sq_part = scratch; a.n_sq_part = kSqParts; a.d_count = count; a.grad_scale = grad_scale;
  a.d_step = reinterpret_cast<const long long*>(step_counter); a.d_step_out = reinterpret_cast<long long*>(step_counter);
  a.init_lr = hp->init_lr; a.end_lr = hp->end_lr; a.num_train_steps = hp->num_train_steps; a.num_warmup_steps = hp->num_warmup_steps;
  a.wd = hp->weight_decay_rate; a.beta1 = hp->beta_1; a.beta2 = hp->beta_2; a.eps = hp->epsilon; a.clip = hp->clip_norm;
  a.d_lr_out = lr_out;
  cudaStream_t st = (cudaStream_t)stream;
  KL_((void)0, "sqnorm+adamw", launch_adamw(a, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------ ranking
extern "C" int b4r_rank_candidates(b4r_session* s, const int64_t* cand, const int64_t* gt, int n_slots, int C,
                                   int64_t* ranking_out, float* scores_out, int32_t* rank_out, uint64_t* hist,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !cand) return fail("null argument");
  if (C < 1 || C > 2048) return fail("candidate count %d unsupported (1..2048)", C);
  if (n_slots > s->Mcap) return fail("n_slots %d exceeds session capacity %d", n_slots, s->Mcap);
  if (join_select(s, st)) return 1;
  KL("rank_candidates", launch_rank_candidates(s->t, s->H, s->shadow + s->lay.find("word_embeddings"), s->params + s->lay.find("head/output_bias"),
                            cand, gt, n_slots, C, s->H, s->V, s->counts, reinterpret_cast<int64_t*>(ranking_out), scores_out, rank_out,
                            reinterpret_cast<unsigned long long*>(hist), st));
  return 0;
}

extern "C" int b4r_rank_full(b4r_session* s, int v_begin, int v_end, int32_t* beat_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !beat_out) return fail("null argument");
  if (v_begin < 0 || v_end > s->V || v_begin >= v_end) return fail("bad vocabulary shard [%d, %d)", v_begin, v_end);
  // ground-truth scores: label logit of the fused CE pass over the FULL vocabulary (b4r_mlm_loss must have run)
  CeArgs c = ce_args(s);
  c.v_begin = v_begin; c.v_end = v_end;
  KL("ce_count", launch_ce_count(c, s->lab, beat_out, st));
  return 0;
}

// Same count for EXTERNAL rows (vocab-sharded evaluation over several GPUs: the hidden rows, labels and ground-truth scores
// of every rank are all-gathered, each rank counts over ITS vocabulary shard, one all-reduce of the counts gives exact ranks).
extern "C" int b4r_rank_full_ext(b4r_session* s, const void* t_rows, const int32_t* labels, const float* gt_scores,
                                 const int32_t* counts2, int rows_cap, int v_begin, int v_end, int32_t* beat_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !t_rows || !labels || !gt_scores || !counts2 || !beat_out) return fail("null argument");
  if (v_begin < 0 || v_end > s->V || v_begin >= v_end) return fail("bad vocabulary shard [%d, %d)", v_begin, v_end);
  if (rows_cap < 1) return fail("rows_cap %d", rows_cap);
  CeArgs c = ce_args(s);
  c.t = reinterpret_cast<const bf16*>(t_rows); c.labels = labels; c.d_counts = counts2; c.M_cap = rows_cap;
  c.v_begin = v_begin; c.v_end = v_end;
  KL("ce_count", launch_ce_count(c, gt_scores, beat_out, st));
  return 0;
}

// Full-catalogue top-k (rank_items(items=None) / apps.Recommender, bert4rec_model.py:235-236, apps/recommender.py:14-63): the K best
// items of every row over the vocabulary shard [v_begin, v_end) as order-preserving keys (see k_ce.cu), best first.
extern "C" size_t b4r_topk_scratch_bytes(int n_rows, int v_begin, int v_end, int K) {
  if (n_rows < 1 || v_end <= v_begin || K < 1) return 0;
  return topk_scratch_bytes(n_rows, v_end - v_begin, K);
}
extern "C" int b4r_topk_full(b4r_session* s, const void* t_rows, int n_rows, int v_begin, int v_end, int K, void* scratch,
                             uint64_t* keys_out, int64_t* ids_out, float* scores_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!s || !scratch) return fail("null argument");
  if (v_begin < 0 || v_end > s->V || v_begin >= v_end) return fail("bad vocabulary shard [%d, %d)", v_begin, v_end);
  if (K < 1 || K > 128) return fail("K %d unsupported (1..128)", K);
  if (n_rows < 1 || (!t_rows && n_rows > s->Mcap)) return fail("n_rows %d", n_rows);
  if (!t_rows && join_select(s, st)) return 1;
  CeArgs c = ce_args(s);
  c.v_begin = v_begin; c.v_end = v_end;
  if (t_rows) { c.t = reinterpret_cast<const bf16*>(t_rows); c.d_counts = nullptr; }
  else c.d_counts = s->counts;
  cudaError_t e = launch_topk_full(c, n_rows, K, reinterpret_cast<unsigned long long*>(scratch), reinterpret_cast<unsigned long long*>(keys_out),
                                   reinterpret_cast<long long*>(ids_out), scores_out, st);
  if (e != cudaSuccess) return fail("topk_full: %s (K %d may not fit shared memory at hidden %d)", cudaGetErrorString(e), K, s->H);
  s->launches += 2;
  return 0;
}
extern "C" int b4r_topk_merge(const uint64_t* keys_in, int nlists, int n_rows, int K, uint64_t* keys_out, int64_t* ids_out,
                              float* scores_out, void* stream) {
  if (!keys_in || nlists < 1 || n_rows < 1 || K < 1 || (long long)nlists * K > 8192) return fail("bad argument");
  cudaError_t e = launch_topk_merge(reinterpret_cast<const unsigned long long*>(keys_in), nlists, n_rows, K, nullptr,
                                    reinterpret_cast<unsigned long long*>(keys_out), reinterpret_cast<long long*>(ids_out), scores_out,
                                    (cudaStream_t)stream);
  if (e != cudaSuccess) return fail("topk_merge: %s", cudaGetErrorString(e));
  return 0;
}

// The step's five int64 inputs travel host -> device in a compact form (ids / positions / labels as int32, the two 0/1 arrays as
// bytes: 2.9x fewer PCIe bytes than the int64 tensors of the reference's batch dict) and are widened on the device into the
// persistent int64 buffers the captured step reads.  packed = [int32 ids n_tok][int32 positions n_pred][int32 mlm_ids n_pred]
// [uint8 mask n_tok][uint8 weights n_pred].
__global__ void __launch_bounds__(256) unpack_inputs_kernel(const int* __restrict__ p32, const unsigned char* __restrict__ p8, int n_tok,
                                                            int n_pred, long long* __restrict__ ids, long long* __restrict__ mask,
                                                            long long* __restrict__ pos, long long* __restrict__ mlm,
                                                            long long* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tok) { ids[i] = p32[i]; mask[i] = p8[i]; }
  if (i < n_pred) { pos[i] = p32[n_tok + i]; mlm[i] = p32[n_tok + n_pred + i]; w[i] = p8[n_tok + i]; }
}
extern "C" size_t b4r_packed_inputs_bytes(int n_tok, int n_pred) { return (size_t)4 * ((size_t)n_tok + 2 * (size_t)n_pred) + n_tok + n_pred; }
extern "C" int b4r_unpack_inputs(const void* packed, int n_tok, int n_pred, int64_t* input_word_ids, int64_t* input_mask,
                                 int64_t* masked_lm_positions, int64_t* masked_lm_ids, int64_t* masked_lm_weights, void* stream) {
  if (!packed || n_tok < 1 || n_pred < 0 || !input_word_ids || !input_mask || (n_pred && (!masked_lm_positions || !masked_lm_ids || !masked_lm_weights)))
    return fail("bad argument");
  const int* p32 = reinterpret_cast<const int*>(packed);
  const unsigned char* p8 = reinterpret_cast<const unsigned char*>(packed) + (size_t)4 * ((size_t)n_tok + 2 * (size_t)n_pred);
  const int n = n_tok > n_pred ? n_tok : n_pred;
  unpack_inputs_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      p32, p8, n_tok, n_pred, reinterpret_cast<long long*>(input_word_ids), reinterpret_cast<long long*>(input_mask),
      reinterpret_cast<long long*>(masked_lm_positions), reinterpret_cast<long long*>(masked_lm_ids),
      reinterpret_cast<long long*>(masked_lm_weights));
  CK(cudaGetLastError());
  return 0;
}

// n host->device copies enqueued back to back (pinned sources: true DMA, the host returns at once).  The step's five int64 inputs
// go to their persistent device views with ~2 us of host time each instead of a packing pass over the batch on the host.
extern "C" int b4r_h2d_copy_many(void* const* dst, const void* const* src, const size_t* nbytes, int n, void* stream) {
  if (!dst || !src || !nbytes || n < 0) return fail("bad argument");
  for (int i = 0; i < n; ++i) CK(cudaMemcpyAsync(dst[i], src[i], nbytes[i], cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int b4r_metrics_from_hist(const uint64_t* hist, int max_rank, const int32_t* ks, int nk, double* out, void* stream) {
  if (!hist || !ks || !out || nk < 0 || nk > 16) return fail("bad argument");
  CK(launch_metrics_from_hist(reinterpret_cast<const unsigned long long*>(hist), max_rank, ks, nk, out, (cudaStream_t)stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ introspection
extern "C" const void* b4r_sequence_output(b4r_session* s, int layer) {
  if (!s) return nullptr;
  int L = s->cfg.num_layers;
  if (layer < 0) layer += L;
  if (layer < 0 || layer >= L) return nullptr;
  return s->layers[layer].out;
}
extern "C" const void* b4r_mlm_hidden(b4r_session* s) { return s ? s->t : nullptr; }
extern "C" const int32_t* b4r_mlm_counts(b4r_session* s) { return s ? s->counts : nullptr; }
extern "C" const int32_t* b4r_mlm_rows(b4r_session* s) { return s ? s->rows : nullptr; }
extern "C" float* b4r_step_stats(b4r_session* s) { return s ? s->step_stats : nullptr; }
extern "C" const uint64_t* b4r_attn_keep_bits(b4r_session* s, int layer, int* words_per_row) {
  if (!s || layer < 0 || layer >= s->cfg.num_layers) return nullptr;
  if (words_per_row) *words_per_row = attn_mask_words(s->S);
  return s->layers[layer].keep;
}
// saved activation of one encoder layer by name (bf16 [T, cols] unless *is_f32): development / parity-test aid
extern "C" const void* b4r_layer_tensor(b4r_session* s, int layer, const char* name, int* cols, int* is_f32) {
  if (!s || !name || layer < 0 || layer >= s->cfg.num_layers) return nullptr;
  const LayerBuf& L = s->layers[layer];
  const std::string n(name);
  int c = s->H, f = 0;
  const void* p = nullptr;
  if (n == "x0") p = s->x0;
  else if (n == "qkv") { p = L.qkv; c = 3 * s->H; }
  else if (n == "ctx") p = L.ctx;
  else if (n == "a_pre") p = L.a_pre;
  else if (n == "y") p = L.y;
  else if (n == "h_pre") { p = L.h_pre; c = s->I; }
  else if (n == "h") { p = L.h; c = s->I; }
  else if (n == "o_pre") p = L.o_pre;
  else if (n == "out") p = L.out;
  else if (n == "mean1") { p = L.mean1; c = 1; f = 1; }
  else if (n == "rstd1") { p = L.rstd1; c = 1; f = 1; }
  else if (n == "mean2") { p = L.mean2; c = 1; f = 1; }
  else if (n == "rstd2") { p = L.rstd2; c = 1; f = 1; }
  else if (n == "lse") { p = L.lse; c = 1; f = 1; }
  if (cols) *cols = c;
  if (is_f32) *is_f32 = f;
  return p;
}
extern "C" int b4r_launch_count(b4r_session* s) { return s ? s->launches : 0; }
extern "C" const void* b4r_debug_buffer(b4r_session* s) { return s ? (const void*)(s->ce_part + (size_t)(s->vsplits_umma - 1) * s->Mcap * 6) : nullptr; }
extern "C" const void* b4r_debug_buffer2(b4r_session* s) { return s ? (const void*)s->dbg_buf : nullptr; }
extern "C" int b4r_session_set_flag(b4r_session* s, int flag, int value) {
  if (!s) return fail("null session");
  if (flag == 1) { s->use_umma = value != 0; return 0; }
  if (flag == 4) { s->overlap_select = value != 0; return 0; }
  if (flag == 6) { s->use_fattn = value != 0; return 0; }
  if (flag == 5) {
    if (value && !s->ce_fused_ok) return fail("one-pass CE backward unavailable for this shape");
    s->use_ce_fused = value != 0;
    return 0;
  }
  if (flag == 3) {
    if (value && !s->fused_bwd_ok) return fail("fused encoder backward unavailable for this shape");
    s->use_fused_bwd = value != 0;
    return 0;
  }
  if (flag == 2) {
    if (value && !s->d_enc_tables) return fail("fused encoder unavailable");
    s->use_fused = value != 0 && enc_fused_supported(s->H, s->N, s->S, s->I);
    return 0;
  }
  return fail("unknown flag %d", flag);
}
extern "C" int b4r_profile_enable(b4r_session* s, int on) {
  (void)s;
  g_prof_on = on != 0;
  if (!g_prof_on) {  // switching off drops every record, including the ones living in captured graphs
    cudaDeviceSynchronize();
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
  }
  return 0;
}
// Synchronises the device, aggregates the recorded per-launch event times by kernel tag and writes lines
// "tag count total_ms" into buf.  Eager records are consumed; records captured into a CUDA graph persist (they are
// re-timed by every replay) until profiling is switched off.
extern "C" int b4r_profile_report(b4r_session* s, char* buf, int cap) {
  (void)s;
  if (!buf || cap < 1) return fail("bad argument");
  CK(cudaDeviceSynchronize());
  std::vector<std::string> tags; std::vector<int> cnt; std::vector<double> tot;
  std::vector<ProfRec> keep;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); ms = -1.f; }
    if (r.in_graph) keep.push_back(r);
    else { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (ms < 0.f) continue;   // a captured record that no replay has timed yet
    size_t i = 0;
    for (; i < tags.size(); ++i) if (tags[i] == r.tag) break;
    if (i == tags.size()) { tags.push_back(r.tag); cnt.push_back(0); tot.push_back(0.0); }
    cnt[i]++; tot[i] += ms;
  }
  g_prof.swap(keep);
  std::string out;
  for (size_t i = 0; i < tags.size(); ++i) {
    char line[160];
    snprintf(line, sizeof(line), "%s %d %.6f\n", tags[i].c_str(), cnt[i], tot[i]);
    out += line;
  }
  snprintf(buf, cap, "%s", out.c_str());
  return 0;
}

// ------------------------------------------------------------------------------------------------ test helpers
extern "C" int b4r_dropout_keep_mask(uint8_t* out, int rows, int cols, float rate, uint64_t seed, int site, int layer,
                                     uint32_t step, void* stream) {
  if (!out) return fail("null argument");
  CK(launch_dropout_mask_dump(out, rows, cols, rate, seed, site_id((uint32_t)site, (uint32_t)layer), step, (cudaStream_t)stream));
  return 0;
}
extern "C" int b4r_embed_ln_fwd(const int64_t* ids, const void* table, const void* pos, const float* gamma, const float* beta,
                                void* out, int batch, int seq_len, int hidden, int vocab, void* stream) {
  if (!ids || !table || !pos || !gamma || !beta || !out) return fail("null argument");
  CK(launch_embed_ln_fwd(ids, (const bf16*)table, (const bf16*)pos, gamma, beta, (bf16*)out, batch, seq_len, hidden, vocab,
                         0.f, 0, 0, nullptr, (cudaStream_t)stream));
  return 0;
}

extern "C" size_t b4r_table_grad_workspace_bytes(int tokens, int hidden) {
  return (size_t)4 * 4 * tokens + 4 * 256 * (size_t)table_grad_sort_blocks(tokens) + 4 * (1 + 4 * (size_t)table_grad_max_long_runs(tokens)) +
         4 * (size_t)table_grad_chunks(tokens) * 2 * hidden + 8 * 256;
}
extern "C" int b4r_table_grad(const int64_t* ids, const float* dx, float* grad_table, int tokens, int vocab, int hidden,
                              void* workspace, void* stream) {
  if (!ids || !dx || !grad_table || !workspace) return fail("null argument");
  if (tokens < 1 || vocab < 1) return fail("empty problem");
  TableGradArgs a{};
  a.ids = ids; a.T = tokens; a.V = vocab; a.H = hidden; a.dx = dx; a.grad_table = grad_table;
  char* w = reinterpret_cast<char*>(workspace);
  auto take = [&](size_t bytes) { char* p = w; w += (bytes + 255) / 256 * 256; return p; };
  for (int i = 0; i < 2; ++i) { a.keys[i] = (uint32_t*)take(4 * (size_t)tokens); a.vals[i] = (uint32_t*)take(4 * (size_t)tokens); }
  a.hist = (uint32_t*)take(4 * 256 * (size_t)table_grad_sort_blocks(tokens));
  a.long_runs = (int*)take(4 * (1 + 4 * (size_t)table_grad_max_long_runs(tokens)));
  a.carry = (float*)take(4 * (size_t)table_grad_chunks(tokens) * 2 * hidden);
  CK(launch_token_sort(a, (cudaStream_t)stream));
  CK(launch_table_grad(a, (cudaStream_t)stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ data-parallel all-reduce
extern "C" int b4r_p2p_allreduce_max_world(void) { return p2p_allreduce_max_world(); }
extern "C" int b4r_p2p_allreduce_f32(const void* buffer_ptrs_dev, const void* flag_ptrs_dev, void* multicast_ptr, size_t offset_floats,
                                     size_t n_floats, int rank, int world, void* state, void* stream) {
  if (!buffer_ptrs_dev || !flag_ptrs_dev || !state) return fail("null argument");
  if (world < 2 || world > p2p_allreduce_max_world()) return fail("world size %d unsupported by the peer-memory all-reduce (2..%d)", world, p2p_allreduce_max_world());
  if (rank < 0 || rank >= world) return fail("rank %d outside the world of %d", rank, world);
  if (offset_floats % 4) return fail("the buffer range must start on a 16-byte boundary");
  CK(launch_p2p_allreduce(reinterpret_cast<float* const*>(buffer_ptrs_dev), reinterpret_cast<uint32_t* const*>(flag_ptrs_dev),
                          reinterpret_cast<float*>(multicast_ptr), offset_floats, n_floats, rank, world, reinterpret_cast<uint32_t*>(state),
                          (cudaStream_t)stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------ DLPack adapter
namespace {
struct DLDevice_ { int32_t device_type; int32_t device_id; };
struct DLDataType_ { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor_ { void* data; DLDevice_ device; int32_t ndim; DLDataType_ dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; };
struct DLManagedTensor_ { DLTensor_ dl_tensor; void* manager_ctx; void (*deleter)(DLManagedTensor_*); };
}  // namespace
extern "C" int b4r_dl_view_of(const void* dl_managed_tensor, b4r_dl_view* out) {
  if (!dl_managed_tensor || !out) return fail("null argument");
  const DLTensor_& t = reinterpret_cast<const DLManagedTensor_*>(dl_managed_tensor)->dl_tensor;
  if (t.device.device_type != 2 /* kDLCUDA */) return fail("tensor is not on a CUDA device (device_type %d)", t.device.device_type);
  if (t.ndim < 0 || t.ndim > 4) return fail("ndim %d unsupported", t.ndim);
  if (t.dtype.lanes != 1) return fail("vector dtypes unsupported");
  int64_t expect = 1;
  for (int i = t.ndim - 1; i >= 0; --i) {
    if (t.strides && t.shape[i] > 1 && t.strides[i] != expect) return fail("tensor is not contiguous row-major");
    expect *= t.shape[i];
  }
  out->data = reinterpret_cast<char*>(t.data) + t.byte_offset;
  out->device_type = t.device.device_type; out->device_id = t.device.device_id; out->ndim = t.ndim;
  out->dtype_code = t.dtype.code; out->dtype_bits = t.dtype.bits;
  for (int i = 0; i < 4; ++i) out->shape[i] = i < t.ndim ? t.shape[i] : 1;
  return 0;
}
