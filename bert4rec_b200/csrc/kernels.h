// Host-callable launchers for every kernel of the path.  Raw device pointers, explicit stream, no allocation,
// no hidden synchronisation.  All launchers return cudaError_t from the launch (cudaGetLastError()).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b4r {
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ generic GEMM (k_gemm.cu)
enum Epi : int {
  EPI_BIAS_BF16 = 0,   // out_bf16 = acc + bias
  EPI_BIAS_GELU = 1,   // out_bf16 = acc + bias (pre-activation) ; out2_bf16 = gelu(pre)
  EPI_GELU_GRAD = 2,   // out_bf16 = acc * gelu'(aux_bf16) ; per-CTA column sums -> colsum_part[mtile][N]
  EPI_BF16 = 3,        // out_bf16 = acc
  EPI_F32_RES = 4,     // out_f32 = acc + res_f32
  EPI_SCATTER_F32 = 5, // out_f32[scatter_rows[m]] = acc
  EPI_F32_PARTIAL = 6, // out_f32[split][m][n] = acc   (split-K partials)
  EPI_BIAS_F32 = 7,    // out_f32 = acc + bias          (materialised logits for the API / tests)
};

struct GemmArgs {
  // operands (see gemm.cuh)
  const bf16* A; int lda; const int* a_rows; bool a_trans;
  const bf16* B; int ldb; bool b_trans;
  int M, N, K;            // logical problem; K is the loop bound (multiple of 8)
  int a_kmax, b_kmax;     // valid k extents of each operand (<= K); 0 -> K
  const int* d_M;         // optional device-side dynamic row count (rows >= *d_M - d_M_off are skipped)
  int d_M_off;
  int splits;             // split-K factor (EPI_F32_PARTIAL only), >= 1
  // epilogue
  const float* bias;
  bf16* out_bf16; int ld_out;
  bf16* out2_bf16;
  const bf16* aux_bf16; int ld_aux;
  float* out_f32; int ld_f32;
  const float* res_f32;
  const int* scatter_rows;
  float* colsum_part;     // [ceil(M/BM)][N]
  size_t split_stride;
};
cudaError_t launch_gemm(int epi, const GemmArgs& a, cudaStream_t st);
int gemm_block_m();  // BM of the generic kernel (for sizing colsum_part)
// generation 2 (k_tgemm.cu): persistent tcgen05 GEMM with the same epilogues; launch_gemm dispatches to it when supported
bool tgemm_supported(int epi, const GemmArgs& a);
cudaError_t launch_tgemm(int epi, const GemmArgs& a, cudaStream_t st);

// GEMM with a full-row epilogue (N == H in {64,128,256}), k_gemm.cu
enum RowMode : int {
  ROW_RES_DROP_LN = 0,  // v = drop(acc+bias) + residual ; pre=v ; y = LN(v)
  ROW_GELU_LN = 1,      // pre = acc+bias ; act = gelu(pre) ; y = LN(act)
};
struct RowLnArgs {
  const bf16* A; int lda; const int* a_rows;   // A [M][K] (optionally row-gathered)
  const bf16* W;                               // [K][H] row-major (TF kernel layout)
  int M, K, H;
  const int* d_M;
  const float* bias; const float* gamma; const float* beta;
  const bf16* residual;                        // [M][H] (ROW_RES_DROP_LN)
  bf16* pre;                                   // [M][H] pre-LN value (RES mode) / pre-activation (GELU mode)
  bf16* act;                                   // [M][H] gelu(pre) (GELU mode only)
  bf16* y;                                     // [M][H] LN output
  float* mean; float* rstd;                    // [M]
  // dropout
  float drop_rate; uint64_t seed; uint32_t site; uint32_t step;
  const long long* d_step;                     // optional device counter added to step (CUDA-graph replays)
};
cudaError_t launch_gemm_rowln(int mode, const RowLnArgs& a, cudaStream_t st);
// generation 2 (k_tgemm.cu): tcgen05 GEMM + residual/dropout/LayerNorm epilogue for hidden 256; launch_gemm_rowln dispatches
bool trowln_supported(int mode, const RowLnArgs& a);
cudaError_t launch_trowln(const RowLnArgs& a, cudaStream_t st);

// Weight-gradient GEMM: out[split][M][N] = sum_{t in split} X[t][m] * dY[t][n]   (both operands "transposed")
struct WgradArgs {
  const bf16* X; int ldx; const int* x_rows;   // [T][M] (rows optionally gathered)
  const bf16* dY; int ldy;                     // [T][N]
  int M, N, T;
  const int* d_T; int d_T_off;                 // optional dynamic contraction length: min(T, *d_T - d_T_off)
  int splits;
  float* out; size_t split_stride;             // partials (accumulate==0) or final (splits==1, accumulate==1: out += )
  int ld_out; int accumulate;
  int x_mmax;                                  // valid columns of X (multiple of 8; defaults to M rounded up)
  float* colsum_part;                          // optional (tcgen05 kernel only): column sums of dY per token split, [splits][N]
};
cudaError_t launch_wgrad(const WgradArgs& a, cudaStream_t st);
// generation 2 (k_tgemm.cu): tcgen05 weight gradient for M, N multiples of 256; launch_wgrad dispatches to it
bool twgrad_shape_ok(int M, int N, int T);
int twgrad_splits(int M, int N, int T);   // token-row splits (= partial buffers) the tcgen05 kernel wants
bool twgrad_supported(const WgradArgs& a);
cudaError_t launch_twgrad(const WgradArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ embedding (k_embed.cu)
cudaError_t launch_embed_ln_fwd(const int64_t* ids, const bf16* table, const bf16* pos, const float* gamma,
                                const float* beta, bf16* out, int B, int S, int H, int V, float drop_rate,
                                uint64_t seed, uint32_t step, const long long* d_step, cudaStream_t st);
// d_out fp32 [T][H] -> dx rows (fp32 [T][H], may alias d_out), dpos partials [bsplits][S*H], dgamma/dbeta partials [nparts][2H].
// The item-table gradient is the per-item sum of the dx rows: launch_table_grad (k_tablegrad.cu).
cudaError_t launch_embed_bwd(const int64_t* ids, const bf16* table, const bf16* pos, const float* gamma,
                             const float* d_out, const bf16* d_branch /* optional bf16 addend of d_out */, float* dx_rows, float* dpos_part, float* dln_part, int B, int S,
                             int H, int V, float drop_rate, uint64_t seed, uint32_t step, const long long* d_step,
                             int bsplits, cudaStream_t st);
int embed_bwd_bsplits(int B);

// ------------------------------------------------------------------ peer-memory gradient all-reduce (k_p2p.cu)
// bufs / flags: device arrays of `world` peer-mapped pointers (this rank's own included) to the fp32 buffer and to a zeroed u32 flag
// array of >= 3 * p2p_allreduce_max_world() entries on every rank; state: local zeroed u32[8]
int p2p_allreduce_max_world();
// multicast (optional): NVLS multicast mapping of the same buffer -> one-pass multimem.ld_reduce / multimem.st variant.  A buffer must
// always be reduced by the same variant (the two keep different arrival counts in `state`).
cudaError_t launch_p2p_allreduce(float* const* bufs, uint32_t* const* flags, float* multicast, size_t off, size_t n, int rank, int world,
                                 uint32_t* state, cudaStream_t st);

// ------------------------------------------------------------------ deterministic item-table gradient (k_tablegrad.cu)
// grad_table[id] += sum over the tokens t with ids[t] == id of dx[t], in a FIXED order: the tokens are sorted by (id, t) with a
// stable LSD radix sort (integer work, depends on the batch only: runs on a side branch of the step), the sorted list is cut
// into chunks of 64 (16 for small batches) tokens that are summed run by run, and runs that cross chunk boundaries are finished from per-chunk
// carries in chunk order.  No floating-point atomics: the result is bit-reproducible and popular items do not serialise.
struct TableGradArgs {
  const int64_t* ids;     // [T] token ids (clamped to [0, V) like the gather)
  int T, V, H;
  uint32_t* keys[2];      // [T] each: sort ping-pong (item ids)
  uint32_t* vals[2];      // [T] each: sort ping-pong (token indices)
  uint32_t* hist;         // [256 * table_grad_sort_blocks(T)] digit-major block histograms
  int* long_runs;         // [1 + 4 * table_grad_max_long_runs(T)]: count, then {first token, end token, id, -} per long run
  float* carry;           // [table_grad_chunks(T)][2][H] head / tail partial rows of every chunk
  const float* dx;        // [T][H] fp32 rows to sum
  float* grad_table;      // [V][H] fp32, accumulated into (one read-modify-write per item, by one thread per column)
};
int table_grad_sort_blocks(int T);
int table_grad_chunk_tokens(int T);   // 16 (small batches) or 64
int table_grad_chunks(int T);
int table_grad_max_long_runs(int T);
cudaError_t launch_token_sort(const TableGradArgs& a, cudaStream_t st);    // keys/vals[table_grad_sorted_buf(V)] = sorted (id, t)
int table_grad_sorted_buf(int V);
int table_grad_sort_launches(int T, int V);   // kernels launch_token_sort enqueues (launch_table_grad: always 2)
cudaError_t launch_table_grad(const TableGradArgs& a, cudaStream_t st);    // 2 launches: chunk sums, boundary runs

// ------------------------------------------------------------------ attention (k_attn.cu)
struct AttnArgs {
  const bf16* qkv;        // [B*S][3H]  (q | k | v), head n occupies columns n*D..n*D+D-1 of each part
  const int64_t* mask;    // [B][S] input_mask (1 = valid key)
  bf16* ctx;              // [B*S][H]
  float* lse;             // [B][N][S]
  uint64_t* keep_bits;    // [B][N][S][words] dropout keep bits (training only)
  int B, S, H, N;
  float drop_rate; uint64_t seed; uint32_t site; uint32_t step; const long long* d_step;
  // backward
  const bf16* dctx;       // [B*S][H]
  bf16* dqkv;             // [B*S][3H]
  bool no_tcgen05;        // force the mma.sync generation (session flag 6 = 0; parity partner of k_fattn.cu)
};
cudaError_t launch_attn_fwd(const AttnArgs& a, cudaStream_t st);
cudaError_t launch_attn_bwd(const AttnArgs& a, cudaStream_t st);
int attn_mask_words(int S);
// generation 2 (k_fattn.cu): tcgen05 + TMEM + TMA, persistent warp-specialised CTAs, seq_len <= 256, head dim 32 / 64;
// launch_attn_fwd / launch_attn_bwd dispatch to it unless no_tcgen05
bool fattn_supported(const AttnArgs& a);
cudaError_t launch_fattn_fwd(const AttnArgs& a, cudaStream_t st);
cudaError_t launch_fattn_bwd(const AttnArgs& a, cudaStream_t st);
// first tcgen05 forward (k_tattn.cu, opt-in B4R_ENABLE_TATTN): one serial chain per CTA, kept for comparison
bool tattn_fwd_supported(const AttnArgs& a);
cudaError_t launch_tattn_fwd(const AttnArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ fused encoder forward (k_enc_fused.cu)
// Whole encoder stack in one launch (tcgen05 + TMEM + TMA), hidden 64 / 2 heads / seq_len <= 128 / inner_dim % 64 == 0.
struct EncFusedLayerHost {
  const float *bqkv, *bo, *g1, *be1, *b1, *b2, *g2, *be2;
  bf16 *qkv, *ctx, *a_pre, *y, *h_pre, *h, *o_pre, *out;
  float *lse, *mean1, *rstd1, *mean2, *rstd2;
  uint64_t* keep;
};
struct EncFusedArgs {
  const int64_t* ids; const int64_t* mask;
  const bf16* table; const bf16* pos; const float* emb_g; const float* emb_b;
  bf16* x0;
  const void* dev_tables;   // device block built by enc_fused_build_tables
  int B, S, V, L, I, training;
  float out_drop, attn_drop; uint64_t seed; uint32_t step; const long long* d_step;
  void* dbg;                // optional device uint64[256]: phase timestamps of CTA 0
};
bool enc_fused_supported(int H, int N, int S, int I);
size_t enc_fused_smem_bytes(int I);
size_t enc_fused_table_bytes(int L);
// w_ptrs: [L][4] bf16 device pointers (wqkv, wo, w1, w2); host: staging buffer of enc_fused_table_bytes(L)
bool enc_fused_build_tables(const EncFusedLayerHost* layers, int L, int I, const bf16* const* w_ptrs, void* host, void* dev_base);
cudaError_t launch_enc_fwd_fused(const EncFusedArgs& a, cudaStream_t st);

// fused encoder backward (k_enc_fused_bwd.cu): all layers in one launch, same tile ownership as the fused forward
struct EncBwdArgs {
  const int64_t* mask; const bf16* x0;
  const void* dev_tables;   // the forward's device block (layer table + weight tensor maps)
  float* dx;                // [T][64] fp32 in/out
  float* wpart; float* bpart;
  // embedding stage backward (fused tail)
  const int64_t* ids; const bf16* table; const bf16* pos; const float* emb_g;
  float* dx_rows; float* dpos_part; float* embln_part; int V;
  int B, S, L, I;
  float out_drop, attn_drop; uint64_t seed; uint32_t step; const long long* d_step;
  void* dbg;
};
bool enc_bwd_fused_supported(int H, int N, int S, int I);
int enc_fused_ctas(int B, int S);            // CTAs (= partial buffers) of the fused kernels
size_t enc_bwd_wpart_floats(int I);          // per (layer, CTA): wqkv | wo | w1 | w2
size_t enc_bwd_bpart_floats(int I);          // per (layer, CTA): the layer's bias / LayerNorm block
cudaError_t launch_enc_bwd_fused(const EncBwdArgs& a, cudaStream_t st);

// ------------------------------------------------------------------ row kernels (k_rows.cu)
// LayerNorm backward over rows (+ dropout of the branch gradient):
//   d_pre = LNbwd(d_out + branch) ; d_branch = drop(d_pre) (bf16) ; partials[cta] = {dgamma[H], dbeta[H], dbranch_colsum[H]}
// d_out is the fp32 gradient of the residual stream, `branch` (optional) the bf16 gradient that the data-gradient GEMM of the
// sublayer above produced; d_pre may alias d_out (the residual-stream gradient is updated in place).
cudaError_t launch_ln_bwd(const float* d_out, const bf16* branch, const bf16* pre, const float* mean, const float* rstd,
                          const float* gamma, float* d_pre, bf16* d_branch, float* partials, int M, int H,
                          float drop_rate, uint64_t seed, uint32_t site, uint32_t step, const long long* d_step,
                          cudaStream_t st);
int ln_bwd_parts(int M);

struct ReduceJob { const float* src; float* dst; int nparts; int len; long long part_stride; int accumulate; };
cudaError_t launch_grad_reduce(const ReduceJob* d_jobs, int njobs, int max_blocks, cudaStream_t st);
int grad_reduce_blocks(int nparts, int len);  // blocks (grid.x) one job needs; pass the max over the jobs of a launch

// squared L2 norm partials of g[0..n) -> out_part[nblocks] ; deterministic two-stage
cudaError_t launch_sqnorm(const float* g, long long n, float* out_part, int nblocks, cudaStream_t st);
struct AdamWArgs {
  float* p; bf16* shadow; const float* g; float* m; float* v;
  long long n_decay, n;        // [0,n_decay) decayed, [n_decay,n) not decayed
  const float* sq_part; int n_sq_part;   // squared-norm partials of g (before scaling)
  const float* d_count;        // optional device scalar: gradient is divided by max(*d_count,1) (loss normaliser)
  float grad_scale;            // extra host-side scale (e.g. 1/world_size)
  const long long* d_step;     // device step counter (optimizer.iterations) ; incremented by the kernel's last block
  long long* d_step_out;
  float init_lr, end_lr; long long num_train_steps, num_warmup_steps;
  float wd, beta1, beta2, eps, clip;
  float* d_lr_out;             // optional: lr used, global norm (2 floats)
};
cudaError_t launch_adamw(const AdamWArgs& a, cudaStream_t st);
cudaError_t launch_colsum_bf16(const bf16* src, int ld, int M, int N, const int* d_M, int d_M_off, float* part,
                               int splits, cudaStream_t st);
cudaError_t launch_cast_bf16(const float* src, bf16* dst, long long n, cudaStream_t st);

// ------------------------------------------------------------------ MLM head + CE (k_ce.cu)
// Compact the valid masked slots: rows[m] = b*S + pos, labels[m]; then one aux row per sequence that has padded
// slots (label 0, weight 0, mult = #padded slots) for SparseCategoricalAccuracy parity.
// counts = {n_valid, n_rows_total}
cudaError_t launch_mlm_select(const int64_t* positions, const int64_t* ids, const int64_t* weights, int use_weights,
                              int B, int S, int P, int want_aux, int* rows, int* labels, float* row_w, int* row_mult,
                              int* counts, cudaStream_t st);
struct CeArgs {
  const bf16* t; int ldt;            // [M][H] transformed hidden rows
  const bf16* E;                     // [V][H] tied table (bf16 shadow)
  const float* vbias;                // [V]
  const int* labels; const float* row_w; const int* row_mult;
  const int* d_counts;               // {n_valid, n_rows}
  int M_cap, H, V;
  int v_begin, v_end;                // vocabulary shard handled here (whole vocab: 0, V)
  int vsplits;
  int target_ctas, max_splits;       // generation 2: device-side split choice (vsplits < 0 in finalize = dynamic)
  int batch;                         // batch size (weight of the per-batch loss in the running mean)
  float* part;                       // [vsplits][M_cap][6] partial (max, sum, label_logit, best_val, best_idx, unused)
  float* lse;                        // [M_cap]
  float* lab_out;                    // optional [M_cap]: logit of the label column (ground-truth score)
  float* fin_part; int* ticket;      // finalize scratch: [64][5] floats + one zero-initialised int
  float* stats;                      // optional float[16] running accumulators: {loss_sum, n_valid, n_correct_masked,
                                     // n_correct_all, n_all, sum(batch_loss*batch), sum(batch), sum(batch_masked_acc), n_steps}
  float* step_stats;                 // this step only (same layout)
  // backward
  bf16* dlogits; int ld_dl;          // [rows_chunk][Vp]
  int row_begin, row_count;          // chunk of rows for dlogits
};
cudaError_t launch_ce_fwd(const CeArgs& a, cudaStream_t st);
constexpr int kCeFinalizeMaxBlocks = 4 * 148;   // fin_part holds 5 floats per CTA
cudaError_t launch_ce_finalize(const CeArgs& a, cudaStream_t st);
cudaError_t launch_ce_dlogits(const CeArgs& a, cudaStream_t st);
// full-catalogue rank counting (see k_ce.cu): beat[m] += #items of the shard that rank ahead of the ground truth
cudaError_t launch_ce_count(const CeArgs& a, const float* s_gt, int* beat, cudaStream_t st);
int ce_block_m();
// full-catalogue top-k (k_ce.cu): the K best (score, id) of every row over [a.v_begin, a.v_end) as order-preserving u64 keys
// (larger = ranks earlier: higher score, then lower id), sorted best first; K <= 128.  d_counts[0] (optional) bounds the live rows.
int topk_splits(int n_rows, int v_len, int K);
size_t topk_scratch_bytes(int n_rows, int v_len, int K);
cudaError_t launch_topk_full(const CeArgs& a, int n_rows, int K, unsigned long long* scratch, unsigned long long* keys_out,
                             long long* ids_out, float* scores_out, cudaStream_t st);
// merge nlists sorted or unsorted key lists [nlists][n_rows][K] -> the K best per row (also the cross-rank merge of a sharded catalogue)
cudaError_t launch_topk_merge(const unsigned long long* keys_in, int nlists, int n_rows, int K, const int* d_counts,
                              unsigned long long* keys_out, long long* ids_out, float* scores_out, cudaStream_t st);
// generation 2 (tcgen05 + TMA) forward; maps are created once per session on the host
struct CeUmmaMaps {   // a: t rows, b: E rows (128-row boxes); a64 / b64: 64-row boxes (streamed operand of the hidden-256 backward)
  alignas(64) unsigned char a[128]; alignas(64) unsigned char b[128]; alignas(64) unsigned char a64[128]; alignas(64) unsigned char b64[128];
};
bool ce_umma_make_maps(CeUmmaMaps* maps, const bf16* t, int M_cap, const bf16* E, int V, int H);
cudaError_t launch_ce_fwd_umma(const CeUmmaMaps& maps, const CeArgs& a, cudaStream_t st);
int ce_umma_block_m();
// generation 2 backward (k_ce_bwd_umma.cu): row_is_m = true -> dT partials, false -> dE partials + d(output bias)
struct CeBwdArgs {
  const float* vbias; const float* lse; const float* row_w; const int* labels; const int* d_counts;
  int M_cap, V, H;
  int target_ctas, max_splits;   // dT pass: device-side vocabulary split choice
  int msplits;                   // dE pass: static row splits
  float* out; float* dbias_out;
};
bool ce_bwd_umma_supported(int H);
int ce_bwd_umma_xtile(int H);   // rows of a streamed tile (128, or 64 for hidden 256): the unit of the vocabulary split count
cudaError_t launch_ce_bwd_umma(const CeUmmaMaps& maps, const CeBwdArgs& a, bool row_is_m, cudaStream_t st);
// generation 3 (k_ce_bwd_fused.cu, hidden 64): dT and dE from ONE recompute pass; partial slots: dT per vocabulary range
// (ce_bwd_fused_vranges(V), static), dE / bias per row chunk (<= ce_bwd_fused_max_chunks(M_cap), count chosen on the device)
struct CeBwdFusedArgs {
  const float* vbias; const float* lse; const float* row_w; const int* labels; const int* d_counts;
  int M_cap, V, ctas;
  float* dt_part; float* dE_part; float* db_part;
  unsigned long long* dbg;   // optional timestamps (development aid)
};
bool ce_bwd_fused_supported(int H);
int ce_bwd_fused_max_chunks(int M_cap);
// decomposition shared by the kernel and its consumers: row chunks of <= 5 tiles, vocabulary ranges of 2 tiles; the `ctas`
// resident CTAs are split into nch (row chunks) x b (groups of consecutive vocabulary ranges); one dT partial slot per group
constexpr int CF_MR_MAX = 5, CF_VR = 2;
struct CfSplit { int mt, vt, nvr, nch, MR, b, per; };
__host__ __device__ inline CfSplit cf_split(int n_valid, int V, int ctas) {
  CfSplit s;
  s.mt = (n_valid + 127) / 128; s.vt = (V + 127) / 128;
  s.nvr = (s.vt + CF_VR - 1) / CF_VR;
  s.nch = (s.mt + CF_MR_MAX - 1) / CF_MR_MAX;
  s.MR = s.nch ? (s.mt + s.nch - 1) / s.nch : 0;   // balanced chunks, every chunk non-empty
  int b = s.nch ? ctas / s.nch : 0;
  if (b < 1) b = 1;
  if (b > s.nvr) b = s.nvr;
  s.per = b ? (s.nvr + b - 1) / b : 0;
  s.b = s.per ? (s.nvr + s.per - 1) / s.per : 0;   // every group non-empty
  if (s.nch == 0) s.b = 0;
  return s;
}
inline int ce_bwd_fused_dt_slot_cap(int V, int ctas) { const int nvr = ((V + 127) / 128 + CF_VR - 1) / CF_VR; return nvr < ctas ? nvr : ctas; }
cudaError_t launch_ce_bwd_fused(const CeUmmaMaps& maps, const CeBwdFusedArgs& a, cudaStream_t st);
cudaError_t launch_ce_bwd_fused_reduce(const CeBwdFusedArgs& a, float* g_table, float* g_bias, cudaStream_t st);
// Split count the generation-2 CE passes choose on the device from the number of live rows (row tiles of 128).  The grid is
// (row tiles x splits) CTAs of equal work with `target_ctas` resident at a time.
//  * mt <= target: floor(target / mt) splits = one wave (a launch grid of target + mtiles_cap covers it; CTAs without work exit at
//    once, but every launched CTA of these kernels costs ~75 ns of scheduling, so the grid must stay close to the work).
//  * mt > target (C4 / C5: more row tiles than resident CTAs): one split would be ceil(mt / target) FULL passes over the vocabulary
//    (2.0 at C4's 320 tiles on 296 CTAs); with vs splits the pass takes ceil(mt * vs / target) / vs -- the vs <= 8 minimising that
//    is taken (C4 forward: 7 splits = 8 waves of 1/7 = 1.14).
// Every producer and consumer of the split partials calls this one function.
__host__ __device__ inline int ce_dyn_splits128(int n_rows, int ntiles, int target_ctas, int max_splits) {
  int mt = (n_rows + 127) / 128;
  if (mt < 1) mt = 1;
  int lim = max_splits < ntiles ? max_splits : ntiles;
  if (lim < 1) lim = 1;
  if (mt <= target_ctas) {
    int vs = target_ctas / mt;
    if (vs > lim) vs = lim;
    return vs < 1 ? 1 : vs;
  }
  if (lim > 8) lim = 8;
  int best = 1;
  float best_t = 3.0e38f;
  for (int vs = 1; vs <= lim; ++vs) {
    const int waves = (mt * vs + target_ctas - 1) / target_ctas;
    const float t = (float)waves / (float)vs + 0.002f * (float)vs;
    if (t < best_t) { best_t = t; best = vs; }
  }
  return best;
}
// launch grid that covers every live-row count up to mtiles_cap row tiles
inline int ce_dyn_grid(int target_ctas, int mtiles_cap, int ntiles, int max_splits) {
  int g = target_ctas + mtiles_cap;
  for (int mt = target_ctas + 1; mt <= mtiles_cap; ++mt) {
    const int need = mt * ce_dyn_splits128(mt * 128, ntiles, target_ctas, max_splits);
    if (need > g) g = need;
  }
  return g;
}

// ------------------------------------------------------------------ vocabulary-sharded projection, row-side helpers (k_shard.cu)
cudaError_t launch_shard_pack(const bf16* rows_in, const int* labels_in, const float* w_in, const int* mult_in, const int* counts_in,
                              int n, int M_cap, int H, int v_begin, bf16* rows, int* lab_local, int* lab_global, float* w, int* mult,
                              int* counts, cudaStream_t st);
cudaError_t launch_shard_part_merge(const float* part, const int* counts, int M_cap, int ntiles, int target_ctas, int max_splits,
                                    int v_begin, float* out, cudaStream_t st);
cudaError_t launch_shard_dt_unpack(const float* dt_part, const int* counts, const int* counts_in, int n, int M_cap, int cap, int H,
                                   int xtiles, int target_ctas, int max_splits, float* out, cudaStream_t st);
// MLM transform backward over rows: dt = sum_s dt_part[s] ; LN bwd ; gelu bwd -> d_tpre (bf16) ; partials {dgamma,dbeta,dbias}
cudaError_t launch_head_bwd_rows(const float* dt_part, int nsplit, size_t split_stride, const bf16* t_pre,
                                 const bf16* act, const float* mean, const float* rstd, const float* gamma,
                                 bf16* d_tpre, float* partials, int M_cap, const int* d_counts, int H,
                                 cudaStream_t st, int dyn_vtiles = 0, int dyn_target = 0, int dyn_max = 0);

// MLM transform backward in one launch (k_head.cu, hidden 64): dt partial sum + LN/GELU backward + dx scatter + dWt partials
bool head_bwd_fused_supported(int H);
int head_bwd_fused_ctas();   // number of per-CTA partials: p_ln [ctas][3H] = {dgamma, dbeta, dbias}, p_wt [ctas][H*H]
cudaError_t launch_head_bwd_fused(const float* dt_part, int nsplit, size_t split_stride, const bf16* t_pre, const bf16* t_act,
                                  const float* mean, const float* rstd, const float* gamma, const bf16* wt, const bf16* x,
                                  const int* rows, const int* d_counts, int M_cap, float* dx_out, float* p_ln, float* p_wt,
                                  cudaStream_t st, int dyn_vtiles = 0, int dyn_target = 0, int dyn_max = 0);

// ------------------------------------------------------------------ ranking (k_rank.cu)
// scores[m][c] = t[m] . E[cand[m][c]] + vbias[cand[m][c]] ; ranking = stable descending order (lower index first on ties)
// rank[m] = 1 + position of the first candidate equal to gt[m] ; hist[r] += 1.  Candidate ids outside [0, V) score -inf (ranked
// last); rows at or beyond d_counts[0] (the number of selected slots, device side; optional) get rank 0 and are not scored.
cudaError_t launch_rank_candidates(const bf16* t, int ldt, const bf16* E, const float* vbias, const int64_t* cand,
                                   const int64_t* gt, int M, int C, int H, int V, const int* d_counts, int64_t* ranking,
                                   float* scores, int* rank, unsigned long long* hist, cudaStream_t st);
// metric sums from a rank histogram: out = {n, ndcg@k.., hr@k.., map} in fp64
cudaError_t launch_metrics_from_hist(const unsigned long long* hist, int max_rank, const int* ks, int nk,
                                     double* out, cudaStream_t st);
// keep-mask dump for tests: out[r][c] = keep ? 1 : 0 for the elementwise dropout sites
cudaError_t launch_dropout_mask_dump(uint8_t* out, int rows, int cols, float rate, uint64_t seed, uint32_t site,
                                     uint32_t step, cudaStream_t st);
}  // namespace b4r
