/* libb4r -- C ABI of the B200-native BERT4Rec train + ranking path.
 *
 * The reference (maneymarkus/BERT4Rec) is pure Python over TensorFlow and has no FFI seam of its own; the boundary
 * below is the set of entry points its Python classes would bind for this path.  Each entry cites the reference
 * interface it replaces.  Conventions:
 *   - every function returns 0 on success, non-zero on error; b4r_last_error() (thread-local) describes the failure;
 *   - all pointers are plain device pointers (cudaMalloc'ed memory, e.g. a torch tensor's data_ptr) unless the name
 *     says host; no torch / DLPack types cross the boundary except in b4r_dl_view (the optional DLPack adapter);
 *   - every compute call takes a cudaStream_t (as void*) and only ENQUEUES work: no allocation, no hidden sync;
 *   - integer model inputs are int64, row-major, exactly the reference's batch dict (bert4rec_model.py:15-22).
 */
#ifndef B4R_H_
#define B4R_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B4R_VERSION 100

/* Bert4RecEncoder(vocab_size, hidden_size, num_layers, num_attention_heads, max_sequence_length, inner_dim,
 * output_dropout, attention_dropout) -- bert4rec_encoder.py:62-80 (post-LN, exact-erf GELU, LN eps 1e-12). */
typedef struct b4r_config {
  int32_t vocab_size, hidden_size, num_layers, num_heads, max_seq_len, inner_dim;
  float output_dropout, attention_dropout;
} b4r_config;

typedef struct b4r_param_entry {
  char name[48];      /* internal segment name, e.g. "word_embeddings", "layer_0/wqkv" */
  int64_t offset;     /* element offset inside the flat parameter buffer */
  int64_t numel;
  int32_t rows, cols; /* 2-D view (cols == 0 for vectors) */
  int32_t group;      /* 0 = weight-decayed, 1 = not decayed (bias / LayerNorm), 2 = frozen (pooler: no gradient) */
} b4r_param_entry;

typedef struct b4r_adamw_hparams {  /* optimizers.create_adam_w_optimizer defaults, optimizers/__init__.py:7-15 */
  float init_lr, end_lr;
  int64_t num_train_steps, num_warmup_steps;
  float weight_decay_rate, beta_1, beta_2, epsilon, clip_norm;
} b4r_adamw_hparams;

typedef struct b4r_session b4r_session;

/* ---- library / device ---------------------------------------------------------------------------------- */
int b4r_version(void);
const char* b4r_last_error(void);
int b4r_device_check(int device);   /* fails unless the device is sm_100 (B200) */

/* ---- flat parameter layout (replaces keras variable creation, bert4rec_encoder.py:102-153, bert4rec_model.py:76-81) */
int b4r_param_entries(const b4r_config* cfg, b4r_param_entry* out, int cap);          /* returns #entries (or <0) */
int b4r_param_counts(const b4r_config* cfg, int64_t* n_decay, int64_t* n_trainable, int64_t* n_total);

/* ---- session: static shapes (batch, seq_len, max_predictions_per_seq) + caller-owned buffers ------------ */
size_t b4r_session_workspace_bytes(const b4r_config* cfg, int batch, int seq_len, int max_pred);
int b4r_session_create(const b4r_config* cfg, int batch, int seq_len, int max_pred, float* params, void* shadow_bf16,
                       float* grads, void* workspace, size_t workspace_bytes, b4r_session** out);
void b4r_session_destroy(b4r_session* s);
int b4r_sync_shadow(b4r_session* s, void* stream);   /* shadow_bf16 = bf16(params) after a host-side weight load */

/* Bert4RecEncoder.call, bert4rec_encoder.py:186-231.  ids/mask: int64 [batch, seq_len].  With training != 0 on a session that has a
 * gradient buffer the call also enqueues the token sort of the embedding gradient on an internal stream branch (forked from `stream`
 * here, joined back inside b4r_backward): a stream capture must therefore contain the matching b4r_backward. */
int b4r_encode(b4r_session* s, const int64_t* input_word_ids, const int64_t* input_mask, int training, uint64_t seed,
               uint32_t step, const int64_t* step_counter, void* stream);
/* Dropout masks are Philox(seed; row, col, site, step + *step_counter): step_counter (optional device int64, e.g. the
 * optimizer's iteration counter) lets a captured CUDA graph draw fresh masks on every replay. */
/* tfm MaskedLM._gather_indexes (bert4rec_model.py:141-143) + tf.boolean_mask of the loss (trainer_utils.py:19-20).
 * mode 0: slots with masked_lm_ids != 0 (training); 1: slots with masked_lm_weights != 0 (rank_items,
 * bert4rec_model.py:218-220); 2: all batch*max_pred slots (BERT4RecModel.call). want_aux adds one row per sequence
 * with padded slots so that SparseCategoricalAccuracy over ALL slots stays exact. */
int b4r_mlm_select(b4r_session* s, const int64_t* masked_lm_positions, const int64_t* masked_lm_ids,
                   const int64_t* masked_lm_weights, int mode, int want_aux, void* stream);
/* MaskedLM transform (dense+gelu+LayerNorm) on the selected rows. */
int b4r_mlm_transform(b4r_session* s, void* stream);
/* tied projection x online-softmax CE + accuracies (trainer_utils.py:12-23,49-60).  stats: optional device float[16]
 * of running accumulators {loss_sum, n_valid, correct_masked, correct_all, n_all, sum(batch_loss*batch), sum(batch),
 * sum(batch_masked_accuracy), n_steps} (the Keras Mean metrics of train_step, bert4rec_model.py:171-173). */
int b4r_mlm_loss(b4r_session* s, float* stats, void* stream);
/* materialised logits [n_rows, vocab] fp32 for BERT4RecModel.call()'s "mlm_logits" (bert4rec_model.py:139-147) */
int b4r_mlm_logits(b4r_session* s, float* out, void* stream);
/* tape.gradient of the SUM loss over all trainable variables (bert4rec_model.py:166-167) -> grads (flat, fp32). */
int b4r_backward(b4r_session* s, uint64_t seed, uint32_t step, const int64_t* step_counter, void* stream);
/* The same backward when the gradient of the transformed rows comes from OUTSIDE (vocabulary-sharded projection below):
 * dt fp32 [batch*max_pred + batch, hidden] in this session's row order (rows behind n_valid are ignored).  The cross-entropy backward is skipped; the tied-table and
 * output-bias gradients already in `grads` (written by b4r_shard_ce_backward) are kept and the embedding backward adds to them. */
int b4r_backward_from_dt(b4r_session* s, const float* dt, uint64_t seed, uint32_t step, const int64_t* step_counter, void* stream);
/* tanh pooler on token 0 (bert4rec_encoder.py:224-226) -> out fp32 [batch, hidden] */
int b4r_pooled_output(b4r_session* s, float* out, void* stream);

/* Batch data-parallel training (one process per GPU; reference: the loss of a step is normalised by the number of valid masked
 * slots, trainer_utils.py:22 -- summed over ranks here): SUM-all-reduce of an fp32 range [offset, offset + n) of a buffer that
 * lives in SYMMETRIC memory, in place, by this library's own two-shot kernel over NVLink peer mappings (rank r adds slice r of all
 * ranks in rank order, then everyone gathers; deterministic, no NCCL launch, CUDA-graph capturable).  buffer_ptrs_dev /
 * flag_ptrs_dev: DEVICE arrays of `world` pointers (this rank's own included) to the buffer and to a zero-initialised uint32 flag
 * array of >= 3 * b4r_p2p_allreduce_max_world() entries on every rank, each valid for peer access from this device (what
 * torch.distributed._symmetric_memory.rendezvous(...).buffer_ptrs_dev holds); state: this rank's own zero-initialised device
 * uint32[8] (call counter; state[7] != 0 after a call = a peer never arrived within the bounded wait).  multicast_ptr (optional, may
 * be NULL): the NVSwitch multicast mapping of the same buffer (symmetric-memory handle's multicast_ptr); with it the reduction is one
 * pass of multimem.ld_reduce / multimem.st -- the switch adds the ranks' values -- and two barriers.  A buffer must always be reduced
 * with the same choice.  Every rank must enqueue the same sequence of calls. */
int b4r_p2p_allreduce_max_world(void);
int b4r_p2p_allreduce_f32(const void* buffer_ptrs_dev, const void* flag_ptrs_dev, void* multicast_ptr, size_t offset_floats,
                          size_t n_floats, int rank, int world, void* state, void* stream);

/* AdamWeightDecay.apply_gradients (adam_w_optimizer.py:100-136): clip_by_global_norm, WarmUp/PolynomialDecay lr,
 * decoupled decay, Adam; also refreshes shadow_bf16.  The gradient is first multiplied by grad_scale / max(*count,1)
 * (count = number of valid masked slots, the loss normaliser of trainer_utils.py:22).  step_counter: device int64
 * (optimizer.iterations), incremented by the call.  lr_out (optional): device float[2] = {lr, global_norm}. */
size_t b4r_adamw_scratch_floats(void);
int b4r_adamw_step(float* params, void* shadow_bf16, const float* grads, float* m, float* v, int64_t n_decay,
                   int64_t n_trainable, const b4r_adamw_hparams* hp, const float* count, float grad_scale,
                   int64_t* step_counter, float* scratch, float* lr_out, void* stream);

/* BERT4RecModel.rank_items with per-slot candidate lists (bert4rec_model.py:224-234) + rank lookup
 * (bert4rec_evaluator.py:112-117).  cand: int64 [n_slots, C]; gt: int64 [n_slots] or NULL.
 * ranking_out int64 [n_slots, C] (optional), scores_out fp32 [n_slots, C] (optional), rank_out int32 [n_slots]
 * (1-based, 0 if gt absent), hist: uint64 [C+1] rank histogram, accumulated (optional).
 * Candidate ids outside [0, vocab_size) are scored -inf (ranked last, never dereferenced); rows at or beyond the number of
 * slots the last b4r_mlm_select selected get rank 0 and are not scored. */
int b4r_rank_candidates(b4r_session* s, const int64_t* cand, const int64_t* gt, int n_slots, int C,
                        int64_t* ranking_out, float* scores_out, int32_t* rank_out, uint64_t* hist, void* stream);
/* rank_items(items=None) for serving (bert4rec_model.py:235-236: tf.argsort(DESCENDING) over the vocabulary, lower id first among
 * equal logits; apps/recommender.py:14-63): the K (<= 128) best items of every row over the vocabulary shard [v_begin, v_end), no
 * logits materialised.  t_rows: bf16 [n_rows, hidden] external hidden rows (e.g. all-gathered from other ranks) or NULL = the rows of
 * the last b4r_mlm_select + b4r_mlm_transform.  keys_out uint64 [n_rows, K] (optional): (score, id) packed so that a LARGER key ranks
 * EARLIER (order-preserving score bits << 32 | 0xFFFFFFFF - id; 0 = empty), best first -- lists of different shards / ranks are
 * merged with b4r_topk_merge after an all-gather.  ids_out int64 [n_rows, K] / scores_out fp32 [n_rows, K] (optional; -1 / -inf = empty).
 * scratch: b4r_topk_scratch_bytes(n_rows, v_begin, v_end, K) bytes of device memory. */
size_t b4r_topk_scratch_bytes(int n_rows, int v_begin, int v_end, int K);
int b4r_topk_full(b4r_session* s, const void* t_rows, int n_rows, int v_begin, int v_end, int K, void* scratch,
                  uint64_t* keys_out, int64_t* ids_out, float* scores_out, void* stream);
/* keys_in uint64 [nlists, n_rows, K] -> the K best keys per row, best first (nlists * K <= 8192) */
int b4r_topk_merge(const uint64_t* keys_in, int nlists, int n_rows, int K, uint64_t* keys_out, int64_t* ids_out,
                   float* scores_out, void* stream);
/* rank_items(items=None) for evaluation: 1-based rank of the label of every selected row over the vocabulary shard
 * [v_begin, v_end): beat_out[i] += #items ranking ahead (caller zeroes; sum over shards + 1 = rank). */
int b4r_rank_full(b4r_session* s, int v_begin, int v_end, int32_t* beat_out, void* stream);
/* The same count for EXTERNAL rows: t_rows bf16 [rows_cap, hidden], labels int32 [rows_cap], gt_scores fp32 [rows_cap],
 * counts2 device int32[2] (counts2[1] = number of rows).  Vocab-sharded evaluation over several GPUs all-gathers the rows
 * of every rank, counts per shard here and all-reduces the counts (rank = 1 + sum over shards). */
int b4r_rank_full_ext(b4r_session* s, const void* t_rows, const int32_t* labels, const float* gt_scores, const int32_t* counts2,
                      int rows_cap, int v_begin, int v_end, int32_t* beat_out, void* stream);
/* ---- vocabulary-sharded tied projection for TRAINING (large catalogues; the reference builds the whole [B,P,V] logits on
 * one device, bert4rec_model.py:139-147 + trainer_utils.py:12-23).  One b4r_shard per rank owns catalogue rows
 * [v_begin, v_end) and works on the masked-slot rows of ALL n_ranks ranks (rows_per_rank = the session's row capacity,
 * batch*max_pred + batch, identical on every rank):
 *   all-gather {b4r_mlm_hidden, b4r_mlm_labels, b4r_mlm_row_weights, b4r_mlm_row_mult, b4r_mlm_counts}  -> b4r_shard_pack
 *   b4r_shard_ce_partial -> all-gather the [rows][6] partials -> b4r_shard_ce_merge (loss / accuracy of the global batch)
 *   b4r_shard_ce_backward -> reduce-scatter(SUM) dt_out -> b4r_backward_from_dt ; all-reduce(SUM) the flat gradients.
 * The collectives are the caller's (torch.distributed / NCCL on the same stream).  table_bf16 / output_bias / grad_table /
 * grad_bias point at the FULL [vocab, hidden] / [vocab] arrays of this rank. */
typedef struct b4r_shard b4r_shard;
size_t b4r_shard_workspace_bytes(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end);
int b4r_shard_create(int hidden, int vocab, int n_ranks, int rows_per_rank, int v_begin, int v_end, const void* table_bf16,
                     const float* output_bias, float* grad_table, float* grad_bias, void* workspace, size_t workspace_bytes,
                     b4r_shard** out);
void b4r_shard_destroy(b4r_shard* s);
/* rows_in bf16 [n_ranks][rows_per_rank][hidden]; labels_in int32, weights_in fp32, mult_in int32 [n_ranks][rows_per_rank];
 * counts_in int32 [n_ranks][2] = each rank's {n_valid, n_rows} (must stay alive until b4r_shard_ce_backward). */
int b4r_shard_pack(b4r_shard* s, const void* rows_in, const int32_t* labels_in, const float* weights_in, const int32_t* mult_in,
                   const int32_t* counts_in, void* stream);
/* part_out fp32 [n_ranks*rows_per_rank][6] = {max, sum exp(x-max), label logit | -inf, best logit, best GLOBAL id (int bits), 0} */
int b4r_shard_ce_partial(b4r_shard* s, float* part_out, void* stream);
/* parts fp32 [n_shards][n_ranks*rows_per_rank][6]; stats: optional running statistics, layout of b4r_mlm_loss */
int b4r_shard_ce_merge(b4r_shard* s, const float* parts, int n_shards, int global_batch, float* stats, void* stream);
/* dt_out fp32 [n_ranks][rows_per_rank][hidden]; zero_all: zero the whole table / bias gradient before writing the shard's slice */
int b4r_shard_ce_backward(b4r_shard* s, float* dt_out, int zero_all, void* stream);
const float* b4r_shard_step_stats(b4r_shard* s);   /* float[8], layout of b4r_step_stats, GLOBAL batch */
const float* b4r_shard_lse(b4r_shard* s);          /* fp32 [n_ranks*rows_per_rank] packed row order */
const int32_t* b4r_shard_counts(b4r_shard* s);     /* int32[2] = {n_valid, n_rows} over all ranks */

/* Compact host->device transfer of a step's inputs (the reference's batch dict holds int64 tensors, bert4rec_model.py:15-22):
 * packed = [int32 input_word_ids n_tok][int32 masked_lm_positions n_pred][int32 masked_lm_ids n_pred][uint8 input_mask n_tok]
 * [uint8 masked_lm_weights n_pred] (b4r_packed_inputs_bytes), copied with ONE H2D copy and widened on the device into the int64
 * buffers the step reads.  n_tok = batch*seq_len, n_pred = batch*max_pred. */
size_t b4r_packed_inputs_bytes(int n_tok, int n_pred);
int b4r_unpack_inputs(const void* packed, int n_tok, int n_pred, int64_t* input_word_ids, int64_t* input_mask,
                      int64_t* masked_lm_positions, int64_t* masked_lm_ids, int64_t* masked_lm_weights, void* stream);
/* n host->device copies enqueued back to back on `stream` (dst[i] device, src[i] host -- pinned for a true asynchronous DMA).
 * Host inputs of a step (the reference feeds host tensors through tf.data, dataloader_utils.py:306-346) reach the persistent
 * device buffers of the captured step without a packing pass on the host. */
int b4r_h2d_copy_many(void* const* dst, const void* const* src, const size_t* nbytes, int n, void* stream);
/* HR@k / NDCG@k / MAP from a rank histogram (evaluation_metrics.py:47-112): out fp64 [2 + 2*nk] =
 * {n, NDCG@k..., HR@k..., MAP}; ks: device int32 [nk]. */
int b4r_metrics_from_hist(const uint64_t* hist, int max_rank, const int32_t* ks, int nk, double* out, void* stream);

/* ---- host data path (no GPU involved, n_threads <= 0: all hardware threads).  Bit-exact with the reference's
 * Python given the same seeds: CPython `random` (MT19937, init_by_array seeding, _randbelow rejection) for the masking, numpy's
 * legacy RandomState (np.random.seed / np.random.choice) for the samplers.  Ragged inputs are CSR: values + offsets [n+1]. */
/* apply_dynamic_masking_task (dataloader_utils.py:186-261) + the padded layout of BERT4RecPreprocessor.process_element
 * (bert4rec_preprocessor.py:94-114) for n already tokenised / windowed sequences (each <= max_seq_len tokens).
 * seeds[i] = the `seed` argument of sequence i.  Outputs int64: labels, input_word_ids, input_mask [n][max_seq_len];
 * masked_lm_ids, masked_lm_positions, masked_lm_weights [n][max_pred]; padding value pad_id everywhere. */
int b4r_host_cloze_mask_batch(const int64_t* tokens, const int64_t* offsets, const uint64_t* seeds, int n, int max_seq_len,
                              int max_pred, int64_t mask_id, int64_t pad_id, const int64_t* special_ids, int n_special,
                              int64_t vocab_size, double selection_rate, double mask_token_rate, double random_token_rate,
                              int64_t* labels, int64_t* input_word_ids, int64_t* input_mask, int64_t* masked_lm_ids,
                              int64_t* masked_lm_positions, int64_t* masked_lm_weights, int n_threads);
/* RandomSampler.sample(seed=seeds[i], without=without_i) (random_sampler.py:63-79): out int64 [n][sample_size] */
int b4r_host_sample_random_batch(const int64_t* vocab, int64_t n_vocab, const int64_t* without, const int64_t* without_off,
                                 const uint32_t* seeds, int n, int sample_size, int allow_duplicates, int64_t* out, int n_threads);
/* PopularSampler.sample(without=without_i) over an already popularity-ranked source (popular_sampler.py:53-71):
 * out int64 [n][sample_size] (zero-filled tail), out_len int32 [n] */
int b4r_host_sample_popular_batch(const int64_t* ranked, int64_t n_ranked, const int64_t* without, const int64_t* without_off, int n,
                                  int sample_size, int64_t* out, int32_t* out_len, int n_threads);
/* PopularRandomSampler.sample(seed=seeds[i], without=without_i) (popular_random_sampler.py:77-117); p = the sampler's
 * probability_distribution, float64 [n_vocab]: out int64 [n][sample_size] (zero-filled tail), out_len int32 [n] */
int b4r_host_sample_pop_random_batch(const int64_t* vocab, const double* p, int64_t n_vocab, const int64_t* without,
                                     const int64_t* without_off, const uint32_t* seeds, int n, int sample_size, int allow_duplicates,
                                     int64_t* out, int32_t* out_len, int n_threads);

/* ---- introspection (device pointers into the workspace; valid for the session lifetime) ------------------ */
const void* b4r_sequence_output(b4r_session* s, int layer);  /* bf16 [batch*seq_len, hidden]; layer -1 = last */
const void* b4r_mlm_hidden(b4r_session* s);                  /* bf16 [n_rows, hidden] transformed rows */
const int32_t* b4r_mlm_counts(b4r_session* s);               /* int32[2] = {n_valid, n_rows} */
const int32_t* b4r_mlm_rows(b4r_session* s);                 /* int32 [n_rows] flat row index b*seq_len+pos */
const int32_t* b4r_mlm_labels(b4r_session* s);               /* int32 [n_rows] label id of the row */
const float* b4r_mlm_row_weights(b4r_session* s);            /* fp32 [n_rows] masked_lm_weights of the row */
const int32_t* b4r_mlm_row_mult(b4r_session* s);             /* int32 [n_rows] slots the row stands for (all-slot accuracy) */
float* b4r_step_stats(b4r_session* s);  /* float[8] {loss_sum, n_valid, correct_masked, correct_all, n_all} of the last step */
const uint64_t* b4r_attn_keep_bits(b4r_session* s, int layer, int* words_per_row);
/* saved activation of encoder layer `layer` by name ("x0","qkv","ctx","a_pre","y","h_pre","h","o_pre","out": bf16
 * [batch*seq_len, *cols]; "mean1","rstd1","mean2","rstd2": fp32 [batch*seq_len]; "lse": fp32 [batch*heads*seq_len]) */
const void* b4r_layer_tensor(b4r_session* s, int layer, const char* name, int* cols, int* is_f32);
int b4r_launch_count(b4r_session* s);   /* kernels launched through this session so far */
/* flag 1: tcgen05/TMA generation of the tied-projection kernels (1, default) or the mma.sync generation (0);
 * flag 2: whole-encoder fused tcgen05 forward (1, default where the shape allows) or the layered kernels (0);
 * flag 3: whole-encoder fused tcgen05 backward (needs flag 2) or the layered backward kernels (0);
 * flag 4: while set, b4r_mlm_select runs as a parallel branch (internal side stream, joined by the next consumer of the
 *         selection) so that it overlaps whatever the caller enqueues next, e.g. b4r_encode (default 0);
 * flag 5: one-pass tcgen05 CE backward (1, default at hidden 64) or the two recompute passes (0);
 * flag 6: tcgen05 attention kernels of the layered encoder path (1, default) or the mma.sync generation (0) */
int b4r_session_set_flag(b4r_session* s, int flag, int value);
/* ---- DLPack adapter: validates a (borrowed) DLManagedTensor (CUDA device, row-major contiguous, scalar dtype) and returns its
 * device pointer; the Python host hands every int64 input of b4r_encode / b4r_mlm_select / b4r_rank_candidates over through it.
 * Development / test helpers (profiling, debug buffers, standalone kernels) are declared in include/b4r_debug.h. */
typedef struct b4r_dl_view { void* data; int32_t device_type, device_id, ndim, dtype_code, dtype_bits; int64_t shape[4]; } b4r_dl_view;
int b4r_dl_view_of(const void* dl_managed_tensor, b4r_dl_view* out);

#ifdef __cplusplus
}
#endif
#endif /* B4R_H_ */
