/* libb4r.so -- development and test helpers (NOT part of the drop-in boundary; see include/b4r.h for the product ABI).
 * Per-kernel profiling of a session, debug buffers of the fused kernels, the keep-mask dump the oracle tests replay, and a
 * standalone embedding kernel for unit parity tests. */
#ifndef B4R_DEBUG_H_
#define B4R_DEBUG_H_
#include "b4r.h"
#ifdef __cplusplus
extern "C" {
#endif

const void* b4r_debug_buffer(b4r_session* s);
const void* b4r_debug_buffer2(b4r_session* s); /* uint64[512]: fused-kernel phase timestamps (ns) when B4R_FUSED_DEBUG is set */  /* kernel-internal timestamps when B4R_CE_DEBUG&8 (development aid) */
/* per-kernel CUDA-event timing of everything launched through the session (events on the launching stream);
 * report: lines "tag count total_ms", synchronises, clears the records. */
int b4r_profile_enable(b4r_session* s, int on);
int b4r_profile_report(b4r_session* s, char* buf, int cap);

/* ---- test helpers ------------------------------------------------------------------------------------------ */
/* keep mask (1 byte / element) of an elementwise dropout site: site 1 = embedding, 2 = attention output, 3 = FFN output */
int b4r_dropout_keep_mask(uint8_t* out, int rows, int cols, float rate, uint64_t seed, int site, int layer, uint32_t step,
                          void* stream);
/* standalone kernels for unit parity tests */
int b4r_embed_ln_fwd(const int64_t* ids, const void* table_bf16, const void* pos_bf16, const float* gamma,
                     const float* beta, void* out_bf16, int batch, int seq_len, int hidden, int vocab, void* stream);
/* deterministic item-table gradient of the embedding gather, standalone: grad_table[ids[t]] += dx[t] (fp32 rows of `hidden`
 * columns, hidden in {64, 128, 256}) in a fixed summation order; workspace of b4r_table_grad_workspace_bytes() device bytes */
size_t b4r_table_grad_workspace_bytes(int tokens, int hidden);
int b4r_table_grad(const int64_t* ids, const float* dx, float* grad_table, int tokens, int vocab, int hidden, void* workspace,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B4R_DEBUG_H_ */
