"""ctypes binding of libb4r.so (the C ABI declared in include/b4r.h).

There is NO CPU fallback: if the shared library is missing or the device is not an sm_100 GPU the product path
raises.  ``load()`` is lazy so that host-only pieces (samplers, tokenizer, metrics) import without a GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb4r.so")


class B4RError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden_size", C.c_int32), ("num_layers", C.c_int32),
                ("num_heads", C.c_int32), ("max_seq_len", C.c_int32), ("inner_dim", C.c_int32),
                ("output_dropout", C.c_float), ("attention_dropout", C.c_float)]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("offset", C.c_int64), ("numel", C.c_int64),
                ("rows", C.c_int32), ("cols", C.c_int32), ("group", C.c_int32)]


class AdamWHParams(C.Structure):
    _fields_ = [("init_lr", C.c_float), ("end_lr", C.c_float), ("num_train_steps", C.c_int64),
                ("num_warmup_steps", C.c_int64), ("weight_decay_rate", C.c_float), ("beta_1", C.c_float),
                ("beta_2", C.c_float), ("epsilon", C.c_float), ("clip_norm", C.c_float)]


class DLView(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device_type", C.c_int32), ("device_id", C.c_int32), ("ndim", C.c_int32),
                ("dtype_code", C.c_int32), ("dtype_bits", C.c_int32), ("shape", C.c_int64 * 4)]


_P = C.c_void_p
_SIGNATURES = {
    "b4r_version": (C.c_int, []),
    "b4r_last_error": (C.c_char_p, []),
    "b4r_device_check": (C.c_int, [C.c_int]),
    "b4r_param_entries": (C.c_int, [C.POINTER(Config), C.POINTER(ParamEntry), C.c_int]),
    "b4r_param_counts": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b4r_session_workspace_bytes": (C.c_size_t, [C.POINTER(Config), C.c_int, C.c_int, C.c_int]),
    "b4r_session_create": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_size_t, C.POINTER(_P)]),
    "b4r_session_destroy": (None, [_P]),
    "b4r_sync_shadow": (C.c_int, [_P, _P]),
    "b4r_encode": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint64, C.c_uint32, _P, _P]),
    "b4r_mlm_select": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "b4r_mlm_transform": (C.c_int, [_P, _P]),
    "b4r_mlm_loss": (C.c_int, [_P, _P, _P]),
    "b4r_mlm_logits": (C.c_int, [_P, _P, _P]),
    "b4r_backward": (C.c_int, [_P, C.c_uint64, C.c_uint32, _P, _P]),
    "b4r_pooled_output": (C.c_int, [_P, _P, _P]),
    "b4r_p2p_allreduce_max_world": (C.c_int, []),
    "b4r_p2p_allreduce_f32": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_size_t, C.c_int, C.c_int, _P, _P]),
    "b4r_adamw_scratch_floats": (C.c_size_t, []),
    "b4r_adamw_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.POINTER(AdamWHParams), _P, C.c_float, _P, _P, _P, _P]),
    "b4r_rank_candidates": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "b4r_rank_full": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "b4r_rank_full_ext": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "b4r_backward_from_dt": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "b4r_mlm_labels": (_P, [_P]),
    "b4r_mlm_row_weights": (_P, [_P]),
    "b4r_mlm_row_mult": (_P, [_P]),
    "b4r_shard_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "b4r_shard_create": (C.c_int, [C.c_int] * 6 + [_P, _P, _P, _P, _P, C.c_size_t, C.POINTER(_P)]),
    "b4r_shard_destroy": (None, [_P]),
    "b4r_shard_pack": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "b4r_shard_ce_partial": (C.c_int, [_P, _P, _P]),
    "b4r_shard_ce_merge": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "b4r_shard_ce_backward": (C.c_int, [_P, _P, C.c_int, _P]),
    "b4r_shard_step_stats": (_P, [_P]),
    "b4r_shard_lse": (_P, [_P]),
    "b4r_shard_counts": (_P, [_P]),
    "b4r_host_cloze_mask_batch": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, C.c_int, C.c_int64,
                                            C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P, _P, C.c_int]),
    "b4r_host_sample_random_batch": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
    "b4r_host_sample_popular_batch": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int, C.c_int, _P, _P, C.c_int]),
    "b4r_host_sample_pop_random_batch": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int]),
    "b4r_packed_inputs_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "b4r_unpack_inputs": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "b4r_h2d_copy_many": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "b4r_metrics_from_hist": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "b4r_sequence_output": (_P, [_P, C.c_int]),
    "b4r_mlm_hidden": (_P, [_P]),
    "b4r_mlm_counts": (_P, [_P]),
    "b4r_mlm_rows": (_P, [_P]),
    "b4r_step_stats": (_P, [_P]),
    "b4r_attn_keep_bits": (_P, [_P, C.c_int, C.POINTER(C.c_int)]),
    "b4r_layer_tensor": (_P, [_P, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b4r_launch_count": (C.c_int, [_P]),
    "b4r_session_set_flag": (C.c_int, [_P, C.c_int, C.c_int]),
    "b4r_debug_buffer": (_P, [_P]),
    "b4r_debug_buffer2": (_P, [_P]),
    "b4r_profile_enable": (C.c_int, [_P, C.c_int]),
    "b4r_profile_report": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "b4r_dropout_keep_mask": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_int, C.c_int, C.c_uint32, _P]),
    "b4r_embed_ln_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "b4r_table_grad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "b4r_table_grad": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "b4r_dl_view_of": (C.c_int, [_P, C.POINTER(DLView)]),
    "b4r_topk_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "b4r_topk_full": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "b4r_topk_merge": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Loads libb4r.so (once).  Raises B4RError if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B4RError(f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                       "bert4rec_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise B4RError(load().b4r_last_error().decode())
