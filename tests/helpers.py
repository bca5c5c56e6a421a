"""Shared synthetic-input builders for the parity tests (seeded; same on CPU oracle and CUDA path)."""
import numpy as np
import torch

from oracle import host_ops
from oracle import model as om


def make_batch(B, S, P, V, p_mask=0.2, ragged=True, seed=0, eval_mode=False):
    """Sequences of item ids in [3, V), right-padded with 0; Cloze masking through the oracle restatement of
    apply_dynamic_masking_task (training) or mask_last_token_only (eval)."""
    rng = np.random.RandomState(seed)
    out = {k: [] for k in ("labels", "input_word_ids", "input_mask", "masked_lm_ids", "masked_lm_positions", "masked_lm_weights")}
    for b in range(B):
        n = int(rng.randint(min(5, S), S + 1)) if ragged else S
        toks = rng.randint(3, V, size=n).tolist()
        el = host_ops.layout_element(toks, S, P, True, eval_mode, masked_lm_rate=p_mask, mask_token_rate=1.0,
                                     random_token_rate=0.0, vocab_size=V, seed=seed * 100003 + b)
        for k in out:
            out[k].append(el[k])
    return {k: torch.from_numpy(np.stack(v).astype(np.int64)) for k, v in out.items()}


def to_cuda(batch, device="cuda:0"):
    return {k: v.to(device).contiguous() for k, v in batch.items()}


def oracle_cfg(store_kwargs):
    return om.Config(vocab_size=store_kwargs["vocab_size"], hidden_size=store_kwargs["hidden_size"],
                     num_layers=store_kwargs["num_layers"], num_attention_heads=store_kwargs["num_attention_heads"],
                     max_sequence_length=store_kwargs["max_sequence_length"], inner_dim=store_kwargs["inner_dim"],
                     output_dropout=store_kwargs.get("output_dropout", 0.1),
                     attention_dropout=store_kwargs.get("attention_dropout", 0.1))


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
