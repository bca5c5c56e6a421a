"""CPU suite: invariants of the floating-point oracle (its parity with TF is unpinned -- see oracle/__init__.py --
so these guard the restatement itself): fp64 twin, loss at init, tied-table gradient structure, optimizer facts."""
import math

import torch

from oracle import model as om
from tests.helpers import make_batch


def _setup(dtype=torch.float32):
    cfg = om.Config(vocab_size=211, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=24, inner_dim=128)
    params = om.init_params(cfg, 0, dtype=dtype)
    batch = make_batch(6, 24, 5, 211, seed=3)
    return cfg, params, batch


def test_loss_at_init_is_ln_vocab_and_fp64_twin_agrees():
    cfg, p32, batch = _setup()
    p64 = {k: v.double() for k, v in p32.items()}
    o32 = om.model_forward(p32, cfg, batch)
    o64 = om.model_forward(p64, cfg, batch)
    l32 = float(om.masked_sparse_ce(batch["masked_lm_ids"], o32["mlm_logits"]))
    l64 = float(om.masked_sparse_ce(batch["masked_lm_ids"], o64["mlm_logits"]))
    assert abs(l32 - math.log(211)) < 0.1 and abs(l32 - l64) < 1e-5
    assert (o32["mlm_logits"].double() - o64["mlm_logits"]).abs().max() < 1e-4
    assert tuple(o32["mlm_logits"].shape) == (6, 5, 211) and tuple(o32["pooled_output"].shape) == (6, 64)


def test_padded_keys_do_not_influence_valid_tokens():
    cfg, params, batch = _setup()
    out1 = om.model_forward(params, cfg, batch)["sequence_output"]
    b2 = dict(batch)
    ids = batch["input_word_ids"].clone()
    ids[batch["input_mask"] == 0] = 7           # garbage under the padding mask
    b2["input_word_ids"] = ids
    out2 = om.model_forward(params, cfg, b2)["sequence_output"]
    m = batch["input_mask"].bool()
    assert torch.allclose(out1[m], out2[m], atol=1e-6)


def test_train_step_facts():
    cfg, params, batch = _setup()
    before = {k: v.clone() for k, v in params.items()}
    opt = om.AdamW({k: v for k, v in params.items() if not k.startswith("pooler")})
    m0, grads, lr0 = om.train_step(params, cfg, batch, opt, training=False)
    assert lr0 == 0.0 and all(torch.equal(before[k], params[k]) for k in params)      # lr(0) == 0
    assert "pooler_transform/kernel" not in grads                                      # no gradient through the pooler
    kb = grads["transformer/layer_0/self_attention/key/bias"]
    assert float(kb.abs().max()) < 1e-6                                                # softmax shift invariance
    m1, _, lr1 = om.train_step(params, cfg, batch, opt, training=False)
    assert lr1 > 0 and not torch.equal(before["word_embeddings/embeddings"], params["word_embeddings/embeddings"])
    assert torch.equal(before["pooler_transform/kernel"], params["pooler_transform/kernel"])
    # tied table gradient = gather part + projection part
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    E2 = leaves["word_embeddings/embeddings"].detach().clone().requires_grad_(True)   # separate leaf for the projection
    out = om.encoder_forward(leaves, cfg, batch["input_word_ids"], batch["input_mask"])
    head = dict(leaves)
    head["word_embeddings/embeddings"] = E2
    logits = om.masked_lm(head, out["sequence_output"], batch["masked_lm_positions"])
    loss = om.masked_sparse_ce(batch["masked_lm_ids"], logits)
    g_gather, g_proj = torch.autograd.grad(loss, [leaves["word_embeddings/embeddings"], E2])
    leaves2 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss2 = om.masked_sparse_ce(batch["masked_lm_ids"], om.model_forward(leaves2, cfg, batch)["mlm_logits"])
    (g_all,) = torch.autograd.grad(loss2, [leaves2["word_embeddings/embeddings"]])
    assert torch.allclose(g_all, g_gather + g_proj, atol=1e-7)


def test_clip_and_rank_semantics():
    g = {"a": torch.full((10,), 3.0), "b": torch.full((6,), -4.0)}
    clipped, gn = om.clip_by_global_norm(g, 5.0)
    assert abs(float(gn) - math.sqrt(90 + 96)) < 1e-5
    assert abs(float(torch.sqrt(sum((v ** 2).sum() for v in clipped.values()))) - 5.0) < 1e-5
    small, _ = om.clip_by_global_norm({"a": torch.tensor([0.3, 0.4])}, 5.0)
    assert torch.allclose(small["a"], torch.tensor([0.3, 0.4]))
    from oracle import host_ops as ho
    import numpy as np
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], dtype=np.float32)
    assert ho.stable_desc_argsort(s).tolist() == [1, 3, 0, 2, 4]        # lower index first on ties
    assert ho.rank_of(ho.rank_candidates(s, [4, 2, 0]), 0) == 2          # ground truth ranks behind a tied earlier negative
    assert om.first_argmax(torch.tensor([[1.0, 3.0, 3.0]])).tolist() == [1]
