"""Generation-2 (tcgen05) generic GEMM against the mma.sync generation it replaces, through a whole train step at a
hidden size where it is active (K, N multiples of 64, M >= 256).  Run with `-m gpu` on a B200."""
import os

import pytest
import torch

from tests.helpers import make_batch, to_cuda

pytestmark = pytest.mark.gpu


def _step(disable, B, S, P, kw, dropout):
    from bert4rec_b200.engine import ParamStore
    if disable:
        os.environ["B4R_DISABLE_TGEMM"] = "1"
    else:
        os.environ.pop("B4R_DISABLE_TGEMM", None)
    try:
        store = ParamStore(device="cuda:0", output_dropout=dropout, attention_dropout=dropout, **kw)
        store.init_weights(7)
        g = torch.Generator().manual_seed(8)
        with torch.no_grad():
            for k, v in store.tf_views().items():
                if k.endswith("bias") or k.endswith("beta"):
                    v.copy_((torch.randn(v.shape, generator=g) * 0.05).to(v.device))
        store.sync_shadow()
        store.ensure_training_buffers()
        cb = to_cuda(make_batch(B, S, P, kw["vocab_size"], seed=31))
        sess = store.session(B, S, P)
        n0 = sess.launch_count()
        sess.encode(cb["input_word_ids"], cb["input_mask"], training=True, seed=5, step=2)
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.transform(); sess.loss(); sess.backward(seed=5, step=2)
        torch.cuda.synchronize()
        out = {"seq": sess.sequence_output().float().clone(), "loss": sess.step_stats().clone()[:2],
               "grads": {k: v.clone() for k, v in store.grad_dict().items()}}
        return out
    finally:
        os.environ.pop("B4R_DISABLE_TGEMM", None)


@pytest.mark.parametrize("name,kw,B,S,P", [
    ("h256_persistent", dict(vocab_size=2001, hidden_size=256, num_layers=2, num_attention_heads=4, max_sequence_length=72,
                             inner_dim=1024), 64, 72, 10),      # 36 x 8 tiles > 148 CTAs: several tiles per CTA
    ("h128_ragged_m", dict(vocab_size=515, hidden_size=128, num_layers=1, num_attention_heads=4, max_sequence_length=40,
                           inner_dim=512), 11, 37, 6),          # M = 407: partial last row tile
])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_tcgen05_gemm_matches_mma_sync_generation(name, kw, B, S, P, dropout):
    ref = _step(True, B, S, P, kw, dropout)
    got = _step(False, B, S, P, kw, dropout)
    scale = float(ref["seq"].abs().max())
    assert float((got["seq"] - ref["seq"]).abs().max()) <= 2.5e-2 * scale
    assert torch.allclose(ref["loss"], got["loss"], rtol=2e-3)
    gmax = max(float(g.norm()) for g in ref["grads"].values())
    bad = []
    for k, g in ref["grads"].items():
        d = float((got["grads"][k] - g).norm())
        if not d <= 2e-2 * float(g.norm()) + 1e-6 * gmax:
            bad.append((k, d, float(g.norm())))
    assert not bad, (name, bad)


@pytest.mark.parametrize("name,kw,B,S,P", [
    ("h64_s200_two_key_tiles", dict(vocab_size=1203, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=200,
                                    inner_dim=256), 7, 200, 20),
    ("h256_d64_s72", dict(vocab_size=2001, hidden_size=256, num_layers=1, num_attention_heads=4, max_sequence_length=72,
                          inner_dim=512), 5, 72, 8),
    ("h128_d32_s130", dict(vocab_size=515, hidden_size=128, num_layers=1, num_attention_heads=4, max_sequence_length=130,
                           inner_dim=256), 4, 130, 9),
])
@pytest.mark.parametrize("dropout", [0.0, 0.25])
def test_tcgen05_attention_forward_matches_mma_sync_generation(name, kw, B, S, P, dropout):
    """Layered path: tcgen05 attention forward (k_tattn.cu) against attn_fwd_kernel: context, log-sum-exp, keep bits."""
    from bert4rec_b200.engine import ParamStore
    outs = []
    for disable in (True, False):
        if disable:
            os.environ.pop("B4R_ENABLE_TATTN", None)
        else:
            os.environ["B4R_ENABLE_TATTN"] = "1"     # opt-in kernel (see k_tattn.cu)
        os.environ["B4R_DISABLE_FUSED"] = "1"
        try:
            store = ParamStore(device="cuda:0", output_dropout=dropout, attention_dropout=dropout, **kw)
            store.init_weights(11)
            store.ensure_training_buffers()
            cb = to_cuda(make_batch(B, S, P, kw["vocab_size"], seed=37))
            sess = store.session(B, S, P)
            sess.encode(cb["input_word_ids"], cb["input_mask"], training=True, seed=9, step=4)
            torch.cuda.synchronize()
            o = {"ctx": sess.layer_tensor(0, "ctx").float().clone(), "lse": sess.layer_tensor(0, "lse").float().clone(),
                 "out": sess.sequence_output().float().clone()}
            if dropout > 0:
                o["keep"] = sess.attn_keep_mask(0).clone()
            outs.append(o)
        finally:
            os.environ.pop("B4R_ENABLE_TATTN", None)
            os.environ.pop("B4R_DISABLE_FUSED", None)
    ref, got = outs
    if dropout > 0:
        assert torch.equal(ref["keep"], got["keep"])
    for k in ("ctx", "lse", "out"):
        scale = float(ref[k].abs().max()) + 1e-6
        err = float((got[k] - ref[k]).abs().max())
        assert err <= 2.5e-2 * scale, (name, k, err, scale)
