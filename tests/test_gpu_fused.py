"""The whole-encoder fused tcgen05 forward (k_enc_fused.cu) against the layered kernels it replaces, tensor by
tensor, and through a full training step.  Run with `-m gpu` on a B200."""
import pytest
import torch

from tests.helpers import make_batch, to_cuda

pytestmark = pytest.mark.gpu

SHAPES = {
    # name: (B, S, P, inner_dim, layers)   hidden 64, 2 heads
    "s50_two_per_tile": (25, 50, 8, 64, 2),       # odd batch: the last tile holds one sequence
    "s20_six_per_tile": (13, 20, 4, 128, 3),
    "s128_one_per_tile": (3, 128, 20, 64, 1),
    "s64_i192": (5, 64, 10, 192, 2),
    "s33_slot64": (9, 33, 6, 64, 2),
    "s7_tiny": (40, 7, 2, 64, 1),
}
SAVED = ("x0", "qkv", "ctx", "a_pre", "y", "h_pre", "h", "o_pre", "out", "mean1", "rstd1", "mean2", "rstd2", "lse")


def _store(S, I, L, dropout):
    from bert4rec_b200.engine import ParamStore
    store = ParamStore(vocab_size=977, hidden_size=64, num_layers=L, num_attention_heads=2, max_sequence_length=S,
                       inner_dim=I, output_dropout=dropout, attention_dropout=dropout, device="cuda:0")
    store.init_weights(3)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for k, v in store.tf_views().items():
            if k.endswith("bias") or k.endswith("beta"):
                v.copy_((torch.randn(v.shape, generator=g) * 0.05).to(v.device))
            elif k.endswith("gamma"):
                v.copy_((1.0 + torch.randn(v.shape, generator=g) * 0.1).to(v.device))
    store.sync_shadow()
    return store


def _run(sess, cb, fused, training, backward, store):
    sess.set_flag(2, int(fused))
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=training, seed=99, step=7)
    out = {}
    L = store.L
    for l in range(L):
        names = SAVED if training else ("out",)
        for n in names:
            if n == "x0" and l > 0:
                continue
            out[f"{l}.{n}"] = sess.layer_tensor(l, n).float().clone()
        if training and store.cfg.attention_dropout > 0:
            out[f"{l}.keep"] = sess.attn_keep_mask(l).clone()
    if backward:
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.transform(); sess.loss(); sess.backward(seed=99, step=7)
        torch.cuda.synchronize()
        out["loss"] = sess.step_stats().clone()[:2]
        out["grads"] = {k: v.clone() for k, v in store.grad_dict().items()}
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_fused_forward_matches_layered(name, dropout):
    B, S, P, I, L = SHAPES[name]
    store = _store(S, I, L, dropout)
    store.ensure_training_buffers()
    cb = to_cuda(make_batch(B, S, P, 977, seed=17))
    sess = store.session(B, S, P)
    sess.set_flag(2, 1)   # raises if the fused kernel does not support the shape
    for training in (False, True):
        ref = _run(sess, cb, False, training, training, store)
        got = _run(sess, cb, True, training, training, store)
        for k in ref:
            if k in ("grads", "loss"):
                continue
            if k.endswith(".keep"):
                assert torch.equal(ref[k], got[k]), f"{name}: attention keep bits differ in {k}"
                continue
            a, b = got[k], ref[k]
            scale = float(b.abs().max()) + 1e-6
            err = float((a - b).abs().max())
            # same math, different summation order: a bf16 ulp or two of the tensor's range
            assert err <= 2.5e-2 * scale, f"{name} training={training}: {k} max abs diff {err} (range {scale})"
            assert float((a - b).norm() / (b.norm() + 1e-20)) < 4e-3, (name, k)
        if training:
            assert torch.allclose(ref["loss"], got["loss"], rtol=2e-3)
            for k, g in ref["grads"].items():
                d = float((got["grads"][k] - g).norm())
                assert d <= 2e-2 * float(g.norm()) + 1e-6 * max(float(x.norm()) for x in ref["grads"].values()), (name, k, d)


def test_fused_is_default_for_beauty_shape_and_counts_one_launch():
    store = _store(50, 64, 2, 0.1)
    sess = store.session(16, 50, 5)
    cb = to_cuda(make_batch(16, 50, 5, 977, seed=1))
    n0 = sess.launch_count()
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=False)
    assert sess.launch_count() - n0 == 1
    sess.set_flag(2, 0)
    n0 = sess.launch_count()
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=False)
    assert sess.launch_count() - n0 == 1 + 5 * 2


BWD_SHAPES = {
    "s50_two_per_tile": (25, 50, 8, 64, 2),
    "s20_four_per_tile": (13, 20, 4, 128, 3),
    "s128_one_per_tile": (3, 128, 20, 64, 1),
    "s33_slot64_i128": (9, 33, 6, 128, 2),
}


@pytest.mark.parametrize("name", list(BWD_SHAPES))
@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_fused_backward_matches_layered(name, dropout):
    """Same fused forward, then the one-launch tcgen05 backward against the 12-launches-per-layer backward."""
    B, S, P, I, L = BWD_SHAPES[name]
    store = _store(S, I, L, dropout)
    store.ensure_training_buffers()
    cb = to_cuda(make_batch(B, S, P, 977, seed=23))
    sess = store.session(B, S, P)
    sess.set_flag(2, 1)
    outs = []
    for fused_bwd in (0, 1):
        sess.set_flag(3, fused_bwd)   # raises if unsupported
        n0 = sess.launch_count()
        sess.encode(cb["input_word_ids"], cb["input_mask"], training=True, seed=99, step=7)
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.transform(); sess.loss(); sess.backward(seed=99, step=7)
        torch.cuda.synchronize()
        outs.append(({k: v.clone() for k, v in store.grad_dict().items()}, sess.launch_count() - n0))
    (ref, n_ref), (got, n_got) = outs
    assert n_got == n_ref - 12 * L, (n_ref, n_got)   # 12 launches per layer + embed_bwd -> one launch
    gmax = max(float(g.norm()) for g in ref.values())
    bad = []
    for k, g in ref.items():
        d = float((got[k] - g).norm())
        if not d <= 2e-2 * float(g.norm()) + 1e-6 * gmax:
            bad.append((k, d, float(g.norm())))
    assert not bad, (name, bad)


def _edge_batch(kind, B, S, P, V=977):
    """Hand-built batches for the corner cases of the reference's dataloader tests (empty / one-item sequences, no
    masked slot, everything masked)."""
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, V, (B, S), generator=g)
    mask = torch.ones(B, S, dtype=torch.int64)
    pos = torch.zeros(B, P, dtype=torch.int64)
    mids = torch.zeros(B, P, dtype=torch.int64)
    if kind == "one_item_sequences":          # every sequence holds a single (masked) item, the rest is padding
        ids[:, 1:] = 0; mask[:, 1:] = 0
        mids[:, 0] = ids[:, 0]; ids[:, 0] = 1
    elif kind == "fully_padded_rows":         # some sequences are entirely padding (input_mask all zero)
        ids[::2] = 0; mask[::2] = 0
        for b in range(1, B, 2):
            mids[b, 0] = ids[b, S - 1]; pos[b, 0] = S - 1; ids[b, S - 1] = 1
    elif kind == "no_masked_slot":            # nothing to predict: the loss has no valid slot
        pass
    elif kind == "all_slots_used":            # max_predictions_per_seq reached in every sequence
        for b in range(B):
            p = torch.randperm(S, generator=g)[:P].sort().values
            pos[b] = p; mids[b] = ids[b, p]; ids[b, p] = 1
    w = (mids != 0).to(torch.int64)
    return {"labels": ids.clone(), "input_word_ids": ids, "input_mask": mask, "masked_lm_ids": mids,
            "masked_lm_positions": pos, "masked_lm_weights": w}


@pytest.mark.parametrize("kind,B,S,P", [("one_item_sequences", 9, 12, 1), ("fully_padded_rows", 6, 50, 4),
                                          ("no_masked_slot", 4, 33, 3), ("all_slots_used", 1, 50, 30)])
def test_fused_edge_cases_match_layered(kind, B, S, P):
    store = _store(S, 64, 2, 0.2)
    store.ensure_training_buffers()
    cb = to_cuda(_edge_batch(kind, B, S, P))
    sess = store.session(B, S, P)
    outs = []
    for fused in (0, 1):
        sess.set_flag(2, fused); sess.set_flag(3, fused)
        sess.encode(cb["input_word_ids"], cb["input_mask"], training=True, seed=3, step=1)
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.transform(); sess.loss(); sess.backward(seed=3, step=1)
        torch.cuda.synchronize()
        outs.append((sess.sequence_output().float().clone(), sess.step_stats().clone()[:5], sess.counts().clone(),
                     {k: v.clone() for k, v in store.grad_dict().items()}))
    (seq0, st0, c0, g0), (seq1, st1, c1, g1) = outs
    assert torch.equal(c0, c1)
    assert torch.isfinite(seq1).all() and all(torch.isfinite(v).all() for v in g1.values())
    assert float((seq0 - seq1).abs().max()) <= 2.5e-2 * (float(seq0.abs().max()) + 1e-6)
    assert torch.allclose(st0, st1, rtol=2e-3, atol=1e-4), (st0, st1)
    gmax = max(float(v.norm()) for v in g0.values())
    for k in g0:
        assert float((g0[k] - g1[k]).norm()) <= 2e-2 * float(g0[k].norm()) + 1e-5 * gmax + 1e-7, k


@pytest.mark.parametrize("case", ["one_tile", "two_row_tiles", "many_items_two_chunks"])
def test_one_pass_ce_backward_matches_two_pass(case):
    """Generation 3 of the CE backward (k_ce_bwd_fused.cu: dT and dE from one recompute pass, session flag 5) against the
    two-pass tcgen05 generation on the same forward state: same dl tiles, different summation order only."""
    from bert4rec_b200.engine import ParamStore
    from tests.helpers import make_batch, to_cuda
    shape = {"one_tile": (1203, 24, 50, 8, 0.2), "two_row_tiles": (3709, 6, 200, 40, 0.2),
             "many_items_two_chunks": (40013, 64, 50, 20, 0.6)}[case]
    V, B, S, P, pm = shape
    store = ParamStore(device="cuda:0", vocab_size=V, hidden_size=64, num_layers=1, num_attention_heads=2, max_sequence_length=S,
                       inner_dim=64, output_dropout=0.0, attention_dropout=0.0)
    store.init_weights(0)
    with torch.no_grad():
        store.seg("head/output_bias").normal_(0, 0.5, generator=None)
    store.sync_shadow()
    store.ensure_training_buffers()
    cb = to_cuda(make_batch(B, S, P, V, p_mask=pm, seed=21))
    sess = store.session(B, S, P)
    grads = {}
    for gen3 in (0, 1):
        sess.set_flag(5, gen3)
        store.grads.zero_()
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.encode(cb["input_word_ids"], cb["input_mask"], training=True)
        sess.transform()
        sess.loss()
        sess.backward()
        torch.cuda.synchronize()
        grads[gen3] = {k: v.clone() for k, v in store.tf_views(store.grads).items()}
    n_valid = int(sess.counts()[0])
    if case == "many_items_two_chunks":
        assert n_valid > 5 * 128 and (V + 127) // 128 // 2 > 148
    bad = []
    gmax = max(float(g.norm()) for g in grads[0].values())
    for k, g in grads[0].items():
        err = float((grads[1][k].double() - g.double()).norm())
        if not err <= 2e-3 * float(g.norm()) + 1e-6 * gmax:
            bad.append((k, err, float(g.norm())))
    assert not bad, f"one-pass vs two-pass CE backward (name, l2 err, ref norm): {bad}"


def test_backward_with_no_valid_slot_is_finite_and_zero():
    """A batch whose masked_lm_weights are all zero (no item to predict): no gradient anywhere, nothing NaN, and the one-pass CE
    backward (zero work items, partial slots never written) agrees with the two-pass generation."""
    from bert4rec_b200.engine import ParamStore
    from tests.helpers import make_batch, to_cuda
    V, B, S, P = 1203, 16, 50, 8
    store = ParamStore(device="cuda:0", vocab_size=V, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=S,
                       inner_dim=64, output_dropout=0.0, attention_dropout=0.0)
    store.init_weights(0)
    store.ensure_training_buffers()
    batch = make_batch(B, S, P, V, seed=3)
    batch["masked_lm_weights"].zero_(); batch["masked_lm_ids"].zero_()
    cb = to_cuda(batch)
    sess = store.session(B, S, P)
    for gen3 in (1, 0):
        sess.set_flag(5, gen3)
        store.grads.normal_()                       # stale garbage must be overwritten, not accumulated
        sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
        sess.encode(cb["input_word_ids"], cb["input_mask"], training=True)
        sess.transform()
        sess.loss()
        sess.backward()
        torch.cuda.synchronize()
        st = sess.step_stats().cpu()
        assert int(sess.counts()[0]) == 0 and float(st[0]) == 0.0 and float(st[1]) == 0.0
        for k, g in store.tf_views(store.grads).items():      # (the flat buffer's padding between segments is never written)
            if not k.startswith("pooler_transform"):
                assert bool(torch.isfinite(g).all()) and float(g.abs().max()) == 0.0, (gen3, k)
    sess.set_flag(5, 1)
