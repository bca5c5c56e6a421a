"""GPU parity at the FULL sizes of BASELINE.json's configurations (C2 headline: V 12 004, S 50, B 256, P 30; C1: V 3 709,
S 200, B 256, P 40): the oracle itself where it finishes in seconds, and size-independent properties elsewhere --
additivity of the SUM-loss gradient over a split batch, agreement of independent kernel generations (fused tcgen05 vs layered
mma.sync encoder, one-pass vs two-pass CE backward), vocabulary-shard additivity of the full-catalogue rank, rank semantics
recomputed from the kernel's own scores."""
import numpy as np
import pytest
import torch

import bench
from tests.helpers import make_batch, to_cuda, oracle_cfg, rel_l2

pytestmark = pytest.mark.gpu


def _store(wl, dropout=0.0):
    from bert4rec_b200.engine import ParamStore
    w = bench.WORKLOADS[wl]
    kw = {k: w[k] for k in bench.ENC_KEYS}
    kw.update(output_dropout=dropout, attention_dropout=dropout)
    store = ParamStore(device="cuda:0", **kw)
    store.init_weights(0)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for k, v in store.tf_views().items():
            if k.endswith(("bias", "beta")):
                v.copy_((torch.randn(v.shape, generator=g) * 0.02).to(v.device))
            elif k.endswith("gamma"):
                v.copy_((1.0 + torch.randn(v.shape, generator=g) * 0.05).to(v.device))
    store.sync_shadow()
    store.ensure_training_buffers()
    return store, kw, w


def _grads(store, sess, cb, flags=()):
    for f, v in flags:
        sess.set_flag(f, v)
    store.grads.zero_()
    sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=True)
    sess.transform()
    sess.loss()
    sess.backward()
    torch.cuda.synchronize()
    return store.grads.clone(), sess.step_stats().clone()


def _compare(store, got, ref, tol, what):
    gd, rd = store.tf_views(got), store.tf_views(ref)
    gmax = max(float(g.norm()) for g in rd.values())
    bad = [(k, float((gd[k].double() - g.double()).norm()), float(g.norm())) for k, g in rd.items()
           if not float((gd[k].double() - g.double()).norm()) <= tol * float(g.norm()) + 1e-5 * gmax]
    assert not bad, f"{what} (name, l2 err, ref norm): {bad}"


def test_c2_full_size_train_step_against_oracle():
    """The headline configuration at full size against the fp32 CPU oracle: loss, counts, every gradient tensor."""
    from oracle import model as om
    store, kw, w = _store("c2")
    B, S, P, V = w["batch"], w["seq_len"], w["max_pred"], w["vocab_size"]
    batch = make_batch(B, S, P, V, p_mask=w["mask_prob"], seed=77)
    sess = store.session(B, S, P)
    got, st = _grads(store, sess, to_cuda(batch))
    sd = {k: v.to(torch.bfloat16).float() if k.endswith(("kernel", "embeddings")) else v for k, v in store.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = om.model_forward(leaves, oracle_cfg(kw), batch, training=False)
    y = batch["masked_lm_ids"]
    loss = om.masked_sparse_ce(y, out["mlm_logits"])
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    n_valid = int((y != 0).sum())
    assert int(st[1]) == n_valid and int(st[4]) == B * P
    loss = loss.detach()
    assert abs(float(st[0]) / n_valid - float(loss)) < 1e-3 * float(loss)          # north_star: loss within 1e-3 relative
    assert abs(float(loss) - np.log(V)) < 0.2                                      # near ln V at initialisation
    gd = store.tf_views(got)
    ref = {k: g for k, g in zip(names, gs) if g is not None}
    gmax = max(float(g.norm()) for g in ref.values())
    bad = []
    for k, g in ref.items():
        err = float((gd[k].cpu().double() / n_valid - g.double()).norm())
        if not err < 4e-2 * float(g.norm()) + 1e-5 * gmax:
            bad.append((k, err, float(g.norm())))
    assert not bad, f"gradient mismatches at full C2 size (name, l2 err, ref norm): {bad}"


@pytest.mark.parametrize("wl", ["c2", "c1"])
def test_full_size_gradient_is_additive_over_a_split_batch(wl):
    """Gradient of the SUM loss of the whole batch == sum over its two halves (different session shapes, tile counts, split
    counts and partial-reduction trees: checks every cross-CTA reduction at production size)."""
    store, kw, w = _store(wl)
    B, S, P, V = w["batch"], w["seq_len"], w["max_pred"], w["vocab_size"]
    batch = make_batch(B, S, P, V, p_mask=w["mask_prob"], seed=5)
    whole, st = _grads(store, store.session(B, S, P), to_cuda(batch))
    half = store.session(B // 2, S, P)
    parts = torch.zeros_like(whole)
    n = 0.0
    for lo in (0, B // 2):
        g, s = _grads(store, half, to_cuda({k: v[lo:lo + B // 2] for k, v in batch.items()}))
        parts += g
        n += float(s[1])
    assert n == float(st[1])
    _compare(store, parts, whole, 5e-3, f"{wl}: halves vs whole batch")


def test_c2_full_size_kernel_generations_agree():
    """At full C2 size: fused tcgen05 encoder (flags 2, 3) vs the layered mma.sync encoder, one-pass vs two-pass CE backward
    (flag 5), tcgen05 vs mma.sync CE (flag 1) -- independent kernels, same mathematics."""
    store, kw, w = _store("c2")
    B, S, P, V = w["batch"], w["seq_len"], w["max_pred"], w["vocab_size"]
    cb = to_cuda(make_batch(B, S, P, V, p_mask=w["mask_prob"], seed=6))
    sess = store.session(B, S, P)
    base, st0 = _grads(store, sess, cb)
    two_pass, st1 = _grads(store, sess, cb, flags=[(5, 0)])
    _compare(store, two_pass, base, 2e-3, "two-pass vs one-pass CE backward")
    layered, st2 = _grads(store, sess, cb, flags=[(5, 1), (3, 0), (2, 0)])
    _compare(store, layered, base, 3e-2, "layered vs fused encoder")
    gen1, st3 = _grads(store, sess, cb, flags=[(1, 0)])
    sess.set_flag(1, 1); sess.set_flag(2, 1); sess.set_flag(3, 1)
    _compare(store, gen1, layered, 2e-2, "mma.sync vs tcgen05 CE (layered encoder)")
    for s in (st1, st2, st3):
        assert int(s[1]) == int(st0[1]) and abs(float(s[0]) - float(st0[0])) < 2e-3 * float(st0[0])


def test_c2_full_size_ranking_properties():
    """100-negative ranking at full size: rank == 1 + #(score > gt score) + #(equal score at a lower candidate index), from the
    kernel's own scores; the full-catalogue rank is additive over vocabulary shards."""
    store, kw, w = _store("c2")
    B, S, P, V = w["batch"], w["seq_len"], w["max_pred"], w["vocab_size"]
    batch = make_batch(B, S, P, V, seed=8, eval_mode=True)
    cb = to_cuda(batch)
    sess = store.session(B, S, P)
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=False)
    sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0)
    sess.transform()
    sess.loss()
    n = int(sess.counts()[0])
    assert n == B
    rng = np.random.RandomState(3)
    gt = batch["masked_lm_ids"][:, 0].numpy()
    cand = rng.randint(3, V, size=(B, 101)).astype(np.int64)
    clash = cand[:, :100] == gt[:, None]                       # negatives never contain the ground truth (evaluator contract)
    cand[:, :100][clash] = np.where(gt + 1 < V, gt + 1, 3)[:, None].repeat(100, 1)[clash]
    cand[:, 100] = gt
    cand[:, 7] = cand[:, 9]                                    # duplicate candidates: ties are broken by the lower index
    ranking, scores, rank = sess.rank_candidates(torch.from_numpy(cand).cuda(), torch.from_numpy(gt.astype(np.int64)).cuda(),
                                                 want_ranking=True, want_scores=True)
    sc = scores.cpu().numpy()
    gs = sc[:, 100:101]
    expect = 1 + (sc[:, :100] > gs).sum(1) + (sc[:, :100] == gs).sum(1)          # gt is the LAST candidate: equal scores rank ahead
    assert np.array_equal(rank.cpu().numpy().astype(np.int64), expect)
    order = np.argsort(-sc, axis=1, kind="stable")
    assert np.array_equal(ranking.cpu().numpy(), np.take_along_axis(cand, order, axis=1))
    whole = sess.rank_full(n)
    a = sess.rank_full(n, 0, 5000)
    b = sess.rank_full(n, 5000, V)
    assert torch.equal(whole, a + b) and int(whole.min()) >= 0 and int(whole.max()) < V


def _oracle_step_parity(wl, B, seed, ragged=True):
    """One training forward + backward at the workload's REAL shape (all of V, S, H, L, N, I, P; batch ``B``) against the fp32
    CPU oracle: logits of all P slots (rel-L2 1e-2), loss (1e-3 relative), counts, every gradient tensor (4e-2 rel-L2)."""
    from oracle import model as om
    store, kw, w = _store(wl)
    S, P, V = w["seq_len"], w["max_pred"], w["vocab_size"]
    batch = make_batch(B, S, P, V, p_mask=w["mask_prob"], ragged=ragged, seed=seed)
    cb = to_cuda(batch)
    sess = store.session(B, S, P)
    # forward: logits over ALL P slots (BERT4RecModel.call semantics)
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=False)
    sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=2)
    sess.transform()
    logits = sess.logits(B * P).cpu().reshape(B, P, V)
    seq = sess.sequence_output().float().cpu()
    got, st = _grads(store, sess, cb)
    sd = {k: v.to(torch.bfloat16).float() if k.endswith(("kernel", "embeddings")) else v for k, v in store.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = om.model_forward(leaves, oracle_cfg(kw), batch, training=False)
    y = batch["masked_lm_ids"]
    loss = om.masked_sparse_ce(y, out["mlm_logits"])
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    n_valid = int((y != 0).sum())
    assert int(st[1]) == n_valid and int(st[4]) == B * P
    loss = loss.detach()
    assert abs(float(st[0]) / n_valid - float(loss)) < 1e-3 * float(loss), (float(st[0]) / n_valid, float(loss))
    err = (seq - out["sequence_output"].detach()).abs().max().item()
    assert err < 8e-2, f"{wl}: sequence_output max abs err {err}"
    rl = rel_l2(logits, out["mlm_logits"].detach())
    assert rl < 1e-2, f"{wl}: logits rel l2 {rl}"                                  # north_star: logits within 1e-2 relative
    gd = store.tf_views(got)
    ref = {k: g for k, g in zip(names, gs) if g is not None}
    gmax = max(float(g.norm()) for g in ref.values())
    bad = []
    for k, g in ref.items():
        e = float((gd[k].cpu().double() / n_valid - g.double()).norm())
        if not e < 4e-2 * float(g.norm()) + 1e-5 * gmax:
            bad.append((k, e, float(g.norm())))
    assert not bad, f"{wl}: gradient mismatches (name, l2 err, ref norm): {bad}"


def test_c1_full_size_train_step_against_oracle():
    """BASELINE config 1 (ML-1m shape) at its full size, B = 256 (152 MB of oracle logits)."""
    _oracle_step_parity("c1", 256, seed=71)


def test_c3_shape_train_step_against_oracle():
    """BASELINE config 3 (ML-20m shape: V 26 732, S 200, H 64) on a 48-sequence slice of the per-GPU batch."""
    _oracle_step_parity("c3", 48, seed=73)


def test_c4_shape_train_step_against_oracle():
    """BASELINE config 4 at its real shape (H 256, 4 layers x 4 heads of 64, S 200, I 1024, V 13 047), 24 sequences."""
    _oracle_step_parity("c4", 24, seed=79)


def test_c4_shape_dense_sequences_against_oracle():
    """The same with full-length sequences (the bench's shape: no padded keys, every 128-row tile full)."""
    _oracle_step_parity("c4", 8, seed=83, ragged=False)
