"""Vocabulary-sharded tied projection (b4r_shard_*): two ranks x two catalogue slices emulated in ONE process
against the unsharded CUDA path of the same library (which `test_gpu_engine.py` pins to the CPU oracle)."""
import pytest
import torch

from tests.helpers import make_batch, to_cuda, rel_l2
from tests.test_gpu_engine import build, CONFIGS

pytestmark = pytest.mark.gpu


def _forward(sess, cb, dropout_seed=3):
    sess.select(cb["masked_lm_positions"], cb["masked_lm_ids"], cb["masked_lm_weights"], mode=0, want_aux=True)
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=True, seed=dropout_seed)
    sess.transform()


@pytest.mark.parametrize("name,dropout", [("h64_s50", 0.0), ("h64_s50", 0.1), ("h256_d64", 0.0), ("h128_s37", 0.0)])
def test_sharded_projection_matches_unsharded(name, dropout):
    from bert4rec_b200.engine import VocabShard, shard_range
    store, kw, B, S, P = build(name, dropout=dropout)
    store.ensure_training_buffers()
    V, H = kw["vocab_size"], kw["hidden_size"]
    sess = store.session(B, S, P)
    M = sess.Mcap
    batches = [to_cuda(make_batch(B, S, P, V, seed=11)), to_cuda(make_batch(B, S, P, V, p_mask=0.1, seed=12))]
    # ---- unsharded: gradient of the SUM loss over both "ranks" = sum of the two backward passes
    ref_grads = torch.zeros_like(store.grads)
    ref_stats = torch.zeros(8, device="cuda:0")
    for cb in batches:
        _forward(sess, cb)
        sess.loss()
        sess.backward(seed=3)
        torch.cuda.synchronize()
        ref_grads += store.grads
        ref_stats += sess.step_stats()
    # ---- sharded: gather the rows of both "ranks"
    rows, labels, weights, mult, counts = [], [], [], [], []
    for cb in batches:
        _forward(sess, cb)
        rows.append(sess.mlm_hidden().clone()); labels.append(sess.labels().clone()); weights.append(sess.row_weights().clone())
        mult.append(sess.row_mult().clone()); counts.append(sess.counts().clone())
    rows, labels, weights = torch.stack(rows), torch.stack(labels), torch.stack(weights)
    mult, counts = torch.stack(mult), torch.stack(counts)
    assert int(counts[0, 0]) != int(counts[1, 0])   # ragged: the two ranks contribute different row counts
    world = 2
    shards = [VocabShard(store, world, M, *shard_range(V, world, r)) for r in range(world)]
    assert shard_range(V, world, 0)[1] == shard_range(V, world, 1)[0] and shard_range(V, world, 1)[1] == V
    parts = []
    for sh in shards:
        sh.pack(rows, labels, weights, mult, counts)
        parts.append(sh.partial())
    parts = torch.stack(parts).contiguous()
    stats = torch.zeros(16, device="cuda:0")
    store.grads.zero_()   # (the unsharded passes above left their gradients behind)
    dt = torch.zeros(world, M, H, device="cuda:0")
    for i, sh in enumerate(shards):
        sh.merge(parts, world * B, stats if i == 0 else None)
        dt += sh.backward(zero_all=(i == 0))
    torch.cuda.synchronize()
    for sh in shards:   # loss / accuracy sums of the global batch, identical on every shard
        st = sh.step_stats()
        assert int(st[1]) == int(ref_stats[1]) and int(st[4]) == int(ref_stats[4])
        assert abs(float(st[0]) - float(ref_stats[0])) < 2e-4 * abs(float(ref_stats[0]))
        assert abs(int(st[2]) - int(ref_stats[2])) <= 1 and abs(int(st[3]) - int(ref_stats[3])) <= 1
        assert int(sh.counts()[0]) == int(counts[:, 0].sum()) and int(sh.counts()[1]) == int(counts[:, 1].sum())
    assert float(stats[6]) == world * B and abs(float(stats[5]) / float(stats[6]) - float(ref_stats[0] / ref_stats[1])) < 1e-3
    got = store.grads.clone()          # the projection part: table slice of each shard + output bias
    for r, cb in enumerate(batches):   # the rest of the backward on the owning "rank"
        _forward(sess, cb)
        store.grads.zero_()
        sess.backward_from_dt(dt[r].contiguous(), seed=3)
        torch.cuda.synchronize()
        got += store.grads
    gd = store.tf_views(got)
    rd = store.tf_views(ref_grads)
    bad = []
    gmax = max(float(g.norm()) for g in rd.values())
    for k, g in rd.items():
        err = float((gd[k].double() - g.double()).norm())
        if not err < 5e-3 * float(g.norm()) + 1e-5 * gmax:
            bad.append((k, err, float(g.norm())))
    assert not bad, f"sharded vs unsharded gradient mismatches (name, l2 err, ref norm): {bad}"


def test_shard_argument_checks():
    from bert4rec_b200.engine import VocabShard
    store, kw, B, S, P = build("h64_s50")
    with pytest.raises(ValueError):
        VocabShard(store, 2, 64, 10, 10)        # empty slice
    with pytest.raises(ValueError):
        VocabShard(store, 2, 64, 0, kw["vocab_size"] + 1)
