"""GPU tests of the training / evaluation API around the step (b): pooler values, test_step,
fit (steps_per_epoch / validation_steps / History), BERT4RecTrainer.train with the best-only checkpoint and resume,
EarlyStopping, the pinned staging ring (host batches are never overwritten before their copy ran), graph invalidation when
compile() re-creates the sessions, candidate-id validation, and the DLPack hand-off of the C ABI.
Reference: bert4rec/models/bert4rec_model.py:151-192, bert4rec/trainers/bert4rec_trainer.py:37-68."""
import numpy as np
import pytest
import torch

from tests.helpers import make_batch, to_cuda, oracle_cfg

pytestmark = pytest.mark.gpu

KW = dict(vocab_size=503, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=32, inner_dim=128,
          output_dropout=0.1, attention_dropout=0.1)


def _model(seed=3, lr=5e-3, **over):
    from bert4rec_b200 import trainers
    from bert4rec_b200.models import BERT4RecModel
    from bert4rec_b200.models.components import networks
    kw = dict(KW, **over)
    model = BERT4RecModel(networks.Bert4RecEncoder(**kw, device="cuda:0", seed=seed))
    trainer = trainers.get("bert4rec", model=model)
    trainer.initialize_model(optimizer=trainers.optimizers.get("adamw", init_lr=lr, num_warmup_steps=2, num_train_steps=1000))
    return model, trainer


@pytest.mark.parametrize("over", [dict(), dict(hidden_size=256, num_attention_heads=4, inner_dim=512, max_sequence_length=72)])
def test_pooled_output_values_match_oracle(over):
    """pooled_output = tanh(x[:, 0] Wp + bp) (bert4rec_encoder.py:149-153,224-226): values, not only the shape."""
    from oracle import model as om
    model, _ = _model(**over)
    kw = dict(KW, **over)
    S = kw["max_sequence_length"]
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        v = model.store.tf_views()
        v["pooler_transform/kernel"].copy_((torch.randn(v["pooler_transform/kernel"].shape, generator=g) * (1.2 / kw["hidden_size"] ** 0.5)).cuda())
        v["pooler_transform/bias"].copy_((torch.randn(v["pooler_transform/bias"].shape, generator=g) * 0.1).cuda())
    model.store.sync_shadow()
    batch = make_batch(6, S, 5, kw["vocab_size"], seed=2)
    out = model(batch, training=False)
    ref = om.model_forward(model.state_dict(), oracle_cfg(kw), batch, training=False)
    assert float((out["pooled_output"].cpu() - ref["pooled_output"]).abs().max()) < 3e-2    # tanh of a bf16 GEMV over H terms
    assert float(ref["pooled_output"].abs().mean()) > 0.2                                       # (not a trivial all-zero comparison)
    assert float((out["sequence_output"].cpu() - ref["sequence_output"]).abs().max()) < 8e-2
    for a, b in zip(out["encoder_outputs"], ref["encoder_outputs"]):
        assert float((a.cpu() - b).abs().max()) < 8e-2


def test_test_step_matches_oracle_and_does_not_update():
    """test_step (bert4rec_model.py:175-192): inference forward, masked CE + both accuracies, no parameter update; evaluate()
    returns the Keras running means over the batches."""
    from oracle import model as om
    model, _ = _model()
    before = model.state_dict()
    batches = [make_batch(9, 32, 6, KW["vocab_size"], seed=s) for s in (4, 5, 6)]
    cfg = oracle_cfg(KW)
    model.reset_metrics("test")
    losses = []
    for b in batches:
        ref = om.model_forward(before, cfg, b, training=False)
        y = b["masked_lm_ids"]
        losses.append((float(om.masked_sparse_ce(y, ref["mlm_logits"])), float(om.masked_accuracy(y, ref["mlm_logits"])),
                       float(om.sparse_categorical_accuracy(y, ref["mlm_logits"])), b["input_word_ids"].shape[0]))
    one = dict(model.test_step(batches[0]))
    assert abs(one["loss"] - losses[0][0]) < 2e-3 * losses[0][0]
    assert abs(one["masked_accuracy"] - losses[0][1]) < 0.05 and abs(one["sparse_categorical_accuracy"] - losses[0][2]) < 0.05
    res = model.evaluate(batches)
    n = sum(l[3] for l in losses)
    assert abs(res["loss"] - sum(l[0] * l[3] for l in losses) / n) < 2e-3 * res["loss"]     # keras Mean weighted by batch size
    after = model.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)
    assert int(model.store.step_counter.item()) == 0
    assert set(res) == {"loss", "sparse_categorical_accuracy", "masked_accuracy"}
    assert model.evaluate(batches, steps=1)["loss"] == pytest.approx(one["loss"], rel=1e-6)


def test_fit_history_steps_and_validation():
    model, _ = _model()
    train = [make_batch(8, 32, 6, KW["vocab_size"], seed=s) for s in range(5)]
    val = [make_batch(8, 32, 6, KW["vocab_size"], seed=100 + s) for s in range(3)]
    events = []

    class Rec:
        def set_model(self, m): events.append("set_model")
        def on_train_begin(self, logs=None): events.append("train_begin")
        def on_epoch_begin(self, e, logs=None): events.append(("epoch_begin", e))
        def on_epoch_end(self, e, logs=None): events.append(("epoch_end", e, sorted(logs)))
        def on_train_end(self, logs=None): events.append("train_end")

    h = model.fit(x=train, validation_data=val, epochs=3, callbacks=[Rec()], steps_per_epoch=4, validation_steps=2)
    assert h.epoch == [0, 1, 2]
    keys = {"loss", "sparse_categorical_accuracy", "masked_accuracy"}
    assert set(h.history) == keys | {"val_" + k for k in keys} and all(len(v) == 3 for v in h.history.values())
    assert int(model.store.step_counter.item()) == 12                        # steps_per_epoch honoured
    assert h.history["loss"][-1] < h.history["loss"][0]                      # it trains
    assert events[0] == "set_model" and events[1] == "train_begin" and events[-1] == "train_end"
    assert [e for e in events if isinstance(e, tuple) and e[0] == "epoch_end"][0][2] == sorted(h.history)
    # validation_steps: the val metrics are those of evaluate(val, steps=2)
    assert model.evaluate(val, steps=2)["loss"] == pytest.approx(h.history["val_loss"][-1], rel=1e-6)


def test_trainer_train_best_only_checkpoint_resume_and_early_stopping(tmp_path):
    """BERT4RecTrainer.train (bert4rec_trainer.py:37-68): ModelCheckpoint(monitor=val_masked_accuracy, save_best_only) writes only
    on improvement; a second train() call resumes from the checkpoint's weights; EarlyStopping stops the loop."""
    from bert4rec_b200.trainers import callbacks as cb
    train = [make_batch(8, 32, 6, KW["vocab_size"], seed=s) for s in range(4)]
    val = [make_batch(8, 32, 6, KW["vocab_size"], seed=50 + s) for s in range(2)]
    model, trainer = _model(seed=7)
    ck = tmp_path / "ckpt" / "weights"
    h = trainer.train(train, val, checkpoint_path=ck, epochs=4)
    assert (tmp_path / "ckpt" / "weights.npz").is_file()
    ckpt_cb = [c for c in trainer.callbacks if isinstance(c, cb.ModelCheckpoint)][0]
    assert ckpt_cb.monitor == "val_masked_accuracy" and ckpt_cb.save_best_only
    best = max(h.history["val_masked_accuracy"])
    assert ckpt_cb.best == best
    # the file holds the weights of the BEST epoch: evaluating them reproduces the best validation accuracy
    m2, t2 = _model(seed=99)
    m2.load_weights(ck)
    assert m2.evaluate(val)["masked_accuracy"] == pytest.approx(best, abs=1e-6)
    # resume: a fresh trainer pointed at the same path starts from those weights (optimizer slots are not restored, reference :57-58)
    sd_ck = m2.state_dict()
    m3, t3 = _model(seed=123)
    seen = {}
    class Peek(cb.Callback):
        def on_train_begin(self, logs=None):
            seen["sd"] = self.model.state_dict()
    t3.append_callback(Peek())
    t3.train(train, val, checkpoint_path=ck, epochs=1)
    assert all(torch.equal(seen["sd"][k], sd_ck[k]) for k in sd_ck)
    # EarlyStopping: a monitor that cannot improve stops after `patience` further epochs
    m4, t4 = _model(seed=5, lr=0.0)
    t4.append_callback(cb.EarlyStopping(monitor="val_loss", patience=1))
    h4 = t4.train(train, val, epochs=10)
    assert len(h4.epoch) == 3 and m4.stop_training


def test_host_batches_are_not_overwritten_before_their_copy_ran():
    """fit() over DISTINCT pageable host batches without reading any metric (the host runs ahead of the stream) must train exactly
    like a run that synchronises after every step (single pinned staging buffer rewritten under a pending copy)."""
    batches = [make_batch(16, 32, 6, KW["vocab_size"], seed=s) for s in range(12)]
    out = []
    for sync in (False, True):
        model, _ = _model(seed=21)
        for i in range(36):
            r = model.train_step(batches[i % 12])
            if sync:
                _ = r["loss"]
        torch.cuda.synchronize()
        out.append(model.state_dict())
    a, b = out
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_rank_graphs_are_dropped_when_compile_recreates_sessions():
    """evaluate (untrained) -> compile() -> evaluate: compile allocates the gradient buffers and re-creates every session; the
    ranking graphs captured before must not be replayed against the freed workspaces."""
    from bert4rec_b200.models import BERT4RecModel
    from bert4rec_b200.models.components import networks
    from bert4rec_b200 import trainers
    model = BERT4RecModel(networks.Bert4RecEncoder(**KW, device="cuda:0", seed=1))
    ev = make_batch(8, 32, 6, KW["vocab_size"], seed=9, eval_mode=True)
    gt = ev["masked_lm_ids"][:, 0].clone()
    cand = torch.from_numpy(np.random.RandomState(0).randint(3, KW["vocab_size"], size=(8, 21)).astype(np.int64))
    cand[:, 20] = gt
    r0 = [model.rank_candidates(ev, cand, gt)[1].cpu().clone() for _ in range(3)]      # eager, capture, replay
    assert torch.equal(r0[0], r0[1]) and torch.equal(r0[0], r0[2])
    gen = model.store.generation
    trainers.get("bert4rec", model=model).initialize_model()
    assert model.store.generation == gen + 1
    junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0").fill_(0xFF)       # reuse the freed blocks
    r1 = [model.rank_candidates(ev, cand, gt)[1].cpu().clone() for _ in range(3)]
    assert all(torch.equal(r0[0], r) for r in r1)
    del junk


def test_rank_candidates_rejects_foreign_ids_and_extra_rows():
    """Candidate ids outside the model's catalogue are ranked last instead of being dereferenced; candidate rows beyond the number
    of selected slots get rank 0."""
    model, _ = _model()
    V = KW["vocab_size"]
    ev = make_batch(6, 32, 6, V, seed=3, eval_mode=True)
    gt = ev["masked_lm_ids"][:, 0].clone()
    cand = torch.from_numpy(np.random.RandomState(1).randint(3, V, size=(6, 11)).astype(np.int64))
    cand[:, 10] = gt
    sess, _ = model._encode_for_ranking(ev)
    good, sc_good, rk_good = sess.rank_candidates(cand.cuda(), gt.cuda(), want_ranking=True, want_scores=True)
    bad = cand.clone()
    bad[:, 3] = V + 12345
    bad[:, 5] = -7
    ranking, scores, rank = sess.rank_candidates(bad.cuda(), gt.cuda(), want_ranking=True, want_scores=True)
    sc = scores.cpu()
    assert torch.isinf(sc[:, 3]).all() and torch.isinf(sc[:, 5]).all() and (sc[:, 3] < 0).all()
    assert torch.equal(ranking.cpu()[:, -2:], torch.tensor([[V + 12345, -7]] * 6))     # last, in candidate order (stable)
    keep = [c for c in range(11) if c not in (3, 5)]
    assert torch.equal(sc[:, keep], sc_good.cpu()[:, keep])
    # more candidate rows than selected slots
    n = int(sess.counts()[0])
    assert n == 6
    many = torch.cat([cand, cand[:2]]).cuda()
    _, _, rank2 = sess.rank_candidates(many, torch.cat([gt, gt[:2]]).cuda(), want_ranking=False)
    assert torch.equal(rank2.cpu()[:6], rk_good.cpu()) and int(rank2[6]) == 0 and int(rank2[7]) == 0


def test_dlpack_capsule_through_the_c_abi():
    """b4r_dl_view_of on a real DLPack capsule (north_star: tensors cross the C ABI zero-copy via DLPack): pointer, dtype, shape;
    non-contiguous and host tensors are refused."""
    from bert4rec_b200 import engine, _lib
    t = torch.arange(3 * 5 * 7, dtype=torch.int64, device="cuda:0").view(3, 5, 7)
    v = engine.dl_view(t)
    assert v.data == t.data_ptr() and v.ndim == 3 and list(v.shape)[:3] == [3, 5, 7]
    assert v.device_type == 2 and v.dtype_bits == 64 and v.dtype_code == 0          # kDLCUDA, int64
    h = torch.zeros(4, 8, dtype=torch.bfloat16, device="cuda:0")
    vh = engine.dl_view(h)
    assert vh.dtype_bits == 16 and vh.dtype_code == 4 and vh.data == h.data_ptr()   # kDLBfloat
    with pytest.raises(_lib.B4RError):
        engine.dl_view(t.transpose(0, 2))
    with pytest.raises(_lib.B4RError):
        engine.dl_view(torch.zeros(3))
    # the session's input hand-off goes through the same adapter
    model, _ = _model()
    b = to_cuda(make_batch(4, 32, 6, KW["vocab_size"], seed=1))
    sess = model.store.session(4, 32, 6)
    with pytest.raises((_lib.B4RError, AssertionError)):
        sess.encode(b["input_word_ids"].t(), b["input_mask"])


@pytest.mark.parametrize("H,N,V,k", [(64, 2, 1203, 10), (128, 4, 5003, 100), (256, 4, 2001, 37), (64, 2, 40007, 128)])
def test_full_catalogue_top_k_is_the_head_of_the_stable_argsort(H, N, V, k):
    """b4r_topk_full: ids / scores of the k best items per slot == the first k of the stable descending argsort of the kernel's own
    logits (tf.argsort(DESCENDING) semantics: lower id first among equal logits, bert4rec_model.py:235-236); duplicate table rows
    force exact ties; vocabulary shards merged with b4r_topk_merge give the same lists (the multi-GPU path)."""
    from bert4rec_b200.engine import ParamStore, topk_merge
    kw = dict(vocab_size=V, hidden_size=H, num_layers=1, num_attention_heads=N, max_sequence_length=24, inner_dim=2 * H)
    store = ParamStore(device="cuda:0", **kw)
    store.init_weights(5)
    with torch.no_grad():
        tv = store.tf_views()
        E, vb = tv["word_embeddings/embeddings"], tv["cls/predictions/output_bias/bias"]
        E.mul_(20.0)
        vb.copy_((torch.randn(V, generator=torch.Generator().manual_seed(3)) * 0.1).cuda())
        for a, b in ((7, 700), (8, 9), (V - 1, 11), (500, 501), (500, 502)):      # exact ties (same row, same bias)
            E[b] = E[a]; vb[b] = vb[a]
    store.sync_shadow()
    B, S, P = 70, 24, 3
    batch = make_batch(B, S, P, V, seed=4)
    cb = to_cuda(batch)
    sess = store.session(B, S, P)
    sess.encode(cb["input_word_ids"], cb["input_mask"], training=False)
    sess.select(cb["masked_lm_positions"], None, cb["masked_lm_weights"], mode=1)
    sess.transform()
    n = int(sess.counts()[0])
    ids, scores = sess.topk_full(k, n_rows=n)
    logits = sess.logits(n)
    order = torch.sort(logits, dim=-1, descending=True, stable=True)
    # the GEMM that materialises the test's logits and the top-k kernel round identically only up to fp32 summation order:
    # compare the ORDER through the kernel's own scores, and the scores against the logits
    got_s = scores.cpu()
    assert float((got_s - torch.gather(logits, 1, ids).cpu()).abs().max()) < 2e-3
    assert bool((got_s[:, :-1] >= got_s[:, 1:]).all())
    tie = got_s[:, :-1] == got_s[:, 1:]
    assert bool((ids.cpu()[:, :-1][tie] < ids.cpu()[:, 1:][tie]).all())                      # lower id first among equal scores
    # membership: everything left out scores no better than the k-th entry
    kth = got_s[:, -1:]
    mask = torch.ones_like(logits, dtype=torch.bool).scatter_(1, ids, False).cpu()
    assert float((logits.cpu()[mask].view(n, V - k) - kth).max()) < 2e-3
    agree = (order.indices[:, :k].cpu() == ids.cpu()).float().mean()
    assert float(agree) > 0.97, float(agree)                                                 # identical up to near-ties of two kernels
    for r in range(n):
        assert len(set(ids[r].tolist())) == k
    # shards + merge == whole
    cuts = [0, V // 3, V // 3 + 1, V]
    keys = torch.stack([sess.topk_full(k, lo, hi, n_rows=n, want_keys=True)[2] for lo, hi in zip(cuts[:-1], cuts[1:])])
    m_ids, m_sc = topk_merge(keys.contiguous())
    assert torch.equal(m_ids, ids) and torch.equal(m_sc, scores)
    # external rows (the all-gathered rows of other ranks) give the same lists
    e_ids, e_sc = sess.topk_full(k, t_rows=sess.mlm_hidden()[:n].clone().contiguous())
    assert torch.equal(e_ids, ids) and torch.equal(e_sc, scores)


def test_top_k_items_api_and_recommender_exclusion():
    from bert4rec_b200.apps import Recommender, InferenceDataloader
    model, _ = _model(seed=6)
    V = KW["vocab_size"]
    ev = make_batch(5, 32, 6, V, seed=12, eval_mode=True)
    ids, sc = model.top_k_items(ev, 7)
    assert tuple(ids.shape) == (5, 7)
    own = model(ev, training=False)["mlm_logits"][:, 0]                      # slot 0 is the only weighted slot of an eval batch
    ref = torch.sort(own, dim=-1, descending=True, stable=True).indices[:, :7]
    assert float((ref == ids).float().mean()) > 0.9
    ban = [ids[r, :3].tolist() for r in range(5)]
    ids2, _ = model.top_k_items(ev, 4, exclude=ban)
    assert torch.equal(ids2, ids[:, 3:7])
    dl = InferenceDataloader(max_seq_len=32)
    dl.tokenizer.tokenize([f"item{j}" for j in range(V - 3)])
    hist = [f"item{j}" for j in (5, 17, 33, 120, 64)]
    got = Recommender(model, dl)(list(hist))
    assert got not in hist
    inp = dl.prepare_inference(list(hist))
    logits = model(inp, training=False)["mlm_logits"][0, 0].clone()
    logits[torch.tensor(dl.tokenizer.tokenize(hist), device=logits.device)] = -float("inf")
    assert float(logits.max() - logits[dl.tokenizer.tokenize(got)]) < 2e-2
