"""GPU tests of the rows SURVEY 8(f) marks next after the hot path: model directory save / load (N2) and the Ranker /
Recommender apps (N4), against the oracle's logits."""
import json

import numpy as np
import pytest
import torch

from tests.helpers import make_batch, oracle_cfg

pytestmark = pytest.mark.gpu

KW = dict(vocab_size=403, hidden_size=64, num_layers=2, num_attention_heads=2, max_sequence_length=24, inner_dim=64,
          output_dropout=0.1, attention_dropout=0.1)


def _model(seed=3):
    from bert4rec_b200 import trainers
    from bert4rec_b200.models import BERT4RecModel
    from bert4rec_b200.models.components import networks
    model = BERT4RecModel(networks.Bert4RecEncoder(**KW, device="cuda:0", seed=seed))
    trainers.get("bert4rec", model=model).initialize_model(
        optimizer=trainers.optimizers.get("adamw", init_lr=5e-3, num_warmup_steps=2, num_train_steps=1000))
    return model


def test_model_directory_roundtrip_and_resume(tmp_path):
    from bert4rec_b200 import tokenizers
    from bert4rec_b200.models import BERT4RecModelWrapper
    batches = [make_batch(8, 24, 5, 403, seed=s) for s in range(4)]
    model = _model()
    for b in batches[:2]:
        model.train_step(b)
    tok = tokenizers.get("simple")
    tok.tokenize(["[PAD]", "[MASK]", "[UNK]"] + [f"item{j}" for j in range(400)])
    w = BERT4RecModelWrapper(model)
    w.update_meta({"trained_on_dataset": "synthetic"})
    assert w.save(tmp_path / "run1", tokenizer=tok, mode=2) is True
    meta = json.load(open(tmp_path / "run1" / "meta_config.json"))
    assert meta["model"] == "BERT4Rec" and meta["tokenizer"] == "simple" and meta["trained_on_dataset"] == "synthetic"
    assert meta["encoder_config"]["vocab_size"] == 403 and (tmp_path / "run1" / "vocab.txt").exists()
    assets = BERT4RecModelWrapper.load(tmp_path / "run1", mode=2)
    m2 = assets["model_wrapper"].model
    assert assets["tokenizer"].get_vocab_size() == tok.get_vocab_size() and assets["tokenizer"].tokenize("item7") == tok.tokenize("item7")
    s1, s2 = model.state_dict(), m2.state_dict()
    assert set(s1) == set(s2) and all(torch.equal(s1[k], s2[k]) for k in s1)
    # resume: the reloaded model continues exactly like the original (same Adam moments, iteration counter, dropout stream)
    m2.compile(optimizer=model.optimizer)
    la, lb = [], []
    for b in batches[2:]:
        for m, acc in ((model, la), (m2, lb)):
            m.reset_metrics("train")          # (the returned loss is the Keras running mean since the last reset)
            acc.append(m.train_step(b)["loss"])
    assert la == lb
    s1, s2 = model.state_dict(), m2.state_dict()
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    with pytest.raises(ValueError):
        BERT4RecModelWrapper.load(tmp_path / "missing", mode=2)


def test_ranker_and_recommender_against_oracle():
    from oracle import model as om
    from bert4rec_b200.apps import Ranker, Recommender, InferenceDataloader
    model = _model(seed=5)
    dl = InferenceDataloader(max_seq_len=24)
    items = [f"item{j}" for j in range(400)]
    dl.tokenizer.tokenize(items)                                   # ids 3..402
    history = [f"item{j}" for j in (5, 17, 33, 5, 120, 399, 64)]
    ranker, rec = Ranker(model, dl), Recommender(model, dl)
    inp = dl.prepare_inference(list(history))
    logits = om.model_forward(model.state_dict(), oracle_cfg(KW), inp, training=False)["mlm_logits"][0, -1]
    own = model(inp, training=False)["mlm_logits"][0, -1].cpu()
    assert float((own - logits).abs().max()) < 5e-2
    # whole vocabulary (reference sign convention: descending order of the NEGATED logits)
    for name in ("item9", "item250"):
        rank, text = ranker(list(history), name)
        t = dl.tokenizer.tokenize(name)
        expect_own = int((torch.argsort(-own, descending=True, stable=True) == t).nonzero()[0, 0]) + 1
        expect_ref = int((torch.argsort(-logits, descending=True, stable=True) == t).nonzero()[0, 0]) + 1
        assert rank == expect_own and abs(rank - expect_ref) <= 8 and "whole vocabulary" in text and name in text
    # relative to a candidate list
    cands = [f"item{j}" for j in (9, 250, 17, 300, 301, 12)]
    rank, text = ranker(list(history), "item300", cands)
    ct = torch.tensor(dl.tokenizer.tokenize(cands))
    order = torch.argsort(-logits[ct], descending=True, stable=True)
    assert abs(rank - (int((ct[order] == dl.tokenizer.tokenize("item300")).nonzero()[0, 0]) + 1)) <= 1
    assert "relative to 6 other elements" in text
    with pytest.raises(IndexError):
        ranker(list(history), "item8", cands)
    # recommendation: arg-max over the vocabulary minus the history
    got = rec(list(history))
    seen = set(dl.tokenizer.tokenize(list(history)))
    masked = logits.clone(); masked[list(seen)] = -float("inf")
    assert got not in history and dl.tokenizer.tokenize(got) not in seen
    assert float(masked.max() - masked[dl.tokenizer.tokenize(got)]) < 2e-2      # the oracle's arg-max up to bf16 near-ties
