"""CPU suite: oracle AND product host-side code against the golden vectors generated from the reference's own
functions (tests/golden/host_golden.json, oracle/gen_golden.py) and the reference's known-answer tests."""
import numpy as np
import pytest

from oracle import host_ops as ho
from bert4rec_b200.dataloaders import dataloader_utils as du, samplers
from bert4rec_b200.dataloaders.preprocessors import BERT4RecPreprocessor
from bert4rec_b200 import tokenizers, evaluation
from bert4rec_b200.evaluation import evaluation_metrics as em


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_cloze_masking_bit_exact(golden, impl):
    fn = ho.cloze_mask if impl == "oracle" else du.apply_dynamic_masking_task
    for c in golden["masking"]["dynamic"]:
        ids, pos, lab = fn(np.array(c["seq"], dtype=np.int64), c["P"], c["mask_id"], c["special"], c["vocab"],
                           c["selection_rate"], c["mask_token_rate"], c["random_token_rate"], c["seed"])
        assert ids.tolist() == c["out_ids"] and pos.tolist() == c["out_pos"] and lab.tolist() == c["out_lab"]
        assert ids.dtype == np.int64 and pos.dtype == np.int64


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_mask_last_token(golden, impl):
    fn = ho.mask_last if impl == "oracle" else du.mask_last_token_only
    for c in golden["masking"]["last"]:
        ids, pos, lab = fn(np.array(c["seq"], dtype=np.int64), 1)
        assert ids.tolist() == c["out_ids"] and pos.tolist() == c["out_pos"] and list(lab) == c["out_lab"]


def test_masking_invariants_like_reference_tests():
    """Structural properties asserted by the reference (dataloader_utils_tests.py:180-248)."""
    rng = np.random.RandomState(0)
    for seed in range(20):
        seq = rng.randint(3, 500, size=rng.randint(5, 120)).astype(np.int64)
        ids, pos, lab = du.apply_dynamic_masking_task(seq, 25, 1, [2, 0], 500, 0.2, 0.8, 0.1, seed=seed)
        assert len(ids) == len(seq) and 0 < len(pos) <= 25 and len(pos) == len(lab)
        assert set(lab.tolist()) <= set(seq.tolist()) and pos.min() >= 0 and pos.max() < len(seq)
        assert np.array_equal(seq[pos], lab) and list(pos) == sorted(pos)


def test_popularity_ranking(golden):
    for c in golden["popularity"]:
        assert du.rank_items_by_popularity(c["items"]) == c["out"] == ho.popularity_order(c["items"])
    # the reference test's fixed list (dataloader_utils_tests.py:19-29): most popular 6, least popular 0
    items = [6, 6, 6, 6, 3, 3, 3, 5, 5, 0]
    r = du.rank_items_by_popularity(items)
    assert r[0] == 6 and r[-1] == 0


def test_samplers_bit_exact(golden):
    s = golden["samplers"]
    src, vocab = s["source"], list(range(3, 400))
    for c in s["random"]:
        assert samplers.get("random", vocab=vocab, sample_size=c["size"], seed=c["seed"]).sample(without=c["without"]) == c["out"]
        assert ho.sample_random(vocab, c["size"], c["seed"], c["without"]) == c["out"]
    for c in s["pop_random"]:
        for ident in ("pop_random", "popular_random"):
            sm = samplers.get(ident, source=src, vocab=vocab, sample_size=c["size"], seed=c["seed"])
            assert sm.sample(without=c["without"]) == c["out"]
        assert ho.sample_pop_random(src, vocab, c["size"], c["seed"], c["without"]) == c["out"]
    for c in s["popular"]:
        assert samplers.get("popular", source=src, sample_size=c["size"]).sample(without=c["without"]) == c["out"]
        assert ho.sample_popular(src, c["size"], c["without"]) == c["out"]
    d = s["random_dup"]
    assert samplers.RandomSampler().sample(d["size"], vocab=vocab, allow_duplicates=True, seed=d["seed"]) == d["out"]
    pr = samplers.get("pop_random", source=src, vocab=vocab, sample_size=5, seed=0)
    assert pr.probability_distribution[:16] == s["prob_first16"] == ho.popularity_probabilities(src, vocab)[:16]


def test_sampler_contract_errors():
    with pytest.raises(ValueError):
        samplers.get("nope")
    with pytest.raises(ValueError):
        samplers.RandomSampler(sample_size=-1)
    with pytest.raises(ValueError):
        samplers.RandomSampler(vocab=[1, 2, 3]).sample()             # no sample size
    with pytest.raises(ValueError):
        samplers.RandomSampler(vocab=[1, 2, 3], sample_size=5).sample()  # more than the vocab, no duplicates
    with pytest.raises(ValueError):
        samplers.PopularSampler(sample_size=3).sample()              # no source
    with pytest.raises(ValueError):
        samplers.PopularRandomSampler(sample_size=3).sample()        # no source / vocab
    inst = samplers.RandomSampler(vocab=[1, 2, 3], sample_size=2)
    assert samplers.get(inst) is inst
    a = samplers.RandomSampler(vocab=list(range(100)), sample_size=10, seed=1).sample()
    b = samplers.RandomSampler(vocab=list(range(100)), sample_size=10, seed=2).sample()
    assert a != b and all(x not in [5, 6] for x in samplers.RandomSampler(vocab=list(range(10)), sample_size=8, seed=0).sample(without=[5, 6]))


def test_metrics_known_answers_from_reference_tests():
    """evaluation_metrics_tests.py:12-103 of the reference."""
    r1, r2, r3 = [1, 2, 3, 4, 5], [1, 5, 10, 15, 20], [2, 8, 4, 13, 20, 6, 3, 11, 2, 5]

    def run(metric, ranks):
        metric.reset()
        for r in ranks:
            metric.update(r)
        return metric.result()

    assert [run(em.HR(k), r1) for k in (1, 5, 10)] == [0.2, 1, 1]
    assert [run(em.HR(k), r2) for k in (1, 5, 10)] == [0.2, 0.4, 0.6]
    assert [run(em.HR(k), r3) for k in (1, 5, 10)] == [0, 0.5, 0.7]
    assert [round(run(em.NDCG(k), r1), 2) for k in (1, 5, 10)] == [0.2, 0.59, 0.59]
    assert [round(run(em.NDCG(k), r2), 2) for k in (1, 5, 10)] == [0.2, 0.28, 0.34]
    assert [round(run(em.NDCG(k), r3), 2) for k in (1, 5, 10)] == [0, 0.26, 0.33]
    assert [round(run(em.MAP(), r), 2) for r in (r1, r2, r3)] == [0.46, 0.28, 0.23]
    assert [run(em.Counter(), r) for r in (r1, r2, r3)] == [5, 5, 10]
    assert em.HR(5).name == "HR@5" and em.NDCG(10).name == "NDCG@10" and em.MAP().name == "MAP"


def test_metrics_bit_exact_vs_reference_classes(golden):
    for c in golden["metrics"]:
        acc = ho.MetricAccumulator()
        for r in c["ranks"]:
            acc.update(r)
        assert acc.results() == c["results"]
        for vectorised in (False, True):
            ms = evaluation.default_bert4rec_metrics()
            for m in ms:
                if vectorised:
                    m.update_many(np.array(c["ranks"]))
                else:
                    for r in c["ranks"]:
                        m.update(r)
            assert {m.name: float(m.result()) for m in ms} == c["results"]
            for m in ms:
                m.reset()
                assert m.result() == 0


def test_tokenizer_ids(golden):
    g = golden["tokenizer"]
    t = tokenizers.get("simple")
    assert [t.tokenize(w) for w in g["words"]] == g["ids"]          # first id is 0: PAD=0, MASK=1, UNK=2
    assert t.tokenize(g["list_in"]) == g["list_out"] and t.get_vocab_size() == g["vocab_size"]
    assert t.detokenize(0) == "[PAD]" and t.detokenize([1, 2]) == ["[MASK]", "[UNK]"]
    t.disable_extensibility()
    with pytest.raises(RuntimeError):
        t.tokenize("never seen")
    with pytest.raises(ValueError):
        tokenizers.get("nope")
    assert tokenizers.get(t) is t


def test_tokenizer_vocab_file_roundtrip(tmp_path):
    t = tokenizers.get("simple")
    t.tokenize(["[PAD]", "[MASK]", "[UNK]", "x", "y"])
    f = tmp_path / "vocab.txt"
    t.export_vocab_to_file(f)
    assert f.read_text().splitlines()[3] == "x|3"
    t2 = tokenizers.get("simple", vocab_file_path=f)
    assert t2.get_vocab() == t.get_vocab() and t2.tokenize("y") == 4


def test_preprocessor_layout(golden):
    g = golden["preprocessor"]
    tok = tokenizers.get("simple")
    tok.tokenize(["[PAD]", "[MASK]", "[UNK]"])
    BERT4RecPreprocessor.set_properties(tokenizer=tok, max_seq_len=g["max_seq_len"],
                                        max_predictions_per_seq=g["max_predictions_per_seq"], mask_token_id=1,
                                        unk_token_id=2, pad_token_id=0, masked_lm_rate=0.2, mask_token_rate=1.0,
                                        random_token_rate=0.0)
    for c in g["cases"]:
        out = BERT4RecPreprocessor.process_element(list(c["seq"]), c["apply_mlm"], c["finetuning"])
        if len(c["seq"]) > g["max_seq_len"] and not c["finetuning"]:
            # random window start (python random, unseeded in the reference): contract = a contiguous window
            ids = np.asarray(out["input_word_ids"]).tolist()
            assert len(ids) == g["max_seq_len"] and ids == list(range(ids[0], ids[0] + g["max_seq_len"]))
            continue
        if c.get("structural"):
            assert {k: list(np.asarray(v).shape) for k, v in out.items()} == c["out_shapes"]
            assert 1 in out["input_word_ids"] and all(v.dtype == np.int64 for v in out.values())
        else:
            assert {k: np.asarray(v).tolist() for k, v in out.items()} == c["out"]
            toks = tok.tokenize(list(c["seq"]))
            ora = ho.layout_element(toks, g["max_seq_len"], g["max_predictions_per_seq"], c["apply_mlm"], c["finetuning"])
            assert {k: np.asarray(v).tolist() for k, v in ora.items()} == c["out"]
    inf = BERT4RecPreprocessor.prepare_inference([f"i{j}" for j in range(5)])
    assert tuple(inf["input_word_ids"].shape) == (1, 12) and int(inf["input_word_ids"][0, 5]) == 1
    assert int(inf["masked_lm_weights"].sum()) == 1 and int(inf["masked_lm_positions"][0, 0]) == 5
    with pytest.raises(ValueError):
        BERT4RecPreprocessor.prepare_inference("not a list")


def test_make_batches_layout():
    tok = tokenizers.get("simple")
    tok.tokenize(["[PAD]", "[MASK]", "[UNK]"])
    BERT4RecPreprocessor.set_properties(tokenizer=tok, max_seq_len=8, max_predictions_per_seq=3, mask_token_id=1,
                                        unk_token_id=2, pad_token_id=0, masked_lm_rate=0.3, mask_token_rate=1.0,
                                        random_token_rate=0.0)
    els = BERT4RecPreprocessor.process_dataset([[f"a{i}", f"b{i}", "c", "d", "e"] for i in range(10)], True, False)
    ds = du.make_batches(els, batch_size=4, seed=3)
    assert ds.cardinality() == 3 and [b["input_word_ids"].shape[0] for b in ds] == [4, 4, 2]   # last batch is partial
    b0 = ds[0]
    assert set(b0) == {"labels", "input_word_ids", "input_mask", "masked_lm_ids", "masked_lm_positions", "masked_lm_weights"}
    assert tuple(b0["input_word_ids"].shape) == (4, 8) and tuple(b0["masked_lm_ids"].shape) == (4, 3)
    assert all(v.dtype.is_floating_point is False for v in b0.values())
