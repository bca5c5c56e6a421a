"""CPU suite: the C++ batch host data path (b4r_host_*, csrc/host_data.cu) against golden vectors generated from the
reference's own functions (tests/golden/host_batch_golden.json <- oracle/gen_golden_batch.py, tests/golden/host_golden.json)
and against the per-sequence Python product code on random cases.  Integer work: bit-exact."""
import json
import os

import numpy as np
import pytest

from bert4rec_b200.dataloaders import host_native as hn, dataloader_utils as du, samplers

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def batch_golden():
    with open(os.path.join(HERE, "golden", "host_batch_golden.json")) as f:
        return json.load(f)


def _check_masking(out, i, seq, ids, pos, lab, S, P):
    n, k = len(seq), len(pos)
    assert out["input_word_ids"][i, :n].tolist() == ids and not out["input_word_ids"][i, n:].any()
    assert out["labels"][i, :n].tolist() == seq and not out["labels"][i, n:].any()
    assert out["input_mask"][i].tolist() == [1] * n + [0] * (S - n)
    assert out["masked_lm_positions"][i, :k].tolist() == pos and not out["masked_lm_positions"][i, k:].any()
    assert out["masked_lm_ids"][i, :k].tolist() == lab and not out["masked_lm_ids"][i, k:].any()
    assert out["masked_lm_weights"][i].tolist() == [1] * k + [0] * (P - k)


@pytest.mark.parametrize("threads", [1, 4])
def test_cloze_mask_batch_vs_reference_golden(batch_golden, threads):
    for g in batch_golden["masking"]:
        out = hn.cloze_mask_batch([np.array(s, dtype=np.int64) for s in g["seqs"]], g["S"], g["P"], g["mask_id"], g["special"],
                                  g["vocab"], g["selection_rate"], g["mask_token_rate"], g["random_token_rate"],
                                  seeds=np.array(g["seeds"], dtype=np.uint64), n_threads=threads)
        for i, (seq, o) in enumerate(zip(g["seqs"], g["outs"])):
            _check_masking(out, i, seq, o["ids"], o["pos"], o["lab"], g["S"], g["P"])


def test_cloze_mask_batch_vs_per_call_golden(golden):
    for c in golden["masking"]["dynamic"]:
        S = max(len(c["seq"]), 1)
        out = hn.cloze_mask_batch([np.array(c["seq"], dtype=np.int64)], S, c["P"], c["mask_id"], c["special"], c["vocab"],
                                  c["selection_rate"], c["mask_token_rate"], c["random_token_rate"],
                                  seeds=np.array([c["seed"]], dtype=np.uint64))
        _check_masking(out, 0, c["seq"], c["out_ids"], c["out_pos"], c["out_lab"], S, c["P"])


def test_cloze_mask_batch_random_cases_vs_python_product():
    rng = np.random.RandomState(0)
    V, S, P = 997, 64, 20
    seqs = [rng.randint(0, V, size=rng.randint(0, S + 1)).astype(np.int64) for _ in range(400)]   # ids 0 / 2 are special
    seeds = (rng.randint(0, 2**31, size=len(seqs)).astype(np.uint64) << np.uint64(rng.randint(0, 33)))
    for rates in ((0.15, 1.0, 0.0), (0.2, 0.8, 0.1), (0.9, 0.3, 0.3)):
        out = hn.cloze_mask_batch(seqs, S, P, 1, [2, 0, 2], V, *rates, seeds=seeds)
        for i, s in enumerate(seqs):
            m, pos, ids = du.apply_dynamic_masking_task(s, P, 1, [2, 0, 2], V, *rates, seed=int(seeds[i]))
            _check_masking(out, i, s.tolist(), m.tolist(), pos.tolist(), ids.tolist(), S, P)


def test_cloze_mask_batch_errors():
    with pytest.raises(ValueError, match="max_seq_len"):
        hn.cloze_mask_batch([np.arange(3, 20)], 8, 4, 1, [2, 0], 100, seeds=np.array([1], dtype=np.uint64))
    with pytest.raises(ValueError):   # random.choice on an empty selectable vocabulary
        hn.cloze_mask_batch([np.array([5, 6, 7])], 8, 4, 1, [0, 1, 2], 3, 1.0, 0.0, 1.0, seeds=np.array([1], dtype=np.uint64))
    out = hn.cloze_mask_batch([], 8, 4, 1, [2, 0], 100, seeds=np.zeros(0, dtype=np.uint64))
    assert out["labels"].shape == (0, 8) and out["masked_lm_ids"].shape == (0, 4)


def _w(withouts):
    return [[] if w is None else w for w in withouts]


def test_samplers_batch_vs_reference_golden(batch_golden):
    s = batch_golden["samplers"]
    vocab, withouts = s["vocab"], s["withouts"]
    for c in s["random"]:
        got = hn.sample_random_batch(vocab, _w(withouts), c["size"], False, c["seed"])
        assert got.tolist() == c["outs"]
    for c in s["random_dup"]:
        got = hn.sample_random_batch(vocab, _w(withouts), c["size"], True, c["seed"], n_threads=3)
        assert got.tolist() == c["outs"]
    pr = samplers.get("pop_random", source=s["source"], vocab=vocab, sample_size=5, seed=0)
    probs = pr.probability_distribution
    assert float(np.sum(np.asarray(probs) * np.arange(len(probs)))) == s["probs_checksum"]
    for key, dup in (("pop_random", False), ("pop_random_dup", True)):
        for c in s[key]:
            got, lens = hn.sample_pop_random_batch(vocab, probs, _w(withouts), c["size"], dup, c["seed"])
            assert [got[i, :lens[i]].tolist() for i in range(len(withouts))] == c["outs"]
    ranked = du.rank_items_by_popularity(s["source"])
    for c in s["popular"]:
        got, lens = hn.sample_popular_batch(ranked, _w(withouts), c["size"])
        assert [got[i, :lens[i]].tolist() for i in range(len(withouts))] == c["outs"]


def test_samplers_batch_vs_per_call_golden(golden):
    s = golden["samplers"]
    src, vocab = s["source"], list(range(3, 400))
    for c in s["random"]:
        assert hn.sample_random_batch(vocab, _w([c["without"]]), c["size"], False, c["seed"])[0].tolist() == c["out"]
    probs = samplers.get("pop_random", source=src, vocab=vocab, sample_size=5, seed=0).probability_distribution
    for c in s["pop_random"]:
        got, lens = hn.sample_pop_random_batch(vocab, probs, _w([c["without"]]), c["size"], False, c["seed"])
        assert got[0, :lens[0]].tolist() == c["out"]
    ranked = du.rank_items_by_popularity(src)
    for c in s["popular"]:
        got, lens = hn.sample_popular_batch(ranked, _w([c["without"]]), c["size"])
        assert got[0, :lens[0]].tolist() == c["out"]
    d = s["random_dup"]
    assert hn.sample_random_batch(vocab, [[]], d["size"], True, d["seed"])[0].tolist() == d["out"]


def test_sampler_batch_methods_match_per_call_product():
    rng = np.random.RandomState(3)
    vocab = list(range(3, 700))
    source = (rng.zipf(1.3, size=3000) % 697 + 3).tolist() + vocab
    withouts = [rng.randint(3, 700, size=rng.randint(0, 80)).tolist() for _ in range(40)]
    for ident, kw in (("random", dict(vocab=vocab, seed=4)), ("popular", dict(source=source)),
                      ("pop_random", dict(source=source, vocab=vocab, seed=4))):
        sm = samplers.get(ident, sample_size=50, **kw)
        got = sm.sample_batch(withouts)
        assert [list(map(int, g)) for g in got] == [sm.sample(without=w) for w in withouts]


def test_sampler_batch_errors_mirror_python():
    with pytest.raises(ValueError, match="larger sample than population"):
        hn.sample_random_batch([3, 4, 5, 6], [[3, 4]], 3, False, 0)          # pool shrinks below the sample size
    with pytest.raises(ValueError, match="can not be greater"):
        hn.sample_random_batch([3, 4, 5], [[]], 5, False, 0)
    with pytest.raises(ValueError, match="do not sum to 1"):
        hn.sample_pop_random_batch([3, 4, 5], [0.5, 0.2, 0.2], [[]], 1, False, 0)
    with pytest.raises(ValueError, match="Fewer non-zero"):
        hn.sample_pop_random_batch([3, 4, 5], [1.0, 0.0, 0.0], [[]], 2, False, 0)
    with pytest.raises(ValueError, match="Seed must be between"):
        hn.sample_random_batch([3, 4, 5], [[]], 1, False, 2**32)


def test_preprocessor_process_batch_matches_process_element():
    from bert4rec_b200 import tokenizers
    from bert4rec_b200.dataloaders.preprocessors import BERT4RecPreprocessor as P
    tok = tokenizers.get("simple")
    tok.tokenize(["[PAD]", "[MASK]", "[UNK]"] + [f"i{j}" for j in range(300)])
    P.set_properties(tokenizer=tok, max_seq_len=20, max_predictions_per_seq=5, mask_token_id=1, unk_token_id=2, pad_token_id=0,
                     masked_lm_rate=0.3, mask_token_rate=0.8, random_token_rate=0.1)
    rng = np.random.RandomState(1)
    seqs = [[f"i{j}" for j in rng.randint(0, 300, size=rng.randint(1, 21))] for _ in range(50)]
    seeds = rng.randint(0, 2**31, size=len(seqs)).astype(np.uint64)
    got = P.process_batch(seqs, True, False, seeds=seeds)
    for i, s in enumerate(seqs):
        ref = P.process_element(list(s), True, False, seed=int(seeds[i]))
        for k, v in ref.items():
            assert got[k][i].tolist() == np.asarray(v).tolist(), (i, k)
    for mode in ((True, True), (False, False)):
        got = P.process_batch(seqs, *mode)
        for i, s in enumerate(seqs):
            ref = P.process_element(list(s), *mode)
            assert set(got) == set(ref) and all(got[k][i].tolist() == np.asarray(ref[k]).tolist() for k in ref)


def test_evaluator_candidates_batched_equal_per_slot_calls():
    """BERT4RecEvaluator.build_candidates (one native sample_batch call) == the reference's per-slot loop
    (bert4rec_evaluator.py:98-104) for a seeded sampler."""
    from bert4rec_b200 import evaluation
    rng = np.random.RandomState(2)
    V, B, S, P = 500, 12, 16, 3
    vocab = list(range(3, V))
    labels = rng.randint(3, V, size=(B, S)); labels[:, 12:] = 0
    w = np.zeros((B, P), dtype=np.int64); w[:, 0] = 1; w[3, 1] = 1
    ids = rng.randint(3, V, size=(B, P)) * w
    batch = {"labels": labels, "masked_lm_ids": ids, "masked_lm_weights": w}
    for sm in (samplers.get("random", vocab=vocab, sample_size=100, seed=7),
               samplers.get("pop_random", source=labels[labels > 0].tolist() + vocab, vocab=vocab, sample_size=100, seed=7)):
        ev = evaluation.get("bert4rec", sampler=sm)
        cands, gts = ev.build_candidates(batch)
        ref = []
        for b in range(B):
            for p in np.nonzero(w[b])[0]:
                gt = int(ids[b, p])
                ref.append(sm.sample(without=labels[b].tolist() + [gt]) + [gt])
        assert cands == ref and list(gts) == [r[-1] for r in ref]
        arr, g2 = ev.build_candidates(batch, as_array=True)
        assert np.asarray(arr).tolist() == ref


def test_random_sampler_batch_sparse_ids_use_the_hash_path():
    """Ids outside the small non-negative range (negative / huge) take the hash-set exclusion: same draws as numpy."""
    vocab = [-5, 10**12, 7, 3, 2**40, 11, -1, 99]
    withouts = [[], [7, -1], [10**12, 3, 3]]
    for seed in (0, 8):
        got = hn.sample_random_batch(vocab, withouts, 3, False, seed)
        for i, w in enumerate(withouts):
            np.random.seed(seed)
            assert got[i].tolist() == np.random.choice([v for v in vocab if v not in w], size=3, replace=False).tolist()


def test_sampler_batch_per_request_seeds_take_the_general_path():
    """Different seeds per request (what seed=None means: fresh entropy per call) cannot share a shuffle: every request does its
    own full legacy permutation / weighted draw -- still the numpy streams, checked against numpy."""
    rng = np.random.RandomState(5)
    vocab = list(range(3, 900)); rng.shuffle(vocab)
    src = (rng.zipf(1.3, size=4000) % 897 + 3).tolist() + vocab
    p = [src.count(v) / len(src) for v in vocab]
    withouts = [rng.randint(3, 900, size=rng.randint(0, 60)).tolist() for _ in range(12)]
    seeds = rng.randint(0, 2**32 - 1, size=len(withouts)).astype(np.uint32)
    got = hn.sample_random_batch(vocab, withouts, 50, False, seeds)
    gp, lens = hn.sample_pop_random_batch(vocab, p, withouts, 50, False, seeds)
    for i, w in enumerate(withouts):
        np.random.seed(int(seeds[i]))
        ex = set(w)
        assert got[i].tolist() == np.random.choice([v for v in vocab if v not in ex], size=50, replace=False).tolist()
        np.random.seed(int(seeds[i]))
        drawn = np.random.choice(vocab, 50 + len(ex), False, p).tolist()
        assert gp[i, :lens[i]].tolist() == [v for v in drawn if v not in ex][:50]
    # a vocabulary with a repeated id is not eligible for the shared-shuffle path either
    dup_vocab = vocab + [vocab[0]]
    got = hn.sample_random_batch(dup_vocab, withouts, 20, False, 9)
    for i, w in enumerate(withouts):
        np.random.seed(9)
        assert got[i].tolist() == np.random.choice([v for v in dup_vocab if v not in set(w)], size=20, replace=False).tolist()


def test_host_path_is_safe_under_concurrent_callers_and_after_fork():
    """The worker pool serialises jobs of concurrent callers, and a fork()ed child gets its own workers."""
    import multiprocessing as mp
    import threading
    rng = np.random.RandomState(4)
    V, S, P, n = 997, 32, 8, 600
    seqs = [rng.randint(3, V, size=rng.randint(1, S + 1)).astype(np.int64) for _ in range(n)]
    seeds = np.arange(n, dtype=np.uint64)
    ref = hn.cloze_mask_batch(seqs, S, P, 1, [2, 0], V, 0.3, 0.8, 0.1, seeds=seeds, n_threads=1)
    outs = [None] * 4

    def work(i):
        for _ in range(5):
            outs[i] = hn.cloze_mask_batch(seqs, S, P, 1, [2, 0], V, 0.3, 0.8, 0.1, seeds=seeds, n_threads=0)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not any(t.is_alive() for t in ts)
    for o in outs:
        assert all(np.array_equal(o[k], ref[k]) for k in ref)
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def child():
        o = hn.cloze_mask_batch(seqs, S, P, 1, [2, 0], V, 0.3, 0.8, 0.1, seeds=seeds, n_threads=0)
        q.put(all(np.array_equal(o[k], ref[k]) for k in ref))

    p = ctx.Process(target=child)
    p.start()
    assert q.get(timeout=60) is True
    p.join(30)
    assert p.exitcode == 0
