"""Generate golden input/output vectors FROM THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/host_golden.json

The reference functions are executed unmodified from /root/reference under the
stub ``tensorflow`` of ``oracle/ref_loader.py``; python ``random`` (Mersenne
Twister, unchanged 3.10 -> 3.12) and numpy's legacy ``RandomState`` (unchanged
1.24 -> 2.3) make the seeded outputs identical to what the reference produces in
its own pinned environment.
"""
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "host_golden.json")


def _l(a):
    return np.asarray(a).astype(np.int64).tolist()


def gen_masking(ref):
    """dataloader_utils.py:186-261 (apply_dynamic_masking_task), :264-269."""
    cases = []
    rng = np.random.RandomState(1234)
    specs = [
        # (len, P, vocab, sel_rate, mask_rate, rand_rate)
        (50, 30, 12004, 0.15, 1.0, 0.0),
        (200, 40, 3709, 0.2, 1.0, 0.0),
        (37, 40, 500, 0.2, 0.8, 0.1),
        (5, 40, 100, 0.2, 0.8, 0.1),
        (1, 5, 50, 0.6, 1.0, 0.0),
        (120, 20, 1000, 0.4, 0.5, 0.5),
        (64, 3, 1000, 0.6, 0.0, 1.0),
        (30, 30, 200, 0.6, 0.0, 0.0),
    ]
    for ci, (n, P, V, sel, mr, rr) in enumerate(specs):
        for seed in (0, 1, 7 + ci):
            seq = rng.randint(3, V, size=n).astype(np.int64)
            if ci == 2:  # sprinkle special tokens that must never be selected
                seq[3] = 2
                seq[10] = 0
            ids, pos, lab = ref.dataloader_utils.apply_dynamic_masking_task(
                seq.copy(), P, 1, [2, 0], V, selection_rate=sel,
                mask_token_rate=mr, random_token_rate=rr, seed=seed)
            cases.append(dict(seq=_l(seq), P=P, mask_id=1, special=[2, 0], vocab=V,
                              selection_rate=sel, mask_token_rate=mr,
                              random_token_rate=rr, seed=seed,
                              out_ids=_l(ids), out_pos=_l(pos), out_lab=_l(lab)))
    last = []
    for n in (1, 2, 17, 200):
        seq = rng.randint(3, 999, size=n).astype(np.int64)
        ids, pos, lab = ref.dataloader_utils.mask_last_token_only(seq.copy(), 1)
        last.append(dict(seq=_l(seq), out_ids=_l(ids), out_pos=_l(pos), out_lab=_l(lab)))
    return dict(dynamic=cases, last=last)


def gen_popularity(ref):
    """dataloader_utils.py:14-18 + the reference test's fixed list (tests/.../dataloader_utils_tests.py:19-29)."""
    rng = np.random.RandomState(5)
    cases = []
    for n, hi in ((30, 7), (200, 20), (1000, 50)):
        items = rng.randint(0, hi, size=n).tolist()
        cases.append(dict(items=items, out=ref.dataloader_utils.rank_items_by_popularity(items)))
    return cases


def gen_samplers(ref):
    """random_sampler.py:63-79, popular_sampler.py:53-71, popular_random_sampler.py:77-126."""
    S = ref.samplers
    rng = np.random.RandomState(99)
    out = dict(random=[], popular=[], pop_random=[])
    vocab = list(range(3, 400))
    source = (rng.zipf(1.3, size=4000) % 397 + 3).tolist()
    # make sure every vocab item occurs at least once (non-zero probability for each)
    source += vocab
    for seed in (0, 3, 11):
        for without in (None, [5, 7, 9, 250, 250, 3], list(range(3, 120))):
            for size in (1, 10, 100):
                r = S.get("random", vocab=vocab, sample_size=size, seed=seed)
                out["random"].append(dict(vocab_lo=3, vocab_hi=400, size=size, seed=seed,
                                          without=without, out=r.sample(without=without)))
                pr = S.get("pop_random", source=source, vocab=vocab, sample_size=size, seed=seed)
                out["pop_random"].append(dict(size=size, seed=seed, without=without,
                                              out=pr.sample(without=without)))
    for size in (1, 10, 100):
        for without in (None, [5, 7, 9, 250, 3]):
            p = S.get("popular", source=source, sample_size=size)
            out["popular"].append(dict(size=size, without=without, out=p.sample(without=without)))
    # allow_duplicates variant of the random sampler (used by the reference's test fixture,
    # tests/test_utils.py:38-56)
    r = S.RandomSampler()
    np_out = r.sample(25, vocab=vocab, allow_duplicates=True, seed=4)
    out["random_dup"] = dict(size=25, seed=4, out=np_out)
    out["source"] = source
    # probability distribution (popular_random_sampler.py:119-126)
    pr = S.get("pop_random", source=source, vocab=vocab, sample_size=5, seed=0)
    out["prob_first16"] = [float(x) for x in pr.probability_distribution[:16]]
    return out


def gen_metrics(ref):
    """evaluation_metrics.py:47-112 on the reference tests' rank lists and on random ranks."""
    M = ref.metrics
    rng = np.random.RandomState(3)
    lists = [[1, 2, 3, 4, 5], [1, 5, 10, 15, 20], [2, 8, 4, 13, 20, 6, 3, 11, 2, 5],
             rng.randint(1, 102, size=1000).tolist(), rng.randint(1, 12, size=257).tolist()]
    cases = []
    for ranks in lists:
        ms = [M.Counter(name="Valid Ranks"), M.NDCG(1), M.NDCG(5), M.NDCG(10),
              M.HR(1), M.HR(5), M.HR(10), M.MAP()]
        for r in ranks:
            for m in ms:
                m.update(r)
        cases.append(dict(ranks=ranks, results={m.name: float(m.result()) for m in ms}))
    return cases


def gen_tokenizer(ref):
    """simple_tokenizer.py:119-138 -- first id is 0, ids increment in first-seen order."""
    t = ref.tokenizers.get("simple")
    words = ["[PAD]", "[MASK]", "[UNK]", "b", "a", "b", "zz", "a", "10", 10 and "10"]
    ids = [t.tokenize(w) for w in words]
    lst = t.tokenize(["q", "a", "r", "q"])
    return dict(words=words, ids=ids, list_in=["q", "a", "r", "q"], list_out=lst,
                vocab_size=t.get_vocab_size())


def gen_preprocessor(ref):
    """bert4rec_preprocessor.py:47-116 -- tensor layout contract (finetuning / no-mlm are deterministic)."""
    P = ref.preprocessor.BERT4RecPreprocessor
    tok = ref.tokenizers.get("simple")
    for w in ("[PAD]", "[MASK]", "[UNK]"):
        tok.tokenize(w)
    P.set_properties(tokenizer=tok, max_seq_len=12, max_predictions_per_seq=4,
                     mask_token_id=1, unk_token_id=2, pad_token_id=0,
                     masked_lm_rate=0.2, mask_token_rate=1.0, random_token_rate=0.0)
    cases = []
    seqs = [[f"i{j}" for j in range(5)], [f"i{j}" for j in range(12)], [f"i{j}" for j in range(30)]]
    for s in seqs:
        for apply_mlm, finetuning in ((True, True), (False, False), (False, True)):
            out = P.process_element(list(s), apply_mlm, finetuning)
            cases.append(dict(seq=s, apply_mlm=apply_mlm, finetuning=finetuning,
                              out={k: _l(v) for k, v in out.items()}))
    # training mode: structural contract only (the reference reseeds from entropy), record sizes
    out = P.process_element(seqs[1], True, False)
    cases.append(dict(seq=seqs[1], apply_mlm=True, finetuning=False, structural=True,
                      out_shapes={k: list(np.asarray(v).shape) for k, v in out.items()}))
    return dict(max_seq_len=12, max_predictions_per_seq=4, cases=cases)


def main():
    ref = load_reference()
    golden = dict(
        _generated_by="oracle/gen_golden.py from /root/reference (maneymarkus/BERT4Rec) under a tensorflow stub",
        masking=gen_masking(ref),
        popularity=gen_popularity(ref),
        samplers=gen_samplers(ref),
        metrics=gen_metrics(ref),
        tokenizer=gen_tokenizer(ref),
        preprocessor=gen_preprocessor(ref),
    )
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(golden, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
