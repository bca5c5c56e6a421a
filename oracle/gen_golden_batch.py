"""Golden vectors for the BATCH host data path (b4r_host_*), FROM THE REFERENCE'S OWN FUNCTIONS.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden_batch.py      # writes tests/golden/host_batch_golden.json

Same mechanism as oracle/gen_golden.py (reference files executed unmodified under the tensorflow stub).  Covers what the
per-call fixture does not: seeds beyond 32 bits (multi-word init_by_array), empty / all-special / full-length sequences,
the rate grid of the shipped configs, larger catalogues for the samplers, duplicate draws.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "host_batch_golden.json")


def _l(a):
    return np.asarray(a).astype(np.int64).tolist()


def gen_masking(ref):
    rng = np.random.RandomState(4321)
    groups = []
    for (S, P, V, sel, mr, rr) in ((50, 30, 12004, 0.15, 1.0, 0.0), (200, 40, 3709, 0.2, 0.8, 0.1), (24, 6, 97, 0.6, 0.5, 0.3)):
        seqs, seeds, outs = [], [], []
        lens = [0, 1, 2, S, S, S - 1] + rng.randint(1, S + 1, size=34).tolist()
        for i, n in enumerate(lens):
            seq = rng.randint(3, V, size=n).astype(np.int64)
            if i % 7 == 3 and n >= 4:
                seq[1] = 2; seq[n - 1] = 0      # special tokens inside the sequence
            if i == 8 and n:
                seq[:] = 2                      # nothing selectable
            seed = [0, 1, 2**32 - 1, 2**32, 2**32 + 12345, 2**63 - 1][i] if i < 6 else int(rng.randint(0, 2**31)) * int(rng.randint(1, 2**31))
            ids, pos, lab = ref.dataloader_utils.apply_dynamic_masking_task(
                seq.copy(), P, 1, [2, 0], V, selection_rate=sel, mask_token_rate=mr, random_token_rate=rr, seed=seed)
            seqs.append(_l(seq)); seeds.append(seed); outs.append(dict(ids=_l(ids), pos=_l(pos), lab=_l(lab)))
        groups.append(dict(S=S, P=P, vocab=V, selection_rate=sel, mask_token_rate=mr, random_token_rate=rr, mask_id=1, special=[2, 0],
                           seqs=seqs, seeds=seeds, outs=outs))
    return groups


def gen_samplers(ref):
    S = ref.samplers
    rng = np.random.RandomState(77)
    V = 3000
    vocab = list(range(3, V))
    rng.shuffle(vocab)                                # stored order matters: the pool keeps it
    source = (rng.zipf(1.2, size=6000) % (V - 3) + 3).tolist() + vocab
    withouts = [None, [], [5, 5, 9]] + [rng.randint(3, V, size=int(k)).tolist() for k in (1, 20, 51, 200, 700)]
    out = dict(vocab=vocab, source=source, withouts=withouts, random=[], random_dup=[], pop_random=[], pop_random_dup=[], popular=[])
    for seed in (0, 5, 2**32 - 1):
        for size in (1, 100):
            out["random"].append(dict(seed=seed, size=size, outs=[S.get("random", vocab=vocab, sample_size=size, seed=seed)
                                                                   .sample(without=w) for w in withouts]))
            out["random_dup"].append(dict(seed=seed, size=size,
                                          outs=[S.RandomSampler(vocab=vocab, sample_size=size, seed=seed, allow_duplicates=True)
                                                .sample(without=w) for w in withouts]))
    pr = S.get("pop_random", source=source, vocab=vocab, sample_size=100, seed=0)
    pr.sample()                                        # builds the probability distribution once (O(V * |source|))
    probs = [float(x) for x in pr.probability_distribution]
    for seed in (0, 9):
        for size in (1, 100):
            for dup, key in ((False, "pop_random"), (True, "pop_random_dup")):
                s = S.PopularRandomSampler(source=source, vocab=vocab, sample_size=size, seed=seed, allow_duplicates=dup)
                s.probability_distribution = probs
                out[key].append(dict(seed=seed, size=size, outs=[s.sample(without=w) for w in withouts]))
    p = S.get("popular", source=source, sample_size=100)
    out["popular"] = [dict(size=size, outs=[p.sample(sample_size=size, without=w) for w in withouts]) for size in (1, 100)]
    out["probs_checksum"] = float(np.sum(np.asarray(probs) * np.arange(len(probs))))
    return out


def main():
    ref = load_reference()
    golden = dict(_generated_by="oracle/gen_golden_batch.py from /root/reference (maneymarkus/BERT4Rec) under a tensorflow stub",
                  masking=gen_masking(ref), samplers=gen_samplers(ref))
    with open(OUT, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
