"""Oracle restatement of the reference's integer / host-side algorithms.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain Python + numpy, written
for clarity, not speed.  Every function cites the reference lines it follows and
is pinned against tests/golden/host_golden.json (generated from the reference's
own code by oracle/gen_golden.py).
"""
import collections
import math
import random

import numpy as np

PAD_ID, MASK_ID, UNK_ID = 0, 1, 2  # bert4rec_dataloader.py:35-43 with simple_tokenizer.py:119-138


# ----------------------------------------------------------------------------- masking
def cloze_mask(sequence, max_selections, mask_id, special_ids, vocab_size,
               selection_rate=0.2, mask_token_rate=0.8, random_token_rate=0.1, seed=None):
    """Cloze masking; follows dataloader_utils.py:186-261.

    RNG consumption order (python ``random``): seed -> shuffle(range(n_non_special))
    -> per selected index: random(), then choice(selectable_vocab) iff
    rn < mask_rate + random_rate (drawn even when MASK then overrides it, :249-255).
    NB (reference quirk kept): candidate indexes are 0..n_non_special-1, i.e.
    positions in the sequence *with specials removed*, but are applied to the
    original sequence (:230-233,:258-260).
    """
    seq = np.asarray(sequence)
    random.seed(seed)
    n_plain = int(np.count_nonzero(~np.isin(seq, special_ids)))
    n_pred = min(max_selections, max(1, int(n_plain * selection_rate)))
    selectable = [v for v in range(vocab_size) if v not in special_ids]
    order = list(range(n_plain))
    random.shuffle(order)
    chosen = sorted(order[:n_pred])
    out = seq.copy()
    positions, labels = [], []
    for idx in chosen:
        if len(labels) >= n_pred:
            break
        token = seq[idx]
        rn = random.random()
        if rn < mask_token_rate + random_token_rate:
            token = random.choice(selectable)
        if rn < mask_token_rate:
            token = mask_id
        out[idx] = token
        labels.append(seq[idx])
        positions.append(idx)
    return (out, np.asarray(positions, dtype=seq.dtype), np.asarray(labels, dtype=seq.dtype))


def mask_last(sequence, mask_id):
    """dataloader_utils.py:264-269."""
    seq = np.array(sequence, dtype=np.int64)
    label = np.array([seq[-1]], dtype=np.int64)
    seq[-1] = mask_id
    return seq, np.array([len(seq) - 1], dtype=np.int64), label


def layout_element(tokens, max_seq_len, max_pred, apply_mlm, finetuning,
                   masked_lm_rate=0.2, mask_token_rate=1.0, random_token_rate=0.0,
                   vocab_size=None, seed=None):
    """Model-input layout of one tokenised sequence; follows bert4rec_preprocessor.py:47-116.

    Only the deterministic branches are restated exactly (finetuning / len <= max);
    the random-window branch (:64-67) consumes ``random.randint`` before masking.
    """
    tokens = list(tokens)
    if finetuning or len(tokens) <= max_seq_len:
        seg = tokens[-max_seq_len:]
    else:
        start = random.randint(0, len(tokens) - max_seq_len)
        seg = tokens[start:start + max_seq_len]
    ids = np.array(seg, dtype=np.int64)
    mask = np.ones_like(ids)
    labels = ids.copy()
    out = {}
    if apply_mlm:
        if finetuning:
            ids, pos, lab = mask_last(ids, MASK_ID)
        else:
            ids, pos, lab = cloze_mask(ids, max_pred, MASK_ID, [UNK_ID, PAD_ID], vocab_size,
                                       masked_lm_rate, mask_token_rate, random_token_rate, seed)
        w = np.ones_like(lab)
        k = max_pred - lab.shape[0]
        if k > 0:
            lab, pos, w = (np.pad(a, (0, k)) for a in (lab, pos, w))
        out.update(masked_lm_ids=lab, masked_lm_positions=pos, masked_lm_weights=w)
    k = max_seq_len - ids.shape[0]
    if k > 0:
        ids, mask, labels = (np.pad(a, (0, k)) for a in (ids, mask, labels))
    out.update(labels=labels, input_word_ids=ids, input_mask=mask)
    return out


# ----------------------------------------------------------------------------- samplers
def popularity_order(items):
    """dataloader_utils.py:14-18: stable sort by count (desc), first-seen order on ties, de-duplicated."""
    cnt = collections.Counter(items)
    ordered = sorted(items, key=cnt.get, reverse=True)
    return list(dict.fromkeys(ordered))


def sample_random(vocab, size, seed=None, without=None, allow_duplicates=False):
    """random_sampler.py:48-50,63-79: reseed numpy's global RandomState on every call, filter, choice."""
    np.random.seed(seed)
    pool = list(vocab)
    if without is not None:
        pool = [v for v in pool if v not in without]
    return np.random.choice(pool, size=size, replace=allow_duplicates).tolist()


def sample_popular(source, size, without=None):
    """popular_sampler.py:19-22,53-71 (source pre-ranked at construction)."""
    ranked = popularity_order(source)
    if without is not None:
        ranked = [v for v in ranked if v not in without]
    return ranked[:size]


def popularity_probabilities(source, vocab):
    """popular_random_sampler.py:119-126."""
    cnt = collections.Counter(source)
    n = len(source)
    return [cnt.get(v, 0) / n for v in vocab]


def sample_pop_random(source, vocab, size, seed=None, without=None, allow_duplicates=False, probs=None):
    """popular_random_sampler.py:53-55,77-117: draw size+|set(without)| ids with popularity
    probabilities, drop the excluded ones, keep the first ``size``."""
    np.random.seed(seed)
    if probs is None:
        probs = popularity_probabilities(source, vocab)
    n = size
    if without is not None:
        without = list(set(without))
        n += len(without)
    if not allow_duplicates and n > len(vocab):
        raise ValueError("without list reduces the vocab too much")
    drawn = np.random.choice(vocab, n, allow_duplicates, probs).tolist()
    if without is not None:
        drawn = [v for v in drawn if v not in without]
    return drawn[:size]


# ----------------------------------------------------------------------------- ranking + metrics
def stable_desc_argsort(scores):
    """tf.argsort(x, direction='DESCENDING') == top_k(x, k=len).indices: lower index first on ties
    (bert4rec_model.py:232-236)."""
    s = np.asarray(scores)
    return np.argsort(-s, kind="stable") if s.dtype.kind == "f" else np.argsort(-s.astype(np.float64), kind="stable")


def rank_candidates(scores_row, candidates):
    """bert4rec_model.py:229-234: gather the candidates' logits, sort descending (stable), return ids."""
    cand = np.asarray(candidates, dtype=np.int64)
    order = stable_desc_argsort(np.asarray(scores_row)[cand])
    return cand[order]


def rank_of(ranking, gt):
    """bert4rec_evaluator.py:112-117: 1-based position of the first occurrence of gt."""
    return int(np.where(np.asarray(ranking) == gt)[0][0]) + 1


class MetricAccumulator:
    """evaluation_metrics.py:47-112 (+ defaults bert4rec_evaluator.py:12-21): sequential python-float sums."""

    def __init__(self, ks=(1, 5, 10)):
        self.ks = tuple(ks)
        self.n = 0
        self.hr = {k: 0.0 for k in self.ks}
        self.ndcg = {k: 0.0 for k in self.ks}
        self.ap = 0.0

    def update(self, rank):
        self.n += 1
        for k in self.ks:
            if rank <= k:
                self.hr[k] += 1
                self.ndcg[k] += 1 if rank == 1 else 1 / np.log2(rank + 1)
        self.ap += 1 / rank

    def results(self):
        d = float(self.n)
        out = {"Valid Ranks": self.n}
        for k in self.ks:
            out[f"NDCG@{k}"] = self.ndcg[k] / d
        for k in self.ks:
            out[f"HR@{k}"] = self.hr[k] / d
        out["MAP"] = self.ap / d
        return out
