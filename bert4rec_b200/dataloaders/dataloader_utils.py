"""Sequence -> model-input utilities of the hot path: popularity ranking, Cloze masking, last-token masking,
batching.  Integer results are bit-exact with the reference (bert4rec/dataloaders/dataloader_utils.py), pinned by
tests/golden/host_golden.json; tensors are numpy / torch instead of tf.data."""
import collections
import random

import numpy as np
import torch


def rank_items_by_popularity(items: list) -> list:
    """Unique items ordered by occurrence count (descending); ties keep first-seen order
    (reference: dataloader_utils.py:14-18 -- stable sort of all occurrences, then de-duplication)."""
    counts = collections.Counter(items)
    return sorted(counts, key=counts.get, reverse=True)  # Counter keeps first-seen order; sorted() is stable


class _SelectableVocab:
    """Virtual ``[i for i in range(vocab_size) if i not in special]`` -- ``random.choice`` only needs len + getitem."""

    def __init__(self, vocab_size, special_ids):
        self._special = sorted({int(s) for s in special_ids if 0 <= int(s) < vocab_size})
        self._n = vocab_size - len(self._special)

    def __len__(self):
        return self._n

    def __getitem__(self, k):
        if k < 0:
            k += self._n
        v = k
        for s in self._special:
            if s <= v:
                v += 1
            else:
                break
        return v


def apply_dynamic_masking_task(sequence: np.ndarray, max_selections_per_seq: int, mask_token_id: int,
                               special_token_ids: list, vocab_size: int, selection_rate: float = 0.2,
                               mask_token_rate: float = 0.8, random_token_rate: float = 0.1, seed: int = None):
    """Cloze masking with the reference's exact python-``random`` stream (dataloader_utils.py:186-261):
    seed -> shuffle of the candidate indexes -> per selected index one ``random()`` and, iff it falls below
    mask_rate + random_rate, one ``choice`` over the non-special vocab (drawn even when MASK overrides it).
    Returns (masked_token_ids, masked_lm_positions, masked_lm_ids), all of ``sequence.dtype``."""
    seq = np.asarray(sequence)
    random.seed(seed)
    n_plain = int(seq.shape[0] - np.count_nonzero(np.isin(seq, special_token_ids)))
    n_pred = min(max_selections_per_seq, max(1, int(n_plain * selection_rate)))
    order = list(range(n_plain))
    random.shuffle(order)
    chosen = sorted(order[:n_pred])
    selectable = _SelectableVocab(vocab_size, special_token_ids)
    masked = seq.copy()
    threshold = mask_token_rate + random_token_rate
    for idx in chosen:
        rn = random.random()
        token = seq[idx]
        if rn < threshold:
            token = random.choice(selectable)
        if rn < mask_token_rate:
            token = mask_token_id
        masked[idx] = token
    positions = np.asarray(chosen, dtype=seq.dtype)
    return masked, positions, seq[chosen].astype(seq.dtype) if len(chosen) else np.asarray([], dtype=seq.dtype)


def mask_last_token_only(sequence: np.ndarray, mask_token_id: int):
    """Leave-one-out evaluation masking (reference: dataloader_utils.py:264-269); mutates ``sequence`` like the
    reference does."""
    label = np.array([sequence[-1]], dtype=np.int64)
    sequence[-1] = mask_token_id
    return np.array(sequence, dtype=np.int64), np.array([len(sequence) - 1], dtype=np.int64), label


class BatchedDataset:
    """What ``make_batches`` returns: an in-memory (i.e. "cached") list of dict batches of int64 torch tensors."""

    def __init__(self, batches):
        self._batches = batches

    def __iter__(self):
        return iter(self._batches)

    def __len__(self):
        return len(self._batches)

    def __getitem__(self, i):
        return self._batches[i]

    def cardinality(self):
        return len(self._batches)

    def take(self, n):
        return BatchedDataset(self._batches[:n])

    def pin_memory(self):
        self._batches = [{k: v.pin_memory() for k, v in b.items()} for b in self._batches]
        return self

    def to(self, device):
        return BatchedDataset([{k: v.to(device, non_blocking=True) for k, v in b.items()} for b in self._batches])


def make_batches(dataset, buffer_size: int = None, batch_size: int = 64, squeeze_tensors: bool = False,
                 reshuffle_each_iteration: bool = False, seed: int = None) -> BatchedDataset:
    """shuffle -> batch -> [squeeze] -> cache (reference: dataloader_utils.py:306-346).  ``dataset`` is a sequence /
    iterable of per-sequence feature dicts (``BERT4RecPreprocessor.process_element`` outputs); the last batch may be
    partial.  The shuffle is a seeded full permutation (tf.data's shuffle stream is not reproducible outside TF)."""
    elements = list(dataset)
    if buffer_size is None:
        buffer_size = len(elements)
    order = np.arange(len(elements))
    if buffer_size and buffer_size > 1:
        np.random.RandomState(seed).shuffle(order)
    batches = []
    for i in range(0, len(order), batch_size):
        chunk = [elements[j] for j in order[i:i + batch_size]]
        batch = {}
        for k in chunk[0]:
            arr = np.stack([np.asarray(e[k]) for e in chunk]).astype(np.int64)
            if squeeze_tensors:
                arr = np.squeeze(arr)
            batch[k] = torch.from_numpy(arr)
        batches.append(batch)
    return BatchedDataset(batches)
