"""Batch versions of the per-sequence host work of the hot loop on the C++ threads of libb4r.so
(``b4r_host_*``, csrc/host_data.cu): Cloze masking + padded layout, and the three negative samplers.  Given the same seeds
the results are bit-identical to the per-sequence Python of the reference (``random`` / ``np.random`` streams restated in
C++); no GPU is involved.  Reference: dataloader_utils.py:186-261, bert4rec_preprocessor.py:47-116,
random_sampler.py:63-79, popular_sampler.py:53-71, popular_random_sampler.py:77-117."""
import ctypes as C

import numpy as np

from bert4rec_b200 import _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _csr(lists, dtype=np.int64):
    """list of int sequences -> (values, offsets[n+1]); a ready ``(values, offsets)`` pair passes through"""
    if isinstance(lists, tuple) and len(lists) == 2 and isinstance(lists[1], np.ndarray):
        return np.ascontiguousarray(lists[0], dtype=dtype), np.ascontiguousarray(lists[1], dtype=np.int64)
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    if len(lists):
        np.cumsum([len(x) for x in lists], out=off[1:])
    vals = np.empty(int(off[-1]), dtype=dtype)
    for i, x in enumerate(lists):
        vals[off[i]:off[i + 1]] = x
    return vals, off


def fresh_seeds(n, bits=32):
    """What ``seed=None`` means in the reference: fresh entropy for every call."""
    ss = np.random.SeedSequence()
    return ss.generate_state(n, dtype=np.uint32 if bits == 32 else np.uint64)


def _check(rc):
    if rc != 0:
        raise ValueError(_lib.load().b4r_last_error().decode())


def cloze_mask_batch(sequences, max_seq_len, max_predictions_per_seq, mask_token_id, special_token_ids, vocab_size,
                     selection_rate=0.2, mask_token_rate=0.8, random_token_rate=0.1, seeds=None, pad_token_id=0, n_threads=0):
    """``process_element(apply_mlm=True, finetuning=False)`` of n tokenised, windowed sequences at once.
    Returns the reference's dict of int64 arrays (labels / input_word_ids / input_mask [n, S]; masked_lm_* [n, P])."""
    lib = _lib.load()
    toks, off = _csr(sequences)
    n = len(off) - 1
    seeds = fresh_seeds(n, 64) if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
    assert seeds.shape == (n,)
    sp = np.ascontiguousarray(list(special_token_ids), dtype=np.int64)
    S, P = int(max_seq_len), int(max_predictions_per_seq)
    out = {k: np.empty((n, S), dtype=np.int64) for k in ("labels", "input_word_ids", "input_mask")}
    out.update({k: np.empty((n, P), dtype=np.int64) for k in ("masked_lm_ids", "masked_lm_positions", "masked_lm_weights")})
    _check(lib.b4r_host_cloze_mask_batch(_p(toks), _p(off), _p(seeds), n, S, P, int(mask_token_id), int(pad_token_id), _p(sp),
                                         len(sp), int(vocab_size), float(selection_rate), float(mask_token_rate),
                                         float(random_token_rate), _p(out["labels"]), _p(out["input_word_ids"]),
                                         _p(out["input_mask"]), _p(out["masked_lm_ids"]), _p(out["masked_lm_positions"]),
                                         _p(out["masked_lm_weights"]), int(n_threads)))
    return out


def _seeds32(seed, n):
    if seed is None:
        return fresh_seeds(n, 32)
    if np.ndim(seed) == 0:
        if not 0 <= int(seed) <= 0xFFFFFFFF:
            raise ValueError("Seed must be between 0 and 2**32 - 1")
        return np.full(n, int(seed), dtype=np.uint32)   # the reference re-seeds with the same seed on every call
    s = np.ascontiguousarray(seed, dtype=np.uint32)
    assert s.shape == (n,)
    return s


def sample_random_batch(vocab, withouts, sample_size, allow_duplicates=False, seed=None, n_threads=0):
    """n x ``RandomSampler.sample(without=withouts[i], seed=seed)`` -> int64 [n, sample_size]."""
    lib = _lib.load()
    v = np.ascontiguousarray(vocab, dtype=np.int64)
    w, off = _csr(withouts)
    n = len(off) - 1
    out = np.empty((n, int(sample_size)), dtype=np.int64)
    seeds = _seeds32(seed, n)
    _check(lib.b4r_host_sample_random_batch(_p(v), len(v), _p(w), _p(off), _p(seeds), n, int(sample_size),
                                            int(bool(allow_duplicates)), _p(out), int(n_threads)))
    return out


def sample_popular_batch(ranked_source, withouts, sample_size, n_threads=0):
    """n x ``PopularSampler.sample(without=withouts[i])`` over the ranked source -> (int64 [n, sample_size], lengths [n])."""
    lib = _lib.load()
    r = np.ascontiguousarray(ranked_source, dtype=np.int64)
    w, off = _csr(withouts)
    n = len(off) - 1
    out = np.empty((n, int(sample_size)), dtype=np.int64)
    lens = np.empty(n, dtype=np.int32)
    _check(lib.b4r_host_sample_popular_batch(_p(r), len(r), _p(w), _p(off), n, int(sample_size), _p(out), _p(lens), int(n_threads)))
    return out, lens


def sample_pop_random_batch(vocab, probabilities, withouts, sample_size, allow_duplicates=False, seed=None, n_threads=0):
    """n x ``PopularRandomSampler.sample(without=withouts[i], seed=seed)`` -> (int64 [n, sample_size], lengths [n])."""
    lib = _lib.load()
    v = np.ascontiguousarray(vocab, dtype=np.int64)
    p = np.ascontiguousarray(probabilities, dtype=np.float64)
    assert p.shape == v.shape
    w, off = _csr(withouts)
    n = len(off) - 1
    out = np.empty((n, int(sample_size)), dtype=np.int64)
    lens = np.empty(n, dtype=np.int32)
    seeds = _seeds32(seed, n)
    _check(lib.b4r_host_sample_pop_random_batch(_p(v), _p(p), len(v), _p(w), _p(off), _p(seeds), n, int(sample_size),
                                                int(bool(allow_duplicates)), _p(out), _p(lens), int(n_threads)))
    return out, lens
