"""Sampler base: holds an optional ``source`` (item occurrences, with duplicates), ``vocab`` (unique ids) and
``sample_size``; per-call arguments override the stored ones (reference contract: base_sampler.py:14-45)."""
import abc


class BaseSampler(abc.ABC):
    def __init__(self, source: list = None, vocab: list = None, sample_size: int = None):
        if sample_size is not None and sample_size < 0:
            raise ValueError(f"The sample size shouldn't be negative to avoid unexpected outputs (Given: {sample_size})")
        self.source = list(source) if source is not None else None
        self.vocab = list(vocab) if vocab is not None else None
        self.sample_size = sample_size

    def _resolve(self, source=None, vocab=None, sample_size=None):
        source = self.source if source is None else source
        vocab = self.vocab if vocab is None else vocab
        if sample_size is None:
            sample_size = self.sample_size
            if sample_size is None:
                raise ValueError("The sample size has to be given either during the initialization of the sampler "
                                 "or as an argument in the sample() method call.")
        if sample_size < 0:
            raise ValueError(f"A negative sample size is not allowed (Given: {sample_size})")
        return source, vocab, sample_size

    @abc.abstractmethod
    def sample(self, sample_size: int = None, source: list = None, vocab: list = None, without: list = None) -> list:
        ...

    @abc.abstractmethod
    def is_fully_prepared(self) -> bool:
        ...

    def sample_batch(self, withouts: list, sample_size: int = None, as_array: bool = False):
        """One ``sample(without=w)`` per entry of ``withouts`` (the evaluator's per-slot calls, bert4rec_evaluator.py:98-104).
        The concrete samplers run the whole batch on the C++ threads of libb4r.so (bit-exact with the per-call results);
        this generic version is the per-call loop.  ``as_array``: (int64 [n, sample_size] zero-padded, lengths [n])."""
        if isinstance(withouts, tuple):
            vals, off = withouts
            withouts = [vals[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]
        outs = [self.sample(sample_size=sample_size, without=w) for w in withouts]
        return self._pack(outs, sample_size) if as_array else outs

    @staticmethod
    def _withouts(withouts):
        """list of exclusion lists (None = nothing excluded), or a ready CSR pair (values, offsets[n+1]) of numpy arrays"""
        if isinstance(withouts, tuple):
            return withouts
        return [w if w is not None else [] for w in withouts]

    def _pack(self, outs, sample_size):
        import numpy as np
        size = self.sample_size if sample_size is None else sample_size
        arr = np.zeros((len(outs), size), dtype=np.int64)
        lens = np.zeros(len(outs), dtype=np.int32)
        for i, o in enumerate(outs):
            arr[i, :len(o)] = o
            lens[i] = len(o)
        return arr, lens

    @staticmethod
    def _unpack(arr, lens, as_array):
        return (arr, lens) if as_array else [arr[i, :lens[i]].tolist() for i in range(arr.shape[0])]

    def set_source(self, source: list):
        self.source = list(source)

    def set_vocab(self, vocab: list):
        self.vocab = list(vocab)

    def set_sample_size(self, sample_size: int):
        self.sample_size = sample_size
