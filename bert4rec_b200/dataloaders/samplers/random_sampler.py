"""Uniform sampler.  Bit-exact with the reference (random_sampler.py:48-50,63-79): numpy's *global* legacy
RandomState is re-seeded with ``seed`` on every call (None -> fresh entropy), the pool is the vocab in its stored
order minus ``without``, and one ``np.random.choice`` draws the sample.  The exclusion uses a hash set instead of
the reference's O(V*len(without)) list scans -- same pool, same order, same draw."""
import numpy as np

from .base_sampler import BaseSampler


class RandomSampler(BaseSampler):
    def __init__(self, source: list = None, vocab: list = None, sample_size: int = None,
                 allow_duplicates: bool = False, seed: int = None):
        super().__init__(source, vocab, sample_size)
        if self.vocab is None and self.source is not None:
            self.vocab = list(set(self.source))
        self.allow_duplicates = allow_duplicates
        self.seed = seed

    def is_fully_prepared(self) -> bool:
        return self.vocab is not None and self.sample_size is not None

    def sample(self, sample_size: int = None, source: list = None, vocab: list = None, allow_duplicates: bool = None,
               seed: int = None, without: list = None) -> list:
        src, vocab, sample_size = self._resolve(source, vocab, sample_size)
        if vocab is None and src is not None and self.source is None:
            vocab = list(set(src))
        if vocab is None:
            raise ValueError("No vocab or any other source has been given to the random sampler.")
        np.random.seed(self.seed if seed is None else seed)
        if allow_duplicates is None:
            allow_duplicates = self.allow_duplicates
        if allow_duplicates is False and sample_size > len(vocab):
            raise ValueError(f"When no duplicates are allowed in the final sample then the sample size (given sample "
                             f"size: {sample_size})) can not be greater than the length length of the vocab (length "
                             f"of the vocab: {len(vocab)})")
        pool = vocab
        if without is not None:
            excluded = set(without)
            pool = [v for v in vocab if v not in excluded]
        return np.random.choice(pool, size=sample_size, replace=allow_duplicates).tolist()

    def sample_batch(self, withouts: list, sample_size: int = None, as_array: bool = False, seed: int = None):
        """Batch of ``sample(without=w)`` calls on the native host path (``b4r_host_sample_random_batch``): same pools, same
        numpy legacy streams, same draws as the per-call method (every call re-seeds with the same seed; None = entropy)."""
        from bert4rec_b200.dataloaders import host_native
        _, vocab, sample_size = self._resolve(None, None, sample_size)
        if vocab is None:
            raise ValueError("No vocab or any other source has been given to the random sampler.")
        arr = host_native.sample_random_batch(vocab, self._withouts(withouts), sample_size, self.allow_duplicates,
                                              self.seed if seed is None else seed)
        import numpy as np
        return self._unpack(arr, np.full(arr.shape[0], arr.shape[1], dtype=np.int32), as_array)

    def set_source(self, source: list):
        super().set_source(source if self.allow_duplicates else list(set(source)))
