"""Popularity-weighted random sampler -- the evaluator's default ("pop_random").

Bit-exact with the reference (popular_random_sampler.py:53-55,77-126): numpy's global RandomState is re-seeded on
every call; ``sample_size + len(set(without))`` ids are drawn without replacement with probability
count(item)/len(source); excluded ids are dropped and the first ``sample_size`` survivors returned.  The probability
table is built with one Counter pass instead of the reference's O(V*len(source)) ``list.count`` loop (same floats),
and the exclusion uses a hash set."""
import collections

import numpy as np

from .base_sampler import BaseSampler


class PopularRandomSampler(BaseSampler):
    def __init__(self, source: list = None, vocab: list = None, sample_size: int = None,
                 allow_duplicates: bool = False, seed: int = None):
        super().__init__(source, vocab, sample_size)
        self.vocab = vocab  # the reference keeps the caller's list object here
        self.probability_distribution = []
        self.allow_duplicates = allow_duplicates
        self.seed = seed
        if self.source is not None and self.vocab is not None:
            self._determine_probability_distribution(self.source, self.vocab)

    def is_fully_prepared(self) -> bool:
        return (self.vocab is not None and len(self.vocab) == len(self.probability_distribution)
                and self.sample_size is not None)

    def _determine_probability_distribution(self, source: list, vocab: list):
        counts = collections.Counter(source)
        total = len(source)
        self.probability_distribution = [counts.get(item, 0) / total for item in vocab]
        self._p_array = np.asarray(self.probability_distribution, dtype=np.float64)

    def sample(self, sample_size: int = None, source: list = None, vocab: list = None, allow_duplicates: bool = None,
               seed: int = None, without: list = None) -> list:
        src, vocab, sample_size = self._resolve(source, vocab, sample_size)
        np.random.seed(self.seed if seed is None else seed)
        if src is None:
            raise ValueError("The source argument has to be given either during the initialization of the sampler or "
                             "as an argument in the sample method call when working with the popular random sampler.")
        if vocab is None:
            raise ValueError("The vocab argument has to be given either during the initialization of the sampler or "
                             "as an argument in the sample method call when working with the popular random sampler.")
        if allow_duplicates is None:
            allow_duplicates = self.allow_duplicates
        if allow_duplicates is False and sample_size > len(vocab):
            raise ValueError(f"When no duplicates are allowed in the final sample then the sample size (given sample "
                             f"size: {sample_size})) can not be greater than the length of the vocab (length of the "
                             f"vocab: {len(vocab)})")
        if not self.probability_distribution:
            self._determine_probability_distribution(src, vocab)
        n_draw = sample_size
        excluded = None
        if without is not None:
            excluded = set(without)
            n_draw += len(excluded)
        if not allow_duplicates and n_draw > len(vocab):
            raise ValueError(f"The given without list (length: {len(excluded)} reduces the vocab (length: "
                             f"{len(vocab)}) too much to take a sample of size {sample_size} (since no duplicates "
                             f"are allowed).")
        drawn = np.random.choice(vocab, n_draw, allow_duplicates, self._p_array).tolist()
        if excluded is not None:
            drawn = [v for v in drawn if v not in excluded]
        return drawn[:sample_size]

    def sample_batch(self, withouts: list, sample_size: int = None, as_array: bool = False, seed: int = None):
        """Batch of ``sample(without=w)`` calls on the native host path (``b4r_host_sample_pop_random_batch``): the
        weighted ``np.random.choice`` of the legacy RandomState restated in C++, bit-exact with the per-call method."""
        from bert4rec_b200.dataloaders import host_native
        src, vocab, sample_size = self._resolve(None, None, sample_size)
        if src is None or vocab is None:
            raise ValueError("The source and vocab arguments have to be given during the initialization of the sampler "
                             "when working with the popular random sampler.")
        if not self.probability_distribution:
            self._determine_probability_distribution(src, vocab)
        arr, lens = host_native.sample_pop_random_batch(vocab, self._p_array, self._withouts(withouts), sample_size,
                                                        self.allow_duplicates, self.seed if seed is None else seed)
        return self._unpack(arr, lens, as_array)

    def set_source(self, source: list):
        super().set_source(source)
        if self.vocab is not None:
            self._determine_probability_distribution(self.source, self.vocab)

    def set_vocab(self, vocab: list):
        super().set_vocab(vocab)
        if self.source is not None:
            self._determine_probability_distribution(self.source, self.vocab)
