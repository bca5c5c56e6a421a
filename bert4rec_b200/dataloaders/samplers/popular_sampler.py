"""Deterministic top-popularity sampler (reference: popular_sampler.py:16-75): the source is ranked by occurrence
count once, ``sample`` returns the first ``sample_size`` ranked ids not in ``without`` (possibly fewer)."""
from absl import logging

from .base_sampler import BaseSampler
from .. import dataloader_utils


class PopularSampler(BaseSampler):
    def __init__(self, source: list = None, vocab: list = None, sample_size: int = None):
        super().__init__(source, vocab, sample_size)
        if self.source is not None:
            self.source = dataloader_utils.rank_items_by_popularity(self.source)

    def is_fully_prepared(self) -> bool:
        return self.source is not None and self.sample_size is not None

    def sample(self, sample_size: int = None, source: list = None, vocab: list = None, without: list = None) -> list:
        src, vocab, sample_size = self._resolve(source, vocab, sample_size)
        if src is None:
            raise ValueError("The source argument has to be provided to the popular sampler but None was given.")
        if sample_size >= len(src):
            logging.info("popular sampler: sample size >= len(source); the sample will be shorter than requested")
        ranked = src
        if without is not None:
            excluded = set(without)
            ranked = [v for v in src if v not in excluded]
        if self.source is None:  # ad-hoc source given per call: rank it now
            ranked = dataloader_utils.rank_items_by_popularity(ranked)
        return ranked[:sample_size]

    def sample_batch(self, withouts: list, sample_size: int = None, as_array: bool = False):
        """Batch of ``sample(without=w)`` calls over the stored (already ranked) source on the native host path."""
        from bert4rec_b200.dataloaders import host_native
        src, _, sample_size = self._resolve(None, None, sample_size)
        if src is None:
            raise ValueError("The source argument has to be provided to the popular sampler but None was given.")
        arr, lens = host_native.sample_popular_batch(src, self._withouts(withouts), sample_size)
        return self._unpack(arr, lens, as_array)

    def set_source(self, source: list):
        self.source = dataloader_utils.rank_items_by_popularity(list(source))
