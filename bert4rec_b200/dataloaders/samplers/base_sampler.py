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

    def set_source(self, source: list):
        self.source = list(source)

    def set_vocab(self, vocab: list):
        self.vocab = list(vocab)

    def set_sample_size(self, sample_size: int):
        self.sample_size = sample_size
