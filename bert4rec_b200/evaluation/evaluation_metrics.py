"""Rank-based evaluation metrics (reference: bert4rec/evaluation/evaluation_metrics.py:10-112): running
nominator/denominator python floats updated one rank at a time.  ``update_many`` applies a whole array of ranks with
``np.cumsum`` (strict left-to-right float64 accumulation), which reproduces the sequential sums bit for bit."""
import abc

import numpy as np


class EvaluationMetric(abc.ABC):
    def __init__(self, name: str, initial_value: int = 0):
        self._name = name
        self._initial_value = initial_value
        self._value = initial_value

    @property
    def name(self):
        return self._name

    @abc.abstractmethod
    def update(self, rank: int):
        ...

    def update_many(self, ranks):
        for r in np.asarray(ranks).tolist():
            self.update(r)

    def reset(self):
        self._value = self._initial_value

    def result(self):
        return self._value

    # ---- data-parallel evaluation: every rank evaluates its own sequences, the sums are merged at the end
    def _partials(self):
        """float64 partial sums that add across ranks."""
        return [float(self._value)]

    def _set_partials(self, p):
        self._value = type(self._initial_value)(p[0]) if isinstance(self._initial_value, int) else p[0]


class Counter(EvaluationMetric):
    """Counts update() calls (used as "Valid Ranks")."""

    def __init__(self, name: str = "Counter", initial_value: int = 0):
        super().__init__(name, initial_value)

    def update(self, rank: int):
        self._value += 1

    def update_many(self, ranks):
        self._value += int(np.asarray(ranks).size)


class RatioEvaluationMetric(EvaluationMetric):
    def __init__(self, name: str, initial_value: int = 0):
        super().__init__(name, initial_value)
        self._nominator = 0.0
        self._denominator = 0.0

    def _gain(self, rank):  # contribution of one rank to the nominator
        raise NotImplementedError

    def _gains(self, ranks: np.ndarray) -> np.ndarray:
        return np.array([self._gain(int(r)) for r in ranks], dtype=np.float64)

    def update(self, rank: int):
        self._denominator += 1
        self._nominator += self._gain(rank)
        self._value = self._nominator / self._denominator
        return self._value

    def update_many(self, ranks):
        ranks = np.asarray(ranks).reshape(-1)
        if ranks.size == 0:
            return self._value
        gains = self._gains(ranks)
        self._nominator = float(np.cumsum(np.concatenate(([self._nominator], gains)))[-1])
        self._denominator += float(ranks.size)
        self._value = self._nominator / self._denominator
        return self._value

    def _partials(self):
        return [self._nominator, self._denominator]

    def _set_partials(self, p):
        self._nominator, self._denominator = float(p[0]), float(p[1])
        self._value = self._nominator / self._denominator if self._denominator else self._initial_value

    def reset(self):
        super().reset()
        self._nominator = 0.0
        self._denominator = 0.0


class HitRatio(RatioEvaluationMetric):
    def __init__(self, k: int, name: str = "HitRatio", initial_value: int = 0):
        super().__init__(f"{name}@{k}", initial_value)
        self._k = k

    def _gain(self, rank):
        return 1 if rank <= self._k else 0

    def _gains(self, ranks):
        return (ranks <= self._k).astype(np.float64)


class NormalizedDiscountedCumulativeGain(RatioEvaluationMetric):
    def __init__(self, k: int, name: str = "NormalizedDiscountedCumulativeGain", initial_value: int = 0):
        super().__init__(f"{name}@{k}", initial_value)
        self._k = k

    def _gain(self, rank):
        if rank > self._k:
            return 0
        return 1 if rank == 1 else 1 / np.log2(rank + 1)

    def _gains(self, ranks):
        r = ranks.astype(np.int64)
        g = np.where(r == 1, 1.0, 1.0 / np.log2(r + 1))
        return np.where(r <= self._k, g, 0.0)


class MeanAveragePrecision(RatioEvaluationMetric):
    def __init__(self, name: str = "MeanAveragePrecision", initial_value: int = 0):
        super().__init__(name, initial_value)

    def _gain(self, rank):
        return 1 / rank

    def _gains(self, ranks):
        return 1.0 / ranks.astype(np.float64)


class HR(HitRatio):
    def __init__(self, k: int, name: str = "HR", initial_value: int = 0):
        super().__init__(k, name, initial_value)


class NDCG(NormalizedDiscountedCumulativeGain):
    def __init__(self, k: int, name: str = "NDCG", initial_value: int = 0):
        super().__init__(k, name, initial_value)


class MAP(MeanAveragePrecision):
    def __init__(self, name: str = "MAP", initial_value: int = 0):
        super().__init__(name, initial_value)
