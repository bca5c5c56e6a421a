"""Evaluator base (reference: bert4rec/evaluation/base_evaluator.py:15-79)."""
import abc
import json
import pathlib

from absl import logging

from bert4rec_b200.dataloaders import samplers
from .evaluation_metrics import EvaluationMetric


class BaseEvaluator(abc.ABC):
    def __init__(self, metrics, sampler="popular", dataloader=None):
        self.sampler = samplers.get(sampler)
        if self.sampler.sample_size is None:
            logging.warning(f"The sampler used in the evaluator {self} does not have a sample size set. "
                            "This might lead to problems during evaluation.")
        if self.sampler.source is None:
            logging.info(f"The sampler used in the evaluator {self} does not have a source set.")
        self._metrics = metrics
        self.dataloader = dataloader
        self.reset_metrics()

    def reset_metrics(self) -> None:
        for m in self._metrics:
            m.reset()

    @abc.abstractmethod
    def evaluate(self, model, test_data):
        ...

    def get_metrics(self):
        return self._metrics

    def all_reduce_metrics(self, group=None, device=None) -> dict:
        """Data-parallel evaluation: every rank evaluates its own sequences; ONE all-reduce(SUM) of the metrics'
        float64 partial sums (counter, nominators, denominators) merges them, after which every rank holds the global
        results.  ``device``: where the reduction buffer lives (a CUDA device for NCCL, None / "cpu" for gloo)."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.get_metrics_results()
        parts = [m._partials() for m in self._metrics]
        flat = torch.tensor([x for p in parts for x in p], dtype=torch.float64, device=device)
        dist.all_reduce(flat, group=group)
        vals, k = flat.cpu().tolist(), 0
        for m, p in zip(self._metrics, parts):
            m._set_partials(vals[k:k + len(p)])
            k += len(p)
        return self.get_metrics_results()

    def get_metrics_results(self) -> dict:
        return {m.name: m.result() for m in self._metrics}

    def save_results(self, save_path: pathlib.Path) -> pathlib.Path:
        save_path = pathlib.Path(save_path)
        if save_path.is_dir():
            save_path = save_path.joinpath("eval_results.json")
        with open(save_path, "w") as f:
            json.dump(self.get_metrics_results(), f, indent=4)
        return save_path
