"""Evaluator factory (reference: bert4rec/evaluation/__init__.py:11-22)."""
from .base_evaluator import BaseEvaluator
from .bert4rec_evaluator import BERT4RecEvaluator, bert4rec_evaluation_metrics, default_bert4rec_metrics
from .evaluation_metrics import *  # noqa: F401,F403

evaluators_map = {"bert4rec": BERT4RecEvaluator}


def get(identifier: str = "bert4rec", **kwargs) -> BaseEvaluator:
    if isinstance(identifier, str) and identifier in evaluators_map:
        return evaluators_map[identifier](**kwargs)
    raise ValueError(f"{identifier} is not known!")
