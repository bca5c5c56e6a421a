"""Trainer factory (reference: bert4rec/trainers/__init__.py:10-21)."""
from .base_trainer import BaseTrainer
from .bert4rec_trainer import BERT4RecTrainer
from . import optimizers, trainer_utils, callbacks  # noqa: F401

trainers_map = {"bert4rec": BERT4RecTrainer}


def get(identifier: str = "bert4rec", **kwargs) -> BaseTrainer:
    if isinstance(identifier, str) and identifier in trainers_map:
        return trainers_map[identifier](**kwargs)
    raise ValueError(f"{identifier} is not known!")
