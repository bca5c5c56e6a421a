"""Negative-sampler factory (reference: bert4rec/dataloaders/samplers/__init__.py:17-30)."""
from typing import Union

from .base_sampler import BaseSampler
from .random_sampler import RandomSampler
from .popular_sampler import PopularSampler
from .popular_random_sampler import PopularRandomSampler

samplers_map = {
    "random": RandomSampler,
    "popular": PopularSampler,
    "pop_random": PopularRandomSampler,
    "popular_random": PopularRandomSampler,
}


def get(identifier: Union[str, BaseSampler] = "popular", **kwargs) -> BaseSampler:
    if isinstance(identifier, BaseSampler):
        return identifier
    if isinstance(identifier, str) and identifier in samplers_map:
        return samplers_map[identifier](**kwargs)
    raise ValueError(f"{identifier} is not known!")
