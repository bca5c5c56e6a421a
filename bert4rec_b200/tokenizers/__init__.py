"""Tokenizer factory (reference: bert4rec/tokenizers/__init__.py:12-25)."""
from typing import Union

from .base_tokenizer import BaseTokenizer
from .simple_tokenizer import SimpleTokenizer

tokenizers_map = {"simple": SimpleTokenizer}


def get(identifier: Union[str, BaseTokenizer] = "simple", **kwargs) -> BaseTokenizer:
    """String id -> new instance; an instance passes through; anything else raises ValueError."""
    if isinstance(identifier, BaseTokenizer):
        return identifier
    if isinstance(identifier, str) and identifier in tokenizers_map:
        return tokenizers_map[identifier](**kwargs)
    raise ValueError(f"{identifier} is not known!")
