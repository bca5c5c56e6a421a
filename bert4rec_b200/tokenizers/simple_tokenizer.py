"""SimpleTokenizer: value -> incrementing integer id in first-seen order (first id is 0), ``key|id`` vocab files.

Behavioural contract from the reference (bert4rec/tokenizers/simple_tokenizer.py:15-202): ids start at 0 and grow by
one per unseen key (:119-138); iterables map element-wise to lists (:140-152); a non-extensible tokenizer raises
RuntimeError on unknown keys; vocab files are ``key|id`` lines (:76-116).  torch / numpy integer tensors are accepted
where the reference accepts tf tensors.
"""
import numbers
import os
import pathlib
from collections.abc import Iterable

import numpy as np

from .base_tokenizer import BaseTokenizer

try:  # pandas is optional on this path
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None


class SimpleTokenizer(BaseTokenizer):
    DELIMITER = "|"

    def __init__(self, vocab_file_path: pathlib.Path = None, extensible: bool = True):
        self._vocab = {}
        self._delimiter = self.DELIMITER
        self._inverse = None
        super().__init__(vocab_file_path=vocab_file_path, extensible=extensible)
        if self._vocab is None:
            self._vocab = {}

    @property
    def identifier(self):
        return "simple"

    def clear_vocab(self):
        self._vocab = {}
        self._vocab_size = 0
        self._inverse = None

    # ------------------------------------------------------------------ tokenize
    def _one(self, key) -> int:
        if isinstance(key, bytes):
            key = key.decode("utf-8")
        tok = self._vocab.get(key)
        if tok is None:
            if not self._extensible:
                raise RuntimeError(f"\"{key}\" is not known!")
            tok = self._vocab_size
            self._vocab[key] = tok
            self._vocab_size += 1
            self._inverse = None
        return tok

    def tokenize(self, input, progress_bar: bool = False):
        if isinstance(input, (bytes, str)):
            return self._one(input)
        if pd is not None and isinstance(input, pd.Series):
            return input.map(self.tokenize)
        if hasattr(input, "detach") and hasattr(input, "cpu"):  # torch tensor of strings is impossible; ints pass through
            input = input.detach().cpu().tolist()
        if isinstance(input, np.ndarray):
            input = input.tolist()
        if isinstance(input, Iterable):
            return [self.tokenize(v) for v in input]
        raise ValueError("The provided argument is not of a supported type")

    # ------------------------------------------------------------------ detokenize
    def _inv(self):
        if self._inverse is None:
            self._inverse = {v: k for k, v in self._vocab.items()}
        return self._inverse

    def detokenize(self, token, drop_tokens=None, progress_bar: bool = False):
        if isinstance(token, numbers.Number):
            value = self._inv().get(int(token))
            if drop_tokens and value in drop_tokens:
                value = None
            return value
        if pd is not None and isinstance(token, pd.Series):
            return token.map(lambda t: self.detokenize(t, drop_tokens))
        if hasattr(token, "detach") and hasattr(token, "cpu"):
            token = token.detach().cpu().tolist()
        if isinstance(token, np.ndarray):
            token = token.tolist()
        if isinstance(token, Iterable):
            out = []
            for t in token:
                v = self.detokenize(t, drop_tokens)
                if v is not None:
                    out.append(v)
            return out
        raise ValueError("The provided argument is not of a supported type")

    # ------------------------------------------------------------------ vocab files
    def import_vocab_from_file(self, vocab_file: pathlib.Path) -> bool:
        vocab_file = pathlib.Path(vocab_file)
        if not vocab_file.is_file():
            raise RuntimeError(f"The vocab file does not exist (yet) or is not located at {vocab_file}.")
        self.clear_vocab()
        with open(vocab_file, "rb") as f:
            lines = [ln.decode() for ln in f.readlines()]
        if not lines:
            raise ValueError(f"The given vocab file ({vocab_file}) is empty.")
        if self._delimiter not in lines[0] or len(lines[0].split(self._delimiter)) != 2:
            raise ValueError(f"The given vocab file ({vocab_file}) should contain \"{self._delimiter}\"-separated "
                             f"key-value-pairs per individual line.")
        for ln in lines:
            key, tok = ln.split(self._delimiter)
            self._vocab[key] = int(tok)
        self._vocab_size = len(self._vocab)
        return True

    def export_vocab_to_file(self, file_path: pathlib.Path) -> bool:
        if not self._vocab:
            raise ValueError("The vocab of the tokenizer is empty and therefore can't be written to a file.")
        with open(file_path, "wb") as f:
            for key, tok in self._vocab.items():
                f.write(f"{key}{self._delimiter}{tok}{os.linesep}".encode("utf-8"))
        return True
