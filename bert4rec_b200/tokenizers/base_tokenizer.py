"""Tokenizer base class (reference contract: bert4rec/tokenizers/base_tokenizer.py)."""
import abc
import pathlib


class BaseTokenizer(abc.ABC):
    def __init__(self, vocab_file_path: pathlib.Path = None, extensible: bool = True):
        self._vocab = None
        self._vocab_size = 0
        self._extensible = extensible
        if vocab_file_path is not None and pathlib.Path(vocab_file_path).is_file():
            self._extensible = False
            self.import_vocab_from_file(pathlib.Path(vocab_file_path))

    @property
    @abc.abstractmethod
    def identifier(self):
        ...

    @abc.abstractmethod
    def tokenize(self, input, progress_bar: bool = False):
        ...

    @abc.abstractmethod
    def detokenize(self, token, drop_tokens=None, progress_bar: bool = False):
        ...

    @abc.abstractmethod
    def import_vocab_from_file(self, vocab_file: pathlib.Path) -> bool:
        ...

    @abc.abstractmethod
    def export_vocab_to_file(self, file_path: pathlib.Path) -> bool:
        ...

    def get_vocab(self):
        return self._vocab

    def get_vocab_size(self) -> int:
        return self._vocab_size

    def enable_extensibility(self):
        self._extensible = True
        return True

    def disable_extensibility(self):
        self._extensible = False
        return True
