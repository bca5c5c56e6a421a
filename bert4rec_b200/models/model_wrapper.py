"""Administrative wrapper around a model: meta information + extended save / load (reference: bert4rec/models/model_wrapper.py)."""
import abc
from typing import Union


class ModelWrapper(abc.ABC):
    _model = None
    _custom_objects: dict = {}

    def __init__(self, model):
        self._model = model
        self._meta_config = {"model": getattr(model, "name", None), "tokenizer": None, "last_trained": None,
                             "trained_on_dataset": None}

    @property
    def model(self):
        return self._model

    def get_meta_config(self) -> dict:
        return self._meta_config

    def update_meta(self, updated_info: dict) -> True:
        self._meta_config.update(updated_info)
        return True

    def delete_keys_from_meta(self, keys: Union[list, str]) -> True:
        for key in ([keys] if isinstance(keys, str) else keys):
            self._meta_config.pop(key, None)
        return True
