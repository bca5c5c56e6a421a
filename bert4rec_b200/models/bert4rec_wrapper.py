"""BERT4RecModelWrapper: save / load a model directory (reference: bert4rec/models/bert4rec_wrapper.py:46-124).

Same directory contract as the reference for everything that is not a TensorFlow SavedModel: ``meta_config.json`` (model,
tokenizer identifier, encoder_config, last_trained, trained_on_dataset) and ``vocab.txt`` (tokenizer vocabulary).  The weights
are ``model_weights.npz`` keyed by the reference's TF variable names and shapes instead of a SavedModel,
so a TF2 run of the reference can exchange weights with this path offline; ``optimizer_state.npz`` (Adam moments + iteration
counter, flat layout) makes the directory a resumable checkpoint."""
import json
import logging
import pathlib

import numpy as np
import torch

from bert4rec_b200 import tokenizers
from bert4rec_b200.models import model_utils as utils
from bert4rec_b200.models.bert4rec_model import BERT4RecModel
from bert4rec_b200.models.components import networks
from bert4rec_b200.models.model_wrapper import ModelWrapper

_META_CONFIG_FILE_NAME = "meta_config.json"
_TOKENIZER_VOCAB_FILE_NAME = "vocab.txt"
_MODEL_WEIGHTS_FILE_NAME = "model_weights.npz"
_OPTIMIZER_STATE_FILE_NAME = "optimizer_state.npz"


class BERT4RecModelWrapper(ModelWrapper):
    def __init__(self, model: BERT4RecModel):
        super().__init__(model)
        enc_cfg = {k: v for k, v in model.encoder.get_config().items() if isinstance(v, (int, float, str, bool, type(None)))}
        self.update_meta({"model": "BERT4Rec", "encoder_config": enc_cfg})

    def save(self, save_path: pathlib.Path, tokenizer: tokenizers.BaseTokenizer = None, mode: int = 0) -> True:
        save_path = utils.determine_model_path(save_path, mode)
        if self.model.loss is None:
            raise RuntimeError("The model can't be saved without a loss. The model needs to be compiled first.")
        logging.info(f"Saving {self.model} to {save_path}")
        save_path.mkdir(parents=True, exist_ok=True)
        self.model.save_weights(save_path.joinpath(_MODEL_WEIGHTS_FILE_NAME))
        st = self.model.store
        if st.grads is not None:   # training buffers exist: Adam moments + optimizer.iterations
            np.savez(save_path.joinpath(_OPTIMIZER_STATE_FILE_NAME), m=st.m.cpu().numpy(), v=st.v.cpu().numpy(),
                     iterations=st.step_counter.cpu().numpy())
        if tokenizer:
            tokenizer.export_vocab_to_file(save_path.joinpath(_TOKENIZER_VOCAB_FILE_NAME))
            self.update_meta({"tokenizer": tokenizer.identifier})
        with open(save_path.joinpath(_META_CONFIG_FILE_NAME), "w") as f:
            json.dump(self._meta_config, f, indent=4)
        return True

    @classmethod
    def load(cls, save_path: pathlib.Path, mode: int = 0, device=None) -> dict:
        save_path = utils.determine_model_path(save_path, mode)
        if not save_path.exists():
            raise ValueError(f"The given path {save_path} does not exist.")
        logging.info(f"Loading model from {save_path}")
        try:
            with open(save_path.joinpath(_META_CONFIG_FILE_NAME)) as jf:
                meta_config = json.load(jf)
        except FileNotFoundError:
            raise ValueError(f"The meta configuration json file ({_META_CONFIG_FILE_NAME}) could not be found in the "
                             f"supposed model directory: {save_path} (it holds the encoder configuration)")
        enc_kwargs = dict(meta_config["encoder_config"])
        if device is not None:
            enc_kwargs["device"] = device
        model = BERT4RecModel(networks.Bert4RecEncoder(**enc_kwargs))
        model.load_weights(save_path.joinpath(_MODEL_WEIGHTS_FILE_NAME))
        model.compile()
        opt_file = save_path.joinpath(_OPTIMIZER_STATE_FILE_NAME)
        if opt_file.exists():
            with np.load(opt_file) as z:
                st = model.store
                st.m.copy_(torch.from_numpy(z["m"])); st.v.copy_(torch.from_numpy(z["v"]))
                st.step_counter.copy_(torch.from_numpy(z["iterations"]))
        wrapper = cls(model)
        wrapper._meta_config = meta_config
        loaded_assets = {"model_wrapper": wrapper}
        if meta_config.get("tokenizer") is not None:
            tokenizer = tokenizers.get(meta_config["tokenizer"])
            tokenizer.import_vocab_from_file(save_path.joinpath(_TOKENIZER_VOCAB_FILE_NAME))
            loaded_assets["tokenizer"] = tokenizer
        return loaded_assets
