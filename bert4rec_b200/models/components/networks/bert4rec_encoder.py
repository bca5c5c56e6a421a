"""Bi-directional Transformer encoder of BERT4Rec on the B200 path.

Same constructor signature, input dict and output dict as the reference's ``Bert4RecEncoder``
(bert4rec/models/components/networks/bert4rec_encoder.py:62-80,186-231).  The layers it composes there
(OnDeviceEmbedding, PositionEmbedding, LayerNormalization(1e-12), SelfAttentionMask, L x post-LN
TransformerEncoderBlock, tanh pooler) are executed here by the CUDA kernels behind ``b4r_encode`` /
``b4r_pooled_output``; the weights live in a flat :class:`~bert4rec_b200.engine.ParamStore`.
"""
from typing import Any, Callable, Optional, Union

import torch

from bert4rec_b200.engine import ParamStore

_Activation = Union[str, Callable[..., Any]]


class Bert4RecEncoder:
    def __init__(self,
                 vocab_size: int,
                 hidden_size: int = 768,
                 num_layers: int = 12,
                 num_attention_heads: int = 12,
                 max_sequence_length: int = 512,
                 inner_dim: int = 3072,
                 inner_activation: _Activation = "gelu",
                 output_dropout: float = 0.1,
                 attention_dropout: float = 0.1,
                 initializer="truncated_normal",
                 output_range: Optional[int] = None,
                 embedding_width: Optional[int] = None,
                 embedding_layer=None,
                 norm_first: bool = False,
                 with_dense_inputs: bool = False,
                 **kwargs):
        # V1-style aliases accepted by the reference (bert4rec_encoder.py:82-93)
        kwargs.pop("dict_outputs", None)
        kwargs.pop("return_all_encoder_outputs", None)
        inner_dim = kwargs.pop("intermediate_size", inner_dim)
        inner_activation = kwargs.pop("activation", inner_activation)
        output_dropout = kwargs.pop("dropout_rate", output_dropout)
        attention_dropout = kwargs.pop("attention_dropout_rate", attention_dropout)
        device = kwargs.pop("device", "cuda:0")
        seed = kwargs.pop("seed", 0)
        self.name = kwargs.pop("name", "bert4_rec_encoder")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        if embedding_width is None:
            embedding_width = hidden_size
        # features of the generic Model-Garden encoder that no shipped BERT4Rec config uses are not on the hot path
        unsupported = []
        if inner_activation not in ("gelu",) and getattr(inner_activation, "__name__", "") != "gelu":
            unsupported.append(f"inner_activation={inner_activation!r} (only exact-erf 'gelu')")
        if embedding_width != hidden_size:
            unsupported.append("embedding_width != hidden_size (factorised embedding projection)")
        if embedding_layer is not None:
            unsupported.append("custom embedding_layer")
        if norm_first:
            unsupported.append("norm_first=True (pre-LN)")
        if with_dense_inputs:
            unsupported.append("with_dense_inputs=True")
        if output_range is not None and (int(output_range) < 1 or int(output_range) > max_sequence_length):
            raise ValueError(f"output_range must lie in [1, max_sequence_length], got {output_range}")
        if unsupported:
            raise NotImplementedError("Bert4RecEncoder (B200 path) does not implement: " + "; ".join(unsupported))
        self._config = {
            "vocab_size": vocab_size, "hidden_size": hidden_size, "num_layers": num_layers,
            "num_attention_heads": num_attention_heads, "max_sequence_length": max_sequence_length,
            "inner_dim": inner_dim, "inner_activation": "gelu", "output_dropout": output_dropout,
            "attention_dropout": attention_dropout, "initializer": initializer, "output_range": output_range,
            "embedding_width": embedding_width, "embedding_layer": embedding_layer, "norm_first": norm_first,
            "with_dense_inputs": with_dense_inputs,
        }
        self.store = ParamStore(vocab_size, hidden_size, num_layers, num_attention_heads, max_sequence_length,
                                inner_dim, output_dropout, attention_dropout, device=device)
        self.store.init_weights(seed)
        self.inputs = dict(input_word_ids=None, input_mask=None)  # placeholder for keras.Input specs

    # ------------------------------------------------------------------ call
    def _prep(self, t):
        t = torch.as_tensor(t)
        if t.dim() != 2:
            raise ValueError(f"expected a [batch, seq_len] tensor, got shape {tuple(t.shape)}")
        return t.to(device=self.store.device, dtype=torch.int64).contiguous()

    def __call__(self, inputs, training=None):
        return self.call(inputs, training=training)

    def call(self, inputs, training=None):
        if not isinstance(inputs, dict):
            raise ValueError("Unexpected inputs type to %s." % self.__class__)
        if inputs.get("input_word_embeddings") is not None or inputs.get("dense_inputs") is not None:
            raise NotImplementedError("input_word_embeddings / dense_inputs are not on the B200 hot path")
        ids, mask = self._prep(inputs.get("input_word_ids")), self._prep(inputs.get("input_mask"))
        B, S = ids.shape
        sess = self.store.session(B, S, 0)
        sess.encode(ids, mask, training=bool(training), seed=torch.seed() & 0x7FFFFFFFFFFF if training else 0)
        L = self._config["num_layers"]
        outs = [sess.sequence_output(l).float() for l in range(L)]
        # output_range (bert4rec_encoder.py:45-48,144): the LAST layer's target sequence is sliced to [0, output_range).  The kernels
        # compute the whole layer (the slice only saves work in the reference); the values of the kept positions are the same.
        k = self._config["output_range"]
        if k is not None:
            outs[-1] = outs[-1][:, :int(k)].contiguous()
        return dict(sequence_output=outs[-1], pooled_output=sess.pooled_output(), encoder_outputs=outs)

    # ------------------------------------------------------------------ accessors
    def get_embedding_table(self):
        return self.store.seg("word_embeddings")

    def get_embedding_layer(self):
        return self

    @property
    def embeddings(self):
        return self.get_embedding_table()

    def get_config(self):
        return dict(self._config)

    @property
    def transformer_layers(self):
        """One dict of weight views per Transformer layer."""
        views = self.store.tf_views()
        return [{k.split(f"layer_{i}/", 1)[1]: v for k, v in views.items() if f"transformer/layer_{i}/" in k}
                for i in range(self._config["num_layers"])]

    @property
    def pooler_layer(self):
        return {"kernel": self.store.seg("pooler/w"), "bias": self.store.seg("pooler/b")}

    def get_weights(self):
        return self.store.state_dict()

    def set_weights(self, sd):
        self.store.load_state_dict(sd)

    @classmethod
    def from_config(cls, config, custom_objects=None):
        return cls(**config)
