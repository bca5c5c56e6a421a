"""Import the reference's *own* host-side Python modules under a stub ``tensorflow``.

Used ONLY by ``oracle/gen_golden.py`` in the build container (``/root/reference``
does not exist on the GPU box).  The reference's masking / sampler / metric /
tokenizer code is pure Python + numpy but does ``import tensorflow as tf`` at
module scope; TensorFlow is not installable here, so a minimal stub module is
placed in ``sys.modules`` first.  Only attributes those modules touch at import
time or on the code paths we execute are provided.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, name="stub"):
        self.__name = name

    def __getattr__(self, item):
        return _Anything(f"{self.__name}.{item}")

    def __call__(self, *a, **k):
        return _Anything(f"{self.__name}()")

    def __mro_entries__(self, bases):
        return (object,)


class _Tensor:  # tf.Tensor stand-in (isinstance checks in simple_tokenizer.py:44,65)
    pass


class _RaggedTensor:
    pass


def _make_tf_stub():
    tf = types.ModuleType("tensorflow")
    tf.Tensor = _Tensor
    tf.RaggedTensor = _RaggedTensor
    tf.constant = lambda v, dtype=None: np.array(v, dtype=dtype)  # dataloader_utils.py:265
    tf.int64 = np.int64
    tf.int32 = np.int32
    tf.float32 = np.float32
    tf.string = str
    tf.bool = bool

    def _mod_getattr(name):
        return _Anything(f"tf.{name}")

    tf.__getattr__ = _mod_getattr

    class _Layer:  # so that class statements deriving from keras bases import fine
        def __init__(self, *a, **k):
            pass

    keras = types.ModuleType("tensorflow.keras")
    keras.__getattr__ = lambda name: _Anything(f"tf.keras.{name}")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer = _Layer
    layers.__getattr__ = lambda name: _Anything(f"tf.keras.layers.{name}")
    keras.layers = layers
    keras.Model = _Layer
    losses = types.ModuleType("tensorflow.keras.losses")
    losses.Loss = _Layer
    losses.__getattr__ = lambda name: _Anything(f"tf.keras.losses.{name}")
    keras.losses = losses
    metrics = types.ModuleType("tensorflow.keras.metrics")
    metrics.Metric = _Layer
    metrics.__getattr__ = lambda name: _Anything(f"tf.keras.metrics.{name}")
    keras.metrics = metrics
    tf.keras = keras

    pf = types.ModuleType("tensorflow.python")
    fw = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")
    ops.Tensor = _Tensor
    ops.EagerTensor = _Tensor
    mods = {
        "tensorflow": tf,
        "tensorflow.keras": keras,
        "tensorflow.python": pf,
        "tensorflow.python.framework": fw,
        "tensorflow.python.framework.ops": ops,
    }
    for name in ("tensorflow_models", "wget", "zstandard"):
        m = types.ModuleType(name)
        m.__getattr__ = (lambda n: (lambda attr: _Anything(f"{n}.{attr}")))(name)
        mods[name] = m
    return mods


_loaded = None


def load_reference():
    """Returns a namespace with the reference's host-side modules imported."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} not present (only available in the build container)")
    os.environ.setdefault("VIRTUAL_ENV", "/tmp")  # bert4rec/utils/utils.py:10
    for name, mod in _make_tf_stub().items():
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib

        ns = types.SimpleNamespace()
        ns.dataloader_utils = importlib.import_module("bert4rec.dataloaders.dataloader_utils")
        ns.samplers = importlib.import_module("bert4rec.dataloaders.samplers")
        ns.metrics = importlib.import_module("bert4rec.evaluation.evaluation_metrics")
        ns.tokenizers = importlib.import_module("bert4rec.tokenizers")
        ns.preprocessor = importlib.import_module(
            "bert4rec.dataloaders.preprocessors.bert4rec_preprocessor")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _loaded = ns
    return ns


if __name__ == "__main__":
    ns = load_reference()
    print("loaded:", [k for k in vars(ns)])
