"""Helpers around the model (reference: bert4rec/models/model_utils.py)."""
import pathlib

import numpy as np


def determine_model_path(path: pathlib.Path, mode: int = 0) -> pathlib.Path:
    """mode 0: path relative to the default save directory, 1: relative to the project root, 2: absolute."""
    from bert4rec_b200 import utils
    path = pathlib.Path(path)
    if mode == 0:
        return utils.get_project_root().joinpath(utils.get_default_model_save_path()).joinpath(path)
    if mode == 1:
        return utils.get_project_root().joinpath(path)
    if mode == 2:
        return path
    raise ValueError(f"The mode argument must be 0, 1 or 2 (given: {mode})")


def rank_items(logits, embeddings, items):
    """Standalone helper of the reference (model_utils.py:41-64): scores ``items`` for every row of ``logits``
    [n, H] against ``embeddings`` [V, H], softmax over the candidates, stable descending order.
    Returns (rankings [n, C] of item ids, probabilities [n, C]).  Small host-side utility (numpy)."""
    logits = np.asarray(logits, dtype=np.float32)
    emb = np.asarray(embeddings, dtype=np.float32)
    items = np.asarray(items, dtype=np.int64)
    assert logits.ndim == 2 and emb.ndim == 2 and logits.shape[1] == emb.shape[1], "shape mismatch"
    assert items.ndim == 1 or (items.ndim == 2 and items.shape[0] == logits.shape[0]), "items shape mismatch"
    cand = np.broadcast_to(items, (logits.shape[0], items.shape[-1])) if items.ndim == 1 else items
    scores = np.einsum("nh,nch->nc", logits, emb[cand])
    e = np.exp(scores - scores.max(-1, keepdims=True))
    probs = e / e.sum(-1, keepdims=True)
    order = np.argsort(-probs, axis=-1, kind="stable")
    return np.take_along_axis(cand, order, -1), probs
