"""Loss / metric objects of the training step (reference: bert4rec/trainers/trainer_utils.py:4-60).

Inside ``BERT4RecModel.train_step`` these are fused into the projection kernel (``b4r_mlm_loss``); the callables
below evaluate the same definitions on materialised logits (torch tensors, any device) for tests and ad-hoc use."""
import torch


def _first_argmax(x):
    mx = x.max(dim=-1, keepdim=True).values
    V = x.shape[-1]
    idx = torch.arange(V, device=x.device).expand_as(x)
    return torch.where(x == mx, idx, torch.full_like(idx, V)).min(dim=-1).values


class MaskedSparseCategoricalCrossentropy:
    """sum over slots with y_true != pad_token of (logsumexp(logits) - logits[y]) / #such slots."""

    def __init__(self, pad_token: int = 0, reduction=None, name: str = None):
        self.pad_token, self.reduction, self.name = pad_token, reduction, name or "masked_sparse_categorical_crossentropy"

    def __call__(self, y_true, y_pred):
        return self.call(y_true, y_pred)

    def call(self, y_true, y_pred):
        y_true = torch.as_tensor(y_true).to(y_pred.device)
        mask = (y_true != self.pad_token).to(y_pred.dtype)
        per = torch.logsumexp(y_pred, dim=-1) - torch.gather(y_pred, -1, y_true.long().unsqueeze(-1)).squeeze(-1)
        return (per * mask).sum() / mask.sum()


def masked_accuracy(y_true, y_pred):
    y_true = torch.as_tensor(y_true).to(y_pred.device).long()
    pred = _first_argmax(y_pred)
    mask = y_true != 0
    return ((pred == y_true) & mask).float().sum() / mask.float().sum()


class SparseCategoricalAccuracy:
    """keras.metrics.SparseCategoricalAccuracy stand-in (name only matters to compile())."""
    name = "sparse_categorical_accuracy"

    def __call__(self, y_true, y_pred):
        y_true = torch.as_tensor(y_true).to(y_pred.device).long()
        return (_first_argmax(y_pred) == y_true).float().mean()


class MaskedAccuracyMetric:
    name = "masked_accuracy"

    def __init__(self, pad_token: int = 0):
        self.pad_token, self.total = pad_token, None

    def update_state(self, y_true, y_pred, sample_weight=None):
        self.total = masked_accuracy(y_true, y_pred)

    def result(self):
        return self.total
