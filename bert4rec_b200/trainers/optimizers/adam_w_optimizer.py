"""AdamWeightDecay + WarmUp schedule objects (reference: bert4rec/trainers/optimizers/adam_w_optimizer.py:6-168).

These objects only carry hyper-parameters: the update itself (global-norm clip 5.0, WarmUp/PolynomialDecay learning
rate evaluated on the device from the step counter, decoupled weight decay on everything whose name does not match
``exclude_from_weight_decay``, Adam in the epsilon-hat form) is ONE fused multi-tensor CUDA launch over the flat
parameter buffer (``b4r_adamw_step``)."""
import re

import numpy as np

from bert4rec_b200 import _lib


class PolynomialDecay:
    """keras PolynomialDecay with power 1 and no cycling: linear from initial to end over decay_steps."""

    def __init__(self, initial_learning_rate, decay_steps, end_learning_rate=0.0, power=1.0):
        if power != 1.0:
            raise NotImplementedError("only power=1 (linear decay)")
        self.initial_learning_rate, self.decay_steps, self.end_learning_rate, self.power = \
            initial_learning_rate, decay_steps, end_learning_rate, power

    def __call__(self, step):
        f32 = np.float32
        s = min(f32(step), f32(self.decay_steps))
        return float((f32(self.initial_learning_rate) - f32(self.end_learning_rate)) * (f32(1) - s / f32(self.decay_steps))
                     + f32(self.end_learning_rate))


class WarmUp:
    """Linear warm-up ``init_lr * step / warmup_steps`` for step < warmup_steps (so lr(0) == 0), then the wrapped
    decay schedule evaluated at the SAME step (adam_w_optimizer.py:22-36)."""

    def __init__(self, initial_learning_rate, decay_schedule_fn, warmup_steps, power=1.0, name=None):
        if power != 1.0:
            raise NotImplementedError("only power=1 (linear warm-up)")
        self.initial_learning_rate, self.decay_schedule_fn, self.warmup_steps, self.power, self.name = \
            initial_learning_rate, decay_schedule_fn, warmup_steps, power, name

    def __call__(self, step):
        f32 = np.float32
        if f32(step) < f32(self.warmup_steps):
            return float(f32(self.initial_learning_rate) * (f32(step) / f32(self.warmup_steps)))
        return self.decay_schedule_fn(step)

    def get_config(self):
        return {"initial_learning_rate": self.initial_learning_rate, "decay_schedule_fn": self.decay_schedule_fn,
                "warmup_steps": self.warmup_steps, "power": self.power, "name": self.name}


_DEFAULT_EXCLUDE = ["LayerNorm", "layer_norm", "bias"]


class AdamWeightDecay:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False,
                 weight_decay_rate=0.0, include_in_weight_decay=None, exclude_from_weight_decay=None,
                 gradient_clip_norm=5.0, name="AdamWeightDecay", **kwargs):
        if amsgrad:
            raise NotImplementedError("amsgrad")
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.weight_decay_rate = weight_decay_rate
        self.gradient_clip_norm = gradient_clip_norm
        self._include_in_weight_decay = include_in_weight_decay
        self._exclude_from_weight_decay = exclude_from_weight_decay
        self.name = name
        self.iterations = 0  # host mirror; the authoritative counter lives on the device (ParamStore.step_counter)

    def _do_use_weight_decay(self, param_name):
        """Name-regex rule of the reference (adam_w_optimizer.py:154-168)."""
        if self.weight_decay_rate == 0:
            return False
        for r in (self._include_in_weight_decay or []):
            if re.search(r, param_name) is not None:
                return True
        for r in (self._exclude_from_weight_decay or []):
            if re.search(r, param_name) is not None:
                return False
        return True

    def check_layout(self, tf_names_by_group):
        """The flat layout decays exactly {embedding tables, kernels}; verify the configured regex rule agrees."""
        for name, group in tf_names_by_group.items():
            if group == 2:
                continue
            want = self._do_use_weight_decay(name)
            if self.weight_decay_rate != 0 and want != (group == 0):
                raise NotImplementedError(f"weight-decay rule for {name!r} differs from the fused layout "
                                          f"(exclude_from_weight_decay={self._exclude_from_weight_decay})")

    def hparams_struct(self):
        lr = self.learning_rate
        if isinstance(lr, WarmUp):
            init_lr, warm = lr.initial_learning_rate, lr.warmup_steps
            dec = lr.decay_schedule_fn
        elif isinstance(lr, PolynomialDecay):
            init_lr, warm, dec = lr.initial_learning_rate, 0, lr
        elif isinstance(lr, (int, float)):
            init_lr, warm, dec = float(lr), 0, None
        else:
            raise NotImplementedError(f"learning-rate schedule {type(lr).__name__}")
        if dec is None:
            steps, end = 1 << 60, float(init_lr)  # constant
        else:
            steps, end = dec.decay_steps, dec.end_learning_rate
        return _lib.AdamWHParams(init_lr=float(init_lr), end_lr=float(end), num_train_steps=int(steps),
                                 num_warmup_steps=int(warm), weight_decay_rate=float(self.weight_decay_rate),
                                 beta_1=float(self.beta_1), beta_2=float(self.beta_2), epsilon=float(self.epsilon),
                                 clip_norm=float(self.gradient_clip_norm))

    def get_config(self):
        return {"name": self.name, "beta_1": self.beta_1, "beta_2": self.beta_2, "epsilon": self.epsilon,
                "weight_decay_rate": self.weight_decay_rate, "gradient_clip_norm": self.gradient_clip_norm}
