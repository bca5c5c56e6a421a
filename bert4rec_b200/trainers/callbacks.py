"""The two Keras callbacks the reference's trainer / examples rely on (bert4rec_trainer.py:46-52)."""
import pathlib


class Callback:
    def set_model(self, model):
        self.model = model


class ModelCheckpoint(Callback):
    def __init__(self, filepath, save_weights_only=True, monitor="val_loss", save_best_only=False, mode="auto"):
        self.filepath = pathlib.Path(str(filepath))
        self.monitor, self.save_best_only = monitor, save_best_only
        self.maximize = mode == "max" or (mode == "auto" and ("acc" in monitor or monitor.startswith("fmeasure")))
        self.best = None

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if self.save_best_only:
            if cur is None:
                return
            if self.best is not None and not (cur > self.best if self.maximize else cur < self.best):
                return
            self.best = cur
        self.model.save_weights(self.filepath)


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", patience=0, min_delta=0.0, mode="auto", restore_best_weights=False):
        self.monitor, self.patience, self.min_delta = monitor, patience, min_delta
        self.maximize = mode == "max" or (mode == "auto" and "acc" in monitor)
        self.best, self.wait = None, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        better = self.best is None or (cur > self.best + self.min_delta if self.maximize else cur < self.best - self.min_delta)
        if better:
            self.best, self.wait = cur, 0
        else:
            self.wait += 1
            if self.wait > self.patience:
                self.model.stop_training = True
