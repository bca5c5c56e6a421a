"""BERT4RecTrainer: compile with the BERT4Rec defaults + fit with an optional best-only weight checkpoint
(reference: bert4rec/trainers/bert4rec_trainer.py:9-71)."""
import pathlib

from absl import logging

from . import optimizers, trainer_utils, callbacks as _cb
from .base_trainer import BaseTrainer


class BERT4RecTrainer(BaseTrainer):
    def __init__(self, model):
        super().__init__(model)

    def initialize_model(self, optimizer=None, loss=None, metrics: list = None):
        if optimizer is None:
            optimizer = optimizers.get("adamw")
        if loss is None:
            loss = trainer_utils.MaskedSparseCategoricalCrossentropy()
        if metrics is None:
            metrics = [trainer_utils.SparseCategoricalAccuracy(), trainer_utils.masked_accuracy]
        self.optimizer, self.loss, self.metrics = optimizer, loss, metrics
        self.model.compile(optimizer=optimizer, loss=loss, metrics=metrics)

    def train(self, train_ds, val_ds, checkpoint_path: pathlib.Path = None, epochs: int = 50,
              steps_per_epoch: int = None, validation_steps: int = None):
        if checkpoint_path:
            checkpoint_path = pathlib.Path(checkpoint_path)
            self.append_callback(_cb.ModelCheckpoint(filepath=checkpoint_path, save_weights_only=True,
                                                     monitor="val_masked_accuracy", save_best_only=True))
            # resume from the latest weights; optimizer slots are deliberately not restored (reference :57-58)
            existing = checkpoint_path if checkpoint_path.suffix == ".npz" else checkpoint_path.with_name(checkpoint_path.name + ".npz")
            if existing.is_file():
                self.model.load_weights(existing)
        logging.info("Start training")
        return self.model.fit(x=train_ds, validation_data=val_ds, epochs=epochs, callbacks=self.callbacks,
                              steps_per_epoch=steps_per_epoch, validation_steps=validation_steps)

    def validate(self):
        pass
