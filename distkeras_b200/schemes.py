"""Optimisation schemes: hyper-parameter meta loops around a trainer
(``distkeras/schemes.py``: ``Scheme``, ``Emperor``)."""
from __future__ import annotations


class Scheme:
    """Base scheme (``schemes.py:13-42``)."""

    def __init__(self, optimizer, num_epoch=15, evaluation_frequency=5):
        self.optimizer = optimizer
        self.num_epoch = int(num_epoch)
        self.evaluation_frequency = int(evaluation_frequency)
        self.optimizer.set_num_epoch(self.evaluation_frequency)

    def initialize(self):
        pass

    def get_epoch_over_evaluation_frequency(self) -> int:
        return self.num_epoch // self.evaluation_frequency

    def optimize(self, training_set, validation_set):
        raise NotImplementedError


class Emperor(Scheme):
    """Divide the trainer's learning rate by 10 whenever the validation loss plateaus
    (``|delta loss| <= loss_threshold``; ``schemes.py:45-88``)."""

    def __init__(self, optimizer, evaluate_loss, num_epoch=15, evaluation_frequency=5, loss_threshold=0.005):
        super().__init__(optimizer, num_epoch, evaluation_frequency)
        self.loss = evaluate_loss
        self.loss_threshold = float(loss_threshold)
        self.log = []

    def optimize(self, training_set, validation_set):
        trained_model = None
        previous_loss = float("inf")
        for i in range(self.get_epoch_over_evaluation_frequency() + 1):
            trained_model = self.optimizer.train(training_set)
            self.optimizer.set_model(trained_model)
            loss = float(self.loss(trained_model, validation_set))
            loss_delta = abs(loss - previous_loss)
            lr = self.optimizer.get_learning_rate()
            self.log.append({"round": i, "loss": loss, "learning_rate": lr})
            if loss_delta <= self.loss_threshold:
                self.optimizer.set_learning_rate(lr / 10.0)
            previous_loss = loss
        return trained_model
