"""Evaluators (``distkeras/evaluators.py``): ``Evaluator``, ``AccuracyEvaluator``."""
from __future__ import annotations

import torch

from .data import Dataset


class Evaluator:
    def __init__(self, label_col="label", prediction_col="prediction"):
        self.label_column = label_col
        self.prediction_column = prediction_col

    def evaluate(self, dataframe: Dataset) -> float:
        raise NotImplementedError


def _to_index(t: torch.Tensor) -> torch.Tensor:
    if t.dim() > 1 and t.shape[-1] > 1:
        return t.argmax(dim=-1)
    return t.reshape(-1).round().long() if t.dtype.is_floating_point else t.reshape(-1).long()


class AccuracyEvaluator(Evaluator):
    """``count(prediction == label) / count`` (``evaluators.py:28-48``) as one reduction."""

    def evaluate(self, dataframe: Dataset) -> float:
        n = dataframe.count()
        if n == 0:
            return 0.0
        pred = _to_index(dataframe[self.prediction_column])
        label = _to_index(dataframe[self.label_column])
        return float((pred == label).sum().item()) / float(n)


class F1Evaluator(Evaluator):
    """Binary F1 (the metric the reference's Higgs notebooks report,
    ``examples/example_1_analysis.ipynb:449-450``)."""

    def __init__(self, label_col="label", prediction_col="prediction", positive=1):
        super().__init__(label_col, prediction_col)
        self.positive = positive

    def evaluate(self, dataframe: Dataset) -> float:
        pred = _to_index(dataframe[self.prediction_column]) == self.positive
        label = _to_index(dataframe[self.label_column]) == self.positive
        tp = float((pred & label).sum())
        fp = float((pred & ~label).sum())
        fn = float((~pred & label).sum())
        return 0.0 if tp == 0 else 2 * tp / (2 * tp + fp + fn)
