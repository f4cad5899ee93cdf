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
    """``count(prediction == label) / count`` (``evaluators.py:28-48``) as one reduction.  A column of probability
    vectors on a GPU box is reduced by ``dk_label_index`` (arg-max + compare + count in one pass over HBM)."""

    def evaluate(self, dataframe: Dataset) -> float:
        n = dataframe.count()
        if n == 0:
            return 0.0
        pred = dataframe[self.prediction_column]
        label = _to_index(dataframe[self.label_column])
        if torch.cuda.is_available() and pred.dim() == 2 and pred.shape[1] > 1 and pred.dtype.is_floating_point:
            return self._evaluate_cuda(pred, label) / float(n)
        return float((_to_index(pred) == label).sum().item()) / float(n)

    @staticmethod
    def _evaluate_cuda(pred: torch.Tensor, label: torch.Tensor) -> int:
        from . import _native as N

        p = pred.to("cuda", torch.float32).contiguous()
        y = label.to("cuda", torch.int32).contiguous()
        idx = torch.empty(p.shape[0], dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        # a threshold no probability reaches: the rule degenerates to the arg-max; rows without a positive entry
        # come back as -1 and are resolved with a plain arg-max (never happens for softmax / sigmoid outputs)
        N.check(N.lib().dk_label_index(p.data_ptr(), p.shape[0], p.shape[1], float("inf"), -1, idx.data_ptr(),
                                       y.data_ptr(), cnt.data_ptr(), N.current_stream()), "dk_label_index")
        correct = int(cnt.item())
        odd = idx < 0
        if bool(odd.any()):
            correct += int((p[odd].argmax(dim=1).to(torch.int32) == y[odd]).sum().item())
        return correct


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
