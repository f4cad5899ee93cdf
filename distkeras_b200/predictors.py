"""Predictors: append a prediction column to a dataset.

Capability parity with ``distkeras/predictors.py`` (``Predictor``, ``ModelPredictor``).  The
reference calls ``model.predict`` once per row inside ``mapPartitions`` (``predictors.py:56-62``);
here inference is batched: on CUDA the forward pass runs through the native engine (tcgen05 GEMMs,
``parallel/engine.py``), otherwise through the autograd executor.
"""
from __future__ import annotations

import torch

from .data import Dataset
from .models.core import Sequential
from .utils import deserialize_keras_model, serialize_keras_model


class Predictor:
    def __init__(self, keras_model: Sequential):
        self.model = serialize_keras_model(keras_model)

    def predict(self, dataframe: Dataset) -> Dataset:
        raise NotImplementedError


class ModelPredictor(Predictor):
    def __init__(self, keras_model, features_col="features", output_col="prediction", batch_size=8192,
                 device=None):
        super().__init__(keras_model)
        self.features_column = features_col
        self.output_column = output_col
        self.batch_size = int(batch_size)
        self.device = device

    def _predict_tensor(self, x: torch.Tensor) -> torch.Tensor:
        model = deserialize_keras_model(self.model)
        device = self.device or ("cuda" if torch.cuda.is_available() else "cpu")
        if str(device).startswith("cuda"):
            from .parallel.engine import native_predict

            out = native_predict(model, x, self.batch_size, device)
            if out is not None:
                return out
        model.to(device)
        outs = []
        with torch.no_grad():
            for i in range(0, x.shape[0], self.batch_size):
                outs.append(model.forward(x[i:i + self.batch_size], training=False).float().cpu())
        return torch.cat(outs, dim=0) if outs else torch.empty(0)

    def predict(self, dataframe: Dataset) -> Dataset:
        return dataframe.with_column(self.output_column, self._predict_tensor(dataframe[self.features_column]))
