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
    """``index_col`` (an extension): also write the class index of every prediction -- the
    :class:`~distkeras_b200.transformers.LabelIndexTransformer` rule with ``activation_threshold`` /
    ``default_index`` -- computed by ``dk_label_index`` in the same pass on the GPU."""

    def __init__(self, keras_model, features_col="features", output_col="prediction", batch_size=8192,
                 device=None, index_col=None, activation_threshold=0.55, default_index=0):
        super().__init__(keras_model)
        self.features_column = features_col
        self.output_column = output_col
        self.batch_size = int(batch_size)
        self.device = device
        self.index_column = index_col
        self.activation_threshold = float(activation_threshold)
        self.default_index = int(default_index)
        self._indices = None

    def _predict_tensor(self, x: torch.Tensor) -> torch.Tensor:
        model = deserialize_keras_model(self.model)
        device = self.device or ("cuda" if torch.cuda.is_available() else "cpu")
        if str(device).startswith("cuda") and isinstance(model, Sequential):
            from .parallel.engine import native_predict

            rule = (self.activation_threshold, self.default_index) if self.index_column else None
            out = native_predict(model, x, self.batch_size, device, label_index=rule)
            if out is not None:
                if rule is not None:
                    out, self._indices = out
                return out
        model.to(device)
        out = model.predict(x, batch_size=self.batch_size)
        return [torch.as_tensor(o) for o in out] if isinstance(out, list) else torch.as_tensor(out)

    def predict(self, dataframe: Dataset) -> Dataset:
        """One prediction column per model output: a multi-output functional model takes a list ``output_col`` (or
        gets ``<output_col>_<k>``); several feature columns feed a multi-input model one per input and a
        single-input model concatenated (the trainers' rule)."""
        cols = self.features_column
        if isinstance(cols, (list, tuple)):
            model = deserialize_keras_model(self.model)
            xs = [dataframe[c] for c in cols]
            if getattr(model, "num_inputs", 1) > 1:
                x = xs
            else:
                x = xs[0] if len(xs) == 1 else torch.cat([t.reshape(t.shape[0], -1).float() for t in xs], dim=1)
        else:
            x = dataframe[cols]
        self._indices = None
        out = self._predict_tensor(x)
        if self.index_column and not isinstance(out, (list, tuple)):
            if self._indices is None:    # no native lowering on this device: the transformer computes the same rule
                from .transformers import LabelIndexTransformer

                self._indices = LabelIndexTransformer(out.shape[-1], default_index=self.default_index,
                                                      activation_threshold=self.activation_threshold).indices(out)
            dataframe = dataframe.with_column(self.index_column, self._indices.to(torch.float32))
        if isinstance(out, (list, tuple)):
            names = self.output_column if isinstance(self.output_column, (list, tuple)) \
                else [f"{self.output_column}_{k}" for k in range(len(out))]
            for name, o in zip(names, out):
                dataframe = dataframe.with_column(name, o)
            return dataframe
        return dataframe.with_column(self.output_column, out)
