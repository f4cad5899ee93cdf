"""Dataset transformers.

Capability parity with ``distkeras/transformers.py``: the seven transformers with the same
constructor arguments and formulas.  The reference maps a Python lambda over every Spark Row
(``rdd.map(self._transform).toDF()``, one RDD pass per column for ``StandardTransformer``); here
each ``transform`` is one vectorised tensor expression over the whole column.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .data import Dataset


class Transformer:
    """Interface: ``transform(dataset) -> dataset`` (``transformers.py:26-32``)."""

    def transform(self, dataframe: Dataset) -> Dataset:
        raise NotImplementedError


class MinMaxTransformer(Transformer):
    """``x' = scale * (x - o_max) + n_max`` with ``scale = (n_max - n_min) / (o_max - o_min)``
    (``transformers.py:35-86``)."""

    def __init__(self, o_min, o_max, n_min, n_max, input_col, output_col, is_vector=True):
        self.o_min, self.o_max = float(o_min), float(o_max)
        self.n_min, self.n_max = float(n_min), float(n_max)
        self.scale = (self.n_max - self.n_min) / (self.o_max - self.o_min)
        self.input_column = input_col
        self.output_column = output_col
        self.is_vector = is_vector

    def affine(self):
        """(scale, shift) such that ``x' = scale * x + shift``; lets the native engine fuse this
        transformer into its input stage and ship raw uint8 pixels over PCIe."""
        return self.scale, self.n_max - self.scale * self.o_max

    def transform(self, dataframe: Dataset) -> Dataset:
        x = dataframe[self.input_column].to(torch.float32)
        return dataframe.with_column(self.output_column, self.scale * (x - self.o_max) + self.n_max)


class BinaryLabelTransformer(Transformer):
    """``value == label -> [1, 0] else [0, 1]`` (``transformers.py:89-125``)."""

    def __init__(self, input_column, output_column, label):
        self.input_column = input_column
        self.output_column = output_column
        self.label = label

    def transform(self, dataframe: Dataset) -> Dataset:
        col = dataframe[self.input_column]
        hit = (col == self.label).reshape(-1).to(torch.float32)
        return dataframe.with_column(self.output_column, torch.stack([hit, 1.0 - hit], dim=1))


class StandardTransformer(Transformer):
    """Per-column ``(x - mean) / stddev_pop`` into ``<col><suffix>`` (``transformers.py:128-194``)."""

    def __init__(self, columns: Sequence[str], suffix="_normalized"):
        self.columns = list(columns)
        self.column_suffix = suffix
        self.means = {}
        self.stddevs = {}

    def transform(self, dataframe: Dataset) -> Dataset:
        for c in self.columns:
            x = dataframe[c].to(torch.float32)
            mean = x.mean(dim=0)
            std = x.std(dim=0, unbiased=False)
            self.means[c], self.stddevs[c] = mean, std
            dataframe = dataframe.with_column(c + self.column_suffix, (x - mean) / std)
        return dataframe


class DenseTransformer(Transformer):
    """Sparse -> dense vectors (``transformers.py:197-225``).  Accepts torch sparse tensors or
    scipy sparse matrices; dense input passes through."""

    def __init__(self, input_col, output_col):
        self.input_column = input_col
        self.output_column = output_col

    def transform(self, dataframe: Dataset) -> Dataset:
        x = dataframe[self.input_column]
        if isinstance(x, torch.Tensor) and x.is_sparse:
            x = x.to_dense()
        return dataframe.with_column(self.output_column, x)


class ReshapeTransformer(Transformer):
    """Vector -> tensor of ``shape`` (e.g. ``(28, 28, 1)``; ``transformers.py:228-263``)."""

    def __init__(self, input_col, output_col, shape):
        self.input_column = input_col
        self.output_column = output_col
        self.shape = tuple(int(s) for s in shape)

    def transform(self, dataframe: Dataset) -> Dataset:
        x = dataframe[self.input_column]
        return dataframe.with_column(self.output_column, x.reshape((x.shape[0],) + self.shape))


class OneHotTransformer(Transformer):
    """Integer index -> one-hot vector of ``output_dim`` (``transformers.py:266-299``)."""

    def __init__(self, output_dim, input_col, output_col):
        self.output_dimensionality = int(output_dim)
        self.input_column = input_col
        self.output_column = output_col

    def transform(self, dataframe: Dataset) -> Dataset:
        idx = dataframe[self.input_column].reshape(-1).long()
        return dataframe.with_column(
            self.output_column, torch.nn.functional.one_hot(idx, self.output_dimensionality).to(torch.float32))


class LabelIndexTransformer(Transformer):
    """Prediction vector -> class index: the FIRST element >= ``activation_threshold``, otherwise
    the arg-max, otherwise ``default_index`` (``transformers.py:302-350``).  With a GPU present the
    column goes through the one-pass kernel ``dk_label_index`` (``csrc/loss_kernels.cu``); the PyTorch
    expression below is the CPU path and the oracle the kernel is tested against."""

    def __init__(self, output_dim, input_col="prediction", output_col="prediction_index", default_index=0,
                 activation_threshold=0.55):
        self.input_column = input_col
        self.output_column = output_col
        self.output_dimensionality = int(output_dim)
        self.activation_threshold = float(activation_threshold)
        self.default_index = int(default_index)

    def get_index(self, vector) -> int:
        return int(self.indices(torch.as_tensor(np.asarray(vector, dtype=np.float32)).reshape(1, -1))[0])

    def indices(self, p: torch.Tensor, device=None) -> torch.Tensor:
        p = p[:, :self.output_dimensionality].to(torch.float32)
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        if str(device).startswith("cuda") and p.shape[0] > 0:
            from . import _native as N

            pd = p.to(device).contiguous()
            out = torch.empty(pd.shape[0], dtype=torch.int32, device=pd.device)
            with torch.cuda.device(pd.device):
                N.check(N.lib().dk_label_index(pd.data_ptr(), pd.shape[0], pd.shape[1], self.activation_threshold,
                                               self.default_index, out.data_ptr(), None, None, N.current_stream()),
                        "dk_label_index")
            return out.to(torch.int64).cpu()
        above = p >= self.activation_threshold
        first_above = torch.where(above.any(dim=1), above.float().argmax(dim=1), torch.full((p.shape[0],), -1))
        best, arg = p.max(dim=1)
        # the reference's running max starts at 0.0: with no positive entry the default index wins
        fallback = torch.where(best > 0.0, arg, torch.full_like(arg, self.default_index))
        return torch.where(first_above >= 0, first_above, fallback)

    def transform(self, dataframe: Dataset) -> Dataset:
        return dataframe.with_column(self.output_column, self.indices(dataframe[self.input_column]).to(torch.float32))
