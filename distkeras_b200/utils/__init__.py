"""Utility functions: model (de)serialisation, history helpers, dataset helpers.

Capability parity with ``distkeras/utils.py`` (all public names kept), re-designed around the
flat-buffer model and the columnar :class:`~distkeras_b200.data.Dataset`.
"""
from __future__ import annotations

import getpass
import json
import os
import pickle
from typing import Dict, List, Sequence

import numpy as np
import torch

from ..data import Dataset, Row
from ..models.core import Sequential, model_from_json

_BASE_DIR = None


def get_os_username() -> str:
    """User name of the calling process (``distkeras/utils.py:28-33``)."""
    try:
        return getpass.getuser()
    except Exception:  # pragma: no cover - containers without passwd entries
        return str(os.getuid())


def set_keras_base_directory(base_dir: str = None) -> str:
    """Directory for framework scratch files; Keras' ``~/.keras`` analogue (``utils.py:36-38``)."""
    global _BASE_DIR
    _BASE_DIR = base_dir or os.path.join("/tmp", get_os_username())
    os.makedirs(_BASE_DIR, exist_ok=True)
    return _BASE_DIR


def get_base_directory() -> str:
    return _BASE_DIR or set_keras_base_directory()


def to_one_hot_encoded_dense(value, n_dim: int = 2) -> np.ndarray:
    """One-hot vector of length ``n_dim`` with a 1 at ``int(value)`` (``utils.py:41-52``)."""
    vector = np.zeros(n_dim)
    vector[int(value)] = 1.0
    return vector


def new_dataframe_row(old_row: Dict, column_name: str, column_value) -> Row:
    """Copy of ``old_row`` with one more column (``utils.py:55-59``)."""
    row = Row(old_row)
    row[column_name] = column_value
    return row


def json_to_dataframe_row(string: str) -> Row:
    """JSON record -> Row (streaming-inference path, ``utils.py:62-67``)."""
    return Row(json.loads(string))


def pickle_object(o) -> bytes:
    return pickle.dumps(o, -1)


def unpickle_object(string: bytes):
    return pickle.loads(string)


def serialize_keras_model(model: Sequential) -> dict:
    """``{'model': json spec, 'weights': [arrays]}`` -- same wire format idea as ``utils.py:80-86``;
    additionally carries the flat buffer so deserialisation is one memcpy."""
    model.build()
    return {"model": model.to_json(), "weights": model.get_weights(),
            "flat": model.get_flat_weights().detach().cpu().numpy().copy(),
            "compile": {"loss": model.loss, "metrics": list(model.metrics),
                        "optimizer": _serialize_optimizer(model.optimizer)}}


def _serialize_optimizer(opt):
    if opt is None or isinstance(opt, (str, dict)):
        return opt
    return opt.serialize()


def deserialize_keras_model(dictionary: dict) -> Sequential:
    """Inverse of :func:`serialize_keras_model` (``utils.py:121-128``)."""
    model = model_from_json(dictionary["model"])
    model.build()
    if dictionary.get("flat") is not None:
        model.set_flat_weights(torch.from_numpy(np.asarray(dictionary["flat"], dtype=np.float32)))
    else:
        model.set_weights(dictionary["weights"])
    comp = dictionary.get("compile") or {}
    if comp.get("loss"):
        model.compile(loss=comp["loss"], optimizer=comp.get("optimizer") or "sgd", metrics=comp.get("metrics") or [])
    return model


def history_executor(history: Sequence[dict], id: int) -> List[dict]:
    """History records of one worker, ordered by iteration (``utils.py:113-118``)."""
    executor_history = [h for h in history if h["worker_id"] == id]
    executor_history.sort(key=lambda x: x["iteration"])
    return executor_history


def history_executors_average(history: Sequence[dict]) -> List[np.ndarray]:
    """Per-iteration metrics averaged over all workers that reached the iteration.

    Intended behaviour of ``utils.py:89-110`` (the reference drops the last worker and assumes
    exactly two metrics; both fixed here).
    """
    if not history:
        return []
    workers = sorted({h["worker_id"] for h in history})
    histories = [history_executor(history, w) for w in workers]
    longest = max(len(h) for h in histories)
    averaged = []
    for i in range(longest):
        rows = [np.asarray(h[i]["history"], dtype=np.float64) for h in histories if len(h) > i]
        averaged.append(np.mean(rows, axis=0))
    return averaged


def uniform_weights(model: Sequential, constraints=(-0.5, 0.5)) -> None:
    """Re-initialise every parameter U(lo, hi) (``utils.py:131-158``), one vectorised fill."""
    model.build()
    lo, hi = constraints
    with torch.no_grad():
        model.flat.uniform_(float(lo), float(hi))


def shuffle(dataset: Dataset, seed=None) -> Dataset:
    """Random row permutation (``orderBy(rand())`` in ``utils.py:161-170``)."""
    return dataset.shuffle(seed)


def precache(dataset: Dataset, num_workers: int) -> Dataset:
    """Partition for ``num_workers`` and pin the host memory (``utils.py:173-186``)."""
    dataset = dataset.repartition(num_workers)
    dataset.pin_memory()
    dataset.count()
    return dataset
