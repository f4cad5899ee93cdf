"""Checkpoint / resume (absent from the reference, README TODO "Save/Load Keras model").

A checkpoint is one ``torch.save`` file holding the model spec + flat center variable, the PS
counters and, optionally, per-worker optimizer state and the iteration count; ``resume`` gives the
trainer a warm start (``trainer.set_model`` + counters), exactly how ``Emperor`` chains runs
(``distkeras/schemes.py:73``).
"""
from __future__ import annotations

import os
import time
from typing import Optional

import torch

from . import deserialize_keras_model, serialize_keras_model

FORMAT_VERSION = 1


def save_checkpoint(path: str, model, num_updates: int = 0, iteration: int = 0, optimizer_state: Optional[dict] = None,
                    history=None, extra: Optional[dict] = None) -> str:
    payload = {"version": FORMAT_VERSION, "time": time.time(), "model": serialize_keras_model(model),
               "num_updates": int(num_updates), "iteration": int(iteration), "optimizer_state": optimizer_state,
               "history": history, "extra": extra or {}, "rng": torch.get_rng_state()}
    tmp = path + ".tmp"
    torch.save(payload, tmp)
    os.replace(tmp, path)  # atomic: a crash never leaves a torn checkpoint
    return path


def load_checkpoint(path: str) -> dict:
    try:
        payload = torch.load(path, weights_only=True)   # plain containers + tensors only
    except Exception:  # checkpoints carrying user objects in `extra` / `history` (trusted, locally written)
        payload = torch.load(path, weights_only=False)
    if payload.get("version") != FORMAT_VERSION:
        raise ValueError(f"unsupported checkpoint version {payload.get('version')}")
    payload["model"] = deserialize_keras_model(payload["model"])
    return payload


def resume_trainer(trainer, path: str) -> dict:
    """Warm-start ``trainer`` from a checkpoint; returns the payload (counters, history)."""
    payload = load_checkpoint(path)
    trainer.set_model(payload["model"])
    trainer.resumed_from = {"path": path, "num_updates": payload["num_updates"], "iteration": payload["iteration"]}
    return payload
