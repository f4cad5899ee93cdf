"""Model replicas: the object a worker trains (``model.train_on_batch`` in the reference,
``distkeras/workers.py:199-202, 327-342``).

``TorchReplica`` is the portable executor (autograd; CPU oracle, and the GPU path for layer types
the native engine does not lower yet).  ``NativeReplica`` (``parallel/engine.py``) is the sm_100a
executor.  Both expose the same surface -- a flat fp32 weight buffer ``W`` plus
``train_on_batch`` -- so every parameter-server algorithm is written once against it.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..models.core import Sequential, apply_deferred, compute_accuracy, compute_loss
from ..ops.flat_optim import FlatOptimizer


class Replica:
    """Interface shared by the executors."""

    model: Sequential
    W: torch.Tensor  # flat fp32 parameters (the `get_weights()` analogue)

    def train_on_batch(self, x, y) -> Tuple[float, float]:  # pragma: no cover - interface
        raise NotImplementedError

    def get_flat(self) -> torch.Tensor:
        return self.W.detach()

    def set_flat(self, flat: torch.Tensor) -> None:
        with torch.no_grad():
            self.W.copy_(flat.to(self.W.device))

    def weights_changed(self) -> None:
        """Hook: call after writing ``W`` from outside (refreshes derived copies)."""


class TorchReplica(Replica):
    def __init__(self, model: Sequential, optimizer, loss: str, device=None, share_model_buffer: bool = False,
                 seed: Optional[int] = None):
        model.build()
        self.model = model
        self.loss = loss
        self.device = torch.device(device) if device is not None else model.get_flat_weights().device
        if share_model_buffer:
            self.W = model.flat
        else:
            self.W = model.get_flat_weights().detach().clone().to(self.device)
        self.W.requires_grad_(True)
        self.opt = FlatOptimizer(optimizer, self.W.numel(), self.W.device, mask=model.trainable_mask())
        self.from_logits = model.logits_tail() is not None
        self.iteration = 0
        if seed is not None:
            torch.manual_seed(seed)

    def train_on_batch(self, x, y) -> Tuple[float, float]:
        y = y if isinstance(y, torch.Tensor) else torch.as_tensor(y)
        y = y.to(self.W.device)
        ctx: dict = {}
        out = self.model.forward(x, flat=self.W, training=True, logits=True, ctx=ctx)
        loss = compute_loss(self.loss, out, y, self.from_logits)
        if self.W.grad is not None:
            self.W.grad = None
        loss.backward()
        with torch.no_grad():
            acc = compute_accuracy(out.detach(), y)
            apply_deferred(ctx)
            self.opt.step(self.W.data, self.W.grad)
        self.iteration += 1
        return float(loss.detach()), float(acc)

    @torch.no_grad()
    def predict(self, x) -> torch.Tensor:
        return self.model.forward(x, flat=self.W, training=False)

    def set_learning_rate(self, lr: float) -> None:
        self.opt.set_learning_rate(lr)
