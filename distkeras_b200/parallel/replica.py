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


def _per_output(spec, n: int, names, default):
    """Keras accepts one value, a list (one per output) or a dict keyed by output name."""
    if spec is None:
        return [default] * n
    if isinstance(spec, dict):
        return [spec.get(nm, default) for nm in names]
    if isinstance(spec, (list, tuple)):
        if len(spec) != n:
            raise ValueError(f"expected {n} entries (one per model output), got {len(spec)}")
        return list(spec)
    return [spec] * n


class TorchReplica(Replica):
    """Autograd executor.  Single-output models: ``train_on_batch -> (loss, accuracy)``.  Models with several
    outputs (``models.functional.Model``; the reference's list-valued ``label_col`` with ``loss_weights`` /
    ``metrics``, ``distkeras/workers.py:75-76, 117-118``): ``[total, loss_1 .. loss_k, metric_1 .. metric_k]`` with
    ``total = sum_i loss_weights[i] * loss_i`` -- the Keras ``train_on_batch`` record."""

    def __init__(self, model, optimizer, loss, device=None, share_model_buffer: bool = False,
                 seed: Optional[int] = None, loss_weights=None, metrics=("accuracy",)):
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
        self.n_out = int(getattr(model, "num_outputs", 1))
        names = [getattr(o, "name", f"output_{i}") for i, o in enumerate(getattr(model, "outputs", [None]))]
        self.losses = _per_output(loss, self.n_out, names, "categorical_crossentropy")
        self.loss_weights = [float(w) for w in _per_output(loss_weights, self.n_out, names, 1.0)]
        self.metrics = list(metrics or [])
        tails = model.logits_tails() if hasattr(model, "logits_tails") else [model.logits_tail()]
        self.from_logits_each = [t is not None for t in tails]
        self.from_logits = self.from_logits_each[0]
        self.iteration = 0
        if seed is not None:
            torch.manual_seed(seed)

    def train_on_batch(self, x, y):
        ys = list(y) if (self.n_out > 1 and isinstance(y, (list, tuple))) else [y]
        if len(ys) != self.n_out:
            raise ValueError(f"the model has {self.n_out} outputs, got {len(ys)} label arrays")
        ys = [(t if isinstance(t, torch.Tensor) else torch.as_tensor(t)).to(self.W.device) for t in ys]
        ctx: dict = {}
        out = self.model.forward(x, flat=self.W, training=True, logits=True, ctx=ctx)
        outs = out if isinstance(out, list) else [out]
        parts = [compute_loss(l, o, t, fl) for l, o, t, fl in zip(self.losses, outs, ys, self.from_logits_each)]
        loss = parts[0] * self.loss_weights[0]
        for p_, w in zip(parts[1:], self.loss_weights[1:]):
            loss = loss + p_ * w
        if self.W.grad is not None:
            self.W.grad = None
        loss.backward()
        with torch.no_grad():
            accs = [compute_accuracy(o.detach(), t) for o, t in zip(outs, ys)]
            apply_deferred(ctx)
            self.opt.step(self.W.data, self.W.grad)
        self.iteration += 1
        if self.n_out == 1:
            return float(loss.detach()), float(accs[0])
        rec = [float(loss.detach())] + [float(p_.detach()) for p_ in parts]
        if any(m in ("accuracy", "acc") for m in self.metrics):
            rec += [float(a) for a in accs]
        return rec

    @torch.no_grad()
    def predict(self, x) -> torch.Tensor:
        return self.model.forward(x, flat=self.W, training=False)

    def set_learning_rate(self, lr: float) -> None:
        self.opt.set_learning_rate(lr)
