"""Worker optimizers on the flat parameter buffer.

The reference hands ``worker_optimizer`` (a Keras optimizer name or object) to ``model.compile`` in
every worker (``distkeras/workers.py:57-61, 103-119``); the optimizer state lives in the worker and
is never reset by a pull.  :class:`FlatOptimizer` keeps that contract on one flat
buffer: ``step`` runs the fused sm_100a kernel (``csrc/optim_kernels.cu``) on CUDA tensors and the
identical math in PyTorch ops on CPU (the oracle the kernel is tested against).
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _native

# Keras defaults (keras/optimizers.py of the reference's era)
_DEFAULTS = {
    "sgd": dict(lr=0.01, momentum=0.0, decay=0.0, nesterov=False),
    "adagrad": dict(lr=0.01, epsilon=1e-7, decay=0.0),
    "rmsprop": dict(lr=0.001, rho=0.9, epsilon=1e-7, decay=0.0),
    "adam": dict(lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0),
    "adadelta": dict(lr=1.0, rho=0.95, epsilon=1e-7, decay=0.0),
    "adamax": dict(lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0),
    # Nesterov Adam (Dozat 2016) with a constant beta_1: Keras' momentum-decay schedule (schedule_decay) is
    # accepted and ignored
    "nadam": dict(lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0, schedule_decay=0.004),
}


class OptimizerSpec:
    """Serializable optimizer description (the ``{'class_name', 'config'}`` dict of Keras)."""

    def __init__(self, name: str, **config):
        name = name.lower()
        if name not in _DEFAULTS:
            raise ValueError(f"unsupported worker optimizer {name!r}; choose from {sorted(_DEFAULTS)}")
        self.name = name
        cfg = dict(_DEFAULTS[name])
        if "learning_rate" in config:
            config["lr"] = config.pop("learning_rate")
        unknown = set(config) - set(cfg)
        if unknown:
            raise ValueError(f"unknown {name} options: {sorted(unknown)}")
        cfg.update(config)
        self.config = cfg

    def serialize(self) -> dict:
        return {"class_name": self.name, "config": dict(self.config)}

    @classmethod
    def parse(cls, spec) -> "OptimizerSpec":
        if isinstance(spec, OptimizerSpec):
            return spec
        if isinstance(spec, str):
            return cls(spec)
        if isinstance(spec, dict):
            return cls(spec["class_name"], **spec.get("config", {}))
        if spec is None:
            return cls("sgd")
        raise TypeError(f"cannot interpret optimizer {spec!r}")

    def __repr__(self):
        return f"OptimizerSpec({self.name!r}, {self.config})"


def SGD(lr=0.01, momentum=0.0, decay=0.0, nesterov=False):
    return OptimizerSpec("sgd", lr=lr, momentum=momentum, decay=decay, nesterov=nesterov)


def Adagrad(lr=0.01, epsilon=1e-7, decay=0.0):
    return OptimizerSpec("adagrad", lr=lr, epsilon=epsilon, decay=decay)


def RMSprop(lr=0.001, rho=0.9, epsilon=1e-7, decay=0.0):
    return OptimizerSpec("rmsprop", lr=lr, rho=rho, epsilon=epsilon, decay=decay)


def Adam(lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0):
    return OptimizerSpec("adam", lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, decay=decay)


def Adadelta(lr=1.0, rho=0.95, epsilon=1e-7, decay=0.0):
    return OptimizerSpec("adadelta", lr=lr, rho=rho, epsilon=epsilon, decay=decay)


def Nadam(lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, schedule_decay=0.004):
    return OptimizerSpec("nadam", lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, schedule_decay=schedule_decay)


def Adamax(lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.0):
    return OptimizerSpec("adamax", lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, decay=decay)


class FlatOptimizer:
    """Optimizer state + step for one flat fp32 buffer."""

    def __init__(self, spec, numel: int, device, mask: Optional[torch.Tensor] = None):
        self.spec = OptimizerSpec.parse(spec)
        c = self.spec.config
        name = self.spec.name
        self.kernel_kind = name
        if name == "sgd" and c["momentum"] != 0.0:
            self.kernel_kind = "momentum"
        self.lr = float(c["lr"])
        self.decay = float(c.get("decay", 0.0))
        self.eps = float(c.get("epsilon", 1e-7))
        self.nesterov = bool(c.get("nesterov", False))
        if self.kernel_kind == "momentum":
            self.p0, self.p1 = float(c["momentum"]), 0.0
        elif name in ("rmsprop", "adadelta"):
            self.p0, self.p1 = float(c["rho"]), 0.0
        elif name in ("adam", "adamax", "nadam"):
            self.p0, self.p1 = float(c["beta_1"]), float(c["beta_2"])
        else:
            self.p0 = self.p1 = 0.0
        need0 = self.kernel_kind != "sgd"
        need1 = self.kernel_kind in ("adam", "adadelta", "adamax", "nadam")
        self.s0 = torch.zeros(numel, dtype=torch.float32, device=device) if need0 else None
        self.s1 = torch.zeros(numel, dtype=torch.float32, device=device) if need1 else None
        self.t = 0
        # elements that must not move (non-trainable BN statistics): gradient is masked to zero
        self.mask = mask.to(device) if mask is not None and not bool(mask.all()) else None

    def set_learning_rate(self, lr: float) -> None:
        self.lr = float(lr)

    def state_dict(self) -> dict:
        return {"spec": self.spec.serialize(), "t": self.t, "lr": self.lr,
                "s0": None if self.s0 is None else self.s0.detach().cpu(),
                "s1": None if self.s1 is None else self.s1.detach().cpu()}

    def load_state_dict(self, sd: dict) -> None:
        self.t = int(sd["t"])
        self.lr = float(sd.get("lr", self.lr))
        if self.s0 is not None and sd["s0"] is not None:
            self.s0.copy_(sd["s0"])
        if self.s1 is not None and sd["s1"] is not None:
            self.s1.copy_(sd["s1"])

    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, w: torch.Tensor, g: torch.Tensor, wb: Optional[torch.Tensor] = None) -> None:
        """In-place update of ``w`` from gradient ``g`` (and refresh of the bf16 shadow ``wb``)."""
        self.t += 1
        if self.mask is not None:
            g = g * self.mask
        if w.is_cuda:
            step_dev = torch.tensor([self.t], dtype=torch.int32, device=w.device)
            _native.check(_native.lib().dk_optim_step(
                _native.OPT_KINDS[self.kernel_kind], _native.ptr(w), _native.ptr(g), _native.ptr(self.s0),
                _native.ptr(self.s1), _native.ptr(wb), w.numel(), self.lr, self.p0, self.p1, self.eps,
                self.decay, int(self.nesterov), _native.ptr(step_dev), 1.0, _native.current_stream()),
                "dk_optim_step")
            return
        self.reference_step(w, g, self.t)
        if wb is not None:
            wb.copy_(w.to(torch.bfloat16))

    @torch.no_grad()
    def reference_step(self, w: torch.Tensor, g: torch.Tensor, t: int) -> None:
        """Plain PyTorch implementation of every update rule (oracle for the CUDA kernel)."""
        lr = self.lr / (1.0 + self.decay * (t - 1)) if self.decay > 0 else self.lr
        k = self.kernel_kind
        if k == "sgd":
            w.sub_(lr * g)
        elif k == "momentum":
            v = self.p0 * self.s0 - lr * g
            self.s0.copy_(v)
            w.add_(self.p0 * v - lr * g if self.nesterov else v)
        elif k == "adagrad":
            self.s0.add_(g * g)
            w.sub_(lr * g / (self.s0.sqrt() + self.eps))
        elif k == "rmsprop":
            self.s0.mul_(self.p0).add_((1 - self.p0) * g * g)
            w.sub_(lr * g / (self.s0.sqrt() + self.eps))
        elif k == "adam":
            self.s0.mul_(self.p0).add_((1 - self.p0) * g)
            self.s1.mul_(self.p1).add_((1 - self.p1) * g * g)
            corr = (1 - self.p1 ** t) ** 0.5 / (1 - self.p0 ** t)
            w.sub_(lr * corr * self.s0 / (self.s1.sqrt() + self.eps))
        elif k == "adadelta":
            self.s0.mul_(self.p0).add_((1 - self.p0) * g * g)
            upd = g * (self.s1 + self.eps).sqrt() / (self.s0 + self.eps).sqrt()
            w.sub_(lr * upd)
            self.s1.mul_(self.p0).add_((1 - self.p0) * upd * upd)
        elif k == "adamax":
            self.s0.mul_(self.p0).add_((1 - self.p0) * g)
            torch.maximum(self.p1 * self.s1, g.abs(), out=self.s1)
            w.sub_(lr / (1 - self.p0 ** t) * self.s0 / (self.s1 + self.eps))
        elif k == "nadam":
            self.s0.mul_(self.p0).add_((1 - self.p0) * g)
            self.s1.mul_(self.p1).add_((1 - self.p1) * g * g)
            mhat = self.p0 * self.s0 / (1 - self.p0 ** (t + 1)) + (1 - self.p0) * g / (1 - self.p0 ** t)
            w.sub_(lr * mhat / ((self.s1 / (1 - self.p1 ** t)).sqrt() + self.eps))
        else:  # pragma: no cover
            raise AssertionError(k)
