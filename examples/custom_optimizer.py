"""Writing your own distributed optimizer (the reference's extension point, ``docs/optimizers.md:80-94``).

Two routes:

1. **Two classes, pure Python** -- a ``NetworkWorker`` subclass with ``optimize()`` and a trainer that
   overrides ``allocate_worker`` / ``allocate_parameter_server``.  Runs on the thread / socket backends
   (replicas on the GPUs, commits over the reference's length-prefixed TCP protocol).
2. **A device-side exchange rule** -- ``algorithm()`` returns ``{"kind": "custom", "window": tau,
   "exchange": fn}``; ``fn(ctx)`` runs on the worker's CUDA stream every ``tau`` mini-batches and talks to
   the center variable in the parameter server's HBM with one-kernel NVLink operations
   (``ctx.pull()``, ``ctx.commit_delta(scale)``, ``ctx.add_to_center(t, alpha)``, ``ctx.read_center()``).

    python examples/custom_optimizer.py [--route python|fabric]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200.data import synthetic_mnist
from distkeras_b200.models import mnist_mlp
from distkeras_b200.parameter_servers import DeltaParameterServer
from distkeras_b200.trainers import AsynchronousDistributedTrainer
from distkeras_b200.utils import deserialize_keras_model
from distkeras_b200.workers import NetworkWorker


# ---- route 1: worker + parameter server in Python -------------------------------------------------
class ClippedDeltaServer(DeltaParameterServer):
    """Clips every committed delta to a maximum L2 norm before adding it to the center."""

    max_norm = 5.0

    def apply_commit(self, data):
        delta = torch.as_tensor(data["delta"])
        norm = float(delta.norm())
        if norm > self.max_norm:
            data = dict(data, delta=(delta * (self.max_norm / norm)).numpy())
        super().apply_commit(data)


class WindowedWorker(NetworkWorker):
    def __init__(self, *a, communication_window=8, **kw):
        super().__init__(*a, **kw)
        self.communication_window = communication_window

    def optimize(self):
        w1 = self._W().clone()
        while True:
            self._train_batch()
            if self.iteration % self.communication_window == 0:
                self.commit(self._W() - w1)
                self.pull()
                self.set_weights_from_center()
                w1 = self._W().clone()
            self.iteration += 1


class ClippedDownpour(AsynchronousDistributedTrainer):
    def allocate_worker(self):
        return WindowedWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                              metrics=self.metrics, features_col=self.features_column, label_col=self.label_column,
                              batch_size=self.batch_size, num_epoch=self.num_epoch, master_host=self.master_host,
                              master_port=self.master_port)

    def allocate_parameter_server(self):
        return ClippedDeltaServer(deserialize_keras_model(self.master_model), self.master_port)


# ---- route 2: device-side exchange rule -----------------------------------------------------------
def topk_exchange(ctx, keep=0.1):
    """Push only the largest 10% of the delta's coordinates (error feedback stays in the local weights)."""
    delta = ctx.W - ctx.W1
    k = max(1, int(keep * delta.numel()))
    thresh = delta.abs().kthvalue(delta.numel() - k + 1).values
    sparse = torch.where(delta.abs() >= thresh, delta, torch.zeros_like(delta))
    residual = delta - sparse
    ctx.add_to_center(sparse, alpha=1.0 / ctx.window)
    ctx.pull()
    ctx.W.add_(residual)  # keep what was not sent


class TopKADAG(AsynchronousDistributedTrainer):
    def __init__(self, *a, communication_window=8, **kw):
        super().__init__(*a, **kw)
        self.communication_window = communication_window

    def algorithm(self):
        return {"kind": "custom", "window": self.communication_window, "exchange": topk_exchange}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--route", default="fabric" if torch.cuda.is_available() else "python")
    ap.add_argument("--rows", type=int, default=16384)
    args = ap.parse_args()
    ds = synthetic_mnist(args.rows, as_uint8=args.route == "fabric")  # uint8 is normalised on the device
    if args.route != "fabric":
        from distkeras_b200.transformers import MinMaxTransformer

        ds = MinMaxTransformer(0.0, 255.0, 0.0, 1.0, "features", "features").transform(ds)
    opt = {"class_name": "adam", "config": {"lr": 1e-3}}
    cls = TopKADAG if args.route == "fabric" else ClippedDownpour
    trainer = cls(mnist_mlp(), opt, "categorical_crossentropy", num_workers=1, batch_size=128, num_epoch=1,
                  master_port=0)
    if args.route != "fabric":
        trainer.backend = "thread" if not torch.cuda.is_available() else "socket"
    model = trainer.train(ds)
    h = trainer.get_history()
    print(f"{cls.__name__}: {len(h)} steps in {trainer.get_training_time():.2f}s, "
          f"loss {h[0]['history'][0]:.3f} -> {h[-1]['history'][0]:.3f}, updates {trainer.num_updates()}")
