"""Workers: the per-partition training loops.

Capability parity with ``distkeras/workers.py`` -- ``Worker``, ``SequentialWorker``,
``NetworkWorker`` and the six algorithm workers, contract ``train(worker_id, iterator)`` -- with
the algorithm math of SURVEY 2.6 (A-F) written once on the replica's flat weight buffer.

These classes are the *portable* path: they run against any :class:`Replica` and reach the
parameter server through a client object (in-process or TCP).  They are the semantic oracle and the
CPU / multi-host backend.  On B200 the same algorithms run as fused device programs
(``parallel/engine.py``: window graph + in-kernel commit / pull over NVLink); the trainers pick
that path when CUDA is available and these classes are what it is tested against.

Differences from the reference, on purpose: the prefetch queue ends with an
explicit sentinel instead of a 10 s ``queue.get`` timeout, exceptions propagate instead of being
printed and swallowed, and ``SequentialWorker`` records history for every batch.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import networking
from .data import Partition
from .ops.flat_optim import OptimizerSpec
from .utils import deserialize_keras_model, serialize_keras_model

_SENTINEL = object()


class Worker:
    """Base worker (``workers.py:49-178``)."""

    def __init__(self, model, optimizer, loss, loss_weights=None, metrics=("accuracy",), features_col="features",
                 label_col="label", batch_size: int = 32, num_epoch: int = 1, learning_rate: float = 1.0):
        self.model = model if isinstance(model, dict) else serialize_keras_model(model)
        self.optimizer = OptimizerSpec.parse(optimizer).serialize()
        self.loss = loss
        self.loss_weights = loss_weights
        self.metrics = list(metrics)
        self.features_column = features_col if isinstance(features_col, (list, tuple)) else [features_col]
        self.label_column = label_col if isinstance(label_col, (list, tuple)) else [label_col]
        self.batch_size = int(batch_size)
        self.num_epoch = int(num_epoch)
        self.learning_rate = float(learning_rate)
        self.max_mini_batches = 100
        self.worker_id = 0
        self.device = None
        self.replica = None
        self.mini_batches: Optional[queue.Queue] = None
        self.prefetching_thread: Optional[threading.Thread] = None
        self.training_history: List[dict] = []
        self.iteration = 1
        self._error: Optional[BaseException] = None
        # a functional model with several inputs takes the feature columns one per input (the reference's
        # list-valued features_col, workers.py:65-66); a single-input model gets them concatenated
        self._multi_input = False
        try:
            import json as _json

            spec = _json.loads(self.model["model"]) if isinstance(self.model.get("model"), str) else {}
            self._multi_input = spec.get("class_name") == "Model" and len(spec.get("inputs", [])) > 1
        except (ValueError, AttributeError, TypeError):
            pass

    # -- small accessors ----------------------------------------------------------------------
    def set_max_prefetch(self, max_mini_batches: int) -> None:
        self.max_mini_batches = int(max_mini_batches)

    def set_learning_rate(self, learning_rate: float) -> None:
        self.learning_rate = float(learning_rate)

    def get_learning_rate(self) -> float:
        return self.learning_rate

    def set_worker_id(self, worker_id: int) -> None:
        self.worker_id = int(worker_id)

    def get_worker_id(self) -> int:
        return self.worker_id

    def set_device(self, device) -> None:
        self.device = device

    # -- model ----------------------------------------------------------------------------------
    def prepare_model(self) -> None:
        """Materialise the replica from the serialized model (``workers.py:103-119``)."""
        from .parallel.replica import TorchReplica

        model = deserialize_keras_model(self.model)
        device = torch.device(self.device) if self.device is not None else torch.device("cpu")
        model.to(device)
        self.replica = TorchReplica(model, self.optimizer, self.loss, device=device, loss_weights=self.loss_weights,
                                    metrics=self.metrics)

    # -- data -----------------------------------------------------------------------------------
    def get_next_minibatch(self):
        item = self.mini_batches.get()
        if item is _SENTINEL:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    def start_prefetching_thread(self, iterator) -> None:
        self.mini_batches = queue.Queue(maxsize=max(2, self.max_mini_batches))
        self.prefetching_thread = threading.Thread(target=self.prefetching, args=(iterator,), daemon=True)
        self.prefetching_thread.start()

    def _column_batches(self, iterator) -> Iterable:
        fcols, lcols = self.features_column, self.label_column
        if isinstance(iterator, Partition):
            for batch in iterator.batches(fcols + lcols, self.batch_size, drop_last=True):
                xs, ys = batch[:len(fcols)], batch[len(fcols):]
                yield (list(xs) if self._multi_input else self._merge_columns(xs)), (ys[0] if len(ys) == 1 else list(ys))
            return
        rows = []
        for row in iterator:
            rows.append(row)
            if len(rows) == self.batch_size:
                xs = [torch.as_tensor(np.stack([np.asarray(r[c]) for r in rows])) for c in fcols]
                ys = [torch.as_tensor(np.stack([np.asarray(r[c]) for r in rows])) for c in lcols]
                rows = []
                yield (xs if self._multi_input else self._merge_columns(xs)), (ys[0] if len(ys) == 1 else ys)

    @staticmethod
    def _merge_columns(xs):
        """Several feature columns (the reference's multi-input form, ``workers.py:65-66``) feed a
        ``Sequential`` as one vector: flattened and concatenated along the feature axis."""
        if len(xs) == 1:
            return xs[0]
        return torch.cat([x.reshape(x.shape[0], -1).float() for x in xs], dim=1)

    def prefetching(self, iterator) -> None:
        """Producer: ``num_epoch`` passes over the partition, exact ``batch_size`` batches
        (partial tail dropped like ``workers.py:140-151``), then a sentinel."""
        try:
            if not isinstance(iterator, Partition):
                iterator = list(iterator)  # replayable, like itertools.tee in the reference
            for _ in range(self.num_epoch):
                for item in self._column_batches(iterator):
                    self.mini_batches.put(item)
        except BaseException as exc:  # propagate to the consumer
            self.mini_batches.put(exc)
        finally:
            self.mini_batches.put(_SENTINEL)

    # -- training -------------------------------------------------------------------------------
    def optimize(self) -> None:
        raise NotImplementedError

    def add_history(self, h: Sequence[float]) -> None:
        from .utils.config import fault_injection_point

        fault_injection_point(self.worker_id, self.iteration)
        self.training_history.append({"history": [float(v) for v in h], "worker_id": self.worker_id,
                                      "iteration": self.iteration, "timestamp": time.time()})

    def train(self, worker_id: int, iterator):
        """Partition callback: returns an iterator with the trained (serialized) model."""
        self.set_worker_id(worker_id)
        self.start_prefetching_thread(iterator)
        self.prepare_model()
        try:
            self.optimize()
        except StopIteration:
            pass
        self.prefetching_thread.join(timeout=30)
        model = self.replica.model
        model.set_flat_weights(self.replica.get_flat())
        return iter([serialize_keras_model(model)])


class SequentialWorker(Worker):
    """Plain sequential training on one partition (``workers.py:181-202``)."""

    def optimize(self) -> None:
        while True:
            X, Y = self.get_next_minibatch()
            h = self.replica.train_on_batch(X, Y)
            self.add_history(h)
            self.iteration += 1

    def train(self, worker_id, iterator):
        out = super().train(worker_id, iterator)
        return out


# --------------------------------------------------------------------------------------------
# parameter-server clients
# --------------------------------------------------------------------------------------------
class InProcessClient:
    """Calls the server's handlers directly (thread workers; no socket, same mutex semantics)."""

    def __init__(self, server):
        self.server = server

    def pull(self):
        return self.server.make_pull_payload()

    def commit(self, data: dict) -> None:
        self.server.apply_commit(data)

    def close(self) -> None:
        pass


class SocketClient:
    """The reference's TCP protocol: opcode byte + framed message (``workers.py:224-240``)."""

    def __init__(self, host: str, port: int, disable_nagle: bool = True):
        self.socket = networking.connect(host, port, disable_nagle)

    def pull(self):
        self.socket.sendall(b"p")
        return networking.recv_data(self.socket)

    def commit(self, data: dict) -> None:
        self.socket.sendall(b"c")
        networking.send_data(self.socket, data)

    def close(self) -> None:
        try:
            self.socket.close()
        except OSError:
            pass


class NetworkWorker(Worker):
    """Worker with a parameter-server client (``workers.py:205-298``)."""

    def __init__(self, model, optimizer, loss, loss_weights=None, metrics=("accuracy",), features_col="features",
                 label_col="label", batch_size=32, num_epoch=1, master_host="localhost", master_port=5000,
                 learning_rate=1.0):
        super().__init__(model, optimizer, loss, loss_weights, metrics, features_col, label_col, batch_size,
                         num_epoch, learning_rate)
        self.master_host = master_host
        self.master_port = master_port
        self.disable_nagle = True
        self.client = None
        self._inproc_server = None
        self.center_variable: Optional[torch.Tensor] = None

    def attach(self, server) -> None:
        """Use the in-process transport against ``server`` instead of TCP."""
        self._inproc_server = server

    def connect(self) -> None:
        if self._inproc_server is not None:
            self.client = InProcessClient(self._inproc_server)
        else:
            self.client = SocketClient(self.master_host, self.master_port, self.disable_nagle)

    def _payload_to_flat(self, payload) -> torch.Tensor:
        W = self.replica.W
        if isinstance(payload, torch.Tensor):
            return payload.to(W.device, torch.float32).reshape(-1)
        return torch.from_numpy(np.ascontiguousarray(payload, dtype=np.float32)).reshape(-1).to(W.device)

    def pull(self) -> None:
        """Fetch the center variable (``workers.py:224-229``)."""
        self.center_variable = self._payload_to_flat(self.client.pull())

    def commit(self, residual: torch.Tensor) -> None:
        """Send a delta to the parameter server (``workers.py:231-240``)."""
        self.client.commit({"worker_id": self.worker_id, "delta": residual.detach().cpu().numpy()})

    def set_tcp_no_delay(self, flag: bool) -> None:
        self.disable_nagle = bool(flag)

    def tcp_no_delay(self) -> bool:
        return self.disable_nagle

    def get_master_host(self):
        return self.master_host

    def get_master_port(self):
        return self.master_port

    def set_weights_from_center(self) -> None:
        with torch.no_grad():
            self.replica.W.data.copy_(self.center_variable)
        self.replica.weights_changed()

    def train(self, worker_id: int, iterator):
        """``workers.py:281-298``: prefetch, build, connect, initial pull, optimize, return history."""
        self.set_worker_id(worker_id)
        self.start_prefetching_thread(iterator)
        self.prepare_model()
        self.connect()
        self.pull()
        self.set_weights_from_center()
        try:
            self.optimize()
        except StopIteration:
            pass
        finally:
            self.client.close()
        self.prefetching_thread.join(timeout=30)
        return iter(self.training_history)

    # helpers shared by the algorithm workers
    def _W(self) -> torch.Tensor:
        return self.replica.W.data

    def _train_batch(self, batch=None) -> None:
        X, Y = batch if batch is not None else self.get_next_minibatch()
        h = self.replica.train_on_batch(X, Y)
        self.add_history(h)


class ADAGWorker(NetworkWorker):
    """Algorithm A (``workers.py:301-342``): window-normalised residual."""

    def __init__(self, *args, communication_window: int = 5, **kw):
        super().__init__(*args, **kw)
        self.communication_window = int(communication_window)

    def commit(self, residual: torch.Tensor) -> None:
        self.client.commit({"worker_id": self.worker_id, "residual": residual.detach().cpu().numpy()})

    def optimize(self) -> None:
        W1 = self._W().clone()
        while True:
            self._train_batch()
            if self.iteration % self.communication_window == 0:
                residual = (self._W() - W1) / float(self.communication_window)
                self.commit(residual)
                self.pull()
                self.set_weights_from_center()
                W1 = self._W().clone()
            self.iteration += 1


class DOWNPOURWorker(NetworkWorker):
    """Algorithm B (``workers.py:345-374``): un-normalised delta, check before the batch."""

    def __init__(self, *args, communication_window: int = 3, **kw):
        super().__init__(*args, **kw)
        self.communication_window = int(communication_window)

    def optimize(self) -> None:
        W1 = self._W().clone()
        while True:
            batch = self.get_next_minibatch()  # fetched first: end of data stops before the commit
            if self.iteration % self.communication_window == 0:
                self.commit(self._W() - W1)
                self.pull()
                self.set_weights_from_center()
                W1 = self._W().clone()
            self._train_batch(batch)
            self.iteration += 1


class AEASGDWorker(NetworkWorker):
    """Algorithm C (``workers.py:377-410``): asynchronous elastic averaging."""

    def __init__(self, *args, communication_window: int = 32, rho: float = 5.0, learning_rate: float = 0.01, **kw):
        kw["learning_rate"] = learning_rate
        super().__init__(*args, **kw)
        self.communication_window = int(communication_window)
        self.rho = float(rho)
        self.alpha = self.rho * self.learning_rate

    def elastic_step(self) -> None:
        self.pull()
        W = self._W()
        E = self.alpha * (W - self.center_variable)
        W.sub_(E)
        self.replica.weights_changed()
        self.commit(E)

    def optimize(self) -> None:
        while True:
            batch = self.get_next_minibatch()
            if self.iteration % self.communication_window == 0:
                self.elastic_step()
            self._train_batch(batch)
            self.iteration += 1


class EASGDWorker(AEASGDWorker):
    """Synchronous elastic averaging (Zhang et al.; the reference only documents it,
    ``docs/optimizers.md:22-31``): every ``communication_window`` mini-batches ALL workers meet, read the
    same center, move towards it by ``alpha`` and the center moves towards their mean.  Two rendezvous
    per round: one so every worker reads the old center, one before anybody commits to it.  The
    barrier is shared by the worker copies of one trainer (thread / socket backends); a worker that runs
    out of data breaks the barrier and the others finish their shards asynchronously."""

    barrier = None  # threading.Barrier installed by the trainer
    _alone = False  # set once the rendezvous is broken

    def _meet(self) -> bool:
        import threading

        if self.barrier is None:
            return True
        try:
            self.barrier.wait(timeout=60)
            return True
        except threading.BrokenBarrierError:
            return False

    def optimize(self) -> None:
        try:
            while True:
                batch = self.get_next_minibatch()
                if self.iteration % self.communication_window == 0:
                    if self._alone or not self._meet():
                        # a peer ran out of data (broken barrier): keep exchanging with the center
                        # asynchronously, i.e. AEASGD's elastic step, for the rest of the shard
                        self._alone = True
                        self.elastic_step()
                    else:
                        self.pull()
                        W = self._W()
                        E = self.alpha * (W - self.center_variable)
                        W.sub_(E)
                        self.replica.weights_changed()
                        self._meet()  # everyone has read the old center
                        self.commit(E)
                self._train_batch(batch)
                self.iteration += 1
        finally:
            if self.barrier is not None:
                self.barrier.abort()  # out of data: release the peers


class EAMSGDWorker(AEASGDWorker):
    """Algorithm D (``workers.py:413-458``): elastic averaging + Nesterov-style momentum."""

    def __init__(self, *args, momentum: float = 0.9, **kw):
        super().__init__(*args, **kw)
        self.momentum = float(momentum)

    def optimize(self) -> None:
        r = torch.zeros_like(self._W())
        while True:
            batch = self.get_next_minibatch()
            if self.iteration % self.communication_window == 0:
                self.elastic_step()
            W = self._W()
            r_t = self.momentum * r
            W_copy = W.clone()
            W.add_(r_t)
            self.replica.weights_changed()
            before = W.clone()
            self._train_batch(batch)
            g = self._W() - before
            r = r_t - self.learning_rate * g
            self._W().copy_(W_copy - r)
            self.replica.weights_changed()
            self.iteration += 1


class DynSGDWorker(NetworkWorker):
    """Algorithm E (``workers.py:461-509``): staleness-aware commits."""

    def __init__(self, *args, communication_window: int = 5, **kw):
        super().__init__(*args, **kw)
        self.communication_window = int(communication_window)
        self.last_update = 0

    def pull(self) -> None:
        data = self.client.pull()
        self.center_variable = self._payload_to_flat(data["model"])
        self.last_update = int(data["update"])

    def commit(self, residual: torch.Tensor) -> None:
        self.client.commit({"worker_id": self.worker_id, "residual": residual.detach().cpu().numpy(),
                            "last_update": self.last_update})

    def optimize(self) -> None:
        W1 = self._W().clone()
        while True:
            self._train_batch()
            if self.iteration % self.communication_window == 0:
                self.commit(self._W() - W1)
                self.pull()
                self.set_weights_from_center()
                W1 = self._W().clone()
            self.iteration += 1


class ExperimentalWorker(NetworkWorker):
    """Algorithm F (``workers.py:512-565``): ADAG residual + stale center for per-element damping."""

    def __init__(self, *args, communication_window: int = 5, **kw):
        super().__init__(*args, **kw)
        self.communication_window = int(communication_window)

    def commit(self, residual: torch.Tensor) -> None:
        self.client.commit({"worker_id": self.worker_id, "residual": residual.detach().cpu().numpy(),
                            "stale_center_variable": self.center_variable.detach().cpu().numpy()})

    def optimize(self) -> None:
        W1 = self._W().clone()
        while True:
            self._train_batch()
            if self.iteration % self.communication_window == 0:
                residual = (self._W() - W1) / float(self.communication_window)
                self.commit(residual)
                self.pull()
                self.set_weights_from_center()
                W1 = self._W().clone()
            self.iteration += 1
