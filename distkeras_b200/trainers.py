"""Trainers: the public training API.

Capability parity with ``distkeras/trainers.py`` -- same class tree, constructor keywords and
accessors: ``Trainer``, ``SingleTrainer``, ``AveragingTrainer``, ``EnsembleTrainer``,
``DistributedTrainer``, ``AsynchronousDistributedTrainer``, ``AEASGD``, ``DOWNPOUR``, ``EAMSGD``,
``ADAG``, ``DynSGD``, ``Experimental`` -- with ``trainer.train(dataset) -> model``.

Where the reference fans work out with ``rdd.mapPartitionsWithIndex(worker.train).collect()`` on
Spark and hosts the parameter server in a driver thread behind a TCP socket
(``trainers.py:488-532``), this module has four execution backends selected by ``backend=``:

``"fabric"``  (default when CUDA is present) one process per GPU; the center variable lives in GPU
              0's HBM, workers run CUDA-graph windows of native kernels and commit / pull with
              in-kernel NVLink atomics (``parallel/runtime.py``, ``parallel/engine.py``).
``"thread"``  (default on CPU) worker threads + in-process parameter server: the semantic oracle.
``"socket"``  worker threads + the TCP parameter server: the reference's wire path, kept for
              multi-host control and parity tests.
``"nccl"``    the library-only baseline: autograd replicas (cuBLAS / cuDNN) + bulk-synchronous
              ``torch.distributed`` all-reduces every window (``parallel/nccl_baseline.py``); the
              number the fabric backend is measured against.
"""
from __future__ import annotations

import copy
import os
import queue
import threading
import time
from typing import List, Optional

import torch

from . import utils
from .data import Dataset, Partition
from .models.core import Sequential
from .networking import determine_host_address
from .ops.flat_optim import OptimizerSpec
from .parameter_servers import (ADAGParameterServer, DeltaParameterServer, DynSGDParameterServer,
                                ExperimentalParameterServer)
from .utils import deserialize_keras_model, history_executor, history_executors_average, serialize_keras_model
from .workers import (ADAGWorker, AEASGDWorker, DOWNPOURWorker, DynSGDWorker, EAMSGDWorker, EASGDWorker, ExperimentalWorker,
                      SequentialWorker)


def _default_backend() -> str:
    env = os.environ.get("DK_BACKEND")
    if env:
        return env
    return "fabric" if torch.cuda.is_available() else "thread"


def _run_tasks(worker_proto, partitions: List[Partition], num_threads: int, device_for=None,
               tolerate_failures: bool = False) -> List[list]:
    """Run ``worker.train(index, partition)`` for every partition on a pool of ``num_threads``
    threads pulling from a shared queue (dynamic shard queue = the reference's
    over-partitioning straggler mitigation, ``trainers.py:624-629``).  Each task gets its own copy
    of the worker, exactly like a pickled Spark closure."""
    tasks: "queue.Queue" = queue.Queue()
    for pos, p in enumerate(partitions):
        tasks.put((pos, p))
    results: List[Optional[list]] = [None] * len(partitions)
    workers_out: List[Optional[object]] = [None] * len(partitions)
    errors: List[BaseException] = []
    retries: dict = {}

    def loop(tid: int):
        while True:
            try:
                pos, part = tasks.get_nowait()
            except queue.Empty:
                return
            try:
                w = copy.copy(worker_proto)
                w.training_history = []
                w.iteration = 1
                if device_for is not None:
                    w.set_device(device_for(tid))
                results[pos] = list(w.train(part.index, part))
                workers_out[pos] = w
            except BaseException as exc:  # surfaced to the caller, not swallowed
                errors.append(exc)
                if tolerate_failures:
                    # the shard goes back to the queue for a surviving worker (Spark would re-run the
                    # task; the lost worker's uncommitted window is dropped)
                    retries[pos] = retries.get(pos, 0) + 1
                    if retries[pos] <= 2:
                        tasks.put((pos, part))
                    results[pos] = []
                    continue
                return

    threads = [threading.Thread(target=loop, args=(i,), daemon=True) for i in range(max(1, num_threads))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors and not tolerate_failures:
        raise errors[0]
    worker_proto.failures = errors
    return results, workers_out


class Trainer:
    """Abstract trainer (``trainers.py:39-120``)."""

    def __init__(self, keras_model: Sequential, loss, worker_optimizer, metrics=("accuracy",), loss_weights=None):
        keras_model.build()
        self.master_model = serialize_keras_model(keras_model)
        self.loss = loss
        self.loss_weights = loss_weights
        self.worker_optimizer = OptimizerSpec.parse(worker_optimizer).serialize()
        self.metrics = list(metrics)
        self.history: List[dict] = []
        self.training_time_start = 0.0
        self.training_time_end = 0.0
        self.training_time = 0.0
        self.max_mini_batches_prefetch = 100
        self.backend = _default_backend()

    def set_max_prefetch(self, max_mini_batches: int) -> None:
        self.max_mini_batches_prefetch = int(max_mini_batches)

    def set_model(self, model: Sequential) -> None:
        self.master_model = serialize_keras_model(model)

    def record_training_start(self) -> None:
        self.training_time = 0.0
        self.training_time_start = time.time()

    def record_training_end(self) -> None:
        self.training_time_end = time.time()
        self.training_time = self.training_time_end - self.training_time_start

    def get_training_time(self) -> float:
        return self.training_time

    def get_history(self) -> List[dict]:
        return self.history

    def get_averaged_history(self):
        return history_executors_average(self.history)

    def get_executor_history(self, executor_id: int):
        return history_executor(self.history, executor_id)

    def train(self, dataframe: Dataset, shuffle: bool = False) -> Sequential:
        raise NotImplementedError

    def serialize(self) -> bytes:
        return utils.pickle_object(self)

    def _maybe_shuffle(self, dataframe: Dataset, shuffle: bool) -> Dataset:
        return utils.shuffle(dataframe) if shuffle else dataframe

    def _thread_device(self, tid: int):
        if self.backend != "thread" and torch.cuda.is_available():
            return f"cuda:{tid % torch.cuda.device_count()}"
        return "cpu"


class SingleTrainer(Trainer):
    """Sequential baseline on one partition (``trainers.py:123-189``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), features_col="features",
                 label_col="label", num_epoch=1, batch_size=32, loss_weights=None):
        super().__init__(keras_model, loss, worker_optimizer, metrics, loss_weights)
        self.features_column = features_col
        self.label_column = label_col
        self.num_epoch = int(num_epoch)
        self.batch_size = int(batch_size)

    def allocate_worker(self) -> SequentialWorker:
        w = SequentialWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights, self.metrics,
                             self.features_column, self.label_column, self.batch_size, self.num_epoch)
        w.set_max_prefetch(self.max_mini_batches_prefetch)
        return w

    def train(self, dataframe: Dataset, shuffle: bool = False) -> Sequential:
        dataframe = self._maybe_shuffle(dataframe, shuffle).coalesce(1)
        worker = self.allocate_worker()
        if self.backend == "fabric" and torch.cuda.is_available():
            from .parallel.runtime import train_single_native

            self.record_training_start()
            model, self.history = train_single_native(self, dataframe)
            self.record_training_end()
            return model
        self.record_training_start()
        results, workers = _run_tasks(worker, dataframe.partitions(1), 1, self._thread_device)
        self.record_training_end()
        self.history = workers[0].training_history
        return deserialize_keras_model(results[0][0])


class AveragingTrainer(Trainer):
    """Synchronous data parallelism by per-epoch model averaging (``trainers.py:192-282``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), features_col="features",
                 label_col="label", num_epoch=1, batch_size=32, num_workers=2, loss_weights=None):
        super().__init__(keras_model, loss, worker_optimizer, metrics, loss_weights)
        self.features_column = features_col
        self.label_column = label_col
        self.num_epoch = int(num_epoch)
        self.batch_size = int(batch_size)
        self.num_workers = int(num_workers)
        self.parameter_buffer: Optional[torch.Tensor] = None

    def average_models(self, models: List[Sequential]) -> Sequential:
        """Mean of the replicas' flat buffers (buffer re-zeroed every call; the reference forgets
        to, ``trainers.py:220-236``).  On CUDA the mean is the in-kernel P2P all-reduce
        (``dk_ps_average``); here, one vectorised reduction."""
        flats = torch.stack([m.get_flat_weights().detach().float().cpu() for m in models])
        self.parameter_buffer = flats.mean(dim=0)
        out = models[0].copy()
        out.set_flat_weights(self.parameter_buffer)
        return out

    def allocate_worker(self) -> SequentialWorker:
        w = SequentialWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights, self.metrics,
                             self.features_column, self.label_column, self.batch_size, 1)
        w.set_max_prefetch(self.max_mini_batches_prefetch)
        return w

    def train(self, dataframe: Dataset, shuffle: bool = False) -> Sequential:
        dataframe = self._maybe_shuffle(dataframe, shuffle).repartition(self.num_workers)
        if self.backend == "fabric" and torch.cuda.is_available():
            from .parallel.runtime import train_averaging_native

            self.record_training_start()
            model, self.history = train_averaging_native(self, dataframe)
            self.record_training_end()
            return model
        self.record_training_start()
        self.history = []
        model = None
        for epoch in range(self.num_epoch):
            worker = self.allocate_worker()
            results, workers = _run_tasks(worker, dataframe.partitions(self.num_workers), self.num_workers,
                                          self._thread_device)
            models = [deserialize_keras_model(r[0]) for r in results]
            model = self.average_models(models)
            self.master_model = serialize_keras_model(model)
            for w in workers:
                for h in w.training_history:
                    h = dict(h)
                    h["epoch"] = epoch
                    self.history.append(h)
        self.record_training_end()
        return model


class EnsembleTrainer(Trainer):
    """``num_ensembles`` independent replicas trained in parallel; returns the list of models
    (``trainers.py:285-352``; the undefined ``num_epoch`` / ``num_workers`` of the reference are
    real attributes here)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), features_col="features",
                 label_col="label", batch_size=32, num_ensembles=2, loss_weights=None, num_epoch=1):
        super().__init__(keras_model, loss, worker_optimizer, metrics, loss_weights)
        self.features_column = features_col
        self.label_column = label_col
        self.batch_size = int(batch_size)
        self.num_ensembles = int(num_ensembles)
        self.num_workers = self.num_ensembles
        self.num_epoch = int(num_epoch)

    def allocate_worker(self) -> SequentialWorker:
        w = SequentialWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights, self.metrics,
                             self.features_column, self.label_column, self.batch_size, self.num_epoch)
        w.set_max_prefetch(self.max_mini_batches_prefetch)
        return w

    def train(self, dataframe: Dataset, shuffle: bool = False) -> List[Sequential]:
        dataframe = self._maybe_shuffle(dataframe, shuffle).repartition(self.num_ensembles)
        worker = self.allocate_worker()
        if self.backend == "fabric" and torch.cuda.is_available():
            from .parallel.runtime import train_ensemble_native

            self.record_training_start()
            models, self.history = train_ensemble_native(self, dataframe)
            self.record_training_end()
            return models
        self.record_training_start()
        results, workers = _run_tasks(worker, dataframe.partitions(self.num_ensembles), self.num_ensembles,
                                      self._thread_device)
        self.record_training_end()
        self.history = [h for w in workers for h in w.training_history]
        return [deserialize_keras_model(r[0]) for r in results]


class DistributedTrainer(Trainer):
    """Base of the parameter-server trainers (``trainers.py:355-532``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, master_port=5000, loss_weights=None):
        super().__init__(keras_model, loss, worker_optimizer, metrics, loss_weights)
        self.num_workers = int(num_workers)
        self.batch_size = int(batch_size)
        self.features_column = features_col
        self.label_column = label_col
        self.num_epoch = int(num_epoch)
        self.parameter_server = None
        self.parameter_server_thread: Optional[threading.Thread] = None
        self.master_host = determine_host_address()
        self.master_port = master_port
        self.learning_rate = 1.0
        self.parallelism_factor = 1
        self.communication_window = 1
        self.strict = bool(int(os.environ.get("DK_STRICT", "0")))
        self.checkpoint_path: Optional[str] = None       # final checkpoint (all backends)
        self.checkpoint_interval: Optional[float] = None  # seconds between mid-run snapshots (fabric backend)
        self.tolerate_worker_failures = False            # re-run a failed task instead of failing the job
        self.watchdog_timeout: Optional[float] = None     # seconds without a heartbeat before a worker is flagged
        self.worker_failures: list = []

    # -- accessors (``trainers.py:389-460``) ---------------------------------------------------
    def set_minibatch_size(self, size: int) -> None:
        self.batch_size = int(size)

    def get_minibatch_size(self) -> int:
        return self.batch_size

    def get_features_column(self):
        return self.features_column

    def get_label_column(self):
        return self.label_column

    def get_learning_rate(self) -> float:
        return self.learning_rate

    def set_learning_rate(self, learning_rate: float) -> None:
        self.learning_rate = float(learning_rate)

    def set_num_epoch(self, num_epoch: int) -> None:
        self.num_epoch = int(num_epoch)

    def get_num_epoch(self) -> int:
        return self.num_epoch

    def allocate_worker(self):
        raise NotImplementedError

    def set_master(self, master: str) -> None:
        self.master_host = master

    def determine_new_master(self) -> None:
        self.master_host = determine_host_address()

    def allocate_parameter_server(self):
        """Default: ``DeltaParameterServer`` (``trainers.py:444-452``)."""
        return DeltaParameterServer(self.master_model, self.master_port)

    def set_num_workers(self, num_workers: int) -> None:
        self.num_workers = int(num_workers)

    def get_num_workers(self) -> int:
        return self.num_workers

    def num_updates(self) -> int:
        """Commits applied by the parameter server (``trainers.py:462-464``, minus its bug)."""
        if getattr(self, "fabric_num_updates", None) is not None:
            return int(self.fabric_num_updates)
        return self.parameter_server.get_num_updates() if self.parameter_server is not None else 0

    # -- algorithm description consumed by the fabric backend ----------------------------------
    def algorithm(self) -> dict:
        """Device-program description of the algorithm (kind + hyper-parameters)."""
        raise NotImplementedError

    # -- parameter-server service (``trainers.py:466-486``) ------------------------------------
    def service(self) -> None:
        self.parameter_server.start()
        self.parameter_server.run()

    def start_service(self) -> None:
        self.parameter_server.initialize()
        self.master_port = self.parameter_server.master_port
        self.parameter_server_thread = threading.Thread(target=self.service, daemon=True)
        self.parameter_server_thread.start()

    def stop_service(self) -> None:
        self.parameter_server.stop()
        if self.parameter_server_thread is not None:
            self.parameter_server_thread.join(timeout=10)
            self.parameter_server_thread = None

    # -- training ------------------------------------------------------------------------------
    def _num_partitions(self) -> int:
        return self.num_workers * max(1, int(self.parallelism_factor))

    def train(self, dataframe: Dataset, shuffle: bool = False) -> Sequential:
        dataframe = self._maybe_shuffle(dataframe, shuffle)
        backend = self.backend
        if backend == "fabric" and type(self).algorithm is DistributedTrainer.algorithm:
            # user subclass that only overrides allocate_worker / allocate_parameter_server (the
            # reference's two-class extension story): no device program to run, so the Python worker
            # and parameter-server classes execute over the socket protocol, replicas still on the GPUs
            import warnings

            warnings.warn(f"{type(self).__name__} does not define algorithm(); falling back to backend='socket'")
            backend = "socket"
        if backend == "fabric":
            if not torch.cuda.is_available():
                raise RuntimeError("backend='fabric' needs CUDA; use backend='thread' or 'socket' on CPU")
            from .parallel.runtime import train_distributed_fabric

            self.record_training_start()
            model, self.history = train_distributed_fabric(self, dataframe)
            self.record_training_end()
            self.worker_failures = [f for st in (getattr(self, "fabric_stats", None) or [])
                                    for f in (st or {}).get("failures", [])]
            self._save_final_checkpoint(model, self.num_updates())
            return model
        if backend == "nccl":
            # library-only baseline: autograd replicas + torch.distributed collectives (parallel/nccl_baseline.py)
            from .parallel.runtime import train_distributed_nccl

            self.record_training_start()
            model, self.history = train_distributed_nccl(self, dataframe)
            self.record_training_end()
            self._save_final_checkpoint(model, self.num_updates())
            return model
        if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1 and backend in ("socket", "spmd"):
            from .parallel.runtime import train_distributed_spmd_socket

            self.record_training_start()
            model, self.history = train_distributed_spmd_socket(self, dataframe)
            self.record_training_end()
            return model
        # thread / socket backends: PS in this process, one thread per worker
        self.parameter_server = self.allocate_parameter_server()
        if backend == "socket":
            self.master_host = "127.0.0.1"
            self.start_service()
        else:
            self.parameter_server.initialize_inproc()
            self.parameter_server.start()
        worker = self.allocate_worker()
        worker.set_max_prefetch(self.max_mini_batches_prefetch)
        if backend != "socket":
            worker.attach(self.parameter_server)
        else:
            worker.master_host, worker.master_port = self.master_host, self.master_port
        n_parts = self._num_partitions()
        dataframe = dataframe.repartition(n_parts)
        self.record_training_start()
        try:
            results, _ = _run_tasks(worker, dataframe.partitions(n_parts), self.num_workers, self._thread_device,
                                    tolerate_failures=self.tolerate_worker_failures)
            self.worker_failures = list(getattr(worker, "failures", []))
        finally:
            self.record_training_end()
            if backend == "socket":
                self.stop_service()
            else:
                self.parameter_server.running = False
                self.parameter_server.finalize()
        self.history = [h for r in results for h in (r or [])]
        model = self.parameter_server.get_model()
        self._save_final_checkpoint(model, self.parameter_server.get_num_updates())
        return model

    def _save_final_checkpoint(self, model, num_updates: int) -> None:
        if not self.checkpoint_path or int(os.environ.get("RANK", "0")) != 0:
            return
        from .utils.checkpoint import save_checkpoint

        save_checkpoint(self.checkpoint_path, model, num_updates,
                        max([h["iteration"] for h in self.history], default=0), history=None)


class AsynchronousDistributedTrainer(DistributedTrainer):
    """Adds ``parallelism_factor``: partitions = factor x workers (``trainers.py:535-639``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, master_port=5000, loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.parallelism_factor = 1

    def set_parallelism_factor(self, factor: int) -> None:
        self.parallelism_factor = int(factor)

    def get_parallelism_factor(self) -> int:
        return self.parallelism_factor


class SynchronousDistributedTrainer(DistributedTrainer):
    """Base of the synchronous parameter-server trainers: one partition per worker, all workers alive at
    the same time (the reference names it in ``docs/optimizers.md:22-31`` / ``workflow.ipynb:112`` but ships
    no implementation).  Runs on the thread / socket backends; ``EASGD`` also on the fabric."""

    def _num_partitions(self) -> int:
        return self.num_workers


def _worker_kwargs(t: DistributedTrainer) -> dict:
    return dict(metrics=t.metrics, features_col=t.features_column, label_col=t.label_column,
                batch_size=t.batch_size, num_epoch=t.num_epoch, master_host=t.master_host,
                master_port=t.master_port)


class AEASGD(AsynchronousDistributedTrainer):
    """Asynchronous Elastic Averaging SGD (``trainers.py:642-689``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=32, rho=5.0,
                 learning_rate=0.1, master_port=5000, loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)
        self.rho = float(rho)
        self.learning_rate = float(learning_rate)

    def allocate_worker(self):
        return AEASGDWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                            communication_window=self.communication_window, rho=self.rho,
                            learning_rate=self.learning_rate, **_worker_kwargs(self))

    def algorithm(self) -> dict:
        return {"kind": "aeasgd", "window": self.communication_window, "alpha": self.rho * self.learning_rate}


class EASGD(SynchronousDistributedTrainer):
    """Synchronous Elastic Averaging SGD: ``C += sum_i alpha (W_i - C)``, ``W_i -= alpha (W_i - C)`` every
    ``communication_window`` mini-batches, all workers in lock step (``alpha = rho * learning_rate``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=32, rho=5.0,
                 learning_rate=0.01, master_port=5000, loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)
        self.rho = float(rho)
        self.learning_rate = float(learning_rate)

    def allocate_worker(self):
        w = EASGDWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                        communication_window=self.communication_window, rho=self.rho,
                        learning_rate=self.learning_rate, **_worker_kwargs(self))
        w.barrier = threading.Barrier(self.num_workers)
        return w

    def algorithm(self) -> dict:
        """Fabric backend: the rendezvous is a device-side barrier on the PS control block (``dk_ps_barrier``),
        the elastic read and the center update are two kernels with a barrier between them."""
        return {"kind": "easgd", "window": self.communication_window, "alpha": self.rho * self.learning_rate}


class DOWNPOUR(AsynchronousDistributedTrainer):
    """DOWNPOUR (Dean et al.; ``trainers.py:692-733``).  ``learning_rate`` is accepted for
    README compatibility (``README.md:120-121``) and only used as the worker's nominal rate."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=5, master_port=5000,
                 loss_weights=None, learning_rate=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)
        if learning_rate is not None:
            self.learning_rate = float(learning_rate)

    def allocate_worker(self):
        return DOWNPOURWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                              communication_window=self.communication_window, **_worker_kwargs(self))

    def algorithm(self) -> dict:
        return {"kind": "downpour", "window": self.communication_window}


class EAMSGD(AsynchronousDistributedTrainer):
    """Asynchronous EASGD with momentum (``trainers.py:736-787``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=32, rho=5.0,
                 learning_rate=0.1, momentum=0.9, master_port=5000, loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)
        self.rho = float(rho)
        self.learning_rate = float(learning_rate)
        self.momentum = float(momentum)

    def allocate_worker(self):
        return EAMSGDWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                            communication_window=self.communication_window, rho=self.rho,
                            learning_rate=self.learning_rate, momentum=self.momentum, **_worker_kwargs(self))

    def algorithm(self) -> dict:
        return {"kind": "eamsgd", "window": self.communication_window, "alpha": self.rho * self.learning_rate,
                "momentum": self.momentum, "eta": self.learning_rate}


class ADAG(AsynchronousDistributedTrainer):
    """Asynchronous Distributed Adaptive Gradients (``trainers.py:790-835``): the recommended
    scheme (``README.md:75-82``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=12, master_port=5000,
                 loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)

    def allocate_worker(self):
        return ADAGWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                          communication_window=self.communication_window, **_worker_kwargs(self))

    def allocate_parameter_server(self):
        return ADAGParameterServer(self.master_model, self.master_port)

    def algorithm(self) -> dict:
        return {"kind": "adag", "window": self.communication_window}


class DynSGD(AsynchronousDistributedTrainer):
    """Staleness-aware SGD (``trainers.py:838-885``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=5, master_port=5000,
                 loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)

    def allocate_worker(self):
        return DynSGDWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                            communication_window=self.communication_window, **_worker_kwargs(self))

    def allocate_parameter_server(self):
        return DynSGDParameterServer(self.master_model, self.master_port)

    def algorithm(self) -> dict:
        return {"kind": "dynsgd", "window": self.communication_window}


class Experimental(AsynchronousDistributedTrainer):
    """Development scheme with per-element staleness damping (``trainers.py:888-914``)."""

    def __init__(self, keras_model, worker_optimizer, loss, metrics=("accuracy",), num_workers=2, batch_size=32,
                 features_col="features", label_col="label", num_epoch=1, communication_window=5,
                 learning_rate=1.0, master_port=5000, loss_weights=None):
        super().__init__(keras_model, worker_optimizer, loss, metrics, num_workers, batch_size, features_col,
                         label_col, num_epoch, master_port, loss_weights)
        self.communication_window = int(communication_window)
        self.learning_rate = float(learning_rate)

    def allocate_worker(self):
        return ExperimentalWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                                  communication_window=self.communication_window, **_worker_kwargs(self))

    def allocate_parameter_server(self):
        return ExperimentalParameterServer(self.master_model, self.master_port, self.learning_rate)

    def algorithm(self) -> dict:
        return {"kind": "experimental", "window": self.communication_window,
                "inv_lr": 1.0 / self.learning_rate}
