"""Parameter servers: owners of the center variable.

Capability parity with ``distkeras/parameter_servers.py`` (class names, lifecycle
``initialize / start / run / stop / finalize``, handlers ``handle_commit / handle_pull``,
``get_model``, ``next_update / get_num_updates``), re-designed around a flat fp32 center buffer:

* ``SocketParameterServer`` and its subclasses serve the reference's commit / pull contract over
  TCP *or* in-process (same handler code, no socket) -- the CPU / multi-host control path and the
  semantic oracle.
* ``FabricParameterServer`` is the B200 product path: the center variable and a control block live
  in one GPU's HBM, are exported over CUDA IPC, and workers update / read them with system-scope
  atomics and loads issued from their own kernels (``csrc/ps_kernels.cu``).  The server is
  *passive memory*: it runs no thread, takes no lock and spends no SM time.

Update rules:
  Delta / ADAG   C += payload                                   (parameter_servers.py:232-241, 276-285)
  DynSGD         C += r / (num_updates - last_update + 1)         (parameter_servers.py:342-354)
  Experimental   C += r / (inv_lr (C - C_stale)^2 + 1)            (parameter_servers.py:372-386)
"""
from __future__ import annotations

import socket
import threading
from typing import Optional

import numpy as np
import torch

from . import networking
from .models.core import Sequential
from .utils import deserialize_keras_model


def _to_flat_tensor(x, like: torch.Tensor) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to(like.device, torch.float32).reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).reshape(-1).to(like.device)


class ParameterServer:
    """Abstract parameter server (``parameter_servers.py:26-72``)."""

    def __init__(self, model):
        if isinstance(model, dict):
            model = deserialize_keras_model(model)
        elif isinstance(model, Sequential):
            model = model.copy()
        self.model: Sequential = model
        self.num_updates = 1

    def initialize(self):
        raise NotImplementedError

    def start(self):
        raise NotImplementedError

    def run(self):
        raise NotImplementedError

    def stop(self):
        raise NotImplementedError

    def get_model(self) -> Sequential:
        return self.model

    def next_update(self) -> None:
        self.num_updates += 1

    def reset_update_counter(self) -> None:
        self.num_updates = 0

    def get_num_updates(self) -> int:
        return self.num_updates


class SocketParameterServer(ParameterServer):
    """TCP (or in-process) server; one thread per connection (``parameter_servers.py:75-216``)."""

    def __init__(self, model, port: Optional[int] = 5000):
        super().__init__(model)
        self.master_port = port
        self.socket: Optional[socket.socket] = None
        self.running = False
        self.connections = []
        self.mutex = threading.Lock()
        self.center_variable: Optional[torch.Tensor] = None
        self.commits_by_worker = {}

    # -- lifecycle ---------------------------------------------------------------------------
    def initialize(self) -> None:
        """Bind and listen.  ``master_port`` of ``None`` or 0 lets the OS pick a port."""
        self.center_variable = self.model.get_flat_weights().detach().clone()
        fd = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        fd.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        fd.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        # loopback unless the trainer asked for a routable address (multi-host control path):
        # bind_host = "0.0.0.0" is opt-in
        fd.bind((getattr(self, "bind_host", None) or "127.0.0.1", int(self.master_port or 0)))
        self.master_port = fd.getsockname()[1]
        fd.listen(64)
        self.socket = fd

    def initialize_inproc(self) -> None:
        """Same state, no socket: used by the in-process (thread worker) backend."""
        self.center_variable = self.model.get_flat_weights().detach().clone()

    def start(self) -> None:
        self.running = True

    def run(self) -> None:
        """Accept loop (``parameter_servers.py:179-192``)."""
        while self.running:
            try:
                conn, addr = self.socket.accept()
            except OSError:
                break
            if not self.running:
                conn.close()
                break
            t = threading.Thread(target=self.handle_connection, args=(conn, addr), daemon=True)
            t.start()
            self.connections.append(t)

    def cancel_accept(self) -> None:
        """Unblock ``accept`` by connecting to ourselves (``parameter_servers.py:141-151``)."""
        try:
            fd = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            fd.settimeout(1.0)
            fd.connect(("127.0.0.1", self.master_port))
            fd.close()
        except OSError:
            pass

    def cleanup_connections(self, timeout: float = 5.0) -> None:
        for t in self.connections:
            t.join(timeout)  # bounded: a hung worker must not hang shutdown
        self.connections = []

    def stop(self) -> None:
        self.running = False
        if self.socket is not None:
            self.cancel_accept()
            self.cleanup_connections()
            self.socket.close()
            self.socket = None
        self.finalize()

    def finalize(self) -> None:
        """Write the center variable back into the model (``parameter_servers.py:257-259``)."""
        if self.center_variable is not None:
            self.model.set_flat_weights(self.center_variable)

    # -- protocol ----------------------------------------------------------------------------
    def handle_connection(self, conn: socket.socket, addr) -> None:
        """Opcode loop: ``b'c'`` commit, ``b'p'`` pull, ``b's'`` stop (``parameter_servers.py:153-172``)."""
        try:
            while self.running:
                action = conn.recv(1)
                if not action:
                    break
                if action == b"c":
                    self.handle_commit(conn, addr)
                elif action == b"p":
                    self.handle_pull(conn, addr)
                else:
                    break
        except (ConnectionError, OSError):
            pass
        finally:
            conn.close()

    def handle_commit(self, conn, addr) -> None:
        self.apply_commit(networking.recv_data(conn))

    def handle_pull(self, conn, addr) -> None:
        networking.send_data(conn, self.make_pull_payload())

    # -- algebra (shared by the TCP and in-process transports) --------------------------------
    def apply_commit(self, data: dict) -> None:
        raise NotImplementedError

    def make_pull_payload(self):
        with self.mutex:
            return self.center_variable.detach().cpu().numpy().copy()

    def _count(self, data: dict) -> None:
        wid = data.get("worker_id", -1)
        self.commits_by_worker[wid] = self.commits_by_worker.get(wid, 0) + 1
        self.next_update()


class DeltaParameterServer(SocketParameterServer):
    """``C += delta`` (DOWNPOUR / AEASGD / EAMSGD; ``parameter_servers.py:219-259``)."""

    payload_key = "delta"

    def apply_commit(self, data: dict) -> None:
        delta = _to_flat_tensor(data[self.payload_key], self.center_variable)
        with self.mutex:
            self.center_variable.add_(delta)
            self._count(data)


class ADAGParameterServer(DeltaParameterServer):
    """Same arithmetic, payload key ``'residual'`` (``parameter_servers.py:262-303``)."""

    payload_key = "residual"


class DynSGDParameterServer(SocketParameterServer):
    """Staleness-aware SGD (``parameter_servers.py:306-354``)."""

    def make_pull_payload(self):
        with self.mutex:
            return {"model": self.center_variable.detach().cpu().numpy().copy(), "update": self.num_updates}

    def apply_commit(self, data: dict) -> None:
        r = _to_flat_tensor(data["residual"], self.center_variable)
        with self.mutex:
            staleness = (self.num_updates - int(data["last_update"])) + 1
            self.center_variable.add_(r / float(max(staleness, 1)))
            self._count(data)


class ExperimentalParameterServer(SocketParameterServer):
    """Per-element staleness damping (``parameter_servers.py:357-404``)."""

    def __init__(self, model, master_port: Optional[int] = 5000, learning_rate: float = 1.0):
        super().__init__(model, master_port)
        self.learning_rate = float(learning_rate)
        self.inverse_learning_rate = 1.0 / self.learning_rate

    def apply_commit(self, data: dict) -> None:
        r = _to_flat_tensor(data["residual"], self.center_variable)
        stale = _to_flat_tensor(data["stale_center_variable"], self.center_variable)
        with self.mutex:
            diff = self.center_variable - stale
            d = 1.0 / (self.inverse_learning_rate * diff * diff + 1.0)
            self.center_variable.add_(d * r)
            self._count(data)


# --------------------------------------------------------------------------------------------
# B200 path
# --------------------------------------------------------------------------------------------
class FabricParameterServer(ParameterServer):
    """Center variable + control block resident in one GPU's HBM, shared over CUDA IPC.

    ``initialize`` allocates and seeds the buffers, ``export`` returns the IPC handles a worker
    process opens (``parallel/fabric.py``), ``finalize`` copies the center back into the model.
    There is no ``run`` loop: commits / pulls are device-side atomics and loads.
    """

    def __init__(self, model, device_index: int = 0, kind: str = "delta", learning_rate: float = 1.0):
        super().__init__(model)
        self.device_index = int(device_index)
        self.kind = kind
        self.learning_rate = float(learning_rate)
        self.region = None

    def initialize(self) -> None:
        from .parallel.fabric import FabricRegion

        self.region = FabricRegion.create(self.model.get_flat_weights(), self.device_index)

    def export(self) -> dict:
        return self.region.export()

    def start(self) -> None:
        pass

    def run(self) -> None:
        pass

    def get_num_updates(self) -> int:
        if self.region is None:
            return self.num_updates
        # the control word starts at 0; the reference's counter starts at 1
        return int(self.region.read_ctrl()[0]) + 1

    def heartbeats(self, num_workers: int = 16):
        """Commits applied per worker (control-block counters bumped by the commit kernels): the
        watchdog's view of worker liveness; the PS itself never waits on a worker."""
        from . import _native

        c = self.region.read_ctrl()
        return [int(v) for v in c[_native.CTRL_HEARTBEAT:_native.CTRL_HEARTBEAT + num_workers]]

    def staleness_histogram(self):
        from . import _native

        c = self.region.read_ctrl()
        return c[_native.CTRL_STALENESS_HIST:_native.CTRL_STALENESS_HIST + 32]

    def finalize(self) -> None:
        if self.region is not None:
            self.model.set_flat_weights(self.region.read_center())
            self.num_updates = int(self.region.read_ctrl()[0]) + 1

    def stop(self) -> None:
        self.finalize()
        if self.region is not None:
            self.region.close()
            self.region = None
