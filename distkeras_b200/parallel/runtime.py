"""Multi-GPU runtime: fabric workers (CUDA-graph windows + in-kernel commit / pull) and the
process orchestration behind ``trainer.train`` on B200.

Topology: one process per GPU.  Rank 0 allocates the center variable + control block
in its HBM (:class:`~distkeras_b200.parallel.fabric.FabricRegion`) and exports it over CUDA IPC;
every rank (rank 0 included -- the parameter server is passive memory, so its SMs are free to run
a worker too) maps it and runs a :class:`FabricWorker`.

A fabric worker executes its algorithm as a device program.  For ADAG with window ``tau``:

    graph[parity] = memset(hist) ; tau x {input stage, forward, loss, backward, optimizer} ;
                    exchange kernel (commit (W - W1)/tau with atom.add.sys over NVLink, adopt the
                    returned center) ; D2H of the window's loss / accuracy records

and the host only (a) DMA-copies the next window's mini-batches from pinned memory on a copy
stream and (b) replays the graph.  The reference does the same work with ``train_on_batch`` +
numpy + pickle + TCP per window (``distkeras/workers.py:327-342``).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import tempfile
import time
from typing import List, Optional

import torch

from .. import _native as N
from ..data import Dataset, Partition
from ..utils import deserialize_keras_model, serialize_keras_model
from ..utils.config import fault_injection_point
from ..utils.timing import log_event
from .engine import NativeReplica, UnsupportedByNativeEngine
from .fabric import FabricRegion


_NVTX = os.environ.get("DK_NVTX", "0") == "1"  # NVTX ranges around windows / exchanges (Nsight timelines)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# ================================================================================================
# fabric worker
# ================================================================================================
class FabricWorker:
    """One worker replica driven by CUDA graphs of whole training *regions*.

    A graph holds ``n`` consecutive training steps (``n`` = ``windows_per_graph * tau`` for training; any
    ``n`` for benchmarking) together with every window-boundary exchange that falls inside them, so
    the host touches the device once per ``n`` steps: one DMA of ``n`` mini-batches from pinned memory,
    one graph replay, one D2H node with the region's loss / accuracy records.  Small mini-batches (the
    reference trains at 4-64 rows, ``examples/mnist_analysis.ipynb:387``) pack several communication
    windows into one graph so the per-replay host cost is amortised.  Graphs are cached by
    ``(parity, n, phase)``: ``parity`` selects the staging buffers, ``phase`` is the position of the
    region's first step inside its communication window.
    """

    def __init__(self, model, optimizer, loss: str, algorithm: dict, region: FabricRegion, worker_id: int,
                 batch_size: int, device_index: int, in_dtype: str, input_affine=(1.0, 0.0), comm: str = "exchange",
                 strict: bool = False, dense_labels: bool = False, seed: int = 0, steps_per_graph: Optional[int] = None,
                 trace: bool = False, shards=None, fuse_comm: bool = True):
        self.alg = dict(algorithm)
        self.tau = int(self.alg["window"])
        self.region = region
        # sharded parameter server: [(lo, hi, center_ptr)] slices of the flat center living on
        # different GPUs' HBM (lifts the single-GPU NVLink ingress ceiling; README TODO of the
        # reference: "multiple parameter servers").  None = the whole center is in `region`.
        self.shards = list(shards) if shards else None
        self.worker_id = int(worker_id)
        self.B = int(batch_size)
        self.comm = "commit_pull" if strict else comm
        self.strict = strict
        self.device_index = device_index
        if steps_per_graph is None:
            # ~8k rows per replay: 1 window per graph at large batches, several at the reference's
            g = max(1, min(64, -(-8192 // (self.tau * self.B))))
            steps_per_graph = g * self.tau
        self.n_max = int(steps_per_graph)
        self.fused_pull = (comm == "fused_pull" and not strict and self.alg["kind"] in ("adag", "dynsgd"))
        dev0 = torch.device("cuda", device_index)
        self.last_update = torch.zeros(1, dtype=torch.int32, device=dev0)
        self.scale_dev = torch.ones(1, dtype=torch.float32, device=dev0)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev0)
        # Exchange fused into the backward pass (compact program): the step that closes a window runs the
        # variant of the fused weight-gradient + optimizer kernel whose epilogue also commits the window's
        # displacement to the PS with system-scope atomics and adopts the returned center -- no separate
        # communication launch (the flat kernels of _comm_ops remain for every other case).
        kind = self.alg["kind"]
        comm_spec = None
        if fuse_comm and not strict and self.comm == "exchange" and kind in ("adag", "downpour", "dynsgd", "aeasgd"):
            model.build()
            comm_spec = dict(mode=N.COMM_ELASTIC if kind == "aeasgd" else N.COMM_EXCHANGE,
                             scale=1.0 / self.tau if kind == "adag" else 1.0, alpha=float(self.alg.get("alpha", 0.0)),
                             scale_dev=self.scale_dev.data_ptr() if kind == "dynsgd" else 0,
                             shards=self.shards or [(0, model.num_params, region.center_ptr)], ctrl=region.ctrl_ptr,
                             worker=self.worker_id, last_update=self.last_update.data_ptr())
        self.rep = NativeReplica(model, optimizer, loss, batch_size, device_index, in_dtype=in_dtype,
                                 input_affine=input_affine, hist_slots=2 * self.n_max, dense_labels=dense_labels,
                                 seed=seed + 7 * worker_id,
                                 pull_center_ptr=region.center_ptr if self.fused_pull else 0,
                                 region_steps=self.n_max, comm_spec=comm_spec)
        self.fused_comm = self.rep.compact and self.rep.L_bwd_comm >= 0
        if self.fused_pull and self.rep.L_step_pull < 0:
            self.fused_pull = False
        if comm == "fused_pull" and not self.fused_pull:
            self.comm = "exchange"
        rep = self.rep
        dev = rep.device
        dense_labels = rep.dense_labels  # the replica decides (mse on a linear head needs dense targets)
        self.lib = rep.lib
        self.F = rep._input_feats
        rep.ensure_snapshot()
        if self.alg["kind"] == "eamsgd":
            self.mom = torch.zeros(rep.P, dtype=torch.float32, device=dev)
            self.wcopy = torch.zeros(rep.P, dtype=torch.float32, device=dev)
        if self.alg["kind"] == "easgd":
            _easgd_state(self, rep.P, dev)
        # double-buffered region staging
        n = self.n_max
        self.x_stage = [torch.zeros(n * self.B, self.F, dtype=rep.in_torch_dtype, device=dev) for _ in (0, 1)]
        if dense_labels:
            self.y_stage = [torch.zeros(n * self.B, rep.num_classes, dtype=torch.float32, device=dev) for _ in (0, 1)]
        else:
            self.y_stage = [torch.zeros(n * self.B, dtype=torch.int32, device=dev) for _ in (0, 1)]
        self.hist_host = [torch.zeros(n, 2, dtype=torch.float32).pin_memory() for _ in (0, 1)]
        self.compute = torch.cuda.Stream(device=dev)
        self.copy = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in (0, 1)]
        self.done = [torch.cuda.Event() for _ in (0, 1)]
        self._graphs: dict = {}          # (parity, n, phase) -> CUDAGraph
        self._graph_kernels: dict = {}   # (parity, n, phase) -> kernel launches per replay
        self._graph_exchanges: dict = {}
        self.kernels_per_window = 0      # launches of one full window (tau steps + its exchange)
        self.kernels_per_step = 0
        self.launched = 0                # kernels launched through graph replays so far
        self.exchanges = 0
        self.windows_run = 0             # communication windows completed (graph replays * windows per graph)
        self.replays = 0
        self.history: List[dict] = []
        self.iteration = 0
        self._pending: List[Optional[tuple]] = [None, None]  # (first iteration, n) of the region in flight per parity
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.trace = bool(trace)
        self.trace_events: List[tuple] = []   # (event after replay, steps in it) -- device-time trace
        self._warm = False

    # -- device program pieces -------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(N.current_stream())

    def set_shards(self, shards) -> None:
        if self.fused_comm:
            raise RuntimeError("pass shards= to the constructor: the fused exchange is planned with them")
        self.shards = list(shards)

    def _comm_ops(self) -> int:
        """Enqueue the window-boundary communication of the algorithm on the current stream; returns
        the number of kernels launched."""
        reg, lib = self.region, self.lib
        ctrl = C.c_void_p(reg.ctrl_ptr)
        st = self._stream()
        k = 0
        if self.strict:
            N.check(lib.dk_ps_lock_acquire(ctrl, self.ticket.data_ptr(), st), "lock_acquire")
            k += 1
        if self.alg["kind"] == "dynsgd":
            N.check(lib.dk_ps_ticket(ctrl, self.last_update.data_ptr(), self.scale_dev.data_ptr(), st), "ticket")
            k += 1
        shards = self.shards or [(0, self.rep.P, reg.center_ptr)]
        if self.alg["kind"] == "easgd":
            k += _easgd_round(self, shards, self.rep.W.data_ptr(), self.rep.Wb.data_ptr(), st)
            if self.rep.compact:
                self.rep.refresh_pads()
            return k
        for i, (lo, hi, cptr) in enumerate(shards):
            # the control block (update counter, heartbeat) is bumped once per commit: by shard 0
            k += self._comm_range(lo, hi, cptr, ctrl if i == 0 else None, st)
        if self.strict:
            N.check(lib.dk_ps_lock_release(ctrl, self.ticket.data_ptr(), st), "lock_release")
            k += 1
        if self.rep.compact:
            self.rep.refresh_pads()   # the exchange kernels wrote the flat bf16 shadow: padded copies follow
        return k

    def _comm_range(self, lo: int, hi: int, center_ptr: int, ctrl, st) -> int:
        rep, lib, k = self.rep, self.lib, self.alg["kind"]
        n = hi - lo
        c = C.c_void_p(center_ptr)
        W, W1, Wb = rep.W.data_ptr() + 4 * lo, rep.W1.data_ptr() + 4 * lo, rep.Wb.data_ptr() + 2 * lo
        it = 0
        lu = self.last_update.data_ptr() if ctrl is not None else None
        if k in ("adag", "downpour", "dynsgd"):
            scale = 1.0 / self.tau if k == "adag" else 1.0
            sdev = self.scale_dev.data_ptr() if k == "dynsgd" else None
            if self.fused_pull:
                # commit only: the pull happens inside the next window's first forward GEMM
                N.check(lib.dk_ps_commit(c, W, W1, n, scale, sdev, ctrl, self.worker_id, it, st), "commit")
                return 1
            if self.comm == "exchange":
                N.check(lib.dk_ps_exchange(c, W, W1, Wb, n, scale, sdev, ctrl, self.worker_id, it, lu, st), "exchange")
                return 1
            N.check(lib.dk_ps_commit(c, W, W1, n, scale, sdev, ctrl, self.worker_id, it, st), "commit")
            N.check(lib.dk_ps_pull(c, W, W1, Wb, n, ctrl, lu, st), "pull")
            return 2
        if k in ("aeasgd", "eamsgd"):
            N.check(lib.dk_ps_elastic(c, W, Wb, n, float(self.alg["alpha"]), ctrl, self.worker_id, it, st), "elastic")
            return 1
        if k == "experimental":
            N.check(lib.dk_ps_damped_exchange(c, W, W1, Wb, n, 1.0 / self.tau, float(self.alg["inv_lr"]), ctrl,
                                              self.worker_id, it, st), "damped_exchange")
            return 1
        raise ValueError(f"unknown algorithm {k!r}")

    def _pull_rest(self) -> int:
        """Pull every segment the fused-pull GEMM does not cover (biases, later layers)."""
        rep, reg = self.rep, self.region
        ranges = rep.pull_rest_ranges()
        for lo, hi in ranges:
            N.check(self.lib.dk_ps_pull(C.c_void_p(reg.center_ptr + 4 * lo), rep.W.data_ptr() + 4 * lo,
                                        rep.W1.data_ptr() + 4 * lo, rep.Wb.data_ptr() + 2 * lo, hi - lo,
                                        C.c_void_p(reg.ctrl_ptr), self.last_update.data_ptr(), self._stream()),
                    "pull_rest")
        return len(ranges)

    def _step(self, parity: int, j: int, fused_pull: bool = False, comm: bool = False) -> int:
        rep = self.rep
        xs, ys = self.x_stage[parity], self.y_stage[parity]
        x_ptr = xs.data_ptr() + j * self.B * self.F * xs.element_size()
        y_ptr = ys.data_ptr() + j * self.B * (ys.shape[1] if ys.dim() == 2 else 1) * ys.element_size()
        before = rep.launches()
        extra = 0
        if self.alg["kind"] == "eamsgd":
            N.check(self.lib.dk_eamsgd_pre(rep.W.data_ptr(), self.mom.data_ptr(), self.wcopy.data_ptr(),
                                           rep.Wb.data_ptr(), rep.P, float(self.alg["momentum"]), self._stream()),
                    "eamsgd_pre")
            extra += 1
        if rep.compact:   # the region's batches were staged by one launch (enqueue_region_input)
            rep.enqueue_step(0, y_ptr, comm=comm, staged_row=j * self.B)
        else:
            rep.enqueue_step(x_ptr, y_ptr, fused_pull=fused_pull)
        if self.alg["kind"] == "eamsgd":
            N.check(self.lib.dk_eamsgd_post(rep.W.data_ptr(), self.mom.data_ptr(), self.wcopy.data_ptr(),
                                            rep.Wb.data_ptr(), rep.P, float(self.alg["eta"]), self._stream()),
                    "eamsgd_post")
            extra += 1
        return rep.launches() - before + extra

    def _region_program(self, parity: int, n: int, phase: int) -> tuple:
        """``n`` steps starting ``phase`` steps into a communication window, with the algorithm's
        exchange wherever an iteration count hits a multiple of ``tau`` -- in reference order
        (ADAG / DynSGD / Experimental check after the batch, DOWNPOUR / AEASGD / EAMSGD
        before it).  Returns (kernel launches, exchanges)."""
        rep, tau = self.rep, self.tau
        seg = rep.hist[parity * self.n_max:parity * self.n_max + n]
        seg.zero_()
        pre_batch = self.alg["kind"] in ("downpour", "aeasgd", "eamsgd", "easgd")
        kernels = exchanges = 0
        if rep.compact:
            before = rep.launches()
            rep.enqueue_region_input(self.x_stage[parity].data_ptr())
            kernels += rep.launches() - before
        fused_before = False   # the previous step's fused update already did this boundary's exchange
        for j in range(n):
            it = phase + j + 1                      # iteration number relative to the last window boundary
            boundary = it % tau == 0
            first_of_window = (phase + j) % tau == 0
            if self.fused_pull and first_of_window:
                kernels += self._pull_rest()
            if pre_batch and boundary and not fused_before:
                kernels += self._comm_ops()
                exchanges += 1
            # exchange fused into this step's backward-update kernel: the step that closes the window
            # (ADAG / DynSGD: check after the batch) or the one before the boundary batch (DOWNPOUR / AEASGD:
            # check before the batch, i.e. right after the previous step's update)
            fuse = self.fused_comm and (((it + 1) % tau == 0 and j + 1 < n) if pre_batch else boundary)
            if fuse and self.alg["kind"] == "dynsgd":
                N.check(self.lib.dk_ps_ticket(C.c_void_p(self.region.ctrl_ptr), self.last_update.data_ptr(),
                                              self.scale_dev.data_ptr(), self._stream()), "ticket")
                kernels += 1
            kernels += self._step(parity, j, fused_pull=self.fused_pull and first_of_window, comm=fuse)
            if fuse:
                exchanges += 1
            fused_before = fuse and pre_batch
            if not pre_batch and boundary and not fuse:
                kernels += self._comm_ops()
                exchanges += 1
        self.hist_host[parity][:n].copy_(seg, non_blocking=True)
        return kernels, exchanges

    def initial_pull(self) -> None:
        """``pull(); set_weights(center)`` before the first batch (``workers.py:286-288``)."""
        rep, reg = self.rep, self.region
        shards = self.shards or [(0, rep.P, reg.center_ptr)]
        with torch.cuda.stream(self.compute):
            for lo, hi, cptr in shards:
                N.check(self.lib.dk_ps_pull(C.c_void_p(cptr), rep.W.data_ptr() + 4 * lo, rep.W1.data_ptr() + 4 * lo,
                                            rep.Wb.data_ptr() + 2 * lo, hi - lo, C.c_void_p(reg.ctrl_ptr),
                                            self.last_update.data_ptr(), self._stream()), "pull")
            rep.refresh_pads()
        self.compute.synchronize()

    def _warm_up(self) -> None:
        """Run every kernel of a step once (lazy module load) on scratch state and restore it."""
        rep = self.rep
        torch.cuda.synchronize(rep.device)
        saved = (rep.W.clone(), rep.W1.clone(), rep.step_counter.clone(),
                 None if rep.opt.s0 is None else rep.opt.s0.clone(), None if rep.opt.s1 is None else rep.opt.s1.clone())
        with torch.cuda.stream(self.compute):
            if rep.compact:
                rep.enqueue_region_input(self.x_stage[0].data_ptr())
            self.kernels_per_step = self._step(0, 0)
        self.compute.synchronize()
        rep.W.copy_(saved[0]); rep.W1.copy_(saved[1]); rep.step_counter.copy_(saved[2])
        if saved[3] is not None:
            rep.opt.s0.copy_(saved[3])
        if saved[4] is not None:
            rep.opt.s1.copy_(saved[4])
        if self.alg["kind"] == "eamsgd":
            self.mom.zero_()
        rep.refresh_shadow()
        rep.hist.zero_()
        torch.cuda.synchronize(rep.device)
        self._warm = True

    def graph(self, parity: int, n: int, phase: int = 0) -> torch.cuda.CUDAGraph:
        """The (cached) CUDA graph of an ``n``-step region; captured on first use."""
        key = (parity, n, phase % self.tau)
        g = self._graphs.get(key)
        if g is None:
            if not self._warm:
                self._warm_up()
            if n > self.n_max:
                raise ValueError(f"region of {n} steps exceeds the staging capacity ({self.n_max})")
            torch.cuda.synchronize(self.rep.device)
            g = torch.cuda.CUDAGraph()
            # thread_local: the checkpointer / watchdog threads issue their own copies and synchronisations meanwhile
            with torch.cuda.graph(g, stream=self.compute, capture_error_mode="thread_local"):
                self._graph_kernels[key], self._graph_exchanges[key] = self._region_program(parity, n, key[2])
            torch.cuda.synchronize(self.rep.device)
            self._graphs[key] = g
        return g

    def replay(self, parity: int, n: int, phase: int = 0) -> None:
        """Replay a region graph on the current stream (bookkeeping included)."""
        key = (parity, n, phase % self.tau)
        self.graph(*key).replay()
        self.launched += self._graph_kernels[key]
        self.exchanges += self._graph_exchanges[key]

    def capture(self) -> None:
        """Capture both parities of the full-size training graph ahead of time."""
        for parity in (0, 1):
            self.graph(parity, self.n_max, 0)
        key = (0, self.n_max, 0)
        wins = max(1, self._graph_exchanges[key])
        self.kernels_per_window = self._graph_kernels[key] // wins if self.n_max % self.tau == 0 else 0

    def comm_kernels(self) -> int:
        """Kernels of one window-boundary exchange."""
        if not self.kernels_per_window:
            return 0
        staged = 1 if self.rep.compact else 0   # the region's input-stage launch is not a comm kernel
        return max(0, self.kernels_per_window - self.tau * self.kernels_per_step - staged)

    # -- host loop ----------------------------------------------------------------------------------
    def _collect(self, parity: int) -> None:
        pend = self._pending[parity]
        if pend is None:
            return
        first, n = pend
        self.done[parity].synchronize()
        recs = self.hist_host[parity][:n].tolist()
        now = time.time()
        wid = self.worker_id
        self.history.extend({"history": recs[j], "worker_id": wid, "iteration": first + j, "timestamp": now}
                            for j in range(n))
        self._pending[parity] = None

    def run_region(self, x_host: torch.Tensor, y_host: torch.Tensor, n: Optional[int] = None) -> None:
        """Train ``n`` steps (a multiple of ``tau``, at most ``n_max``) on ``n * B`` rows of (pinned) host
        data.  Asynchronous: returns once the H2D copies and the graph replay are enqueued; at most two
        regions are in flight."""
        n = self.n_max if n is None else int(n)
        p = self.replays & 1
        for it in range(self.iteration + 1, self.iteration + n + 1):
            fault_injection_point(self.worker_id, it)  # DK_FAULT test hook: fails BEFORE anything is enqueued
        if _NVTX:
            torch.cuda.nvtx.range_push(f"dk.region[{self.replays}] w{self.worker_id}")
        self._collect(p)  # region (r - 2) used this parity: wait for it, harvest its history
        with torch.cuda.stream(self.copy):
            self.x_stage[p][:n * self.B].copy_(x_host.reshape(n * self.B, -1), non_blocking=True)
            self.y_stage[p][:n * self.B].copy_(y_host, non_blocking=True)
            self.copied[p].record(self.copy)
        self.h2d_bytes += x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
        self.compute.wait_event(self.copied[p])
        with torch.cuda.stream(self.compute):
            if n != self.n_max or self._misaligned:
                self._realign(p)
            self.replay(p, n, 0)
            self.done[p].record(self.compute)
            if self.trace:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(self.compute)
                self.trace_events.append((ev, n))
        self._misaligned = n != self.n_max
        self.d2h_bytes += n * 8
        self._pending[p] = (self.iteration + 1, n)
        self.iteration += n
        self.windows_run += n // self.tau
        self.replays += 1
        if _NVTX:
            torch.cuda.nvtx.range_pop()

    run_window = run_region  # one-window-per-graph name kept for callers of the round-1 API

    _misaligned = False

    def _realign(self, parity: int) -> None:
        """Point the device step counter at the history slots of this parity (short regions and eager
        tail steps advance it by less than ``n_max``); it stays monotonic (Adam bias correction)."""
        self.rep.step_counter.fill_(self.replays * self.n_max + self.rep.step_base)  # replays & 1 == parity

    def trace_start(self) -> None:
        """Mark the beginning of the device-time trace on the compute stream."""
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(self.compute)
        self.trace_events.append((ev, 0))

    def trace_ms(self) -> List[tuple]:
        """[(device ms, steps)] per replay since ``trace_start`` (call after ``drain``)."""
        out = []
        for (a, _), (b, n) in zip(self.trace_events, self.trace_events[1:]):
            out.append((a.elapsed_time(b), n))
        return out

    def drain(self) -> None:
        for p in ((self.replays & 1), ((self.replays + 1) & 1)):
            self._collect(p)
        self.compute.synchronize()

    def recover(self) -> None:
        """After a failed task: finish what is in flight, drop the uncommitted window and adopt the
        CURRENT center -- what a re-run Spark task does when it reconnects and pulls
        (``workers.py:286-288``).  Optimizer state is kept."""
        try:
            self.drain()
        except Exception:
            torch.cuda.synchronize(self.rep.device)
            self._pending = [None, None]
        self.initial_pull()
        self._misaligned = True

    def train_partition(self, part: Partition, features_col: str, label_col: str, num_epoch: int = 1) -> None:
        """Consume a data partition: whole windows through the graphs (``n_max`` steps per replay, then
        one shorter replay for the remaining full windows); the uncommitted tail (fewer than ``tau``
        batches) is trained eagerly and, like the reference, never committed."""
        x_all, y_all = part.column(features_col), part.column(label_col)
        if y_all.dim() == 2 and not self.rep.dense_labels:
            y_all = y_all.argmax(dim=1)
        if not self.rep.dense_labels and y_all.dtype != torch.int32:
            y_all = y_all.to(torch.int32)
        if self.rep.dense_labels and y_all.dtype != torch.float32:
            y_all = y_all.to(torch.float32)
        if x_all.dtype != self.rep.in_torch_dtype:
            x_all = x_all.to(self.rep.in_torch_dtype)
        B, tau = self.B, self.tau
        n_rows = x_all.shape[0]
        for _ in range(num_epoch):
            lo = 0
            windows = n_rows // (tau * B)
            while windows > 0:
                n = min(self.n_max, windows * tau)
                n -= n % tau
                if n <= 0:
                    break
                self.run_region(x_all[lo:lo + n * B], y_all[lo:lo + n * B], n)
                lo += n * B
                windows -= n // tau
            tail_batches = (n_rows - lo) // B
            if tail_batches:
                self.drain()
                with torch.cuda.stream(self.compute):
                    for j in range(tail_batches):
                        a = lo + j * B
                        loss, acc = self.rep.train_on_batch(x_all[a:a + B], y_all[a:a + B])
                        self.iteration += 1
                        self.history.append({"history": [loss, acc], "worker_id": self.worker_id,
                                             "iteration": self.iteration, "timestamp": time.time()})
                self._misaligned = True
        self.drain()


class FabricEagerWorker:
    """Fabric worker for models the native planner does not lower yet (BatchNorm / residual
    blocks): the autograd executor trains the replica on the GPU, while every commit / pull still
    runs as the in-kernel NVLink program on the flat fp32 buffer (no NCCL, no host socket)."""

    def __init__(self, model, optimizer, loss: str, algorithm: dict, region: FabricRegion, worker_id: int,
                 batch_size: int, device_index: int, in_dtype: str, input_affine=(1.0, 0.0), comm: str = "exchange",
                 strict: bool = False, dense_labels: bool = False, seed: int = 0, loss_weights=None, metrics=("accuracy",)):
        from .replica import TorchReplica

        self.alg = dict(algorithm)
        self.tau = int(self.alg["window"])
        self.region, self.worker_id, self.B = region, int(worker_id), int(batch_size)
        self.comm = "commit_pull" if strict else comm
        self.strict = strict
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        model = model.copy().to(self.device)
        self.rep = TorchReplica(model, optimizer, loss, device=self.device, seed=seed + worker_id,
                                loss_weights=loss_weights, metrics=metrics)
        self.lib = N.lib()
        self.scale, self.shift = float(input_affine[0]), float(input_affine[1])
        self.W = self.rep.W.data
        self.P = self.W.numel()
        self.W1 = self.W.clone()
        self.last_update = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.scale_dev = torch.ones(1, dtype=torch.float32, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.alg["kind"] == "eamsgd":
            self.mom = torch.zeros_like(self.W)
            self.wcopy = torch.zeros_like(self.W)
        if self.alg["kind"] == "easgd":
            _easgd_state(self, self.P, self.device)
        self.compute = torch.cuda.current_stream(self.device)
        self.history: List[dict] = []
        self.iteration = 0
        self.windows_run = 0
        self.kernels_per_window = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def comm_kernels(self) -> int:
        return 1

    def _stream(self):
        return C.c_void_p(N.current_stream())

    def _comm_ops(self) -> None:
        reg, lib, k = self.region, self.lib, self.alg["kind"]
        c, ctrl = C.c_void_p(reg.center_ptr), C.c_void_p(reg.ctrl_ptr)
        W, W1, st = self.W.data_ptr(), self.W1.data_ptr(), self._stream()
        if self.strict:
            N.check(lib.dk_ps_lock_acquire(ctrl, self.ticket.data_ptr(), st), "lock_acquire")
        if k in ("adag", "downpour", "dynsgd"):
            scale = 1.0 / self.tau if k == "adag" else 1.0
            sdev = None
            if k == "dynsgd":
                N.check(lib.dk_ps_ticket(ctrl, self.last_update.data_ptr(), self.scale_dev.data_ptr(), st), "ticket")
                sdev = self.scale_dev.data_ptr()
            if self.comm == "exchange":
                N.check(lib.dk_ps_exchange(c, W, W1, None, self.P, scale, sdev, ctrl, self.worker_id, self.iteration,
                                           self.last_update.data_ptr(), st), "exchange")
            else:
                N.check(lib.dk_ps_commit(c, W, W1, self.P, scale, sdev, ctrl, self.worker_id, self.iteration, st),
                        "commit")
                N.check(lib.dk_ps_pull(c, W, W1, None, self.P, ctrl, self.last_update.data_ptr(), st), "pull")
        elif k in ("aeasgd", "eamsgd"):
            N.check(lib.dk_ps_elastic(c, W, None, self.P, float(self.alg["alpha"]), ctrl, self.worker_id,
                                      self.iteration, st), "elastic")
        elif k == "easgd":
            _easgd_round(self, [(0, self.P, reg.center_ptr)], W, 0, st)
        elif k == "experimental":
            N.check(lib.dk_ps_damped_exchange(c, W, W1, None, self.P, 1.0 / self.tau, float(self.alg["inv_lr"]), ctrl,
                                              self.worker_id, self.iteration, st), "damped_exchange")
        elif k == "custom":
            # user-defined push / pull rule (the reference's "write your own optimizer" extension
            # point, docs/optimizers.md:80-94, at the fabric level)
            if self._exchange_ctx is None:
                self._exchange_ctx = FabricExchange(self)
            self.alg["exchange"](self._exchange_ctx)
        else:
            raise ValueError(f"unknown algorithm kind {k!r}")
        if self.strict:
            N.check(lib.dk_ps_lock_release(ctrl, self.ticket.data_ptr(), st), "lock_release")
        self.windows_run += 1

    _exchange_ctx = None

    def initial_pull(self) -> None:
        reg = self.region
        N.check(self.lib.dk_ps_pull(C.c_void_p(reg.center_ptr), self.W.data_ptr(), self.W1.data_ptr(), None, self.P,
                                    C.c_void_p(reg.ctrl_ptr), self.last_update.data_ptr(), self._stream()), "pull")
        torch.cuda.synchronize(self.device)

    def capture(self) -> None:
        pass

    def drain(self) -> None:
        torch.cuda.synchronize(self.device)

    def recover(self) -> None:
        self.drain()
        self.initial_pull()

    def train_partition(self, part: Partition, features_col, label_col, num_epoch: int = 1) -> None:
        """``features_col`` / ``label_col`` may be lists: one column per model input / output (functional
        models, ``distkeras/workers.py:65-66, 140-148``); several feature columns feed a single-input model
        concatenated."""
        fcols = list(features_col) if isinstance(features_col, (list, tuple)) else [features_col]
        lcols = list(label_col) if isinstance(label_col, (list, tuple)) else [label_col]
        xs_all = [part.column(c) for c in fcols]
        ys_all = [part.column(c) for c in lcols]
        multi_in = int(getattr(self.rep.model, "num_inputs", 1)) > 1
        losses = self.rep.losses
        pre_batch = self.alg["kind"] in ("downpour", "aeasgd", "eamsgd", "easgd")
        n = xs_all[0].shape[0] // self.B
        row_bytes = sum(x[0].numel() * x.element_size() for x in xs_all)
        for _ in range(num_epoch):
            for b in range(n):
                self.iteration += 1
                fault_injection_point(self.worker_id, self.iteration)
                lo, hi = b * self.B, (b + 1) * self.B
                xs = [x[lo:hi].to(self.device, non_blocking=True).float() for x in xs_all]
                if self.scale != 1.0 or self.shift != 0.0:
                    xs = [x * self.scale + self.shift for x in xs]
                if not multi_in:
                    xs = xs[0] if len(xs) == 1 else torch.cat([x.reshape(x.shape[0], -1) for x in xs], dim=1)
                ys = []
                for y_col, ls in zip(ys_all, losses if len(losses) == len(ys_all) else [losses[0]] * len(ys_all)):
                    y = y_col[lo:hi].to(self.device, non_blocking=True)
                    ys.append(y.long() if (y.dim() == 1 and "crossentropy" in str(ls)) else y)
                self.h2d_bytes += row_bytes * self.B
                if pre_batch and self.iteration % self.tau == 0:
                    self._comm_ops()
                if self.alg["kind"] == "eamsgd":
                    N.check(self.lib.dk_eamsgd_pre(self.W.data_ptr(), self.mom.data_ptr(), self.wcopy.data_ptr(), None,
                                                   self.P, float(self.alg["momentum"]), self._stream()), "eamsgd_pre")
                h = self.rep.train_on_batch(xs, ys if len(ys) > 1 else ys[0])
                if self.alg["kind"] == "eamsgd":
                    N.check(self.lib.dk_eamsgd_post(self.W.data_ptr(), self.mom.data_ptr(), self.wcopy.data_ptr(), None,
                                                    self.P, float(self.alg["eta"]), self._stream()), "eamsgd_post")
                self.d2h_bytes += 4 * len(h)
                self.history.append({"history": [float(v) for v in h], "worker_id": self.worker_id,
                                     "iteration": self.iteration, "timestamp": time.time()})
                if not pre_batch and self.iteration % self.tau == 0:
                    self._comm_ops()
        self.drain()


def _easgd_state(worker, numel: int, device) -> None:
    """Device state of the synchronous-EASGD rendezvous: the stored elastic difference, this worker's barrier
    instance counter (advanced by the barrier kernel itself, so captured graphs replay correctly) and the
    broken-rendezvous flag (a peer did not show up within ``DK_BARRIER_TIMEOUT_MS``: stop waiting for it)."""
    worker.E = torch.zeros(numel, dtype=torch.float32, device=device)
    worker.bar_round = torch.zeros(1, dtype=torch.int32, device=device)
    worker.bar_broken = torch.zeros(1, dtype=torch.int32, device=device)
    worker.bar_timeout_ms = int(os.environ.get("DK_BARRIER_TIMEOUT_MS", "10000"))


def _easgd_round(worker, shards, w_ptr: int, wb_ptr: int, st) -> int:
    """One synchronous-EASGD round on the current stream (``docs/optimizers.md:22-31`` of the reference; the
    thread-backend oracle is ``workers.EASGDWorker``): rendezvous (every rank's previous center update has
    landed) -> ``E = alpha (W - C)``, ``W -= E`` reading the quiescent center -> rendezvous (everybody has read)
    -> ``C += E``.  Returns the number of kernels."""
    lib, reg = worker.lib, worker.region
    ctrl = C.c_void_p(reg.ctrl_ptr)
    nw = int(worker.alg.get("workers", 1))
    alpha = float(worker.alg["alpha"])

    def meet():
        N.check(lib.dk_ps_barrier(ctrl, nw, worker.bar_round.data_ptr(), worker.bar_broken.data_ptr(),
                                  worker.bar_timeout_ms, st), "barrier")

    meet()
    for lo, hi, cptr in shards:
        N.check(lib.dk_ps_easgd_read(C.c_void_p(cptr), w_ptr + 4 * lo, (wb_ptr + 2 * lo) if wb_ptr else None,
                                     worker.E.data_ptr() + 4 * lo, hi - lo, alpha, st), "easgd_read")
    meet()
    for i, (lo, hi, cptr) in enumerate(shards):
        N.check(lib.dk_ps_easgd_add(C.c_void_p(cptr), worker.E.data_ptr() + 4 * lo, hi - lo, ctrl if i == 0 else None,
                                    worker.worker_id, st), "easgd_add")
    return 2 + 2 * len(shards)


class FabricExchange:
    """What a custom exchange rule sees at every communication-window boundary.

    A trainer whose ``algorithm()`` returns ``{"kind": "custom", "window": tau, "exchange": fn}``
    (``fn`` a picklable module-level callable) gets ``fn(ctx)`` called on the worker's stream every
    ``tau`` mini-batches.  ``ctx.W`` is the worker's flat fp32 parameter vector (edit it in place),
    ``ctx.W1`` the snapshot taken at the last ``pull()``; the center variable lives in the parameter
    server's HBM and is reached through the helpers below, each ONE kernel over NVLink peer memory.
    """

    def __init__(self, worker: "FabricEagerWorker"):
        self._w = worker
        self.W, self.W1 = worker.W, worker.W1
        self.worker_id, self.window = worker.worker_id, worker.tau
        self._zeros: Optional[torch.Tensor] = None

    @property
    def iteration(self) -> int:
        return self._w.iteration

    def pull(self) -> None:
        """``W <- center`` and ``W1 <- center`` (relaxed system-scope vector loads)."""
        w = self._w
        N.check(w.lib.dk_ps_pull(C.c_void_p(w.region.center_ptr), self.W.data_ptr(), self.W1.data_ptr(), None, w.P,
                                 C.c_void_p(w.region.ctrl_ptr), w.last_update.data_ptr(), w._stream()), "pull")

    def read_center(self) -> torch.Tensor:
        """A local fp32 copy of the center variable (does not touch ``W`` / ``W1``)."""
        w = self._w
        out = torch.empty_like(self.W)
        N.check(w.lib.dk_ps_copy(out.data_ptr(), C.c_void_p(w.region.center_ptr), w.P, w._stream()), "copy")
        return out

    def add_to_center(self, t: torch.Tensor, alpha: float = 1.0) -> None:
        """``center += alpha * t`` with system-scope vector atomics; counts as one PS update."""
        w = self._w
        assert t.is_cuda and t.dtype == torch.float32 and t.numel() == w.P and t.is_contiguous()
        if self._zeros is None:
            self._zeros = torch.zeros_like(self.W)
        N.check(w.lib.dk_ps_commit(C.c_void_p(w.region.center_ptr), t.data_ptr(), self._zeros.data_ptr(), w.P,
                                   float(alpha), None, C.c_void_p(w.region.ctrl_ptr), w.worker_id, w.iteration,
                                   w._stream()), "commit")

    def commit_delta(self, scale: float = 1.0) -> None:
        """``center += scale * (W - W1)`` -- the DOWNPOUR / ADAG commit."""
        w = self._w
        N.check(w.lib.dk_ps_commit(C.c_void_p(w.region.center_ptr), self.W.data_ptr(), self.W1.data_ptr(), w.P,
                                   float(scale), None, C.c_void_p(w.region.ctrl_ptr), w.worker_id, w.iteration,
                                   w._stream()), "commit")


class FabricWatchdog:
    """Liveness monitor on the rank that owns the center.

    Every commit kernel bumps its worker's heartbeat word in the control block and a worker raises its
    done flag when it has consumed its shards, so liveness is observable without the parameter server
    ever waiting on anybody: a worker whose heartbeat has not moved for ``timeout`` seconds and whose
    done flag is still zero is reported as stalled (structured log + ``stats["watchdog"]``).
    """

    def __init__(self, region: FabricRegion, num_workers: int, device_index: int, interval: float = 0.5,
                 timeout: float = 30.0):
        import threading

        self.region, self.n = region, min(int(num_workers), N.CTRL_MAX_WORKERS)
        self.device_index, self.interval, self.timeout = int(device_index), float(interval), float(timeout)
        self.stalled: dict = {}      # worker id -> seconds without progress when it was flagged
        self.heartbeats: List[int] = [0] * self.n
        self.polls = 0
        self._halt = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True, name="dk-watchdog")

    def start(self) -> None:
        self._thread.start()

    def stop(self) -> dict:
        self._halt.set()
        self._thread.join(timeout=30)
        return {"stalled": dict(self.stalled), "heartbeats": list(self.heartbeats), "polls": self.polls}

    def poll(self, ctrl, now: float, last_change: List[float]) -> None:
        for w in range(self.n):
            hb, done = int(ctrl[N.CTRL_HEARTBEAT + w]), int(ctrl[N.CTRL_DONE_FLAGS + w])
            if hb != self.heartbeats[w] or done:
                self.heartbeats[w] = hb
                last_change[w] = now
                continue
            idle = now - last_change[w]
            if idle > self.timeout and w not in self.stalled:
                self.stalled[w] = idle
                log_event("fabric.worker_stalled", worker_id=w, idle_seconds=round(idle, 3), heartbeat=hb)
        self.polls += 1

    def _loop(self) -> None:
        torch.cuda.set_device(self.device_index)
        stream = torch.cuda.Stream(self.device_index)
        ctrl = torch.zeros(N.CTRL_WORDS, dtype=torch.int32).pin_memory()
        lib, st = N.lib(), C.c_void_p(stream.cuda_stream)
        last_change = [time.time()] * self.n
        while not self._halt.wait(self.interval):
            N.check(lib.dk_memcpy_async(C.c_void_p(ctrl.data_ptr()), C.c_void_p(self.region.ctrl_ptr), 4 * N.CTRL_WORDS, 2,
                                        st), "watchdog ctrl D2H")
            stream.synchronize()
            self.poll(ctrl, time.time(), last_change)


class CenterCheckpointer:
    """Periodic snapshots of the center variable while the workers train.

    The fabric parameter server is memory, so a checkpoint is just one more reader: a thread on the
    owning rank copies the center (or its shards) and the control block to pinned host memory on its
    own stream -- never synchronising the workers' streams -- and writes an atomic checkpoint file.
    A snapshot taken mid-run is hogwild-consistent (it may interleave with commits at 16-byte
    granularity), exactly like a worker's pull.
    """

    def __init__(self, model, region: FabricRegion, path: str, every_s: float, device_index: int, shards=None):
        import threading

        self.model, self.region, self.path = model.copy(), region, path
        self.every_s, self.device_index = float(every_s), int(device_index)
        self.shards = shards  # [(lo, hi, ptr)] or None
        self.snapshots = 0
        self._halt = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True, name="dk-checkpointer")

    def start(self) -> None:
        self._thread.start()

    def stop(self) -> None:
        self._halt.set()
        self._thread.join(timeout=60)

    def _loop(self) -> None:
        torch.cuda.set_device(self.device_index)
        stream = torch.cuda.Stream(self.device_index)
        P = self.model.num_params
        host = torch.empty(P, dtype=torch.float32).pin_memory()
        ctrl = torch.zeros(N.CTRL_WORDS, dtype=torch.int32).pin_memory()
        lib, st = N.lib(), C.c_void_p(stream.cuda_stream)
        while not self._halt.wait(self.every_s):
            pieces = self.shards or [(0, P, self.region.center_ptr)]
            for lo, hi, ptr in pieces:
                N.check(lib.dk_memcpy_async(C.c_void_p(host.data_ptr() + 4 * lo), C.c_void_p(ptr), 4 * (hi - lo), 2, st),
                        "checkpoint D2H")
            N.check(lib.dk_memcpy_async(C.c_void_p(ctrl.data_ptr()), C.c_void_p(self.region.ctrl_ptr), 4 * N.CTRL_WORDS, 2,
                                        st), "checkpoint ctrl D2H")
            stream.synchronize()
            from ..utils.checkpoint import save_checkpoint

            self.model.set_flat_weights(host.clone())
            save_checkpoint(self.path, self.model, num_updates=int(ctrl[N.CTRL_NUM_UPDATES]) + 1,
                            extra={"partial": True, "snapshot": self.snapshots})
            log_event("fabric.checkpoint", path=self.path, snapshot=self.snapshots,
                      num_updates=int(ctrl[N.CTRL_NUM_UPDATES]) + 1)
            self.snapshots += 1


# ================================================================================================
# orchestration
# ================================================================================================
def _scratch_i32(worker) -> torch.Tensor:
    if getattr(worker, "_scratch", None) is None:
        worker._scratch = torch.zeros(1, dtype=torch.int32, device=worker.rep.device)
    return worker._scratch


def _affine_for(dataset: Dataset, features_col: str):
    """uint8 features are shipped raw and normalised to [0, 1] on the device (fused MinMax)."""
    x = dataset[features_col[0] if isinstance(features_col, (list, tuple)) else features_col]
    if x.dtype == torch.uint8:
        return "u8", (1.0 / 255.0, 0.0)
    return "f32", (1.0, 0.0)


def setup_center_shards(model, rank: int, world: int, local: int, exchange_obj):
    """Sharded parameter server: slice r of the flat center variable lives in rank r's HBM (the reference's
    "multiple parameter servers" TODO, ``README.md:218``).  Every rank allocates its slice, exports it over CUDA
    IPC and maps all the others.  Returns (regions, [(lo, hi, center_ptr)], bounds); slices are equal-sized
    (a multiple of 8 elements) so a kernel finds the owner of element ``i`` as ``i // per``."""
    P_total = model.num_params
    per = ((P_total + world - 1) // world + 7) // 8 * 8
    bounds = [(min(r * per, P_total), min((r + 1) * per, P_total)) for r in range(world)]
    init_flat = model.get_flat_weights()
    lo, hi = bounds[rank]
    mine = FabricRegion.create(init_flat[lo:hi] if hi > lo else torch.zeros(8), local)
    infos = [exchange_obj(mine.export() if r == rank else None, r) for r in range(world)]
    regions = [mine if r == rank else FabricRegion.open(infos[r], local) for r in range(world)]
    shards = [(bounds[r][0], bounds[r][1], regions[r].center_ptr) for r in range(world) if bounds[r][1] > bounds[r][0]]
    return regions, shards, bounds


def _rank_train(trainer, dataset: Dataset, rank: int, world: int, exchange_obj, barrier,
                control: Optional["LocalControl"] = None) -> dict:
    """Body shared by the SPMD (torchrun) and the spawned modes.  ``exchange_obj(obj, src)``
    broadcasts a picklable object from rank ``src``; ``barrier()`` synchronises all ranks.  ``control`` (spawned
    mode) is the shard table that lets the survivors take over the partitions of a rank that died."""
    from ..parameter_servers import FabricParameterServer

    local = int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count()
    torch.cuda.set_device(local)
    if world > 1:
        from ..utils.numa import bind_to_gpu_numa_node

        bind_to_gpu_numa_node(local)  # pinned staging memory on the GPU's own socket
    alg = trainer.algorithm()
    log_event("fabric.rank_start", rank=rank, world=world, device=local, algorithm=alg.get("kind"),
              window=alg.get("window"), batch_size=trainer.batch_size, rows=len(dataset))
    model = deserialize_keras_model(trainer.master_model)
    in_dtype, affine = _affine_for(dataset, trainer.features_column)
    if getattr(trainer, "input_affine", None) is not None:
        affine = trainer.input_affine
    ps = None
    if rank == 0:
        ps = FabricParameterServer(model, device_index=local, kind=alg["kind"], learning_rate=trainer.learning_rate)
        ps.initialize()
        trainer.parameter_server = ps
        info = ps.export()
    else:
        info = None
    info = exchange_obj(info, 0)
    region = ps.region if rank == 0 else FabricRegion.open(info, local)
    # optional sharded parameter server: slice r of the flat center lives in rank r's HBM
    shard_regions, shards = [], None
    native_supported = True
    try:
        from .engine import _group_layers

        _group_layers(model)
    except UnsupportedByNativeEngine:
        native_supported = False  # same answer on every rank: the eager worker uses the unsharded center
    bounds = []
    if getattr(trainer, "sharded_ps", False) and world > 1 and native_supported:
        shard_regions, shards, bounds = setup_center_shards(model, rank, world, local, exchange_obj)
    checkpointer = None
    if rank == 0 and getattr(trainer, "checkpoint_path", None) and getattr(trainer, "checkpoint_interval", None):
        checkpointer = CenterCheckpointer(model, region, trainer.checkpoint_path, trainer.checkpoint_interval, local,
                                          shards=shards)
        checkpointer.start()
    watchdog = None
    if rank == 0 and getattr(trainer, "watchdog_timeout", None):
        watchdog = FabricWatchdog(region, trainer.num_workers, local, interval=getattr(trainer, "watchdog_interval", 0.5),
                                  timeout=trainer.watchdog_timeout)
        watchdog.start()
    num_workers = min(trainer.num_workers, world)
    dedicated = bool(getattr(trainer, "dedicated_ps", False)) and world > 1
    worker_ranks = list(range(1, world)) if dedicated else list(range(world))
    worker_ranks = worker_ranks[:num_workers]
    n_parts = max(len(worker_ranks), 1) * max(1, int(trainer.parallelism_factor))
    parts = dataset.repartition(n_parts).partitions(n_parts)
    history: List[dict] = []
    stats = {"kernels_per_window": 0, "windows": 0, "h2d_bytes": 0, "d2h_bytes": 0}
    sync_rows = None
    if alg["kind"] == "easgd":
        # lock step: every worker trains the same number of mini-batches (the shortest shard decides), so every
        # rank runs the same number of rendezvous
        alg["workers"] = len(worker_ranks)
        n_parts = len(worker_ranks)
        if getattr(trainer, "data_is_local_shard", False):
            rows = [exchange_obj(len(dataset) if rank == r else None, r) for r in range(world)]
            sync_rows = min(rows[r] for r in worker_ranks)
        else:
            sync_rows = min(len(p) for p in dataset.repartition(n_parts).partitions(n_parts))
        sync_rows = sync_rows // trainer.batch_size * trainer.batch_size
    if rank in worker_ranks:
        dataset.pin_memory()
        parts = dataset.repartition(n_parts).partitions(n_parts)
        wid = worker_ranks.index(rank)
        wkw = dict(comm=getattr(trainer, "comm", "exchange"), strict=trainer.strict, seed=getattr(trainer, "seed", 0))
        trace = bool(getattr(trainer, "trace_windows", False))
        try:
            if alg["kind"] == "custom":  # python exchange rule: cannot live inside a captured graph
                raise UnsupportedByNativeEngine("custom exchange rule")
            worker = FabricWorker(model, trainer.worker_optimizer, trainer.loss, alg, region, wid, trainer.batch_size,
                                  local, in_dtype, affine, steps_per_graph=getattr(trainer, "steps_per_graph", None),
                                  trace=trace, shards=shards, fuse_comm=getattr(trainer, "fuse_comm", True), **wkw)
        except UnsupportedByNativeEngine as exc:
            import warnings

            warnings.warn(f"the native sm_100a engine does not lower this model ({exc}); the replica runs on the "
                          "autograd executor (cuBLAS / cuDNN), the parameter-server program stays in-kernel")
            worker = FabricEagerWorker(model, trainer.worker_optimizer, trainer.loss, alg, region, wid,
                                       trainer.batch_size, local, in_dtype, affine, loss_weights=trainer.loss_weights,
                                       metrics=trainer.metrics, **wkw)
        worker.initial_pull()
        worker.capture()
        static = getattr(trainer, "shard_mode", "dynamic" if trainer.parallelism_factor > 1 else "static") == "static"
        my_parts = [parts[i] for i in range(wid, n_parts, len(worker_ranks))] if static else None
        if getattr(trainer, "data_is_local_shard", False):
            # SPMD data loading (torchrun): the dataset this rank was given IS its shard
            static = True
            f = max(1, int(trainer.parallelism_factor))
            my_parts = dataset.repartition(f).partitions(f)
        if sync_rows is not None:
            static = True
            src = dataset.partitions(1)[0] if getattr(trainer, "data_is_local_shard", False) else parts[wid]
            my_parts = [Partition(src.dataset, src.index, src.start, src.start + sync_rows)]
        use_table = control is not None and sync_rows is None and not getattr(trainer, "data_is_local_shard", False)
        if use_table and static:
            # statically assigned partitions are claimed BEFORE the start barrier: a fast rank that runs out of work
            # must never find a slower peer's partition unclaimed (it would be trained twice)
            for part in my_parts:
                control.claim(part.index, rank)
        torch.cuda.synchronize()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = worker.windows_run
        it0 = worker.iteration
        t0 = time.time()
        with torch.cuda.stream(worker.compute):
            ev0.record(worker.compute)
            if trace and isinstance(worker, FabricWorker):
                worker.trace_start()
        failures: List[dict] = []
        tolerate = bool(getattr(trainer, "tolerate_worker_failures", False))

        def run_task(part) -> None:
            """One data partition = one task.  With ``tolerate_worker_failures`` a failed task is re-run
            (at most twice) after the worker re-pulled the current center; the failed attempt's history
            is discarded, like the result of a failed Spark task."""
            for attempt in range(3):
                mark = len(worker.history)
                try:
                    worker.train_partition(part, trainer.features_column, trainer.label_column, trainer.num_epoch)
                    return
                except Exception as exc:
                    if not tolerate or attempt == 2:
                        raise
                    failures.append({"worker_id": wid, "partition": part.index, "attempt": attempt,
                                     "iteration": worker.iteration, "error": repr(exc)})
                    log_event("fabric.task_failed", **failures[-1])
                    worker.recover()
                    del worker.history[mark:]

        if use_table:
            # spawned ranks: claims go through the shard table so the launcher can re-queue a dead rank's work
            def on_requeue(idx: int) -> None:   # only reachable in static mode after a peer died
                log_event("fabric.requeued_partition", rank=rank, partition=idx)
                worker.recover()

            drain_shard_table(control, rank, parts, my_parts if static else None, run_task, on_requeue)
        elif static:
            for part in my_parts:
                run_task(part)
        else:
            # dynamic shard queue: claim partitions with a fetch-add on the PS control block
            claim = torch.zeros(1, dtype=torch.int32, device=worker.rep.device)
            word = C.c_void_p(region.ctrl_ptr + 4 * N.CTRL_SHARD_NEXT)
            while True:
                N.check(worker.lib.dk_ps_fetch_add(word, 1, claim.data_ptr(), C.c_void_p(N.current_stream())),
                        "fetch_add")
                idx = int(claim.item())
                if idx >= n_parts:
                    break
                run_task(parts[idx])
        with torch.cuda.stream(worker.compute):
            ev1.record(worker.compute)
            if wid < N.CTRL_MAX_WORKERS:  # raise this worker's done flag for the watchdog
                N.check(worker.lib.dk_ps_fetch_add(C.c_void_p(region.ctrl_ptr + 4 * (N.CTRL_DONE_FLAGS + wid)), 1,
                                                   _scratch_i32(worker).data_ptr(),
                                                   C.c_void_p(N.current_stream())), "done flag")
        torch.cuda.synchronize()
        steps_done = worker.iteration - it0
        windows_done = worker.windows_run - w0
        if isinstance(worker, FabricWorker):
            graph_steps = sum(n for _, n in worker.trace_events) if trace else None
            launches = worker.launched + max(0, steps_done - windows_done * worker.tau) * worker.kernels_per_step
        else:
            graph_steps, launches = None, 0
        stats = {"kernels_per_window": worker.kernels_per_window, "windows": windows_done, "steps": steps_done,
                 "gpu_launches": launches, "exchanges": getattr(worker, "exchanges", windows_done),
                 "steps_per_graph": getattr(worker, "n_max", None), "graph_steps": graph_steps,
                 "h2d_bytes": worker.h2d_bytes, "d2h_bytes": worker.d2h_bytes, "seconds": time.time() - t0,
                 "device_ms": ev0.elapsed_time(ev1), "executor": type(worker).__name__, "failures": failures}
        if trace and isinstance(worker, FabricWorker):
            stats["trace_ms"] = worker.trace_ms()
        history = worker.history
        log_event("fabric.worker_done", rank=rank, worker_id=wid, **stats)
        # release the replica's device buffers / graphs before the next job in this process
        try:
            worker.rep.close()
        except Exception as exc:  # teardown only; the result is already collected
            log_event("fabric.replica_close_failed", rank=rank, error=repr(exc))
        del worker
        import gc

        gc.collect()
        torch.cuda.empty_cache()
    else:
        barrier()
    barrier()
    result = {"history": history, "stats": stats}
    if checkpointer is not None:
        checkpointer.stop()
        stats["checkpoint_snapshots"] = checkpointer.snapshots
    if watchdog is not None:
        stats["watchdog"] = watchdog.stop()
    if rank == 0:
        result["num_updates"] = ps.get_num_updates()
        result["staleness_hist"] = ps.staleness_histogram().tolist()
        ps.finalize()
        if shards:  # assemble the center from the shards (rank 0 maps every shard)
            flat = ps.get_model().get_flat_weights().clone()
            for (lo, hi, _), reg_r in zip([(b[0], b[1], 0) for b in bounds if b[1] > b[0]],
                                          [shard_regions[r] for r in range(world) if bounds[r][1] > bounds[r][0]]):
                flat[lo:hi] = reg_r.read_center()[:hi - lo]
            ps.get_model().set_flat_weights(flat)
        result["model"] = serialize_keras_model(ps.get_model())
    barrier()
    for r, reg_r in enumerate(shard_regions):
        if r != rank:
            reg_r.close()
    barrier()
    if shard_regions:
        shard_regions[rank].close()
    if rank != 0:
        region.close()
    else:
        ps.stop()
    return result


def _spmd_env() -> Optional[tuple]:
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        return int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    return None


def _init_pg(backend: str = "cpu:gloo,cuda:nccl"):
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if torch.cuda.is_available() and "nccl" in backend:
            kwargs["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        dist.init_process_group(backend=backend, **kwargs)
    return dist


def _dist_helpers(dist):
    def exchange_obj(obj, src):
        box = [obj]
        dist.broadcast_object_list(box, src=src, device=torch.device("cpu"))
        return box[0]

    def barrier():
        dist.barrier()

    return exchange_obj, barrier


class LocalControl:
    """Control plane of the ranks a driver process spawns (one per GPU): shared-memory flags + files in the job's
    scratch directory instead of a gloo process group, because it has to keep working when a rank disappears.

    * ``alive[r]``: cleared by the launcher when process ``r`` exits abnormally; ``barrier`` skips dead ranks.
    * ``claimed[i]`` (rank + 1, 0 = free) / ``done[i]``: the shard table.  A worker claims a data partition under
      ``lock``; the launcher frees the unfinished claims of a dead rank, and the survivors -- which keep polling the
      table until every partition is done -- pick them up (the Spark scheduler's task re-submission,
      ``/root/reference/distkeras/workers.py:286-288`` runs again on the new executor: pull, then train).
    """

    MAX_PARTS = 8192

    def __init__(self, world: int, scratch: str):
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        self.world, self.scratch = world, scratch
        self.alive = ctx.Array("b", [1] * world, lock=False)
        self.gen = ctx.Array("l", [0] * world, lock=False)
        self.claimed = ctx.Array("i", [0] * self.MAX_PARTS, lock=False)
        self.done = ctx.Array("b", [0] * self.MAX_PARTS, lock=False)
        self.lock = ctx.Lock()
        self._seq = 0

    # -- collectives --------------------------------------------------------------------------------
    def barrier(self, rank: int, timeout: float = 1800.0) -> None:
        self.gen[rank] += 1
        mine = self.gen[rank]
        t0 = time.time()
        while any(self.alive[r] and self.gen[r] < mine for r in range(self.world)):
            if time.time() - t0 > timeout:
                raise RuntimeError("fabric control barrier timed out")
            time.sleep(0.002)

    def exchange_obj(self, rank: int, obj, src: int, timeout: float = 600.0):
        import pickle

        self._seq += 1
        path = os.path.join(self.scratch, f"xchg_{self._seq}_{src}.pkl")
        if rank == src:
            with open(path + ".tmp", "wb") as f:
                pickle.dump(obj, f)
            os.replace(path + ".tmp", path)
            return obj
        t0 = time.time()
        while not os.path.exists(path):
            if not self.alive[src]:
                raise RuntimeError(f"rank {src} died before publishing object {self._seq}")
            if time.time() - t0 > timeout:
                raise RuntimeError("fabric control exchange timed out")
            time.sleep(0.002)
        with open(path, "rb") as f:
            return pickle.load(f)

    # -- shard table --------------------------------------------------------------------------------
    def claim(self, index: int, rank: int) -> None:
        self.claimed[index] = rank + 1

    def try_claim(self, rank: int, n_parts: int) -> Optional[int]:
        with self.lock:
            for i in range(n_parts):
                if not self.done[i] and self.claimed[i] == 0:
                    self.claimed[i] = rank + 1
                    return i
        return None

    def finish(self, index: int) -> None:
        self.done[index] = 1

    def all_done(self, n_parts: int) -> bool:
        return all(self.done[i] for i in range(n_parts))

    def release_claims_of(self, rank: int) -> List[int]:
        """Launcher side: free the unfinished partitions of a dead rank; returns their indices."""
        freed = []
        with self.lock:
            for i in range(self.MAX_PARTS):
                if self.claimed[i] == rank + 1 and not self.done[i]:
                    self.claimed[i] = 0
                    freed.append(i)
        return freed


def drain_shard_table(control: "LocalControl", rank: int, parts: list, my_parts: Optional[list], run_task, on_requeue=None) -> int:
    """A rank's task loop over the shard table.  ``my_parts`` (static assignment; ALREADY claimed by this rank before the
    start barrier) are run first; then the rank keeps polling: any partition that is neither done nor claimed -- the
    dynamic queue, or what the launcher freed after a peer died -- is claimed under the table's lock and run, until
    every partition is done.  Returns the number of partitions this rank ran."""
    n_parts = len(parts)
    ran = 0
    for part in my_parts or ():
        run_task(part)
        control.finish(part.index)
        ran += 1
    while True:
        idx = control.try_claim(rank, n_parts)
        if idx is not None:
            if my_parts is not None and on_requeue is not None:
                on_requeue(idx)
            run_task(parts[idx])
            control.finish(idx)
            ran += 1
            continue
        if control.all_done(n_parts):
            return ran
        time.sleep(0.01)


def _spawn_entry_local(rank: int, world: int, control: LocalControl, payload_path: str, scratch: str) -> None:
    """Body of a rank spawned by the driver-style launcher (fabric entry): no process group at all."""
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world)})
    payload = torch.load(payload_path, weights_only=False)
    trainer, dataset = payload["trainer"], payload["dataset"]
    res = _rank_train(trainer, dataset, rank, world, lambda obj, src: control.exchange_obj(rank, obj, src),
                      lambda: control.barrier(rank), control=control)
    torch.save(res, os.path.join(scratch, f"result_{rank}.pt.tmp"))
    os.replace(os.path.join(scratch, f"result_{rank}.pt.tmp"), os.path.join(scratch, f"result_{rank}.pt"))
    control.barrier(rank)


def _launch_local(trainer, dataset: Dataset, world: int):
    """Spawn ``world`` ranks, watch them, survive the loss of worker ranks when the trainer asks for it
    (``tolerate_worker_failures``): the dead rank's unfinished partitions go back to the shard table, the
    survivors retrain them from the current center.  Losing rank 0 (it owns the center) ends the job, as losing
    the Spark driver does in the reference."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    tmp = tempfile.mkdtemp(prefix="dk_fabric_")
    payload_path = os.path.join(tmp, "payload.pt")
    ps_obj, trainer.parameter_server = trainer.parameter_server, None
    torch.save({"trainer": trainer, "dataset": dataset, "entry": "fabric"}, payload_path)
    trainer.parameter_server = ps_obj
    control = LocalControl(world, tmp)
    procs = [ctx.Process(target=_spawn_entry_local, args=(r, world, control, payload_path, tmp), daemon=False)
             for r in range(world)]
    for p in procs:
        p.start()
    tolerate = bool(getattr(trainer, "tolerate_worker_failures", False))
    lost: List[dict] = []
    fatal = None
    try:
        while any(p.exitcode is None for p in procs):
            for r, p in enumerate(procs):
                if p.exitcode is not None and p.exitcode != 0 and control.alive[r]:
                    control.alive[r] = 0
                    freed = control.release_claims_of(r)
                    lost.append({"rank": r, "exitcode": p.exitcode, "requeued_partitions": freed})
                    log_event("fabric.rank_lost", rank=r, exitcode=p.exitcode, requeued=freed)
                    if r == 0 or not tolerate:
                        fatal = RuntimeError(f"fabric rank {r} exited with code {p.exitcode}"
                                             + ("" if r == 0 else " (set tolerate_worker_failures to survive "
                                                                  "the loss of worker ranks)"))
                        break
            if fatal is not None:
                break
            time.sleep(0.02)
    finally:
        if fatal is not None:
            for p in procs:
                if p.exitcode is None:
                    p.terminate()
        for p in procs:
            p.join(timeout=30)
    if fatal is not None:
        raise fatal
    res = torch.load(os.path.join(tmp, "result_0.pt"), weights_only=False)
    history, all_stats = [], []
    for r in range(world):
        path = os.path.join(tmp, f"result_{r}.pt")
        if os.path.exists(path):
            rr = res if r == 0 else torch.load(path, weights_only=False)
            history += rr["history"]
            all_stats.append(rr["stats"])
        else:
            all_stats.append({"lost": True})
    res["history"], res["all_stats"], res["lost_ranks"] = history, all_stats, lost
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    return res


def _spawn_entry(rank: int, world: int, port: int, payload_path: str, result_path: str) -> None:
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    payload = torch.load(payload_path, weights_only=False)
    trainer, dataset = payload["trainer"], payload["dataset"]
    entry = payload.get("entry", "fabric")
    dist = _init_pg("cpu:gloo,cuda:nccl" if (entry != "fabric" and torch.cuda.is_available()) else "gloo")
    exchange_obj, barrier = _dist_helpers(dist)
    res = _rank_entry(entry)(trainer, dataset, rank, world, exchange_obj, barrier)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"history": res["history"], "stats": res["stats"]})
    if rank == 0:
        res["history"] = [h for g in gathered for h in g["history"]]
        res["all_stats"] = [g["stats"] for g in gathered]
        torch.save(res, result_path)
    dist.barrier()
    dist.destroy_process_group()


def _rank_entry(name: str):
    if name == "nccl":
        from .nccl_baseline import rank_train_nccl

        return rank_train_nccl
    return _rank_train


def train_distributed_nccl(trainer, dataset: Dataset):
    """The library-collectives baseline (``backend="nccl"``): same launcher, ``nccl_baseline.rank_train_nccl``
    as the per-rank body."""
    return train_distributed_fabric(trainer, dataset, entry="nccl")


def train_distributed_fabric(trainer, dataset: Dataset, entry: str = "fabric"):
    """Run a PS trainer on the NVLink fabric.  Under ``torchrun`` every rank calls this (SPMD);
    from a plain driver process the ranks are spawned here, one per GPU."""
    env = _spmd_env()
    rank_fn = _rank_entry(entry)
    if env is not None:
        rank, world = env
        dist = _init_pg()
        exchange_obj, barrier = _dist_helpers(dist)
        res = rank_fn(trainer, dataset, rank, world, exchange_obj, barrier)
        gathered = [None] * world
        dist.all_gather_object(gathered, {"history": res["history"], "stats": res["stats"]})
        history = [h for g in gathered for h in g["history"]]
        box = [res.get("model"), res.get("num_updates"), res.get("staleness_hist")]
        dist.broadcast_object_list(box, src=0, device=torch.device("cpu"))
        trainer.fabric_stats = [g["stats"] for g in gathered]
        trainer.fabric_num_updates, trainer.staleness_histogram = box[1], box[2]
        return deserialize_keras_model(box[0]), history
    if entry == "nccl":   # one rank per worker; on a CPU-only host the collectives run over gloo
        world = min(max(1, trainer.num_workers), torch.cuda.device_count() or trainer.num_workers)
    else:
        # one rank per GPU; ranks_per_gpu > 1 packs several ranks on a device (CUDA IPC works within a device:
        # how the multi-rank logic is exercised on a one-GPU box)
        world = min(max(1, trainer.num_workers + (1 if getattr(trainer, "dedicated_ps", False) else 0)),
                    torch.cuda.device_count() * max(1, int(getattr(trainer, "ranks_per_gpu", 1))))
    if world == 1:
        # single GPU: PS and worker share the device, no extra process needed
        res = rank_fn(trainer, dataset, 0, 1, lambda obj, src: obj, lambda: None)
        trainer.fabric_stats = [res["stats"]]
        trainer.fabric_num_updates, trainer.staleness_histogram = res["num_updates"], res["staleness_hist"]
        return deserialize_keras_model(res["model"]), res["history"]
    if entry == "fabric":
        res = _launch_local(trainer, dataset, world)
        trainer.fabric_stats = res.get("all_stats")
        trainer.fabric_num_updates, trainer.staleness_histogram = res["num_updates"], res["staleness_hist"]
        trainer.lost_ranks = res.get("lost_ranks", [])
        return deserialize_keras_model(res["model"]), res["history"]
    import torch.multiprocessing as mp

    tmp = tempfile.mkdtemp(prefix="dk_fabric_")
    payload_path, result_path = os.path.join(tmp, "payload.pt"), os.path.join(tmp, "result.pt")
    ps_obj, trainer.parameter_server = trainer.parameter_server, None
    torch.save({"trainer": trainer, "dataset": dataset, "entry": entry}, payload_path)
    trainer.parameter_server = ps_obj
    mp.spawn(_spawn_entry, args=(world, _free_port(), payload_path, result_path), nprocs=world, join=True)
    res = torch.load(result_path, weights_only=False)
    trainer.fabric_stats = res.get("all_stats")
    trainer.fabric_num_updates, trainer.staleness_histogram = res["num_updates"], res["staleness_hist"]
    for f in (payload_path, result_path):
        try:
            os.remove(f)
        except OSError:
            pass
    return deserialize_keras_model(res["model"]), res["history"]


# ------------------------------------------------------------------------------------------------
# single-GPU / averaging trainers on the native engine
# ------------------------------------------------------------------------------------------------
def _sequential_native(trainer, part: Partition, model, device_index: int, worker_id: int, num_epoch: int,
                       init_flat: Optional[torch.Tensor] = None, device_result: bool = False):
    """One replica, one partition, ``num_epoch`` passes through the native engine.  ``init_flat`` (a device
    tensor) overrides the model's weights without a host round trip; ``device_result`` returns the trained flat
    buffer as a device tensor instead of writing it back into ``model`` (both: ``train_averaging_native``)."""
    dataset = part.dataset
    in_dtype, affine = _affine_for(dataset, trainer.features_column)
    rep = NativeReplica(model, trainer.worker_optimizer, trainer.loss, trainer.batch_size, device_index,
                        in_dtype=in_dtype, input_affine=affine, hist_slots=4096)
    if init_flat is not None:
        rep.set_flat(init_flat)
    x_all, y_all = part.column(trainer.features_column), part.column(trainer.label_column)
    if rep.dense_labels:   # regression targets (mse on a linear head): fp32 [n, n_out], never argmax'd
        y_all = y_all.to(torch.float32).reshape(y_all.shape[0], -1)
        if y_all.shape[1] != rep.num_classes:
            raise ValueError(f"label column has {y_all.shape[1]} values per row, the model outputs {rep.num_classes}")
        y_stride = trainer.batch_size * rep.num_classes * 4
    else:
        if y_all.dim() == 2:
            y_all = y_all.argmax(dim=1)
        y_all = y_all.to(torch.int32)
        y_stride = trainer.batch_size * 4
    dev = rep.device
    B = trainer.batch_size
    history, it = [], 0
    n_batches = x_all.shape[0] // B
    chunk = 1024  # batches per device-resident chunk / history flush
    for _ in range(num_epoch):
        for c0 in range(0, n_batches, chunk):
            c1 = min(n_batches, c0 + chunk)
            # the replica was planned for rep.in_torch_dtype: cast here (the kernels read raw pointers)
            xd = x_all[c0 * B:c1 * B].to(dev, non_blocking=True).reshape((c1 - c0) * B, -1).to(rep.in_torch_dtype).contiguous()
            yd = y_all[c0 * B:c1 * B].to(dev, non_blocking=True).contiguous()
            rep.hist.zero_()
            base = int(rep.step_counter.item())
            for j in range(c1 - c0):
                rep.enqueue_step(xd.data_ptr() + j * B * xd.shape[1] * xd.element_size(), yd.data_ptr() + j * y_stride)
            steps = (torch.arange(c1 - c0, device=dev) + base) % rep.hist_slots
            recs = rep.hist[steps].cpu().numpy()
            now = time.time()
            for j in range(c1 - c0):
                it += 1
                history.append({"history": [float(recs[j, 0]), float(recs[j, 1])], "worker_id": worker_id,
                                "iteration": it, "timestamp": now})
    if device_result:
        flat = rep.W.detach().clone()
        rep.close()
        return flat, history
    model.set_flat_weights(rep.W.detach().cpu())
    rep.close()
    return model, history


def train_single_native(trainer, dataset: Dataset):
    """``SingleTrainer`` on one GPU through the native engine."""
    model = deserialize_keras_model(trainer.master_model)
    try:
        return _sequential_native(trainer, dataset.partitions(1)[0], model, torch.cuda.current_device(), 0,
                                  trainer.num_epoch)
    except UnsupportedByNativeEngine:
        from ..trainers import _run_tasks

        worker = trainer.allocate_worker()
        results, workers = _run_tasks(worker, dataset.partitions(1), 1, lambda tid: "cuda:0")
        return deserialize_keras_model(results[0][0]), workers[0].training_history


def train_averaging_native(trainer, dataset: Dataset):
    """``AveragingTrainer`` on the GPUs of this process (reference op K13, ``trainers.py:190-247``): one replica per
    device trains its partition (one host thread each, launches are asynchronous), then the epoch's mean is a
    reduce-scatter + all-gather over peer memory -- device ``r`` launches ``dk_ps_average`` on slice ``r`` only:
    it reads that slice from every replica over NVLink, averages, and writes the result into every replica.  The
    replicas of the next epoch start from those device buffers; the host sees the weights once, at the end."""
    ndev = torch.cuda.device_count()
    W = trainer.num_workers
    parts = dataset.repartition(W).partitions(W)
    master = deserialize_keras_model(trainer.master_model)
    history: List[dict] = []
    lib = N.lib()
    import threading

    flats: List[Optional[torch.Tensor]] = [None] * W
    host_synced = True            # does `master` (host) hold the current mean?
    eager = [False]               # a replica fell back to the autograd worker: it restarts from the host model
    for epoch in range(trainer.num_epoch):
        hists: List[list] = [[] for _ in range(W)]
        errors: list = []
        if eager[0] and not host_synced:
            master.set_flat_weights(flats[0].cpu())
            trainer.master_model = serialize_keras_model(master)
            host_synced = True

        def run(w: int) -> None:
            try:
                dev = w % ndev
                torch.cuda.set_device(dev)
                with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                    try:
                        flat, h = _sequential_native(trainer, parts[w], master.copy(), dev, w, 1, init_flat=flats[w],
                                                     device_result=True)
                    except UnsupportedByNativeEngine:
                        # models the planner does not lower (tanh hidden layers, Reshape, ...): the
                        # autograd SequentialWorker trains this replica on the same GPU
                        from ..trainers import _run_tasks

                        eager[0] = True
                        res, ws = _run_tasks(trainer.allocate_worker(), [parts[w]], 1, lambda tid: f"cuda:{dev}")
                        flat = deserialize_keras_model(res[0][0]).get_flat_weights().to(f"cuda:{dev}", torch.float32)
                        h = ws[0].training_history
                    flats[w] = flat.contiguous()
                    torch.cuda.current_stream().synchronize()
                for rec in h:
                    rec["epoch"] = epoch
                hists[w] = h
            except BaseException as exc:
                errors.append(exc)

        if eager[0] and epoch > 0:
            # the autograd workers read trainer.master_model (host), which was refreshed above
            flats = [None] * W
        threads = [threading.Thread(target=run, args=(w,), daemon=True) for w in range(W)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for h in hists:
            history += h
        _average_replicas(flats, lib)
        host_synced = False
    if not host_synced and flats[0] is not None:
        master.set_flat_weights(flats[0].cpu())
    trainer.master_model = serialize_keras_model(master)
    return master, history


def _average_replicas(flats: List[torch.Tensor], lib) -> None:
    """In place: every tensor of ``flats`` becomes the element-wise mean.  One replica per device with full peer
    access: reduce-scatter + all-gather (device r owns slice r); otherwise one kernel on the first device over
    staged copies."""
    W = len(flats)
    n = flats[0].numel()
    devs = [f.device.index for f in flats]
    peer_ok = len(set(devs)) == W and W <= 16 and all(
        a == b or lib.dk_can_access_peer(a, b) for a in devs for b in devs)
    if peer_ok:
        for a in devs:
            for b in devs:
                if a != b:
                    N.check(lib.dk_enable_peer_access(a, b), "peer access")
        arr = (C.c_void_p * W)(*[f.data_ptr() for f in flats])
        per = (n + W - 1) // W
        per = (per + 3) // 4 * 4                       # slice bounds keep float4 alignment
        for r, f in enumerate(flats):
            lo, hi = min(n, r * per), min(n, (r + 1) * per)
            if hi <= lo:
                continue
            with torch.cuda.device(f.device):
                N.check(lib.dk_ps_average(arr, W, lo, hi, C.c_void_p(N.current_stream())), "dk_ps_average")
        for f in flats:
            torch.cuda.synchronize(f.device)
        return
    first = flats[0].device
    with torch.cuda.device(first):
        stack = torch.empty(W, (n + 3) // 4 * 4, dtype=torch.float32, device=first)   # rows stay 16-byte aligned
        for i, f in enumerate(flats):
            stack[i, :n].copy_(f)
        arr = (C.c_void_p * W)(*[stack[i].data_ptr() for i in range(W)])
        N.check(lib.dk_ps_average(arr, W, 0, n, C.c_void_p(N.current_stream())), "dk_ps_average")
        torch.cuda.synchronize(first)
        for f in flats:
            f.copy_(stack[0, :n])
        torch.cuda.synchronize(first)


def train_ensemble_native(trainer, dataset: Dataset):
    """``EnsembleTrainer``: independent replicas, one host thread + GPU each, native engine."""
    import threading

    ndev = torch.cuda.device_count()
    E = trainer.num_ensembles
    parts = dataset.repartition(E).partitions(E)
    master = deserialize_keras_model(trainer.master_model)
    models: List[Optional[object]] = [None] * E
    hists: List[list] = [[] for _ in range(E)]
    errors: list = []

    def run(i: int) -> None:
        try:
            dev = i % ndev
            torch.cuda.set_device(dev)
            m = master.copy()
            m.seed = (master.seed or 0) + 1 + i
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                try:
                    models[i], hists[i] = _sequential_native(trainer, parts[i], m, dev, i, trainer.num_epoch)
                except UnsupportedByNativeEngine:
                    from ..trainers import _run_tasks

                    res, ws = _run_tasks(trainer.allocate_worker(), [parts[i]], 1, lambda tid: f"cuda:{dev}")
                    models[i], hists[i] = deserialize_keras_model(res[0][0]), ws[0].training_history
                torch.cuda.current_stream().synchronize()
        except BaseException as exc:
            errors.append(exc)

    threads = [threading.Thread(target=run, args=(i,), daemon=True) for i in range(E)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return models, [h for hs in hists for h in hs]


# ------------------------------------------------------------------------------------------------
# CPU multi-process (gloo) path: rank 0 hosts the socket PS, every rank runs a worker
# ------------------------------------------------------------------------------------------------
def train_distributed_spmd_socket(trainer, dataset: Dataset):
    import threading

    rank, world = _spmd_env()
    dist = _init_pg("gloo")
    exchange_obj, barrier = _dist_helpers(dist)
    if type(trainer).__name__ in ("EASGD", "SynchronousDistributedTrainer"):
        raise RuntimeError("synchronous EASGD meets its workers on an in-process barrier; under torchrun use "
                           "AEASGD (asynchronous) or run it with backend='thread'")
    if rank == 0:
        trainer.parameter_server = trainer.allocate_parameter_server()
        trainer.parameter_server.master_port = 0
        # ranks may live on other hosts: advertise the rendezvous address, and only then leave loopback
        trainer.master_host = os.environ.get("MASTER_ADDR", "127.0.0.1")
        if trainer.master_host not in ("127.0.0.1", "localhost"):
            trainer.parameter_server.bind_host = "0.0.0.0"
        trainer.start_service()
        addr = (trainer.master_host, trainer.master_port)
    else:
        addr = None
    addr = exchange_obj(addr, 0)
    n_parts = world * max(1, int(trainer.parallelism_factor))
    parts = dataset.repartition(n_parts).partitions(n_parts)
    worker = trainer.allocate_worker()
    worker.master_host, worker.master_port = addr
    history: List[dict] = []
    import copy as _copy

    for idx in range(rank, n_parts, world):
        w = _copy.copy(worker)
        w.training_history, w.iteration = [], 1
        history += list(w.train(idx, parts[idx]))
    barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, history)
    box = [None]
    if rank == 0:
        trainer.stop_service()
        box = [serialize_keras_model(trainer.parameter_server.get_model())]
    dist.broadcast_object_list(box, src=0)
    return deserialize_keras_model(box[0]), [h for g in gathered for h in g]
