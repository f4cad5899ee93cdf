"""Tracing / profiling helpers (the reference only has ``time.time()`` around the Spark job,
``distkeras/trainers.py:75-93``): device timers on the launching stream, max-over-ranks reduction,
NVTX ranges for Nsight, and a structured logger."""
from __future__ import annotations

import contextlib
import json
import logging
import os
import time
from typing import Dict, List

import torch

_LOG = logging.getLogger("distkeras_b200")
if not _LOG.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s"))
    _LOG.addHandler(_h)
    _LOG.setLevel(os.environ.get("DK_LOG", "WARNING").upper())


def log_event(event: str, **fields) -> None:
    """One JSON line per event (structured logging; replaces the reference's bare ``print``)."""
    _LOG.info(json.dumps({"event": event, "t": time.time(), **fields}, default=str))


class DeviceTimer:
    """CUDA-event timer on the current stream (never time a kernel by wall clock)."""

    def __init__(self):
        self.spans: Dict[str, List[float]] = {}
        self._open = {}

    def start(self, name: str) -> None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._open[name] = ev

    def stop(self, name: str) -> None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.spans.setdefault(name, []).append((self._open.pop(name), ev))

    def summary(self) -> Dict[str, float]:
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.spans.items()}


def max_over_ranks(value_ms: float) -> float:
    """Multi-GPU timings are reported as the max over ranks."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range (visible in Nsight Systems / Compute timelines); no-op without CUDA."""
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
