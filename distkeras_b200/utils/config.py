"""Runtime configuration: constructor kwargs stay the primary interface, like the
reference; this dataclass collects the process-wide switches with environment overrides."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional


def _env_bool(name: str, default: bool) -> bool:
    v = os.environ.get(name)
    return default if v is None else v.strip().lower() in ("1", "true", "yes", "on")


@dataclass
class RuntimeConfig:
    backend: Optional[str] = field(default_factory=lambda: os.environ.get("DK_BACKEND"))  # fabric|thread|socket
    strict: bool = field(default_factory=lambda: _env_bool("DK_STRICT", False))           # ticket-lock commits
    comm: str = field(default_factory=lambda: os.environ.get("DK_COMM", "exchange"))      # exchange|commit_pull
    dedicated_ps: bool = field(default_factory=lambda: _env_bool("DK_DEDICATED_PS", False))
    fault: Optional[str] = field(default_factory=lambda: os.environ.get("DK_FAULT"))      # "<worker>:<iteration>"
    log_level: str = field(default_factory=lambda: os.environ.get("DK_LOG", "WARNING"))

    def apply(self, trainer) -> None:
        if self.backend:
            trainer.backend = self.backend
        trainer.strict = self.strict
        trainer.comm = self.comm
        trainer.dedicated_ps = self.dedicated_ps


_FIRED: set = set()


def fault_injection_point(worker_id: int, iteration: int) -> None:
    """Test hook: ``DK_FAULT=<worker>:<iteration>`` makes that worker fail there.  Like
    a real crash the fault fires ONCE per process and spec, so a retried task gets past it.
    ``DK_FAULT_KILL=<worker>:<iteration>`` ends the whole process there instead (``os._exit``): the rank-loss case
    the fabric launcher survives by re-queueing the dead rank's shards."""
    kill = os.environ.get("DK_FAULT_KILL")
    if kill:
        # process-level loss (a rank that disappears: OOM kill, node failure): no exception, no cleanup
        w, it = kill.split(":")
        if int(w) == int(worker_id) and int(it) == int(iteration):
            os._exit(17)
    spec = os.environ.get("DK_FAULT")
    if not spec or spec in _FIRED:
        return
    w, it = spec.split(":")
    if int(w) == int(worker_id) and int(it) == int(iteration):
        _FIRED.add(spec)
        raise RuntimeError(f"injected fault: worker {worker_id} at iteration {iteration}")
