"""NVLink fabric: device memory shared between the PS GPU and the worker GPUs.

One :class:`FabricRegion` = ``[control block | center variable]`` in a single raw ``cudaMalloc``
allocation on the PS GPU.  The owner exports a CUDA-IPC handle; every worker process opens it and
receives a device pointer that is valid *inside its own kernels* -- loads, stores and
``red`` / ``atom`` instructions on it travel over NVLink 5 / NVSwitch.  In single-process
multi-GPU mode (threads) the raw pointer is used directly after enabling peer access.

This is the transport that replaces the reference's TCP star (``distkeras/networking.py``,
SURVEY 2.4): there is no message, no serialisation and no server thread.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .. import _native

_CTRL_BYTES = 1024  # control block padded so the center stays 1 KiB aligned


class FabricRegion:
    def __init__(self, base: int, numel: int, owner_device: int, owns: bool, via_ipc: bool):
        self.base = int(base)
        self.numel = int(numel)
        self.owner_device = int(owner_device)
        self.owns = owns
        self.via_ipc = via_ipc
        self.closed = False

    # -- addresses ----------------------------------------------------------------------------
    @property
    def ctrl_ptr(self) -> int:
        return self.base

    @property
    def center_ptr(self) -> int:
        return self.base + _CTRL_BYTES

    @property
    def nbytes(self) -> int:
        return _CTRL_BYTES + ((self.numel * 4 + 1023) // 1024) * 1024

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def create(cls, init_flat: torch.Tensor, device_index: int) -> "FabricRegion":
        lib = _native.lib()
        numel = int(init_flat.numel())
        torch.cuda.set_device(device_index)
        _native.check(lib.dk_set_device(device_index), "dk_set_device")
        out = C.c_void_p()
        nbytes = _CTRL_BYTES + ((numel * 4 + 1023) // 1024) * 1024
        _native.check(lib.dk_fabric_alloc(nbytes, C.byref(out)), "dk_fabric_alloc")
        region = cls(out.value, numel, device_index, owns=True, via_ipc=False)
        region.write_center(init_flat)
        return region

    def export(self) -> dict:
        handle = (C.c_char * 64)()
        _native.check(_native.lib().dk_ipc_export(C.c_void_p(self.base), handle), "dk_ipc_export")
        return {"handle": bytes(handle), "numel": self.numel, "device": self.owner_device, "pid": os.getpid(),
                "base": self.base}

    @classmethod
    def open(cls, info: dict, local_device: int) -> "FabricRegion":
        """Map an exported region into this process / device."""
        lib = _native.lib()
        torch.cuda.set_device(local_device)
        _native.check(lib.dk_set_device(local_device), "dk_set_device")
        if info["pid"] == os.getpid():
            if local_device != info["device"]:
                if not lib.dk_can_access_peer(local_device, info["device"]):
                    raise RuntimeError(f"GPU {local_device} cannot peer-access GPU {info['device']}")
                _native.check(lib.dk_enable_peer_access(local_device, info["device"]), "enable_peer_access")
            return cls(info["base"], info["numel"], info["device"], owns=False, via_ipc=False)
        out = C.c_void_p()
        buf = (C.c_char * 64).from_buffer_copy(info["handle"])
        _native.check(lib.dk_ipc_open(buf, C.byref(out)), "dk_ipc_open")
        return cls(out.value, info["numel"], info["device"], owns=False, via_ipc=True)

    # -- host access (setup / teardown only; the hot path never goes through the host) -----------
    def write_center(self, flat: torch.Tensor) -> None:
        src = flat.detach().to("cpu", torch.float32).contiguous()
        lib = _native.lib()
        _native.check(lib.dk_memcpy_async(C.c_void_p(self.center_ptr), C.c_void_p(src.data_ptr()),
                                          self.numel * 4, 1, None), "memcpy H2D")
        _native.check(lib.dk_stream_sync(None), "sync")

    def read_center(self) -> torch.Tensor:
        dst = torch.empty(self.numel, dtype=torch.float32)
        lib = _native.lib()
        _native.check(lib.dk_device_sync(), "sync")
        _native.check(lib.dk_memcpy_async(C.c_void_p(dst.data_ptr()), C.c_void_p(self.center_ptr),
                                          self.numel * 4, 2, None), "memcpy D2H")
        _native.check(lib.dk_stream_sync(None), "sync")
        return dst

    def read_ctrl(self) -> np.ndarray:
        dst = np.zeros(_native.CTRL_WORDS, dtype=np.uint32)
        lib = _native.lib()
        _native.check(lib.dk_memcpy_async(C.c_void_p(dst.ctypes.data), C.c_void_p(self.ctrl_ptr),
                                          dst.nbytes, 2, None), "memcpy D2H")
        _native.check(lib.dk_stream_sync(None), "sync")
        return dst

    def close(self) -> None:
        if self.closed:
            return
        lib = _native.lib()
        if self.owns:
            lib.dk_device_sync()
            lib.dk_fabric_free(C.c_void_p(self.base))
        elif self.via_ipc:
            lib.dk_ipc_close(C.c_void_p(self.base))
        self.closed = True


def local_rank_device(rank: Optional[int] = None) -> int:
    if rank is None:
        rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = torch.cuda.device_count()
    return rank % max(n, 1)
