"""NUMA placement for the host side of the input pipeline.

Every rank streams its mini-batches from pinned host memory over its own PCIe link; on a two-socket
8-GPU node a rank whose pinned buffers live on the *other* socket pays a cross-socket hop for every
DMA read.  ``bind_to_gpu_numa_node`` pins the calling process (and therefore its first-touch /
``cudaHostAlloc`` allocations) to the CPUs of the NUMA node the GPU hangs off.
"""
from __future__ import annotations

import os
from typing import Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Returns the NUMA node the process was bound to, or ``None`` if nothing was changed."""
    if os.environ.get("DK_NUMA", "1") == "0":
        return None
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            return node
    except Exception:
        pass
    return None
