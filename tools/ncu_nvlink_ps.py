#!/usr/bin/env python
"""One process, two GPUs: the center variable lives in GPU 1's HBM, the PS kernels run on GPU 0 through the peer
mapping -- the same instructions (`red` / `atom` / `ld` `.sys`) that cross-process workers issue, but profilable by ncu
(which must not wrap a multi-rank launch).  Used for the NVLink byte counters:

    ncu --set full --section Nvlink --section Nvlink_Tables --section Nvlink_Topology -k regex:ps_ -o gpurun_out/r2_ps_nvlink \
        python tools/ncu_nvlink_ps.py --numel 12000000
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200 import _native as N
from distkeras_b200.parallel.fabric import FabricRegion

ap = argparse.ArgumentParser()
ap.add_argument("--numel", type=int, default=12_000_000)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
assert torch.cuda.device_count() >= 2, "needs two GPUs"
n = a.numel
owner = FabricRegion.create(torch.zeros(n), 1)                  # center in GPU 1's HBM
region = FabricRegion.open(owner.export(), 0)                   # same process: peer access from GPU 0
torch.cuda.set_device(0)
lib = N.lib()
w = torch.randn(n, device="cuda:0")
w1 = torch.randn(n, device="cuda:0")
wb = torch.zeros(n, dtype=torch.bfloat16, device="cuda:0")
c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
st = C.c_void_p(N.current_stream())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for name, op in (("commit", lambda: lib.dk_ps_commit(c, w.data_ptr(), w1.data_ptr(), n, 1e-6, None, ctrl, 0, 0, st)),
                 ("pull", lambda: lib.dk_ps_pull(c, w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, ctrl, None, st)),
                 ("exchange", lambda: lib.dk_ps_exchange(c, w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, 1e-6, None, ctrl, 0, 0, None, st)),
                 ("elastic", lambda: lib.dk_ps_elastic(c, w.data_ptr(), wb.data_ptr(), n, 1e-6, ctrl, 0, 0, st))):
    op()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(a.iters):
        op()
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / a.iters
    print(f"{name:9s} n={n} {us:9.1f} us  {4.0 * n / us / 1e3:7.1f} GB/s per direction (bytes each way: {4 * n})", flush=True)
