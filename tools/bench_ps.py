#!/usr/bin/env python
"""Parameter-server push / pull micro-benchmark over NVLink (BASELINE.json: "PS push/pull GB/s vs
900 GB/s").  Launch with torchrun on >= 2 GPUs: rank 0 owns the center, every other rank is a writer.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_ps.py

Device-timed (CUDA events), max over the participating ranks; bytes counted per direction:
commit = 4N sent, pull = 4N received, exchange / elastic = 4N each way.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
import torch.distributed as dist

from distkeras_b200 import _native as N
from distkeras_b200.parallel import runtime
from distkeras_b200.parallel.fabric import FabricRegion


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    d = runtime._init_pg() if world > 1 else None
    exchange_obj, barrier = runtime._dist_helpers(d) if d else ((lambda o, s: o), (lambda: None))
    lib = N.lib()
    results = []
    sizes = [int(v) for v in os.environ.get("DK_PS_SIZES", "1000000,12000000,100000000").split(",")]
    for n in sizes:
        region = FabricRegion.create(torch.zeros(n), local) if rank == 0 else None
        info = exchange_obj(region.export() if rank == 0 else None, 0)
        if rank != 0:
            region = FabricRegion.open(info, local)
        w = torch.randn(n, device="cuda")
        w1 = torch.randn(n, device="cuda")
        wb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
        c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
        st = C.c_void_p(N.current_stream())
        ops = {
            "commit(red.v4)": lambda: lib.dk_ps_commit(c, w.data_ptr(), w1.data_ptr(), n, 1e-6, None, ctrl, rank, 0, st),
            "pull(ld.v4)": lambda: lib.dk_ps_pull(c, w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, ctrl, None, st),
            "exchange(atom.v4)": lambda: lib.dk_ps_exchange(c, w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, 1e-6, None,
                                                            ctrl, rank, 0, None, st),
            "elastic(ld+red)": lambda: lib.dk_ps_elastic(c, w.data_ptr(), wb.data_ptr(), n, 1e-6, ctrl, rank, 0, st),
            "strict(lock+commit+pull)": lambda: strict_exchange(),
        }
        ticket = torch.zeros(1, dtype=torch.int32, device="cuda")

        def strict_exchange():
            # the reference's mutex semantics: whole commit + pull sequences are serialised by the ticket lock
            lib.dk_ps_lock_acquire(ctrl, ticket.data_ptr(), st)
            lib.dk_ps_commit(c, w.data_ptr(), w1.data_ptr(), n, 1e-6, None, ctrl, rank, 0, st)
            lib.dk_ps_pull(c, w.data_ptr(), w1.data_ptr(), wb.data_ptr(), n, ctrl, None, st)
            return lib.dk_ps_lock_release(ctrl, ticket.data_ptr(), st)

        writer_sets = sorted({v for v in (1, 2, 4, world - 1) if 1 <= v <= world - 1}) if world > 1 else [1]
        for name, op in ops.items():
            for nw in writer_sets:
                active = (world == 1) or (1 <= rank <= nw)
                iters = 20 if n <= 12_000_000 else 5
                if active:
                    for _ in range(3):
                        op()
                torch.cuda.synchronize()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if active:
                    for _ in range(iters):
                        op()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters if active else 0.0
                if d:
                    t = torch.tensor([ms], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t)
                barrier()
                if rank == 0:
                    per_writer = 4.0 * n / (ms * 1e-3) / 1e9
                    results.append({"op": name, "n": n, "writers": nw if world > 1 else 0, "us": round(ms * 1e3, 2),
                                    "GBps_per_writer_per_dir": round(per_writer, 1),
                                    "GBps_ps_ingress_total": round(per_writer * (nw if world > 1 else 1), 1)})
                    print(json.dumps(results[-1]), flush=True)
        # ---- NCCL baseline (BASELINE.json: "a path that only calls NCCL is the baseline"): the same
        # commit / pull / exchange built from library calls -- elementwise kernels + ncclSend/ncclRecv
        # between the writer (rank 1) and the server (rank 0), server-side add as its own kernel
        if world > 1:
            buf = torch.empty(n, device="cuda")
            center_t = torch.zeros(n, device="cuda") if rank == 0 else None

            def nccl_commit():
                if rank == 1:
                    torch.sub(w, w1, out=buf)
                    buf.mul_(1e-6)
                    dist.send(buf, 0)
                elif rank == 0:
                    dist.recv(buf, 1)
                    center_t.add_(buf)

            def nccl_pull():
                if rank == 0:
                    dist.send(center_t, 1)
                elif rank == 1:
                    dist.recv(buf, 0)
                    w.copy_(buf)
                    w1.copy_(buf)
                    wb.copy_(buf)

            def nccl_exchange():
                nccl_commit()
                nccl_pull()

            for name, op in (("nccl:commit(sub+send|recv+add)", nccl_commit), ("nccl:pull(send|recv+3 copies)", nccl_pull),
                             ("nccl:exchange(commit+pull)", nccl_exchange)):
                active = rank in (0, 1)
                iters = 20 if n <= 12_000_000 else 5
                if active:
                    for _ in range(3):
                        op()
                torch.cuda.synchronize()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if active:
                    for _ in range(iters):
                        op()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / iters if active else 0.0], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
                barrier()
                if rank == 0:
                    per_writer = 4.0 * n / (ms * 1e-3) / 1e9
                    results.append({"op": name, "n": n, "writers": 1, "us": round(ms * 1e3, 2),
                                    "GBps_per_writer_per_dir": round(per_writer, 1), "GBps_ps_ingress_total": round(per_writer, 1)})
                    print(json.dumps(results[-1]), flush=True)
            del buf, center_t
        barrier()
        region.close()
        del w, w1, wb
        torch.cuda.empty_cache()
        barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"world": world, "results": results}, open(f"gpurun_out/bench_ps_w{world}.json", "w"), indent=1)
    if d:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
