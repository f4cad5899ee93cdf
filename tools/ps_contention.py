#!/usr/bin/env python
"""Parameter-server correctness under real contention: every rank but 0 (and rank 0 too with --server-writes)
hammers ONE center variable in rank 0's HBM through the NVLink kernels, concurrently.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/ps_contention.py \
        --out gpurun_out/ps_contention_8gpu.json

What is asserted (integer-valued fp32 data: every partial sum is exact, so any interleaving gives the same bits):

  * hogwild commit  (red.add.v4.f32):           center == C0 + rounds * sum_w delta_w, bit for bit
  * hogwild exchange (atom.add.v4.f32 returning) same; torn snapshots (a pulled vector that mixes two centers)
                                                 are counted -- they are allowed here (HogWild!, reference
                                                 ``parameter_servers.py:282-292`` takes no lock either)
  * strict (ticket lock + commit + pull):        same sum AND every pulled snapshot is one consistent center
                                                 (the reference's mutex behaviour, ``parameter_servers.py:266-268``)
  * DynSGD tickets:                              sum(staleness histogram) == commits, num_updates == commits
  * elastic (AEASGD):                            sum_w W_w + C is conserved (what leaves a worker enters the center)
  * sharded center (slice r in rank r's HBM):    assembled center == the unsharded result, bit for bit

Rank 0 prints one JSON document; exit code 1 on any violation.
"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--numel", type=int, default=3_000_000)
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--server-writes", action="store_true", help="rank 0 commits too (world writers)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0)) % torch.cuda.device_count()
    torch.cuda.set_device(local)
    d = runtime._init_pg("gloo")   # control plane only (several ranks may share a GPU: NCCL would refuse that)
    exchange_obj, barrier = runtime._dist_helpers(d)
    lib = N.lib()
    n, K = a.numel, a.rounds
    writers = [r for r in range(world) if r != 0 or a.server_writes or world == 1]
    is_writer = rank in writers
    st = C.c_void_p(N.current_stream())
    report = {"world": world, "writers": len(writers), "numel": n, "rounds": K, "gpus": torch.cuda.device_count(),
              "checks": {}}
    ok = True
    C0 = 1024.0

    def fresh_region():
        region = FabricRegion.create(torch.full((n,), C0), local) if rank == 0 else None
        info = exchange_obj(region.export() if rank == 0 else None, 0)
        if rank != 0:
            region = FabricRegion.open(info, local)
        barrier()
        return region

    def gather(value):
        box = [None] * world
        dist.all_gather_object(box, value)
        return box

    def finish(region):
        torch.cuda.synchronize()
        barrier()
        center = region.read_center() if rank == 0 else None
        ctrl = region.read_ctrl() if rank == 0 else None
        barrier()
        region.close()
        barrier()
        return center, ctrl

    # worker w always commits delta = w + 1 in every element (uniform vectors make torn snapshots visible)
    delta = float(rank + 1)
    total_delta = float(sum(r + 1 for r in writers))

    # ---- 1. hogwild commit -------------------------------------------------------------------------------
    region = fresh_region()
    c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
    w = torch.full((n,), delta, device="cuda")
    w1 = torch.zeros(n, device="cuda")
    if is_writer:
        for it in range(K):
            N.check(lib.dk_ps_commit(c, w.data_ptr(), w1.data_ptr(), n, 1.0, None, ctrl, rank, it, st), "commit")
    center, cw = finish(region)
    if rank == 0:
        want = C0 + K * total_delta
        good = bool((center == want).all()) and int(cw[N.CTRL_NUM_UPDATES]) == K * len(writers)
        report["checks"]["hogwild_commit_red_add"] = {"ok": good, "expected": want, "min": float(center.min()),
                                                      "max": float(center.max()), "num_updates": int(cw[N.CTRL_NUM_UPDATES])}
        ok &= good

    # ---- 2. hogwild exchange (returning atomics) and 3. strict commit + pull -------------------------------
    for name, strict in (("hogwild_exchange_atom_add", False), ("strict_lock_commit_pull", True)):
        region = fresh_region()
        c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
        w = torch.full((n,), C0, device="cuda")
        w1 = torch.full((n,), C0, device="cuda")
        ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
        lu = torch.zeros(1, dtype=torch.int32, device="cuda")
        torn = 0
        monotone = True
        last = C0
        if is_writer:
            for it in range(K):
                w.add_(delta)                       # a "window" of training moved every weight by delta
                if strict:
                    N.check(lib.dk_ps_lock_acquire(ctrl, ticket.data_ptr(), st), "lock")
                    N.check(lib.dk_ps_commit(c, w.data_ptr(), w1.data_ptr(), n, 1.0, None, ctrl, rank, it, st), "commit")
                    N.check(lib.dk_ps_pull(c, w.data_ptr(), w1.data_ptr(), None, n, ctrl, lu.data_ptr(), st), "pull")
                    N.check(lib.dk_ps_lock_release(ctrl, ticket.data_ptr(), st), "unlock")
                else:
                    N.check(lib.dk_ps_exchange(c, w.data_ptr(), w1.data_ptr(), None, n, 1.0, None, ctrl, rank, it,
                                               lu.data_ptr(), st), "exchange")
                lo, hi = float(w.min()), float(w.max())
                torn += int(lo != hi)
                monotone &= lo >= last               # the center only grows: a pull never goes back in time
                last = lo
                if not torch.equal(w, w1):
                    monotone = False
        center, cw = finish(region)
        stats = gather({"torn": torn, "monotone": monotone} if is_writer else None)
        if rank == 0:
            want = C0 + K * total_delta
            tor = sum(s["torn"] for s in stats if s)
            mono = all(s["monotone"] for s in stats if s)
            good = bool((center == want).all()) and mono and int(cw[N.CTRL_NUM_UPDATES]) == K * len(writers)
            if strict:
                good &= tor == 0
            report["checks"][name] = {"ok": good, "expected": want, "min": float(center.min()), "max": float(center.max()),
                                      "torn_snapshots": tor, "pulls": K * len(writers), "monotone": mono,
                                      "num_updates": int(cw[N.CTRL_NUM_UPDATES])}
            ok &= good

    # ---- 4. DynSGD tickets -------------------------------------------------------------------------------
    region = fresh_region()
    c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
    w = torch.full((n,), C0, device="cuda")
    w1 = torch.full((n,), C0, device="cuda")
    lu = torch.zeros(1, dtype=torch.int32, device="cuda")
    sc = torch.ones(1, dtype=torch.float32, device="cuda")
    scales = []
    if is_writer:
        for it in range(K):
            w.add_(delta)
            N.check(lib.dk_ps_ticket(ctrl, lu.data_ptr(), sc.data_ptr(), st), "ticket")
            N.check(lib.dk_ps_exchange(c, w.data_ptr(), w1.data_ptr(), None, n, 1.0, sc.data_ptr(), ctrl, rank, it,
                                       lu.data_ptr(), st), "exchange")
            scales.append(float(sc))
    center, cw = finish(region)
    stats = gather(scales if is_writer else None)
    if rank == 0:
        hist = cw[N.CTRL_STALENESS_HIST:N.CTRL_STALENESS_HIST + 32]
        commits = K * len(writers)
        all_scales = [s for ss in stats if ss for s in ss]
        good = int(hist.sum()) == commits and int(cw[N.CTRL_NUM_UPDATES]) == commits and \
            all(0.0 < s <= 1.0 for s in all_scales) and bool(torch.isfinite(center).all())
        report["checks"]["dynsgd_tickets"] = {"ok": good, "commits": commits, "histogram_sum": int(hist.sum()),
                                              "num_updates": int(cw[N.CTRL_NUM_UPDATES]),
                                              "staleness_histogram": [int(v) for v in hist],
                                              "mean_scale": sum(all_scales) / max(1, len(all_scales))}
        ok &= good

    # ---- 5. elastic: conservation ------------------------------------------------------------------------
    region = fresh_region()
    c, ctrl = C.c_void_p(region.center_ptr), C.c_void_p(region.ctrl_ptr)
    torch.manual_seed(rank)
    w = torch.randn(n, device="cuda") * 4 + C0
    before = float(w.double().sum()) if is_writer else 0.0
    if is_writer:
        for it in range(K):
            N.check(lib.dk_ps_elastic(c, w.data_ptr(), None, n, 0.125, ctrl, rank, it, st), "elastic")
    torch.cuda.synchronize()
    after = float(w.double().sum()) if is_writer else 0.0
    center, cw = finish(region)
    sums = gather((before, after))
    if rank == 0:
        tot_before = sum(s[0] for s in sums) + C0 * n
        tot_after = sum(s[1] for s in sums) + float(center.double().sum())
        rel = abs(tot_after - tot_before) / abs(tot_before)
        good = rel < 1e-6
        report["checks"]["elastic_conservation"] = {"ok": good, "relative_drift": rel}
        ok &= good

    # ---- 6. sharded center == unsharded ------------------------------------------------------------------
    per = ((n + world - 1) // world + 7) // 8 * 8
    bounds = [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]
    lo, hi = bounds[rank]
    mine = FabricRegion.create(torch.full((max(hi - lo, 8),), C0), local)
    infos = [exchange_obj(mine.export() if r == rank else None, r) for r in range(world)]
    regions = [mine if r == rank else FabricRegion.open(infos[r], local) for r in range(world)]
    barrier()
    w = torch.full((n,), C0, device="cuda")
    w1 = torch.full((n,), C0, device="cuda")
    if is_writer:
        for it in range(K):
            w.add_(delta)
            for r in range(world):
                slo, shi = bounds[r]
                if shi > slo:
                    N.check(lib.dk_ps_exchange(C.c_void_p(regions[r].center_ptr), w.data_ptr() + 4 * slo,
                                               w1.data_ptr() + 4 * slo, None, shi - slo, 1.0, None, None, rank, it,
                                               None, st), "exchange")
    torch.cuda.synchronize()
    barrier()
    piece = mine.read_center()[:max(hi - lo, 0)]
    pieces = gather(piece)
    barrier()
    for r, reg in enumerate(regions):
        if r != rank:
            reg.close()
    barrier()
    mine.close()
    if rank == 0:
        assembled = torch.cat(pieces)
        want = C0 + K * total_delta
        good = assembled.numel() == n and bool((assembled == want).all())
        report["checks"]["sharded_center"] = {"ok": good, "shards": world, "expected": want,
                                              "min": float(assembled.min()), "max": float(assembled.max())}
        ok &= good

    if rank == 0:
        report["ok"] = bool(ok)
        text = json.dumps(report, indent=1)
        print(text, flush=True)
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            with open(a.out, "w") as f:
                f.write(text + "\n")
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
