"""``backend="nccl"``: the library-only baseline of the parameter-server trainers.

SURVEY 7.2 (M3) asks for an end-to-end comparator built from stock parts only: the replica runs on the autograd
executor (cuBLAS / cuDNN kernels), the parameter exchange is ``torch.distributed`` collectives over NCCL, and no
kernel of this repository is on the path.  NCCL has no one-sided accumulate, so the asynchronous commit of the
reference (``distkeras/parameter_servers.py:266-292``) becomes its bulk-synchronous equivalent: every
``communication_window`` mini-batches ALL ranks all-reduce their update and apply the same center step --

* ADAG / DOWNPOUR / Experimental: ``C += sum_w scale (W_w - W1_w)``, ``W = W1 = C`` (scale = 1/tau for ADAG);
* DynSGD: the same with ``scale = 1 / world`` (in lock step every commit is ``world`` updates stale);
* (A)EASGD / EAMSGD: ``E_w = alpha (W_w - C)``, ``W_w -= E_w``, ``C += sum_w E_w``.

This is what the fabric backend (in-kernel NVLink atomics, graph-captured windows) is measured against in
``bench.py --backend nccl``; it is also a perfectly usable trainer on any NCCL/gloo cluster.
"""
from __future__ import annotations

import os
import time
from typing import List

import torch

from ..data import Dataset, Partition
from ..utils import deserialize_keras_model, serialize_keras_model
from ..utils.timing import log_event


def rank_train_nccl(trainer, dataset: Dataset, rank: int, world: int, exchange_obj, barrier) -> dict:
    from .replica import TorchReplica

    use_cuda = torch.cuda.is_available()
    if use_cuda:
        local = int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count()
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
    else:
        device = torch.device("cpu")
    dist = None
    if world > 1:
        import torch.distributed as dist

    alg = trainer.algorithm()
    kind, tau = alg["kind"], int(alg["window"])
    if kind == "custom":
        raise RuntimeError("backend='nccl' runs the built-in exchange rules only")
    model = deserialize_keras_model(trainer.master_model).to(device)
    rep = TorchReplica(model, trainer.worker_optimizer, trainer.loss, device=device, seed=getattr(trainer, "seed", 0) + rank,
                       loss_weights=trainer.loss_weights, metrics=trainer.metrics)
    W = rep.W.data
    center = W.clone()
    W1 = W.clone()
    buf = torch.empty_like(W)
    B = trainer.batch_size
    if getattr(trainer, "data_is_local_shard", False):
        part = dataset.partitions(1)[0]
        rows = min(exchange_obj(len(dataset) if rank == r else None, r) for r in range(world))
    else:
        parts = dataset.repartition(world).partitions(world)
        part = parts[rank]
        rows = min(len(p) for p in parts)
    rows = rows // B * B                       # lock step: every rank runs the same number of mini-batches
    part = Partition(part.dataset, part.index, part.start, part.start + rows)
    fcols = trainer.features_column if isinstance(trainer.features_column, (list, tuple)) else [trainer.features_column]
    lcols = trainer.label_column if isinstance(trainer.label_column, (list, tuple)) else [trainer.label_column]
    multi_in = int(getattr(model, "num_inputs", 1)) > 1
    if use_cuda:
        dataset.pin_memory()
    xs_all = [part.column(c) for c in fcols]
    ys_all = [part.column(c) for c in lcols]
    u8 = [x.dtype == torch.uint8 for x in xs_all]
    losses = rep.losses if len(rep.losses) == len(ys_all) else [rep.losses[0]] * len(ys_all)
    pre_batch = kind in ("downpour", "aeasgd", "eamsgd", "easgd")
    elastic = kind in ("aeasgd", "eamsgd", "easgd")
    scale = {"adag": 1.0 / tau, "experimental": 1.0 / tau, "dynsgd": 1.0 / world}.get(kind, 1.0)
    alpha = float(alg.get("alpha", 0.0))
    history: List[dict] = []
    num_updates = 1
    h2d = d2h = 0

    def exchange():
        nonlocal num_updates
        if elastic:
            torch.sub(W, center, out=buf)
            buf.mul_(alpha)
            W.sub_(buf)
        else:
            torch.sub(W, W1, out=buf)
            buf.mul_(scale)
        if dist is not None:
            dist.all_reduce(buf)
        center.add_(buf)
        if not elastic:
            W.copy_(center)
            W1.copy_(center)
        num_updates += world

    barrier()
    if use_cuda:
        torch.cuda.synchronize()
    t0 = time.time()
    ev0 = ev1 = None
    if use_cuda:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    it = 0
    n = rows // B
    mom = torch.zeros_like(W) if kind == "eamsgd" else None
    for _ in range(trainer.num_epoch):
        for b in range(n):
            it += 1
            lo, hi = b * B, (b + 1) * B
            xs = [x[lo:hi].to(device, non_blocking=True).float() for x in xs_all]
            xs = [x / 255.0 if is_u8 else x for x, is_u8 in zip(xs, u8)]
            h2d += sum(x[lo:hi].numel() * x.element_size() for x in xs_all)
            xin = xs if multi_in else (xs[0] if len(xs) == 1 else torch.cat([x.reshape(B, -1) for x in xs], dim=1))
            ys = []
            for y_col, ls in zip(ys_all, losses):
                y = y_col[lo:hi].to(device, non_blocking=True)
                ys.append(y.long() if (y.dim() == 1 and "crossentropy" in str(ls)) else y)
            if pre_batch and it % tau == 0:
                exchange()
            if mom is not None:                 # EAMSGD (reference workers.py:413-458): r <- mu r; train at W + r
                mom.mul_(float(alg["momentum"]))
                w_copy = W.clone()
                W.add_(mom)
            h = rep.train_on_batch(xin, ys if len(ys) > 1 else ys[0])
            if mom is not None:
                g = W - (w_copy + mom)          # the worker optimizer's step at the look-ahead point
                mom.sub_(g, alpha=float(alg["eta"]))
                W.copy_(w_copy - mom)
            d2h += 4 * len(h)
            history.append({"history": [float(v) for v in h], "worker_id": rank, "iteration": it, "timestamp": time.time()})
            if not pre_batch and it % tau == 0:
                exchange()
    if use_cuda:
        ev1.record()
        torch.cuda.synchronize()
    steps = it
    stats = {"executor": "TorchReplica+torch.distributed", "steps": steps, "windows": steps // tau, "exchanges": steps // tau,
             "h2d_bytes": h2d, "d2h_bytes": d2h, "seconds": time.time() - t0, "gpu_launches": 0, "kernels_per_window": 0,
             "device_ms": ev0.elapsed_time(ev1) if use_cuda else (time.time() - t0) * 1e3, "failures": []}
    log_event("nccl.worker_done", rank=rank, **stats)
    barrier()
    result = {"history": history, "stats": stats}
    if rank == 0:
        model.set_flat_weights(center.detach().cpu())
        result["model"] = serialize_keras_model(model.to("cpu"))
        result["num_updates"] = num_updates
        result["staleness_hist"] = [0] * 32
    return result
