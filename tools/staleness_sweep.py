#!/usr/bin/env python
"""BASELINE config 5: DynSGD on ResNet-18 (synthetic 224x224) -- staleness sweep.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/staleness_sweep.py

For each communication window the job trains a few hundred mini-batches per worker and reports the
whole-job samples/s (device-timed, max over ranks), the parameter server's staleness histogram
(bucket s = commits that arrived s updates late, `parameter_servers.py:342-354` semantics) and the
loss trajectory.  ResNet-18's BatchNorm / residual blocks are not lowered by the native planner yet,
so replicas run on the autograd executor; every commit / pull is the in-kernel NVLink program.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np
import torch

from distkeras_b200.data import synthetic_imagenet
from distkeras_b200.models import resnet18
from distkeras_b200.trainers import DynSGD

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--classes", type=int, default=1000)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--windows", default="1,5,20")
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
out = []
for window in [int(w) for w in a.windows.split(",")]:
    ds = synthetic_imagenet(a.batch * a.steps, a.size, a.classes, seed=rank)  # this rank's shard
    t = DynSGD(resnet18((a.size, a.size, 3), a.classes, seed=0), {"class_name": "sgd", "config": {"lr": 0.05, "momentum": 0.9}},
               "categorical_crossentropy", num_workers=world, batch_size=a.batch, communication_window=window)
    t.backend = "fabric"
    t.data_is_local_shard = True
    t.bench_warmup_steps = 0
    t.train(ds)
    if rank == 0:
        stats = [s for s in t.fabric_stats if s.get("steps")]
        ms = max(s["device_ms"] for s in stats)
        steps = min(s["steps"] for s in stats)
        h = t.get_history()
        hist = [int(v) for v in t.staleness_histogram]
        rec = {"gpus": world, "window": window, "batch_per_worker": a.batch, "steps_per_worker": steps,
               "samples_per_s": world * a.batch * steps / (ms * 1e-3), "num_updates": t.num_updates(),
               "staleness_hist": {str(i): v for i, v in enumerate(hist) if v},
               "mean_staleness": float(np.average(np.arange(len(hist)), weights=hist)) if sum(hist) else None,
               "loss_first": float(np.mean([r["history"][0] for r in h[:world]])),
               "loss_last": float(np.mean([r["history"][0] for r in h[-world:]])),
               "executor": stats[0].get("executor")}
        out.append(rec)
        print(json.dumps(rec), flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/staleness_sweep_w{world}.json", "w"), indent=1)
