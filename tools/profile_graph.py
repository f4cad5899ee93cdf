#!/usr/bin/env python
"""Warm, in-graph kernel times of a training region: replays the FabricWorker's CUDA graph under the torch
profiler (CUPTI activity records, no kernel serialisation, caches warm) and prints the mean duration of every
kernel, its launches per step and the gaps between consecutive kernels.

    python tools/profile_graph.py --model mnist_mlp --batch 64 --steps 24 [--out profiles/r2/graph_b64.txt]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch
from torch.profiler import ProfilerActivity, profile

from distkeras_b200 import trainers
from distkeras_b200.models import ZOO
from distkeras_b200.parallel.runtime import FabricWorker
from distkeras_b200.parameter_servers import FabricParameterServer

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mnist_mlp")
ap.add_argument("--algo", default="ADAG")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=24, help="steps per graph (a multiple of the window)")
ap.add_argument("--window", type=int, default=12)
ap.add_argument("--replays", type=int, default=10)
ap.add_argument("--optimizer", default="adam")
ap.add_argument("--no-fuse-comm", action="store_true")
ap.add_argument("--out", default=None)
a = ap.parse_args()

torch.cuda.set_device(0)
model = ZOO[a.model](seed=0)
model.build()
t = getattr(trainers, a.algo)(model, a.optimizer, "categorical_crossentropy", num_workers=1, batch_size=a.batch,
                              communication_window=a.window)
alg = t.algorithm()
ps = FabricParameterServer(model, device_index=0, kind=alg["kind"])
ps.initialize()
in_dtype = "f32" if a.model == "higgs_mlp" else "u8"
w = FabricWorker(model, t.worker_optimizer, t.loss, alg, ps.region, 0, a.batch, 0, in_dtype,
                 (1.0, 0.0) if in_dtype == "f32" else (1 / 255.0, 0.0), steps_per_graph=a.steps, fuse_comm=not a.no_fuse_comm)
for p in (0, 1):
    if in_dtype == "u8":
        w.x_stage[p].copy_(torch.randint(0, 256, w.x_stage[p].shape, dtype=torch.uint8))
    else:
        w.x_stage[p].copy_(torch.randn(w.x_stage[p].shape))
    w.y_stage[p].copy_(torch.randint(0, model.output_shape[-1], w.y_stage[p].shape).to(torch.int32))
w.initial_pull()
w.capture()
with torch.cuda.stream(w.compute):
    for r in range(4):
        w.replay(r & 1, a.steps, 0)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(w.compute):
    ev0.record(w.compute)
    for r in range(a.replays):
        w.replay(r & 1, a.steps, 0)
    ev1.record(w.compute)
torch.cuda.synchronize()
plain_us = 1e3 * ev0.elapsed_time(ev1) / (a.replays * a.steps)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.cuda.stream(w.compute):
        for r in range(a.replays):
            w.replay(r & 1, a.steps, 0)
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0),
             key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
gaps = []
for prev, e in zip([None] + evs[:-1], evs):
    d = agg.setdefault(e.name[:70], [0, 0.0])
    d[0] += 1
    d[1] += e.device_time_total
    if prev is not None:
        gaps.append(e.time_range.start - prev.time_range.end)
steps = a.replays * a.steps
lines = [f"# {a.model} {a.algo} batch {a.batch} window {a.window}: {a.steps} steps/graph x {a.replays} replays; "
         f"unprofiled {plain_us:.2f} us/step; kernels/step {w.kernels_per_step}; fused comm {w.fused_comm}",
         f"{'kernel':70s} {'launches/step':>13s} {'mean us':>9s} {'us/step':>9s}"]
tot = 0.0
for name, (n, us) in agg.items():
    lines.append(f"{name:70s} {n / steps:13.2f} {us / n:9.2f} {us / steps:9.2f}")
    tot += us / steps
gaps = [g for g in gaps if g < 1000]
lines.append(f"{'sum of kernel time':70s} {'':13s} {'':9s} {tot:9.2f}")
if gaps:
    lines.append(f"gap between consecutive kernels: mean {sum(gaps) / len(gaps):.2f} us, p50 {sorted(gaps)[len(gaps) // 2]:.2f} us "
                 f"(negative = overlap through programmatic dependent launch)")
txt = "\n".join(lines)
print(txt)
if a.out:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(txt + "\n")
w.rep.close()
ps.stop()
