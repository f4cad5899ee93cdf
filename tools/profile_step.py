#!/usr/bin/env python
"""Run a few eager training steps of a zoo model through the native engine (for ncu / timing).

    ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --model mnist_mlp --batch 4096
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200.models import ZOO
from distkeras_b200.parallel.engine import NativeReplica

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mnist_mlp")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--optimizer", default="adam")
a = ap.parse_args()

model = ZOO[a.model](seed=0)
model.build()
rep = NativeReplica(model, a.optimizer, "categorical_crossentropy", a.batch, 0, in_dtype="u8",
                    input_affine=(1 / 255.0, 0.0))
feat = rep._input_feats
x = torch.randint(0, 256, (a.batch, feat), dtype=torch.uint8, device="cuda")
y = torch.randint(0, model.output_shape[-1], (a.batch,), device="cuda").to(torch.int32)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
for i in range(a.steps):
    ev[i].record()
    rep.enqueue_step(x.data_ptr(), y.data_ptr())
ev[a.steps].record()
torch.cuda.synchronize()
print("kernels/step", rep.step_kernel_count(), "ms/step", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(a.steps)])
