#!/usr/bin/env python
"""In-kernel timelines (DK_TRACE=1): SM-clock stamps of the first CTA of every traced kernel of a training step.

    DK_TRACE=1 python tools/kernel_timeline.py --model mnist_mlp --batch 64
stamps: 0 entry | 1 setup done (barriers, TMEM alloc, sync) | 2 predecessor complete (griddepcontrol.wait)
        3 first operand stage landed | 4 last stage landed (gemm) / accumulator ready (update) | 5 accumulator ready
        (gemm) / state tiles landed (update) | 6 epilogue done | 7 teardown done
"""
import argparse
import os
import sys

os.environ["DK_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200.models import ZOO
from distkeras_b200.parallel.engine import NativeReplica

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mnist_mlp")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--optimizer", default="adam")
a = ap.parse_args()
model = ZOO[a.model](seed=0)
model.build()
f32_in = a.model == "higgs_mlp"
rep = NativeReplica(model, a.optimizer, "categorical_crossentropy", a.batch, 0, in_dtype="f32" if f32_in else "u8",
                    input_affine=(1.0, 0.0) if f32_in else (1 / 255.0, 0.0))
x = (torch.randn(a.batch, rep._input_feats, device="cuda") if f32_in
     else torch.randint(0, 256, (a.batch, rep._input_feats), dtype=torch.uint8, device="cuda"))
y = torch.randint(0, model.output_shape[-1], (a.batch,), device="cuda").to(torch.int32)
g = torch.cuda.CUDAGraph()
for _ in range(3):
    rep.enqueue_step(x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.graph(g, stream=s):
    for _ in range(a.steps):
        rep.enqueue_step(x.data_ptr(), y.data_ptr())
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph of {a.steps} steps: {1e3 * e0.elapsed_time(e1) / a.steps:.2f} us/step (stamps are of the LAST step)")
mhz = 1965.0
tr = rep._trace.cpu()
for i, name in enumerate(rep._trace_names):
    t = tr[i].tolist()
    if t[0] == 0:
        continue
    rel = [(v - t[0]) / mhz if v else float("nan") for v in t]
    print(f"{name:60s} " + " ".join(f"{v:7.2f}" for v in rel) + "  us since entry")
