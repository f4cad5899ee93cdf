#!/usr/bin/env python
"""Roofline fractions for the bench results in profiles/ against MEASURED_PEAKS.json (driver-measured):
model FLOPs per step (forward + dgrad + wgrad of every Dense / Conv2D, no dgrad for the first layer) ->
achieved TFLOP/s per GPU vs the sustained cuBLAS bf16 figure; end-to-end H2D GB/s per GPU.

    python tools/roofline.py > profiles/roofline.md
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from distkeras_b200.models import ZOO  # noqa: E402


def _conv_macs(conv, in_shape) -> float:
    oh, ow, co = conv.output_shape(in_shape)
    return float(oh) * ow * co * conv.kernel_size[0] * conv.kernel_size[1] * int(in_shape[2])


def flops_per_sample(model) -> float:
    model.build()
    total, first, shape = 0.0, True, tuple(model.input_shape)
    for layer in model.layers:
        out = tuple(layer.output_shape(shape))
        macs = []
        if layer.class_name == "Dense":
            macs = [float(shape[-1]) * out[-1]]
        elif layer.class_name in ("Conv2D", "Convolution2D"):
            macs = [_conv_macs(layer, shape)]
        elif layer.class_name == "ResidualBlock":
            mid = tuple(layer.conv1.output_shape(shape))
            macs = [_conv_macs(layer.conv1, shape), _conv_macs(layer.conv2, mid)]
            if getattr(layer, "proj", None) is not None:
                macs.append(_conv_macs(layer.proj, shape))
        for m in macs:
            total += 2.0 * m * (2 if first else 3)  # fwd + wgrad (+ dgrad unless it is the first layer)
            first = False
        shape = out
    return total


def main() -> None:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"bf16_tflops_sustained": 1440.3, "bf16_tflops": 1710.4, "hbm_gbs": 6583.8}
    print("# Roofline fractions of the committed bench results\n")
    print(f"Denominators: cuBLAS bf16 sustained {peaks['bf16_tflops_sustained']} TFLOP/s (burst {peaks['bf16_tflops']}), "
          f"HBM copy {peaks['hbm_gbs']} GB/s, PCIe Gen5 x16 host->device ~57 GB/s practical.  Regenerate with `tools/roofline.py`.\n")
    print("| file | model / trainer | GPUs | batch/worker | us/step | GFLOP/step/GPU | TFLOP/s/GPU | of sustained bf16 | e2e H2D GB/s/GPU |")
    print("|---|---|---|---|---|---|---|---|---|")
    cache = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_*.json"))):
        try:
            d = json.loads(open(path).read().strip().splitlines()[-1])
        except (ValueError, IndexError):
            continue
        cfg = d.get("config", {})
        name = cfg.get("model")
        if name not in ZOO:
            continue
        if name not in cache:
            cache[name] = flops_per_sample(ZOO[name](seed=0))
        B = cfg.get("batch_per_worker") or cfg.get("global_batch", 0) // max(1, d.get("n_gpus", 1))
        gflop = cache[name] * B / 1e9
        us = d["ms_per_step"] * 1e3
        tflops = gflop / us * 1e3  # GFLOP per us = PFLOP/s -> x1e3 for TFLOP/s
        e2e = d.get("e2e") or {}
        h2d = (e2e.get("h2d_bytes_per_step", 0) / max(1, d.get("n_gpus", 1))) / (e2e.get("ms_per_step", 0) * 1e-3) / 1e9 \
            if e2e.get("ms_per_step") else None
        print(f"| `{os.path.basename(path)}` | {name} / {cfg.get('trainer')} | {d.get('n_gpus')} | {B} | {us:.1f} | {gflop:.1f} | "
              f"{tflops:.0f} | {100 * tflops / peaks['bf16_tflops_sustained']:.0f} % | {'%.1f' % h2d if h2d else '-'} |")


if __name__ == "__main__":
    main()
