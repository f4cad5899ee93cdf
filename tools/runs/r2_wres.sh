#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -m gpu -k "tma_im2col or gradients" > gpurun_out/pytest_wres.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_wres.log; tail -5 gpurun_out/pytest_wres.log
timeout 300 python bench.py --model cifar10_cnn --algo downpour --batch 256 --skip-e2e > gpurun_out/bench_wres_cifar.json 2> gpurun_out/bench_wres_cifar.err; echo "rc=$?"; tail -2 gpurun_out/bench_wres_cifar.err
DK_CONV_WRES=0 timeout 300 python bench.py --model cifar10_cnn --algo downpour --batch 256 --skip-e2e > gpurun_out/bench_nowres_cifar.json 2> gpurun_out/bench_nowres_cifar.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_wres_cifar", "bench_nowres_cifar"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), d["kernels_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
DK_PDL=0 timeout 300 python tools/profile_graph.py --model cifar10_cnn --algo DOWNPOUR --batch 256 --steps 12 --replays 4 2>&1 | grep -E "conv_|maxpool|us/step" | cut -c1-120
