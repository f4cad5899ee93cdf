#!/bin/bash
# first GPU pass of round 2: gpu tests, bench at three batch sizes, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for B in 64 256 16384; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "rc=$?" >> gpurun_out/bench_b$B.err
done
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 --batch 64 > gpurun_out/ref_b64.json 2> gpurun_out/ref_b64.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 --batch 16384 > gpurun_out/ref_b16384.json 2> gpurun_out/ref_b16384.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_b*.json gpurun_out/ref_*.json | cut -c1-600
