#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_dense_backward" > gpurun_out/t_kernel.log 2>&1; echo "rc=$?" >> gpurun_out/t_kernel.log
tail -15 gpurun_out/t_kernel.log
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log
tail -25 gpurun_out/t_engine.log
for B in 64 256; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B > gpurun_out/c_bench_b$B.json 2> gpurun_out/c_bench_b$B.err; echo "rc=$?" >> gpurun_out/c_bench_b$B.err
  tail -3 gpurun_out/c_bench_b$B.err
done
cat gpurun_out/c_bench_b*.json | cut -c1-400
