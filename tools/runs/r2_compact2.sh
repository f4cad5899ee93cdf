#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_dense_backward or optimizer" > gpurun_out/t_kernel.log 2>&1; echo "rc=$?" >> gpurun_out/t_kernel.log
tail -5 gpurun_out/t_kernel.log
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -k "compact or fused_into or oracle or learn" > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log
tail -5 gpurun_out/t_engine.log
timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64.txt 2>&1 | tail -20
timeout 300 python tools/profile_graph.py --batch 256 --out gpurun_out/graph_b256.txt 2>&1 | tail -14
for B in 64; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B --skip-e2e > gpurun_out/c_bench_b$B.json 2> gpurun_out/c_bench_b$B.err; echo "rc=$?" >> gpurun_out/c_bench_b$B.err
  tail -3 gpurun_out/c_bench_b$B.err
done
cat gpurun_out/c_bench_b*.json | cut -c1-300
