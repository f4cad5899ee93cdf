#!/bin/bash
# validation of the multi-GPU tooling on N GPUs (N = all visible): full GPU test-suite, contention checks, PS table
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${N}gpu.log
tail -6 gpurun_out/pytest_gpu_${N}gpu.log
port() { echo $((29500 + RANDOM % 400)); }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) \
  tools/ps_contention.py --out gpurun_out/ps_contention_${N}gpu.json > gpurun_out/ps_contention_${N}gpu.log 2>&1; echo "contention rc=$?"
grep -E '"ok"|torn|expected' gpurun_out/ps_contention_${N}gpu.json | head -30
DK_PS_SIZES=${DK_PS_SIZES:-1000000,12000000,100000000} timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) \
  tools/bench_ps.py > gpurun_out/bench_ps_${N}gpu.log 2>&1; echo "bench_ps rc=$?"
grep -c '"op"' gpurun_out/bench_ps_${N}gpu.log
grep -E 'exchange\(atom|strict' gpurun_out/bench_ps_${N}gpu.log | head -20
