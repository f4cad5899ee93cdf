#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_${N}gpu.log
tail -15 gpurun_out/pytest_gpu_${N}gpu.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_${N}gpu.log
tail -5 gpurun_out/pytest_multi_${N}gpu.log
