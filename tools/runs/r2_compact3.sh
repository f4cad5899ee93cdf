#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/t_kernel.log 2>&1; echo "rc=$?" >> gpurun_out/t_kernel.log
tail -5 gpurun_out/t_kernel.log
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log
tail -5 gpurun_out/t_engine.log
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -6 | tee gpurun_out/timeline_b64_nopdl.txt
timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64.txt 2>&1 | grep -v -i warn | tail -12
DK_PDL=0 timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64_nopdl.txt 2>&1 | grep -v -i warn | tail -12
