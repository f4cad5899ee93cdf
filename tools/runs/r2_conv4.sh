#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv_wgrad" > gpurun_out/t_conv.log 2>&1; echo "rc=$?" >> gpurun_out/t_conv.log
tail -30 gpurun_out/t_conv.log
