#!/bin/bash
mkdir -p gpurun_out
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -8 | tee gpurun_out/timeline2_b64_nopdl.txt
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 --optimizer sgd 2>&1 | tail -8 | tee gpurun_out/timeline2_b64_sgd_nopdl.txt
timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -8 | tee gpurun_out/timeline2_b64.txt
