#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -20 | tee gpurun_out/timeline_b64.txt
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -20 | tee gpurun_out/timeline_b64_nopdl.txt
