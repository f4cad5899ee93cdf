#!/bin/bash
mkdir -p gpurun_out
B=${1:-64}
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -s 200 -c 10 -o gpurun_out/prof_b$B -f \
   python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B --reps 6 --skip-e2e > gpurun_out/ncu_full_b$B.log 2>&1
ls -la gpurun_out/prof_b$B.ncu-rep
