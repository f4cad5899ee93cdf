#!/bin/bash
mkdir -p gpurun_out
B=${1:-64}
# per-launch device time of the steady-state step (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/launches_b$B.csv \
   python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B --reps 6 --skip-e2e > gpurun_out/ncu_b$B.log 2>&1
python tools/ncu_launches.py gpurun_out/launches_b$B.csv 2>/dev/null | head -40
