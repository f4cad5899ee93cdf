#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_graph.py --batch 16384 --steps 12 --replays 4 --out gpurun_out/graph_b16384.txt 2>&1 | grep -v Warn | tail -24
DK_PDL=0 timeout 300 python tools/profile_graph.py --batch 16384 --steps 12 --replays 4 --out gpurun_out/graph_b16384_nopdl.txt 2>&1 | grep -v Warn | tail -24
timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64_final.txt 2>&1 | grep -v Warn | tail -12
