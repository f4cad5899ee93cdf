#!/bin/bash
mkdir -p gpurun_out
DK_PDL=0 timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64_nopdl.txt 2>&1 | grep -v Warn | tail -14
DK_PDL=0 DK_COMPACT=0 timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64_classic_nopdl.txt 2>&1 | grep -v Warn | tail -20
