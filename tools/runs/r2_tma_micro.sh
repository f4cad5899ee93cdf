#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_request tools/microbench/tma_request.cu -lcuda && timeout 120 /tmp/tma_request | tee gpurun_out/tma_request.txt
