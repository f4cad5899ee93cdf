#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_kernel|dense_bwd_update_kernel|input_stage" -s 12 -c 8 -o gpurun_out/r2_step_b64 -f \
   python tools/profile_step.py --model mnist_mlp --batch 64 --steps 6 > gpurun_out/ncu_step_b64.log 2>&1; echo "ncu mlp rc=$?"
ls -la gpurun_out/r2_step_b64.ncu-rep
