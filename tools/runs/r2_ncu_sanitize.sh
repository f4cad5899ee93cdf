#!/bin/bash
mkdir -p gpurun_out
# --- ncu --set full: one compact-program step of the flagship (4 kernels/step) and the conv kernels of the CIFAR step
timeout 600 ncu --set full --clock-control none --import-source on -s 12 -c 8 -o gpurun_out/r2_step_b64 -f \
   python tools/profile_step.py --model mnist_mlp --batch 64 --steps 6 > gpurun_out/ncu_step_b64.log 2>&1; echo "ncu mlp rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tma_kernel|conv_wgrad_tma_kernel" -s 14 -c 7 -o gpurun_out/r2_conv_b256 -f \
   python tools/profile_step.py --model cifar10_cnn --batch 256 --steps 4 > gpurun_out/ncu_conv_b256.log 2>&1; echo "ncu conv rc=$?"
ls -la gpurun_out/*.ncu-rep
# --- compute-sanitizer: memcheck + racecheck over the PS kernels, the fused update kernel, the head-in-forward GEMM, TMA conv
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 \
    python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ps_kernels or fused_dense_backward_update_kernel or tma_im2col_conv_forward or optimizer_matches" \
    > gpurun_out/sanitize_${tool}_kernels.log 2>&1; echo "$tool kernels rc=$?"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 \
    python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "head_in_forward_gemm_matches_head_kernel and 64" \
    > gpurun_out/sanitize_${tool}_head.log 2>&1; echo "$tool head rc=$?"
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_*.log
done
