#!/usr/bin/env bash
# Race / memory checking of the native kernels: run on a GPU box, e.g.
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'
# Tools: memcheck (default), racecheck (shared-memory hazards), synccheck (barrier misuse).
set -euo pipefail
tool="${1:-memcheck}"
python build_native.py --tests >/dev/null
echo "== compute-sanitizer --tool $tool: standalone GEMM harness (quick mode) =="
compute-sanitizer --tool "$tool" --error-exitcode 3 build/gemm_test quick | tail -5
echo "== compute-sanitizer --tool $tool: PS / optimizer / loss / layout kernels =="
compute-sanitizer --tool "$tool" --error-exitcode 3 \
  python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ps_kernels or optimizer or xent or input_stage or maxpool or im2col" | tail -5
