#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_multi.py tests/test_functional_models.py -q -m gpu -k "contention" > gpurun_out/pytest_new_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new_1gpu.log
tail -25 gpurun_out/pytest_new_1gpu.log
