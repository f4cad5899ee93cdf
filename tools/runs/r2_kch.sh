#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "head_in_forward or compact_program or not_multiples or multicast or fused_into_backward or adag_matches" > gpurun_out/pytest_kch.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kch.log; tail -6 gpurun_out/pytest_kch.log
timeout 300 python bench.py --skip-e2e > gpurun_out/bench_kch_b64.json 2> gpurun_out/bench_kch_b64.err; echo "rc=$?"; tail -2 gpurun_out/bench_kch_b64.err
DK_GEMM_KCH=0 timeout 300 python bench.py --skip-e2e > gpurun_out/bench_nokch_b64.json 2> gpurun_out/bench_nokch_b64.err; echo "rc=$?"
DK_SPLIT_M=0 timeout 300 python bench.py --skip-e2e > gpurun_out/bench_kch_nosplit_b64.json 2> gpurun_out/bench_kch_nosplit_b64.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_kch_b64", "bench_nokch_b64", "bench_kch_nosplit_b64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", d["kernels_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -6 | cut -c1-200
