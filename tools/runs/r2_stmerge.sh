#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -m gpu -k "fused_dense_backward or compact_program or not_multiples or fused_into_backward or head_in_forward or trainers_learn" > gpurun_out/pytest_stmerge.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stmerge.log; tail -5 gpurun_out/pytest_stmerge.log
for i in 1 2; do timeout 300 python bench.py --skip-e2e > gpurun_out/bench_stmerge_b64_$i.json 2> gpurun_out/bench_stmerge_b64.err; echo "rc=$?"; done
timeout 300 python bench.py --skip-e2e --batch 256 > gpurun_out/bench_stmerge_b256.json 2> gpurun_out/bench_stmerge_b256.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_stmerge_b64_1", "bench_stmerge_b64_2", "bench_stmerge_b256"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", d["kernels_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -3 | cut -c1-200
