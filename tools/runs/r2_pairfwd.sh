#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "gradients or learns_with_dropout" > gpurun_out/pytest_pairfwd.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pairfwd.log; tail -4 gpurun_out/pytest_pairfwd.log
timeout 300 python bench.py --batch 16384 --skip-e2e > gpurun_out/bench_pairfwd.json 2> gpurun_out/bench_pairfwd.err; echo "rc=$?"; tail -2 gpurun_out/bench_pairfwd.err
DK_PAIR_FWD=0 timeout 300 python bench.py --batch 16384 --skip-e2e > gpurun_out/bench_nopairfwd.json 2> gpurun_out/bench_nopairfwd.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_pairfwd", "bench_nopairfwd"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", d["kernels_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
DK_PDL=0 timeout 300 python tools/profile_graph.py --batch 16384 --steps 12 --replays 4 2>&1 | grep -E "gemm|head" | cut -c1-120
