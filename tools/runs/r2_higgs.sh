#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "not_multiples_of_8 or compact_program or gradients or trainers_learn or fused_into_backward or strict_mode or adag_matches" > gpurun_out/pytest_higgs.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_higgs.log; tail -12 gpurun_out/pytest_higgs.log
timeout 300 python bench.py --model higgs_mlp --algo downpour --optimizer adagrad --batch 64 > gpurun_out/higgs_compact_b64.json 2> gpurun_out/higgs_compact_b64.err; echo "rc=$?"; tail -2 gpurun_out/higgs_compact_b64.err
DK_COMPACT=0 timeout 300 python bench.py --model higgs_mlp --algo downpour --optimizer adagrad --batch 64 --skip-e2e > gpurun_out/higgs_classic_b64.json 2> gpurun_out/higgs_classic_b64.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("higgs_compact_b64", "higgs_classic_b64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), d["kernels_per_step"], d["config"]["program"][:20])
    except Exception as e: print(f, "ERR", e)
PY
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --model higgs_mlp --batch 64 --optimizer adagrad 2>&1 | tail -12 | cut -c1-200
