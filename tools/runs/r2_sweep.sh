#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "head_in_forward" > gpurun_out/pytest_smallb.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_smallb.log; tail -4 gpurun_out/pytest_smallb.log
for B in 4 16 32 64 128 256; do
  timeout 300 python bench.py --batch $B > gpurun_out/sweep_mlp_b$B.json 2> gpurun_out/sweep_mlp_b$B.err; echo "b$B rc=$?"
done
timeout 300 python bench.py --model higgs_mlp --algo downpour --optimizer adagrad --batch 64 > gpurun_out/sweep_higgs_downpour_b64.json 2> gpurun_out/sweep_higgs_downpour_b64.err; echo "higgs rc=$?"
timeout 300 python bench.py --model higgs_mlp --algo aeasgd --optimizer adagrad --batch 64 > gpurun_out/sweep_higgs_aeasgd_b64.json 2> gpurun_out/sweep_higgs_aeasgd_b64.err; echo "higgs rc=$?"
timeout 300 python bench.py --impl reference --batch 4 --steps 20 --warmup 5 > gpurun_out/sweep_reference_b4.json 2> gpurun_out/sweep_reference_b4.err; echo "ref b4 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | e2e", d.get("e2e") and int(d["e2e"]["value"]), "| k/step", d.get("kernels_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
