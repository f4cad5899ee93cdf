#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_final_1gpu.log
tail -5 gpurun_out/pytest_gpu_final_1gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_reference_1gpu.json 2> gpurun_out/final_reference_1gpu.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("final_bench_1gpu", "final_reference_1gpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | e2e", d.get("e2e") and int(d["e2e"]["value"]), "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
