#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
timeout 300 python tools/profile_graph.py --batch 64 --out gpurun_out/graph_b64.txt 2>&1 | grep -v -i warn | tail -12
for B in 64 256; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --batch $B > gpurun_out/d_bench_b$B.json 2> gpurun_out/d_bench_b$B.err; echo "rc=$?" >> gpurun_out/d_bench_b$B.err
  tail -2 gpurun_out/d_bench_b$B.err
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --batch 64 --no-fuse-comm --skip-e2e > gpurun_out/d_bench_b64_nofuse.json 2>> gpurun_out/d_bench_b64.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | e2e", d["e2e"] and round(d["e2e"]["ms_per_step"]*1e3, 2), "| x_us", round(d["exchange_us"], 1), "comm_frac", round(d["comm_fraction"], 4), d["config"]["ps_transport"][:60])
    except Exception as e:
        print(f, "ERR", e)
PY
