#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -k "sharded or spawned or multi" > gpurun_out/t_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_2gpu.log
tail -4 gpurun_out/t_2gpu.log
run() { # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus 2 --steps 20 --warmup 5 $2 > gpurun_out/g2_$1.json 2> gpurun_out/g2_$1.err; echo "rc=$?" >> gpurun_out/g2_$1.err; tail -1 gpurun_out/g2_$1.err
}
run b64 "--batch 64"
run b64_nofuse "--batch 64 --no-fuse-comm --skip-e2e"
run b64_sharded "--batch 64 --sharded-ps --skip-e2e"
run b64_dedicated "--batch 64 --dedicated-ps --skip-e2e"
run b16384 "--batch 16384 --skip-e2e"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | per-rank", d["per_rank_ms_per_step"], "| e2e", d["e2e"] and round(d["e2e"]["ms_per_step"]*1e3, 2), "| x_us", round(d["exchange_us"], 1), "ps_gbs", round(d["ps_gbs"]["push"]), "comm_frac", round(d["comm_fraction"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
