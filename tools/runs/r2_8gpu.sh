#!/bin/bash
# the 8-GPU session of round 2: correctness under contention, PS bandwidth table, scaling rows
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
port() { echo $((29500 + RANDOM % 400)); }
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -q -m gpu -k "contention or sharded or spawned or easgd or average_replicas or averaging" > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_${N}gpu.log
tail -4 gpurun_out/pytest_multi_${N}gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) \
  tools/ps_contention.py --out gpurun_out/ps_contention_${N}gpu_$((N-1))writers.json > gpurun_out/ps_contention_${N}gpu.log 2>&1; echo "contention rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) \
  tools/ps_contention.py --server-writes --out gpurun_out/ps_contention_${N}gpu_${N}writers.json >> gpurun_out/ps_contention_${N}gpu.log 2>&1; echo "contention(all write) rc=$?"
grep -E '"ok"|torn' gpurun_out/ps_contention_${N}gpu_*writers.json | head -40
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $(port) \
  tools/bench_ps.py > gpurun_out/bench_ps_${N}gpu.log 2>&1; echo "bench_ps rc=$?"
run() { # gpus, name, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $(port) \
     bench.py --gpus $1 --steps 20 --warmup 5 $3 > gpurun_out/g$1_$2.json 2> gpurun_out/g$1_$2.err; echo "rc=$?" >> gpurun_out/g$1_$2.err; tail -1 gpurun_out/g$1_$2.err
}
run $N b64 ""
run $N b64_sharded "--sharded-ps --skip-e2e"
run $N b64_dedicated "--dedicated-ps --skip-e2e"
run $N b64_nofuse "--no-fuse-comm --skip-e2e"
run $N b64_dynsgd "--algo dynsgd --skip-e2e"
run $N b256 "--batch 256 --skip-e2e"
run $N b16384 "--batch 16384 --skip-e2e"
run 4 b64 "--skip-e2e"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g[48]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | per-rank", d["per_rank_ms_per_step"], "| e2e", d["e2e"] and round(d["e2e"]["ms_per_step"]*1e3, 2), "| x_us", round(d["exchange_us"], 1), "ps_gbs", round(d["ps_gbs"]["push"]), "comm_frac", round(d["comm_fraction"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
