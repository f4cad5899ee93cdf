#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
port() { echo $((29500 + RANDOM % 400)); }
for i in 1 2 3; do timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_final_${N}gpu_run$i.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_final_${N}gpu_run$i.log; tail -2 gpurun_out/pytest_gpu_final_${N}gpu_run$i.log; done
run() { # gpus, name, extra
  if [ "$1" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 $3 > gpurun_out/final_g$1_$2.json 2> gpurun_out/final_g$1_$2.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $(port) \
       bench.py --gpus $1 --steps 20 --warmup 5 $3 > gpurun_out/final_g$1_$2.json 2> gpurun_out/final_g$1_$2.err
  fi
  echo "rc=$?" >> gpurun_out/final_g$1_$2.err; tail -1 gpurun_out/final_g$1_$2.err
}
run 1 b64 ""
run 2 b64 ""
run 4 b64 ""
run 8 b64 ""
run 8 b16384 "--batch 16384 --skip-e2e"
run 8 higgs_b64 "--model higgs_mlp --algo downpour --optimizer adagrad --skip-e2e"
run 8 cifar_b256 "--model cifar10_cnn --algo downpour --batch 256 --skip-e2e"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $(port) bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/final_g8_reference.json 2> gpurun_out/final_g8_reference.err; echo "ref8 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_g*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("n_gpus"), round(d["ms_per_step"]*1e3, 2), "us/step", int(d["value"]), "samples/s | e2e", d.get("e2e") and int(d["e2e"]["value"]), "| per-rank", d.get("per_rank_ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
