#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log
tail -3 gpurun_out/t_engine.log
timeout 300 python tools/profile_graph.py --model cifar10_cnn --batch 256 --steps 12 --replays 4 --out gpurun_out/graph_cifar_tma.txt 2>&1 | grep -v -i warn | tail -32
for m in cifar10_cnn mnist_convnet; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --model $m --batch 256 --skip-e2e --reps 12 > gpurun_out/cv_tma_$m.json 2> gpurun_out/cv_tma_$m.err; echo "rc=$?" >> gpurun_out/cv_tma_$m.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/cv_tma_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3, 1), "us/step", int(d["value"]), "samples/s", d["kernels_per_step"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-600:])
PY
