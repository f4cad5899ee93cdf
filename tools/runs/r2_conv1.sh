#!/bin/bash
mkdir -p gpurun_out
DK_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "implicit" > gpurun_out/t_conv.log 2>&1; echo "rc=$?" >> gpurun_out/t_conv.log
tail -6 gpurun_out/t_conv.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --model cifar10_cnn --batch 256 --skip-e2e --reps 12 > gpurun_out/cv_$name.json 2> gpurun_out/cv_$name.err; echo "rc=$?" >> gpurun_out/cv_$name.err; tail -1 gpurun_out/cv_$name.err
}
run explicit DK_IMPLICIT_CONV=0
run implicit DK_IMPLICIT_CONV=1
run implicit_ldgsts DK_IMPLICIT_CONV=1 DK_CONV_LDGSTS=1
run implicit_ldgsts_wgrad DK_IMPLICIT_CONV=1 DK_CONV_LDGSTS=1 DK_IMPLICIT_WGRAD=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/cv_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3, 1), "us/step", int(d["value"]), "samples/s", d["kernels_per_step"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-400:])
PY
DK_IMPLICIT_CONV=1 DK_CONV_LDGSTS=1 DK_IMPLICIT_WGRAD=1 timeout 300 python tools/profile_graph.py --model cifar10_cnn --batch 256 --steps 12 --replays 4 --out gpurun_out/graph_cifar_implicit.txt 2>&1 | grep -v -i warn | tail -30
DK_IMPLICIT_CONV=0 timeout 300 python tools/profile_graph.py --model cifar10_cnn --batch 256 --steps 12 --replays 4 --out gpurun_out/graph_cifar_explicit.txt 2>&1 | grep -v -i warn | tail -30
