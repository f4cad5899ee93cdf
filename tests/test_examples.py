"""The example workflows (SURVEY appendix C acceptance scenarios) run end to end on CPU at toy sizes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("mnist.py", ["--rows", "1024", "--epochs", "1", "--batch", "32", "--window", "4"], "accuracy="),
    ("mnist.py", ["--rows", "512", "--epochs", "1", "--batch", "32", "--model", "convnet", "--trainer", "DOWNPOUR"],
     "accuracy="),
    ("higgs_workflow.py", ["--rows", "4096"], "DOWNPOUR"),
    ("streaming_inference.py", [], "micro-batch"),
    ("custom_optimizer.py", ["--route", "python", "--rows", "2048"], "ClippedDownpour"),
]


@pytest.mark.parametrize("script,args,expect", CASES, ids=[f"{c[0]}:{i}" for i, c in enumerate(CASES)])
def test_example_runs(script, args, expect):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", DK_BACKEND="thread")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert expect in r.stdout, r.stdout[-2000:]
