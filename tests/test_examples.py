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
    ("kafka_producer.py", ["--sink", "stdout", "--bursts", "2", "--rows", "3", "--interval", "0"], '"features"'),
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


NOTEBOOKS = [
    ("mnist.ipynb", {"synthetic_mnist(60000": "synthetic_mnist(1024", "num_epoch=2": "num_epoch=1"}),
    ("workflow.ipynb", {"synthetic_higgs(200000)": "synthetic_higgs(4096)", "num_epoch=2": "num_epoch=1"}),
    ("data_preparation.ipynb", {"synthetic_cifar10(2000": "synthetic_cifar10(200", "reshape(2000, -1)": "reshape(200, -1)"}),
    ("streaming_inference.ipynb", {"rows=1000": "rows=200"}),
    ("mnist_preprocessing.ipynb", {"synthetic_mnist(4000": "synthetic_mnist(300"}),
    ("mnist_analysis.ipynb", {"synthetic_mnist(60000": "synthetic_mnist(2048", "batch_size=4,": "batch_size=16,"}),
    ("example_0_data_preprocessing.ipynb", {"'--rows', '20000'": "'--rows', '1000'"}),
    ("example_1_analysis.ipynb", {"synthetic_higgs(200000)": "synthetic_higgs(6000)"}),
    ("cifar-10-preprocessing.ipynb", {"synthetic_cifar10(1000": "synthetic_cifar10(100", "synthetic_cifar10(500": "synthetic_cifar10(50"}),
    ("distributed_numpy_parsing.ipynb", {"range(8)": "range(4)"}),
    ("kafka_spark_high_throughput_ml_pipeline.ipynb", {"synthetic_higgs(20000)": "synthetic_higgs(2048)", "'--rows', '1000'": "'--rows', '200'"}),
]


@pytest.mark.parametrize("name,subst", NOTEBOOKS, ids=[n for n, _ in NOTEBOOKS])
def test_notebook_code_cells_run(name, subst, tmp_path):
    """The notebooks are shipped unexecuted; their code cells must run top to bottom (toy sizes, CPU)."""
    import json

    nb = json.load(open(os.path.join(ROOT, "examples", name)))
    src = "\n".join("".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code")
    for a, b in subst.items():
        assert a in src, f"{name}: substitution target {a!r} not found"
        src = src.replace(a, b)
    script = tmp_path / "nb.py"
    script.write_text(src)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", DK_BACKEND="thread")
    r = subprocess.run([sys.executable, str(script)], cwd=os.path.join(ROOT, "examples"), env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
