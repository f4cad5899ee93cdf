"""The same-box comparator: ``bench.py --impl reference`` drives the UNMODIFIED reference package (``baseline/_ref``)
through its own public API on stand-in Keras / Spark modules (``baseline/shims``).  Skipped when the offline install is
absent (``baseline/_ref`` is git-ignored)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "distkeras")),
                    reason="reference package not installed under baseline/_ref")
def test_reference_arm_runs_the_unmodified_package():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    rec = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["unit"] == "samples/s"
    assert rec["config"]["trainer"] == "ADAG" and rec["config"]["batch_per_worker"] == 64
    # nothing of this repository is on that path: the package under test is the reference's own
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_trainers", os.path.join(ROOT, "baseline", "_ref", "distkeras", "trainers.py"))
    assert spec is not None
    src = open(os.path.join(ROOT, "baseline", "_ref", "distkeras", "trainers.py")).read()
    assert "distkeras_b200" not in src


def test_shims_live_outside_the_reference_package():
    shims = os.path.join(ROOT, "baseline", "shims")
    assert {"keras", "pyspark", "tensorflow"} <= set(os.listdir(shims))
