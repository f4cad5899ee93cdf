"""Multi-process / multi-GPU runs of the NVLink parameter server (launched through torchrun).

With >= 2 GPUs every rank gets its own device and the traffic crosses NVLink; on a one-GPU box the same program runs
with three processes sharing the device (CUDA IPC works within a device too), which still exercises the cross-process
atomics, the ticket lock and the sharded center.  Reference behaviour under test: concurrent commits to one
parameter server (``distkeras/parameter_servers.py:266-292``).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script, *args, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, script), *args]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_parameter_server_kernels_under_contention(tmp_path):
    ndev = torch.cuda.device_count()
    nproc = min(ndev, 8) if ndev >= 2 else 3
    out = tmp_path / "contention.json"
    r = _torchrun(nproc, "tools/ps_contention.py", "--numel", "400000", "--rounds", "25", "--server-writes",
                  "--out", str(out))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.loads(out.read_text())
    assert rep["ok"] and rep["writers"] == nproc
    assert set(rep["checks"]) == {"hogwild_commit_red_add", "hogwild_exchange_atom_add", "strict_lock_commit_pull",
                                  "dynsgd_tickets", "elastic_conservation", "sharded_center"}
    assert rep["checks"]["strict_lock_commit_pull"]["torn_snapshots"] == 0
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, f"ps_contention_{ndev}gpu_{nproc}proc.json"), "w") as f:
            f.write(out.read_text())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_runs_on_every_gpu_under_torchrun():
    """The driver's scaling contract: one JSON line from rank 0, value aggregated over the ranks."""
    n = min(torch.cuda.device_count(), 8)
    r = _torchrun(n, "bench.py", "--gpus", str(n), "--steps", "8", "--warmup", "3", "--skip-e2e", timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["n_gpus"] == n and rec["value"] > 0 and rec["gpu_launches"] > 0
    assert len(rec["per_rank_ms_per_step"]) == n
