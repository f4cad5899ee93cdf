"""Aux subsystems: job deployment, checkpoint/resume, fault injection, gloo multi-process (CPU)."""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

from distkeras_b200.data import Dataset
from distkeras_b200.job_deployment import Job, Punchcard
from distkeras_b200.models import Dense, Sequential
from distkeras_b200.trainers import ADAG, SingleTrainer
from distkeras_b200.utils.checkpoint import load_checkpoint, resume_trainer, save_checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_model(seed=0):
    return Sequential([Dense(16, activation="relu", input_shape=(8,)), Dense(3, activation="softmax")], seed=seed)


def tiny_data(n=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 8, generator=g)
    w = torch.randn(8, 3, generator=g)
    return Dataset({"features": x, "label": (x @ w).argmax(1).to(torch.int32)})


def test_checkpoint_roundtrip_and_resume(tmp_path):
    m = tiny_model(3)
    m.build()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, m, num_updates=7, iteration=42, extra={"note": "x"})
    p = load_checkpoint(path)
    assert p["num_updates"] == 7 and p["iteration"] == 42 and p["extra"]["note"] == "x"
    assert torch.equal(p["model"].get_flat_weights(), m.get_flat_weights())
    t = ADAG(tiny_model(0), "sgd", "categorical_crossentropy", num_workers=1, batch_size=16, communication_window=2)
    t.backend = "thread"
    resume_trainer(t, path)
    assert np.allclose(t.master_model["flat"], m.get_flat_weights().numpy())
    t.checkpoint_path = str(tmp_path / "after.pt")
    out = t.train(tiny_data())
    after = load_checkpoint(t.checkpoint_path)
    assert torch.equal(after["model"].get_flat_weights(), out.get_flat_weights()) and after["num_updates"] > 1


def test_fault_injection_is_survivable(monkeypatch):
    monkeypatch.setenv("DK_FAULT", "1:3")  # worker 1 dies at its 3rd batch
    t = ADAG(tiny_model(0), "sgd", "categorical_crossentropy", num_workers=2, batch_size=16, communication_window=2)
    t.backend = "thread"
    with pytest.raises(RuntimeError, match="injected fault"):
        t.train(tiny_data())  # default: failures surface (the reference prints and swallows them)
    monkeypatch.setenv("DK_FAULT", "1:4")  # a fault fires once per spec: arm a new one for the second run
    t2 = ADAG(tiny_model(0), "sgd", "categorical_crossentropy", num_workers=2, batch_size=16, communication_window=2)
    t2.backend = "thread"
    t2.tolerate_worker_failures = True
    model = t2.train(tiny_data())  # PS never waits on a worker; the shard is retried by a survivor
    assert len(t2.worker_failures) >= 1 and torch.isfinite(model.get_flat_weights()).all()


def test_punchcard_job_roundtrip(tmp_path):
    secrets = tmp_path / "secrets.json"
    secrets.write_text(json.dumps([{"secret": "S3CRET", "identity": "tester"}]))
    data = tmp_path / "data.pt"
    torch.save(tiny_data(128), str(data))
    daemon = Punchcard(secrets_path=str(secrets), port=0, host="127.0.0.1")
    port = daemon.start()
    try:
        trainer = SingleTrainer(tiny_model(0), "adam", "categorical_crossentropy", batch_size=16)
        trainer.backend = "thread"
        bad = Job("WRONG", "j", str(data), 1, 1, trainer)
        with pytest.raises(RuntimeError):
            bad.send(f"http://127.0.0.1:{port}")
        job = Job("S3CRET", "unit-test-job", str(data), 1, 1, trainer)
        job.poll_interval = 0.2
        job.send(f"http://127.0.0.1:{port}")
        job.wait_completion()
        assert job.error is None, job.error
        model = job.get_trained_model()
        assert model is not None and len(job.get_history()) == 8
        assert not torch.equal(model.get_flat_weights(), tiny_model(0).build().get_flat_weights())
    finally:
        daemon.shutdown()


def test_generate_secret_script():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "generate_secret.py"), "--identity", "bob"],
                         capture_output=True, text=True, check=True).stdout
    d = json.loads(out)
    assert d["identity"] == "bob" and len(d["secret"]) == 64


SPMD_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
from distkeras_b200.data import Dataset
from distkeras_b200.models import Dense, Sequential
from distkeras_b200.trainers import ADAG
g = torch.Generator().manual_seed(0)
x = torch.randn(256, 8, generator=g); w = torch.randn(8, 3, generator=g)
ds = Dataset({{"features": x, "label": (x @ w).argmax(1).to(torch.int32)}})
m = Sequential([Dense(16, activation="relu", input_shape=(8,)), Dense(3, activation="softmax")], seed=0)
t = ADAG(m, {{"class_name": "adam", "config": {{"lr": 0.02}}}}, "categorical_crossentropy", num_workers=2, batch_size=16,
         communication_window=2, num_epoch=2)
t.backend = "socket"
model = t.train(ds)
model.compile("categorical_crossentropy")
acc = model.evaluate(ds["features"], ds["label"])[1]
line = " ".join(map(str, ("RANK", os.environ["RANK"], "HIST", len(t.get_history()), "ACC", round(acc, 3), "SUM",
                           float(model.get_flat_weights().sum()))))
os.write(1, (line + "\n").encode())   # one write per rank: the two ranks share the pipe, print() would interleave its pieces
"""


def test_spmd_gloo_two_processes(tmp_path):
    """torchrun-style world_size=2 on CPU: rank 0 hosts the TCP parameter server, both ranks train."""
    script = tmp_path / "spmd.py"
    script.write_text(SPMD_SCRIPT.format(root=ROOT))
    import socket

    for attempt in range(3):   # a rendezvous port picked here can be taken by the time torchrun binds it: retry on a new one
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                           capture_output=True, text=True, timeout=240, env={**os.environ, "DK_BACKEND": "socket"})
        if r.returncode == 0:
            break
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("RANK")]
    assert len(lines) == 2
    sums = {l.split("SUM")[1].strip() for l in lines}
    assert len(sums) == 1  # both ranks return the same final model
    assert all("HIST 32" in l for l in lines)
    assert all(float(l.split("ACC")[1].split()[0]) > 0.5 for l in lines)


def test_local_control_shard_table_and_barrier(tmp_path):
    """Control plane of the spawned fabric ranks: a dead rank's unfinished claims go back to the table and the
    barrier stops waiting for it (the launcher flips ``alive``)."""
    import threading

    from distkeras_b200.parallel.runtime import LocalControl

    c = LocalControl(3, str(tmp_path))
    c.claim(0, 0)
    c.claim(1, 1)
    assert c.try_claim(2, 4) == 2 and c.try_claim(0, 4) == 3 and c.try_claim(0, 4) is None
    c.finish(0)
    c.finish(2)
    assert not c.all_done(4)
    c.alive[1] = 0
    assert c.release_claims_of(1) == [1]                     # rank 1 died holding partition 1
    assert c.try_claim(0, 4) == 1
    for i in (1, 3):
        c.finish(i)
    assert c.all_done(4)
    # barrier: ranks 0 and 2 meet, dead rank 1 is skipped
    t = threading.Thread(target=c.barrier, args=(2,))
    t.start()
    c.barrier(0, timeout=10)
    t.join(timeout=10)
    assert not t.is_alive()
    # object exchange through the scratch directory
    out = {}
    c2 = LocalControl(2, str(tmp_path))
    th = threading.Thread(target=lambda: out.setdefault("v", c2.exchange_obj(1, None, 0, timeout=10)))
    th.start()
    c2._seq = 0
    c3 = LocalControl(2, str(tmp_path))
    c3.exchange_obj(0, {"a": 1}, 0)
    th.join(timeout=10)
    assert out["v"] == {"a": 1}


@pytest.mark.parametrize("static", [True, False])
def test_shard_table_runs_every_partition_exactly_once_with_skewed_ranks_and_a_lost_rank(static, tmp_path):
    """The task loop of the spawned fabric ranks (threads stand in for the processes): ranks reach the start barrier at
    very different times, one rank dies holding a partition.  Every partition must be trained exactly once by a live
    rank -- statically assigned ones are claimed BEFORE the barrier, so a fast rank never steals a slow peer's shard."""
    import random
    import threading
    import time
    from types import SimpleNamespace

    from distkeras_b200.parallel.runtime import LocalControl, drain_shard_table

    world, per_rank = 6, 2
    n_parts = world * per_rank
    parts = [SimpleNamespace(index=i) for i in range(n_parts)]
    control = LocalControl(world, str(tmp_path))
    start = threading.Barrier(world)
    runs, lock = [], threading.Lock()
    dead_rank = 4

    def rank_body(rank):
        rng = random.Random(rank)
        time.sleep(rng.random() * 0.3)                          # set-up skew (graph capture, module load)
        mine = [parts[i] for i in range(rank, n_parts, world)] if static else None
        for p in mine or ():
            control.claim(p.index, rank)
        start.wait()

        def run_task(part):
            if rank == dead_rank and (not static or part.index == mine[-1].index):
                raise SystemExit                                # the process vanishes in the middle of this partition
            time.sleep(0.002 * (1 + rank))
            with lock:
                runs.append((part.index, rank))

        try:
            drain_shard_table(control, rank, parts, mine, run_task)
        except SystemExit:
            time.sleep(0.05)                                    # the launcher notices a little later
            control.alive[rank] = 0
            control.release_claims_of(rank)

    threads = [threading.Thread(target=rank_body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
    assert not any(t.is_alive() for t in threads)
    assert sorted(i for i, _ in runs) == list(range(n_parts))   # exactly once each
    assert control.all_done(n_parts)
    if static:   # nobody but the owner (or, for the orphan, a survivor) touched a statically assigned partition
        for idx, r in runs:
            assert r == idx % world or idx % world == dead_rank
