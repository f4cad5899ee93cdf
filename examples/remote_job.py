#!/usr/bin/env python
"""Remote submission (reference ``README.md:142-161``): start ``scripts/punchcard.py`` on the GPU
box, then submit a trainer with a secret and fetch the trained model.

    python scripts/generate_secret.py --identity me > /tmp/secret.json      # put it in secrets.json as a list
    python scripts/punchcard.py --port 8000 --secrets secrets.json &
    python examples/remote_job.py --address http://127.0.0.1:8000 --secret <secret> --data /path/data.pt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from distkeras_b200.job_deployment import Job
from distkeras_b200.models import mnist_mlp
from distkeras_b200.trainers import ADAG

ap = argparse.ArgumentParser()
ap.add_argument("--address", required=True)
ap.add_argument("--secret", required=True)
ap.add_argument("--data", required=True, help="path on the daemon's host: .pt / .npz / .parquet / .csv")
a = ap.parse_args()
trainer = ADAG(mnist_mlp(), "adam", "categorical_crossentropy", num_workers=8, batch_size=1024, communication_window=12)
job = Job(a.secret, "mnist-adag", a.data, num_executors=8, num_processes=1, trainer=trainer)
job.send(a.address)
job.wait_completion()
print("error:", job.error, "history records:", len(job.get_history() or []))
