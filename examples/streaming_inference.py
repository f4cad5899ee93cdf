#!/usr/bin/env python
"""Streaming inference (the reference's Kafka + Spark-Streaming notebook,
``examples/kafka_spark_high_throughput_ml_pipeline.ipynb`` + ``examples/kafka_producer.py``).

A producer thread emits JSON records (the Kafka topic stand-in: no broker is available offline); the
consumer collects micro-batches, converts them with ``json_to_dataframe_row`` / ``Dataset.from_rows``,
normalises, predicts with ``ModelPredictor`` and filters on the predicted index."""
import json
import os
import queue
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200.data import Dataset, synthetic_higgs
from distkeras_b200.models import higgs_mlp
from distkeras_b200.predictors import ModelPredictor
from distkeras_b200.transformers import LabelIndexTransformer
from distkeras_b200.utils import json_to_dataframe_row

topic: "queue.Queue[str]" = queue.Queue()


def producer(n_batches=5, rows=2000):
    data = synthetic_higgs(rows * n_batches, seed=3)["features"]
    for b in range(n_batches):
        for r in data[b * rows:(b + 1) * rows]:
            topic.put(json.dumps({"features": r.tolist()}))
        time.sleep(0.2)
    topic.put(None)


feeder = threading.Thread(target=producer)
feeder.start()
predictor = ModelPredictor(higgs_mlp(seed=0), features_col="features")
indexer = LabelIndexTransformer(output_dim=2)
batch, done = [], False
while not done:
    deadline = time.time() + 0.25  # micro-batch interval (the notebook uses 10 s)
    while time.time() < deadline:
        try:
            msg = topic.get(timeout=0.05)
        except queue.Empty:
            continue
        if msg is None:
            done = True
            break
        batch.append(json_to_dataframe_row(msg))
    if batch:
        ds = Dataset.from_rows(batch)
        out = indexer.transform(predictor.predict(ds))
        signal = out.filter(lambda d: d["prediction_index"] == 1.0).count()
        print(f"micro-batch: {ds.count()} rows, {signal} predicted signal")
        batch = []
feeder.join()
