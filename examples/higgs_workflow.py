#!/usr/bin/env python
"""The reference's ``examples/workflow.ipynb`` / ``example_1_analysis.ipynb``: ATLAS-Higgs-shaped
data, StandardTransformer-normalised features, a 30-500-500-500-2 MLP, SingleTrainer vs the
asynchronous trainers, accuracy / F1 / training time side by side."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200.data import Dataset, synthetic_higgs
from distkeras_b200.evaluators import AccuracyEvaluator, F1Evaluator
from distkeras_b200.models import higgs_mlp
from distkeras_b200.predictors import ModelPredictor
from distkeras_b200.trainers import AEASGD, DOWNPOUR, EAMSGD, SingleTrainer
from distkeras_b200.transformers import LabelIndexTransformer, OneHotTransformer
from distkeras_b200.utils import shuffle

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200_000)
args = ap.parse_args()

raw = shuffle(synthetic_higgs(args.rows), seed=0)
x = raw["features"]
raw = raw.with_column("features_normalized", (x - x.mean(0)) / x.std(0, unbiased=False))  # StandardScaler step
raw = OneHotTransformer(2, "label", "label_encoded").transform(raw)
train, test = raw.randomSplit([0.7, 0.3], seed=1)
workers = max(1, torch.cuda.device_count()) if torch.cuda.is_available() else 2
common = dict(worker_optimizer="adagrad", loss="categorical_crossentropy", features_col="features_normalized",
              label_col="label_encoded", batch_size=64 if not torch.cuda.is_available() else 1024)
runs = {
    "SingleTrainer": SingleTrainer(higgs_mlp(), **common),
    "AEASGD": AEASGD(higgs_mlp(), num_workers=workers, communication_window=32, rho=5.0, learning_rate=0.1, **common),
    "EAMSGD": EAMSGD(higgs_mlp(), num_workers=workers, communication_window=32, rho=5.0, learning_rate=0.1,
                     momentum=0.9, **common),
    "DOWNPOUR": DOWNPOUR(higgs_mlp(), num_workers=workers, communication_window=5, **common),
}
for name, trainer in runs.items():
    model = trainer.train(train)
    pred = LabelIndexTransformer(2).transform(ModelPredictor(model, "features_normalized").predict(test))
    acc = AccuracyEvaluator("label", "prediction_index").evaluate(pred)
    f1 = F1Evaluator("label", "prediction_index").evaluate(pred)
    print(f"{name:14s} time={trainer.get_training_time():7.2f}s accuracy={acc:.4f} f1={f1:.4f}")
