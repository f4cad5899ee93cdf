#!/usr/bin/env python
"""End-to-end MNIST workflow (the reference's ``examples/mnist.py`` on this framework):

    data -> OneHot -> MinMax -> Reshape -> train (ADAG / DOWNPOUR / AEASGD ...) -> predict
         -> LabelIndexTransformer -> AccuracyEvaluator

Real MNIST CSVs are not available offline; by default a learnable synthetic set of the same shape is
generated.  Pass ``--csv path`` (label column ``label`` + 784 pixel columns) to use real data.

    python examples/mnist.py --trainer ADAG --model mlp --workers 1
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/mnist.py --workers 8
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from distkeras_b200 import trainers
from distkeras_b200.data import Dataset, synthetic_mnist
from distkeras_b200.evaluators import AccuracyEvaluator
from distkeras_b200.models import mnist_convnet, mnist_mlp
from distkeras_b200.predictors import ModelPredictor
from distkeras_b200.transformers import LabelIndexTransformer, MinMaxTransformer, OneHotTransformer, ReshapeTransformer

ap = argparse.ArgumentParser()
ap.add_argument("--csv")
ap.add_argument("--trainer", default="ADAG")
ap.add_argument("--model", default="mlp", choices=["mlp", "convnet"])
ap.add_argument("--workers", type=int, default=1)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--window", type=int, default=5)
ap.add_argument("--rows", type=int, default=60000)
args = ap.parse_args()

raw = Dataset.from_csv(args.csv, label_col="label") if args.csv else synthetic_mnist(args.rows, as_uint8=False)
train, test = raw.randomSplit([0.85, 0.15], seed=0)

# the reference's preprocessing chain (examples/mnist.py:110-135); every step is one tensor op here
encoder = OneHotTransformer(10, input_col="label", output_col="label_encoded")
scaler = MinMaxTransformer(o_min=0.0, o_max=250.0, n_min=0.0, n_max=1.0, input_col="features",
                           output_col="features_normalized")
prep = lambda ds: scaler.transform(encoder.transform(ds))
train, test = prep(train), prep(test)
features = "features_normalized"
if args.model == "convnet":
    reshape = ReshapeTransformer("features_normalized", "matrix", (28, 28, 1))
    train, test = reshape.transform(train), reshape.transform(test)
    features = "matrix"
model = mnist_convnet() if args.model == "convnet" else mnist_mlp()

cls = getattr(trainers, args.trainer)
kw = dict(worker_optimizer="adam", loss="categorical_crossentropy", features_col=features, label_col="label_encoded",
          batch_size=args.batch, num_epoch=args.epochs)
if cls in (trainers.SingleTrainer,):
    trainer = cls(model, **kw)
elif cls in (trainers.AveragingTrainer,):
    trainer = cls(model, num_workers=args.workers, **kw)
else:
    trainer = cls(model, num_workers=args.workers, communication_window=args.window, **kw)
trained = trainer.train(train)

if int(os.environ.get("RANK", "0")) == 0:
    predicted = ModelPredictor(trained, features_col=features).predict(test)
    predicted = LabelIndexTransformer(10).transform(predicted)
    acc = AccuracyEvaluator(prediction_col="prediction_index", label_col="label").evaluate(predicted)
    print(f"trainer={args.trainer} backend={trainer.backend} time={trainer.get_training_time():.2f}s "
          f"accuracy={acc:.4f} history={len(trainer.get_history())}"
          + (f" num_updates={trainer.num_updates()}" if hasattr(trainer, "num_updates") else ""))
