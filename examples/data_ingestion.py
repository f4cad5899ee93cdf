#!/usr/bin/env python
"""Data ingestion helpers corresponding to the reference's preprocessing notebooks:
CIFAR-10 pickle -> one-hot + min-max -> Parquet (``cifar-10-preprocessing.ipynb``), ``.npy`` shards ->
Dataset (``distributed_numpy_parsing.ipynb``), dataset enlargement by ``unionAll``
(``mnist_preprocessing.ipynb:405-407``)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np

from distkeras_b200.data import Dataset, synthetic_cifar10
from distkeras_b200.transformers import MinMaxTransformer, OneHotTransformer

tmp = tempfile.mkdtemp(prefix="dk_ingest_")
cifar = synthetic_cifar10(2000, as_uint8=False)
flat = cifar.with_column("features", cifar["features"].reshape(2000, -1))
flat = OneHotTransformer(10, "label", "label_encoded").transform(flat)
flat = MinMaxTransformer(0.0, 255.0, 0.0, 1.0, "features", "features_normalized").transform(flat)
flat.select("features_normalized", "label_encoded").to_parquet(os.path.join(tmp, "cifar.parquet"))
back = Dataset.from_parquet(os.path.join(tmp, "cifar.parquet"))
print("parquet round trip:", back)

shards = []
for i in range(4):
    p = os.path.join(tmp, f"shard{i}.npy")
    np.save(p, np.random.rand(500, 16).astype(np.float32))
    shards.append(p)
print("npy shards:", Dataset.from_numpy_shards(shards))

big = flat
for _ in range(3):
    big = big.unionAll(flat)
print("enlarged x4:", big.count(), "rows")
