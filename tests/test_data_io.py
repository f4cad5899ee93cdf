"""Dataset readers / writers and the remaining utils helpers (SURVEY 2.1 `utils`, 5.9 data pipeline)."""
import numpy as np
import torch

from distkeras_b200.data import Dataset, synthetic_cifar10, synthetic_higgs, synthetic_mnist
from distkeras_b200.models import mnist_mlp
from distkeras_b200.utils import (deserialize_keras_model, get_os_username, pickle_object, serialize_keras_model,
                                  set_keras_base_directory, uniform_weights, unpickle_object)


def test_csv_reader_assembles_feature_vector(tmp_path):
    p = tmp_path / "d.csv"
    p.write_text("label,a,b,c\n1,0.5,1.5,2.5\n0,3,4,5\n")
    ds = Dataset.from_csv(str(p), label_col="label")
    assert ds.columns == ["label", "features"] and ds.count() == 2
    assert ds["features"].tolist() == [[0.5, 1.5, 2.5], [3.0, 4.0, 5.0]] and ds["label"].tolist() == [1.0, 0.0]
    raw = Dataset.from_csv(str(p), header=True)  # no label column: everything is a feature
    assert tuple(raw["features"].shape) == (2, 4)


def test_parquet_roundtrip_and_npy_shards(tmp_path):
    ds = Dataset({"features": np.random.rand(10, 6).astype(np.float32), "label": np.arange(10)})
    path = str(tmp_path / "d.parquet")
    ds.to_parquet(path)
    back = Dataset.from_parquet(path, num_partitions=2)
    assert back.num_partitions == 2 and torch.allclose(back["features"], ds["features"])
    assert back["label"].tolist() == list(range(10))
    shards = []
    for i in range(3):
        f = tmp_path / f"s{i}.npy"
        np.save(f, np.full((4, 2), i, dtype=np.float32))
        shards.append(str(f))
    sh = Dataset.from_numpy_shards(shards)
    assert sh.count() == 12 and sh.num_partitions == 3 and sh["features"][8:].unique().tolist() == [2.0]


def test_rows_partitions_and_batches():
    ds = Dataset.from_rows([{"features": [float(i), 0.0], "label": i} for i in range(10)], num_partitions=3)
    parts = ds.partitions()
    assert [len(p) for p in parts] == [4, 3, 3] or sum(len(p) for p in parts) == 10
    first = next(iter(parts[0]))
    assert first["label"] == 0 and first.features.tolist() == [0.0, 0.0]
    batches = list(parts[0].batches(["features", "label"], 2))
    assert all(b[0].shape[0] == 2 for b in batches) and len(batches) == len(parts[0]) // 2
    tail = list(parts[0].batches(["features", "label"], 3, drop_last=False))
    assert sum(b[0].shape[0] for b in tail) == len(parts[0])
    assert ds.limit(3).count() == 3 and ds.select("label").columns == ["label"]
    assert ds.withColumnRenamed("label", "y").columns == ["features", "y"]
    assert ds.drop("label").columns == ["features"]
    try:
        Dataset({"a": np.zeros(3), "b": np.zeros(4)})
        raise AssertionError("length mismatch must be rejected")
    except ValueError:
        pass


def test_synthetic_generators_have_reference_shapes():
    m, c, h = synthetic_mnist(32), synthetic_cifar10(16), synthetic_higgs(64)
    assert m["features"].dtype == torch.uint8 and tuple(m["features"].shape) == (32, 784)
    assert tuple(c["features"].shape)[0] == 16 and int(np.prod(c["features"].shape[1:])) == 3072
    assert tuple(h["features"].shape) == (64, 30) and set(h["label"].unique().tolist()) <= {0, 1}
    assert tuple(synthetic_mnist(4, flat=False)["features"].shape) == (4, 28, 28, 1)
    assert torch.equal(synthetic_mnist(8, seed=3)["features"], synthetic_mnist(8, seed=3)["features"])


def test_model_serialisation_and_pickle_roundtrip(tmp_path):
    m = mnist_mlp(seed=1)
    m.compile("categorical_crossentropy", {"class_name": "adam", "config": {"lr": 0.01}})
    d = serialize_keras_model(m)
    assert set(d) >= {"model", "weights"}
    m2 = deserialize_keras_model(unpickle_object(pickle_object(d)))
    assert torch.equal(m2.get_flat_weights(), m.get_flat_weights()) and m2.to_json() == m.to_json()
    uniform_weights(m2, (-0.25, 0.25))
    w = m2.get_flat_weights()
    assert float(w.min()) >= -0.25 and float(w.max()) <= 0.25 and not torch.equal(w, m.get_flat_weights())
    assert isinstance(get_os_username(), str) and len(get_os_username()) > 0
    base = set_keras_base_directory(str(tmp_path / "keras_home"))
    assert str(tmp_path) in base


def test_csv_with_string_label_and_dropped_columns(tmp_path):
    """ATLAS-Higgs layout: bookkeeping columns dropped, 's' / 'b' label indexed (the StringIndexer step)."""
    p = tmp_path / "h.csv"
    p.write_text("EventId,a,b,Weight,Label\n1,0.5,1.5,1.0,s\n2,3,4,2.0,b\n3,5,6,1.0,s\n")
    ds = Dataset.from_csv(str(p), label_col="Label", drop_cols=("EventId", "Weight"), label_map={"b": 0, "s": 1})
    assert ds["Label"].tolist() == [1, 0, 1] and ds["features"].tolist() == [[0.5, 1.5], [3.0, 4.0], [5.0, 6.0]]
