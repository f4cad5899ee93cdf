import numpy as np
import torch

from distkeras_b200.data import Dataset
from distkeras_b200.evaluators import AccuracyEvaluator, F1Evaluator
from distkeras_b200.predictors import ModelPredictor
from distkeras_b200.models import mnist_mlp
from distkeras_b200.transformers import (BinaryLabelTransformer, DenseTransformer, LabelIndexTransformer,
                                         MinMaxTransformer, OneHotTransformer, ReshapeTransformer,
                                         StandardTransformer)
from distkeras_b200.utils import (history_executor, history_executors_average, json_to_dataframe_row,
                                  new_dataframe_row, precache, shuffle, to_one_hot_encoded_dense)


def test_minmax_formula():
    x = np.random.RandomState(0).randint(0, 251, (10, 6)).astype(np.float32)
    ds = Dataset({"features": x})
    t = MinMaxTransformer(o_min=0.0, o_max=250.0, n_min=0.0, n_max=1.0, input_col="features",
                          output_col="features_normalized")
    out = t.transform(ds)["features_normalized"].numpy()
    scale = (1.0 - 0.0) / (250.0 - 0.0)
    assert np.allclose(out, scale * (x - 250.0) + 1.0, atol=1e-6)
    s, b = t.affine()
    assert np.allclose(out, s * x + b, atol=1e-6)


def test_standard_binary_onehot_reshape_dense():
    rs = np.random.RandomState(1)
    ds = Dataset({"a": rs.randn(100).astype(np.float32), "label": rs.randint(0, 3, 100), "v": rs.rand(100, 12)})
    out = StandardTransformer(["a"]).transform(ds)["a_normalized"]
    assert abs(float(out.mean())) < 1e-5 and abs(float(out.std(unbiased=False)) - 1) < 1e-4
    b = BinaryLabelTransformer("label", "bin", 1).transform(ds)["bin"].numpy()
    assert np.array_equal(b[:, 0], (ds["label"].numpy() == 1).astype(np.float32))
    assert np.array_equal(b.sum(1), np.ones(100))
    oh = OneHotTransformer(3, "label", "enc").transform(ds)["enc"].numpy()
    assert np.array_equal(oh.argmax(1), ds["label"].numpy()) and oh.shape == (100, 3)
    r = ReshapeTransformer("v", "m", (3, 4, 1)).transform(ds)["m"]
    assert tuple(r.shape) == (100, 3, 4, 1)
    assert torch.equal(DenseTransformer("v", "d").transform(ds)["d"], ds["v"])


def test_label_index_rule():
    t = LabelIndexTransformer(output_dim=3)
    assert t.get_index([0.6, 0.9, 0.0]) == 0          # first element over the 0.55 threshold wins
    assert t.get_index([0.2, 0.5, 0.3]) == 1          # otherwise arg-max
    assert t.get_index([0.0, 0.0, 0.0]) == 0          # otherwise the default index
    assert LabelIndexTransformer(3, default_index=2).get_index([0.0, 0.0, 0.0]) == 2
    ds = Dataset({"prediction": np.array([[0.1, 0.7, 0.2], [0.4, 0.3, 0.3]], dtype=np.float32)})
    assert t.transform(ds)["prediction_index"].tolist() == [1.0, 0.0]


def test_predictor_and_evaluators():
    m = mnist_mlp(seed=0)
    x = torch.rand(50, 784)
    ds = Dataset({"features": x, "label": torch.randint(0, 10, (50,))})
    pred = ModelPredictor(m, features_col="features", output_col="prediction", device="cpu").predict(ds)
    assert tuple(pred["prediction"].shape) == (50, 10)
    assert np.allclose(pred["prediction"].numpy(), m.predict(x), atol=1e-6)
    idx = LabelIndexTransformer(10).transform(pred)
    acc = AccuracyEvaluator(prediction_col="prediction_index", label_col="label").evaluate(idx)
    want = float((idx["prediction_index"].long() == ds["label"]).float().mean())
    assert abs(acc - want) < 1e-6
    f1 = F1Evaluator("l", "p").evaluate(Dataset({"l": np.array([1, 1, 0, 0]), "p": np.array([1, 0, 1, 0])}))
    assert abs(f1 - 0.5) < 1e-9


def test_dataset_surface():
    ds = Dataset({"features": np.arange(40, dtype=np.float32).reshape(20, 2), "label": np.arange(20)})
    assert ds.count() == 20 and ds.columns == ["features", "label"]
    parts = ds.repartition(3).partitions()
    assert sum(len(p) for p in parts) == 20 and len(parts) == 3
    a, b = ds.randomSplit([0.6, 0.4], seed=0)
    assert a.count() + b.count() == 20
    assert ds.unionAll(ds).count() == 40
    assert sorted(shuffle(ds, 0)["label"].tolist()) == list(range(20))
    rows = ds.rdd.mapPartitionsWithIndex(lambda i, it: [(i, len(list(it)))]).collect()
    assert sum(n for _, n in rows) == 20
    assert precache(ds, 4).num_partitions == 4
    assert ds.take(2)[1]["label"] == 1
    assert ds.filter(lambda d: d["label"] % 2 == 0).count() == 10


def test_utils_helpers():
    assert to_one_hot_encoded_dense(2, 4).tolist() == [0, 0, 1, 0]
    r = new_dataframe_row({"a": 1}, "b", 2)
    assert r["a"] == 1 and r.b == 2
    assert json_to_dataframe_row('{"x": [1, 2]}')["x"] == [1, 2]
    hist = [{"history": [1.0, 0.0], "worker_id": 0, "iteration": 1}, {"history": [3.0, 1.0], "worker_id": 1, "iteration": 1},
            {"history": [5.0, 1.0], "worker_id": 1, "iteration": 2}]
    avg = history_executors_average(hist)
    assert np.allclose(avg[0], [2.0, 0.5]) and np.allclose(avg[1], [5.0, 1.0])
    assert [h["iteration"] for h in history_executor(hist, 1)] == [1, 2]
