"""Functional (multi-input / multi-output) models through the trainers.

The reference trains whatever ``keras.models.Model`` it is handed: list-valued ``features_col`` feed the inputs one
column each (``distkeras/workers.py:65-66``), ``loss`` / ``loss_weights`` go to ``model.compile``
(``workers.py:103-119``) and ``train_on_batch`` returns ``[total, loss_1.., metric_1..]`` which lands in the
history record unchanged (``workers.py:162-178``).
"""
import numpy as np
import pytest
import torch

from distkeras_b200.data import Dataset
from distkeras_b200.models import Add, Concatenate, Dense, Input, Model, model_from_json
from distkeras_b200.predictors import ModelPredictor
from distkeras_b200.trainers import ADAG, AEASGD, DOWNPOUR, SingleTrainer
from distkeras_b200.utils import deserialize_keras_model, serialize_keras_model


def two_tower(seed=0):
    torch.manual_seed(seed)
    ia, ib = Input((6,), name="ia"), Input((4,), name="ib")
    h = Dense(32, activation="relu")(Concatenate()([Dense(16, activation="relu")(ia), Dense(16, activation="relu")(ib)]))
    o1 = Dense(3, activation="softmax", name="o1")(h)
    o2 = Dense(2, activation="softmax", name="o2")(h)
    return Model([ia, ib], [o1, o2]).build()


def two_tower_data(n=2048):
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(n, 6, generator=g), torch.randn(n, 4, generator=g)
    z = torch.cat([a, b], 1)
    w1, w2 = torch.randn(10, 3, generator=g), torch.randn(10, 2, generator=g)
    return Dataset({"fa": a, "fb": b, "y1": (z @ w1).argmax(1).to(torch.int32),
                    "y2": torch.nn.functional.one_hot((z @ w2).argmax(1), 2).float()})


def test_graph_round_trips_through_json_and_the_wire_format():
    m = two_tower()
    m2 = model_from_json(m.to_json())
    assert m2.num_inputs == 2 and m2.num_outputs == 2 and m2.num_params == m.num_params
    m3 = deserialize_keras_model(serialize_keras_model(m))
    x = [torch.randn(5, 6), torch.randn(5, 4)]
    for p, q in zip(m.predict(x), m3.predict(x)):
        assert np.allclose(p, q)
    assert all(np.allclose(p.sum(1), 1.0, atol=1e-5) for p in m.predict(x))


def test_residual_add_and_shared_flat_buffer():
    torch.manual_seed(1)
    i = Input((8,))
    h = Dense(8, activation="relu")(i)
    o = Dense(2, activation="softmax")(Add()([h, i]))
    m = Model(i, o).build()
    flat = m.get_flat_weights().clone()
    m.set_flat_weights(flat * 0)
    assert np.allclose(m.predict(torch.randn(3, 8)), 0.5)
    m.set_flat_weights(flat)
    assert m.count_params() == 8 * 8 + 8 + 8 * 2 + 2


@pytest.mark.parametrize("cls,kw", [(SingleTrainer, {}), (ADAG, dict(num_workers=2, communication_window=4)),
                                    (DOWNPOUR, dict(num_workers=2, communication_window=4)),
                                    (AEASGD, dict(num_workers=2, communication_window=4, rho=1.0, learning_rate=0.05))])
def test_two_head_model_trains_and_history_carries_every_loss_and_metric(cls, kw):
    ds = two_tower_data()
    t = cls(two_tower(), {"class_name": "adam", "config": {"lr": 0.01}},
            ["categorical_crossentropy", "categorical_crossentropy"], features_col=["fa", "fb"],
            label_col=["y1", "y2"], batch_size=32, num_epoch=4, loss_weights=[1.0, 0.5], **kw)
    t.backend = "thread"
    out = t.train(ds)
    hist = t.get_history()
    recs = [r["history"] if isinstance(r, dict) else r for r in (hist[0] if isinstance(hist[0], list) else hist)]
    assert all(len(r) == 5 for r in recs)                      # [loss, loss_1, loss_2, acc_1, acc_2]
    first, last = np.mean(recs[:8], axis=0), np.mean(recs[-8:], axis=0)
    assert np.allclose([r[0] for r in recs], [r[1] + 0.5 * r[2] for r in recs], atol=1e-5)  # loss_weights
    assert last[0] < 0.5 * first[0] and last[3] > 0.8 and last[4] > 0.8
    p1, p2 = out.predict([ds["fa"], ds["fb"]])
    assert (p1.argmax(1) == ds["y1"].numpy()).mean() > 0.85
    assert (p2.argmax(1) == ds["y2"].argmax(1).numpy()).mean() > 0.85


def test_model_predictor_writes_one_column_per_output():
    ds = two_tower_data(256)
    m = two_tower()
    d = ModelPredictor(m, features_col=["fa", "fb"], output_col=["p1", "p2"]).predict(ds)
    assert tuple(d["p1"].shape) == (256, 3) and tuple(d["p2"].shape) == (256, 2)
    d = ModelPredictor(m, features_col=["fa", "fb"]).predict(ds)
    assert "prediction_0" in d.columns and "prediction_1" in d.columns


def test_metrics_argument_selects_what_is_recorded():
    ds = two_tower_data(512)
    t = SingleTrainer(two_tower(), "sgd", ["categorical_crossentropy", "categorical_crossentropy"],
                      features_col=["fa", "fb"], label_col=["y1", "y2"], batch_size=32, metrics=[])
    t.backend = "thread"
    t.train(ds)
    assert all(len(r["history"]) == 3 for r in t.get_history())   # [loss, loss_1, loss_2]


@pytest.mark.gpu
def test_two_head_model_trains_on_the_fabric_backend():
    """No native lowering for a DAG: the replica runs on the autograd executor, the ADAG exchange stays in-kernel."""
    ds = two_tower_data(4096)
    t = ADAG(two_tower(), {"class_name": "adam", "config": {"lr": 0.01}},
             ["categorical_crossentropy", "categorical_crossentropy"], features_col=["fa", "fb"],
             label_col=["y1", "y2"], batch_size=32, num_epoch=3, num_workers=1, communication_window=4,
             loss_weights=[1.0, 0.5])
    t.backend = "fabric"
    with pytest.warns(UserWarning, match="does not lower"):
        out = t.train(ds)
    recs = [r["history"] for r in t.get_history()]
    assert all(len(r) == 5 for r in recs)
    assert np.mean([r[0] for r in recs[-8:]]) < 0.5 * np.mean([r[0] for r in recs[:8]])
    p1, _ = out.predict([ds["fa"], ds["fb"]])
    assert (p1.argmax(1) == ds["y1"].numpy()).mean() > 0.85
