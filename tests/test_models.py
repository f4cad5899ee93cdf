import numpy as np
import pytest
import torch

from distkeras_b200.models import (Dense, Sequential, cifar10_cnn, higgs_mlp,
                                   mnist_convnet, mnist_mlp, model_from_json, resnet18)
from distkeras_b200.utils import deserialize_keras_model, serialize_keras_model, uniform_weights


def test_reference_parameter_counts():
    # examples/mnist_analysis.ipynb:287, examples/mnist.py:150-162, examples/example_1_analysis.ipynb:346
    assert mnist_mlp().count_params() == 987210
    assert mnist_convnet().count_params() == 1048853
    assert higgs_mlp().count_params() == 517502
    assert cifar10_cnn().count_params() == 1250858


def test_json_and_weight_roundtrip():
    m = mnist_convnet(seed=1)
    m.build()
    m2 = model_from_json(m.to_json())
    m2.set_weights(m.get_weights())
    assert torch.equal(m.get_flat_weights(), m2.get_flat_weights())
    # Keras layouts: Dense kernel [in, out], Conv kernel [kh, kw, cin, cout]
    shapes = [w.shape for w in m.get_weights()]
    assert shapes[0] == (3, 3, 1, 32) and shapes[4] == (4608, 225)
    d = serialize_keras_model(m)
    m3 = deserialize_keras_model(d)
    x = torch.rand(4, 28, 28, 1)
    assert torch.allclose(m.forward(x), m3.forward(x))


def test_forward_matches_manual_dense():
    m = Sequential([Dense(5, activation="relu", input_shape=(3,)), Dense(2, activation="softmax")], seed=0)
    m.build()
    w = m.get_weights()
    x = np.random.RandomState(0).rand(7, 3).astype(np.float32)
    h = np.maximum(x @ w[0] + w[1], 0)
    z = h @ w[2] + w[3]
    p = np.exp(z - z.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    assert np.allclose(m.predict(x), p, atol=1e-5)


def test_train_on_batch_decreases_loss():
    torch.manual_seed(0)
    m = mnist_mlp(seed=0, dropout=False)
    m.compile("categorical_crossentropy", "adam")
    x = torch.rand(64, 784)
    y = torch.randint(0, 10, (64,))
    first = m.train_on_batch(x, y)
    for _ in range(30):
        last = m.train_on_batch(x, y)
    assert last[0] < first[0] * 0.5 and last[1] > first[1]


def test_one_hot_and_index_labels_agree():
    m = higgs_mlp(seed=3, dropout=False)
    m.compile("categorical_crossentropy", "sgd")
    x = torch.randn(32, 30)
    y = torch.randint(0, 2, (32,))
    a = m.copy()
    la = a.train_on_batch(x, y)
    lb = m.train_on_batch(x, torch.nn.functional.one_hot(y, 2).float())
    assert abs(la[0] - lb[0]) < 1e-5
    assert torch.allclose(a.get_flat_weights(), m.get_flat_weights(), atol=1e-6)


def test_resnet18_builds_and_steps():
    m = resnet18(input_shape=(32, 32, 3), classes=10, seed=0)
    assert m.count_params() > 11_000_000
    m.compile("categorical_crossentropy", "sgd")
    x = torch.rand(4, 32, 32, 3)
    y = torch.randint(0, 10, (4,))
    before = m.get_flat_weights().clone()
    loss, acc = m.train_on_batch(x, y)
    assert np.isfinite(loss)
    assert not torch.equal(before, m.get_flat_weights())


def test_uniform_weights():
    m = mnist_mlp(seed=0)
    uniform_weights(m, (-0.5, 0.5))
    f = m.get_flat_weights()
    assert float(f.min()) >= -0.5 and float(f.max()) <= 0.5 and float(f.std()) > 0.2


def test_native_planner_grouping_cpu():
    """The native planner's block grouping is pure Python: check it without a GPU."""
    from distkeras_b200.parallel.engine import UnsupportedByNativeEngine, _group_layers

    assert [b.kind for b in _group_layers(mnist_mlp())] == ["dense"] * 3
    assert [b.drop_p for b in _group_layers(mnist_mlp())] == [0.2, 0.2, 0.0]
    kinds = [b.kind for b in _group_layers(resnet18((32, 32, 3), 10))]
    assert kinds == ["conv", "bn", "pool"] + ["res"] * 8 + ["gap", "dense"]
    res = [b for b in _group_layers(resnet18((32, 32, 3), 10)) if b.kind == "res"]
    assert [r.proj is not None for r in res] == [False, False, True, False, True, False, True, False]
    with pytest.raises(UnsupportedByNativeEngine):
        _group_layers(Sequential([Dense(8, activation="tanh", input_shape=(4,)), Dense(2, activation="softmax")]))


def test_fit_save_and_load_roundtrip(tmp_path):
    """Keras-style conveniences: ``fit`` (loss goes down), ``save`` / ``load_model``, ``save_weights`` / ``load_weights``."""
    from distkeras_b200.models import load_model

    g = torch.Generator().manual_seed(0)
    x = torch.randn(512, 8, generator=g)
    y = (x @ torch.randn(8, 3, generator=g)).argmax(1)
    m = Sequential([Dense(16, activation="relu", input_shape=(8,)), Dense(3, activation="softmax")], seed=0)
    m.compile("categorical_crossentropy", {"class_name": "adam", "config": {"lr": 0.02}})
    h = m.fit(x, y, batch_size=32, epochs=4, seed=0)
    assert len(h["loss"]) == 4 and h["loss"][-1] < 0.7 * h["loss"][0] and h["accuracy"][-1] > h["accuracy"][0]
    path = str(tmp_path / "model.dk")
    m.save(path)
    m2 = load_model(path)
    assert m2.to_json() == m.to_json() and torch.equal(m2.get_flat_weights(), m.get_flat_weights())
    assert np.allclose(m2.predict(x[:16]), m.predict(x[:16]))
    assert m2.evaluate(x, y) == m.evaluate(x, y)
    wpath = str(tmp_path / "w.pt")
    m.save_weights(wpath)
    m3 = Sequential([Dense(16, activation="relu", input_shape=(8,)), Dense(3, activation="softmax")], seed=5)
    m3.build()
    m3.load_weights(wpath)
    assert torch.equal(m3.get_flat_weights(), m.get_flat_weights())
    with pytest.raises(ValueError):
        Sequential([Dense(4, input_shape=(8,))], seed=0).load_weights(wpath)


@pytest.mark.parametrize("loss", ["mae", "mean_absolute_percentage_error", "msle", "hinge", "squared_hinge", "kld",
                                  "poisson", "cosine_proximity", "binary_crossentropy", "mean_squared_error"])
def test_keras_objectives_are_differentiable_and_train(loss):
    """Every Keras-1 objective name is accepted and a few steps reduce it (autograd executor)."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(128, 6, generator=g)
    target = torch.softmax(x @ torch.randn(6, 4, generator=g), dim=1)  # positive, rows sum to 1
    if loss in ("hinge", "squared_hinge"):
        target = torch.where(target > 0.25, torch.ones_like(target), -torch.ones_like(target))
    head = "tanh" if loss in ("hinge", "squared_hinge") else "softmax"
    m = Sequential([Dense(16, activation="relu", input_shape=(6,)), Dense(4, activation=head)], seed=0)
    m.compile(loss, {"class_name": "adam", "config": {"lr": 0.02}})
    first = m.train_on_batch(x, target)[0]
    for _ in range(40):
        last = m.train_on_batch(x, target)[0]
    assert np.isfinite(first) and np.isfinite(last) and last < first
