"""The reference's extension story (docs/optimizers.md:80-94, examples/workflow.ipynb:112): a user adds
a distributed optimizer with TWO classes -- a trainer that overrides ``allocate_worker`` /
``allocate_parameter_server`` and a worker with ``optimize()`` -- plus, on the fabric backend, a custom
device-side exchange rule."""
import numpy as np
import pytest
import torch

from distkeras_b200.data import Dataset
from distkeras_b200.models import Dense, Sequential
from distkeras_b200.parameter_servers import DeltaParameterServer
from distkeras_b200.trainers import AsynchronousDistributedTrainer
from distkeras_b200.workers import NetworkWorker


def _model(seed=0, in_dim=8, hidden=16, classes=3):
    return Sequential([Dense(hidden, activation="relu", input_shape=(in_dim,)), Dense(classes, activation="softmax")],
                      seed=seed)


def _data(n=1024, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 8, generator=g)
    w = torch.randn(8, 3, generator=g)
    return Dataset({"features": x, "label": (x @ w).argmax(1).to(torch.int32)})


class HalvingServer(DeltaParameterServer):
    """Applies only half of every committed delta (a damped DOWNPOUR)."""

    def apply_commit(self, data):
        data = dict(data)
        data["delta"] = 0.5 * np.asarray(data["delta"])
        super().apply_commit(data)


class EveryOtherWorker(NetworkWorker):
    """Commits the accumulated delta every 2 mini-batches."""

    def optimize(self):
        w1 = self._W().clone()
        while True:
            self._train_batch()
            if self.iteration % 2 == 0:
                self.commit(self._W() - w1)
                self.pull()
                self.set_weights_from_center()
                w1 = self._W().clone()
            self.iteration += 1


class HalvedDownpour(AsynchronousDistributedTrainer):
    def allocate_worker(self):
        return EveryOtherWorker(self.master_model, self.worker_optimizer, self.loss, self.loss_weights,
                                metrics=self.metrics, features_col=self.features_column, label_col=self.label_column,
                                batch_size=self.batch_size, num_epoch=self.num_epoch, master_host=self.master_host,
                                master_port=self.master_port)

    def allocate_parameter_server(self):
        from distkeras_b200.utils import deserialize_keras_model

        return HalvingServer(deserialize_keras_model(self.master_model), self.master_port)


@pytest.mark.parametrize("backend", ["thread", "socket"])
def test_two_class_extension_trains(backend):
    ds = _data()
    t = HalvedDownpour(_model(0), {"class_name": "adam", "config": {"lr": 0.02}}, "categorical_crossentropy",
                       num_workers=2, batch_size=16, num_epoch=3, master_port=0)
    t.backend = backend
    model = t.train(ds)
    model.compile("categorical_crossentropy")
    assert model.evaluate(ds["features"], ds["label"])[1] > 0.55
    # 2 workers x 3 epochs x 32 batches, one commit every 2 batches
    assert t.num_updates() == 1 + 2 * 3 * 32 // 2
    assert isinstance(t.parameter_server, HalvingServer)


def test_extension_without_algorithm_falls_back_from_fabric():
    """A subclass with no device program still trains when the default backend is 'fabric'."""
    ds = _data(256)
    t = HalvedDownpour(_model(0), "sgd", "categorical_crossentropy", num_workers=1, batch_size=16, master_port=0)
    t.backend = "fabric"
    with pytest.warns(UserWarning, match="falling back"):
        t.train(ds)
    assert t.num_updates() == 1 + 16 // 2


# ---------------------------------------------------------------------------------------------------
# fabric-level hook: the exchange rule runs on the worker's stream against the center in the PS's HBM
# ---------------------------------------------------------------------------------------------------
def sign_exchange(ctx):
    """signSGD-style commit: push sign(delta) * mean|delta|, then pull."""
    delta = ctx.W - ctx.W1
    ctx.add_to_center(torch.sign(delta) * delta.abs().mean(), alpha=1.0 / ctx.window)
    ctx.pull()


def plain_exchange(ctx):
    ctx.commit_delta(1.0 / ctx.window)
    ctx.pull()


class CustomFabric(AsynchronousDistributedTrainer):
    def __init__(self, *a, rule=None, communication_window=4, **kw):
        super().__init__(*a, **kw)
        self.rule, self.communication_window = rule, communication_window

    def algorithm(self):
        return {"kind": "custom", "window": self.communication_window, "exchange": self.rule}


def _u8_data(n=4096, seed=0):
    g = torch.Generator().manual_seed(seed)
    proto = torch.randint(0, 200, (10, 64), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    x = (proto[y] + torch.randint(0, 56, (n, 64), generator=g)).clamp(0, 255).to(torch.uint8)
    return Dataset({"features": x, "label": y.to(torch.int32)}), x, y


@pytest.mark.gpu
def test_custom_fabric_exchange_learns():
    ds, x, y = _u8_data()
    t = CustomFabric(_model(0, 64, 128, 10), {"class_name": "adam", "config": {"lr": 0.003}},
                     "categorical_crossentropy", num_workers=1, batch_size=128, num_epoch=2, rule=sign_exchange)
    t.backend = "fabric"
    model = t.train(ds)
    model.compile("categorical_crossentropy")
    assert model.evaluate(x.float() / 255.0, y)[1] > 0.8
    assert t.num_updates() == 1 + 2 * (4096 // 128) // 4


@pytest.mark.gpu
def test_custom_fabric_rule_reproduces_adag():
    """commit_delta(1/tau) + pull IS the ADAG rule: same final center as the built-in trainer."""
    from distkeras_b200.trainers import ADAG

    ds, _, _ = _u8_data(2048)
    kw = dict(num_workers=1, batch_size=128, num_epoch=1, communication_window=4)
    a = CustomFabric(_model(0, 64, 128, 10), "sgd", "categorical_crossentropy", rule=plain_exchange, **kw)
    b = ADAG(_model(0, 64, 128, 10), "sgd", "categorical_crossentropy", **kw)
    a.backend = b.backend = "fabric"
    wa, wb = a.train(ds).get_flat_weights(), b.train(ds).get_flat_weights()
    assert torch.allclose(wa, wb, atol=2e-2, rtol=0), float((wa - wb).abs().max())
