"""Keras optimizer descriptions with Keras-2 default hyper-parameters, realised by ``torch.optim``."""


class Optimizer(object):
    defaults = {}

    def __init__(self, **kwargs):
        cfg = dict(self.defaults)
        for k, v in kwargs.items():
            if k == "learning_rate":
                k = "lr"
            cfg[k] = v
        self.config = cfg

    def get_config(self):
        return dict(self.config)

    @classmethod
    def from_config(cls, cfg):
        return cls(**cfg)

    def build(self, params):
        raise NotImplementedError


class SGD(Optimizer):
    defaults = {"lr": 0.01, "momentum": 0.0, "decay": 0.0, "nesterov": False}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.SGD(params, lr=c["lr"], momentum=c["momentum"], nesterov=bool(c["nesterov"]) and c["momentum"] > 0)


class Adam(Optimizer):
    defaults = {"lr": 0.001, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-7, "decay": 0.0}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.Adam(params, lr=c["lr"], betas=(c["beta_1"], c["beta_2"]), eps=c["epsilon"] or 1e-7)


class Adamax(Optimizer):
    defaults = {"lr": 0.002, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-7, "decay": 0.0}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.Adamax(params, lr=c["lr"], betas=(c["beta_1"], c["beta_2"]), eps=c["epsilon"] or 1e-7)


class Nadam(Optimizer):
    defaults = {"lr": 0.002, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-7, "schedule_decay": 0.004}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.NAdam(params, lr=c["lr"], betas=(c["beta_1"], c["beta_2"]), eps=c["epsilon"] or 1e-7,
                                 momentum_decay=c["schedule_decay"])


class Adagrad(Optimizer):
    defaults = {"lr": 0.01, "epsilon": 1e-7, "decay": 0.0}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.Adagrad(params, lr=c["lr"], eps=c["epsilon"] or 1e-7)


class RMSprop(Optimizer):
    defaults = {"lr": 0.001, "rho": 0.9, "epsilon": 1e-7, "decay": 0.0}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.RMSprop(params, lr=c["lr"], alpha=c["rho"], eps=c["epsilon"] or 1e-7)


class Adadelta(Optimizer):
    defaults = {"lr": 1.0, "rho": 0.95, "epsilon": 1e-7, "decay": 0.0}

    def build(self, params):
        import torch

        c = self.config
        return torch.optim.Adadelta(params, lr=c["lr"], rho=c["rho"], eps=c["epsilon"] or 1e-7)


_BY_NAME = {c.__name__.lower(): c for c in (SGD, Adam, Adamax, Nadam, Adagrad, RMSprop, Adadelta)}


def serialize(optimizer):
    return {"class_name": optimizer.__class__.__name__, "config": optimizer.get_config()}


def deserialize(config, custom_objects=None):
    if isinstance(config, Optimizer):
        return config
    if isinstance(config, str):
        config = {"class_name": config, "config": {}}
    cls = _BY_NAME.get(str(config["class_name"]).lower())
    if cls is None:
        raise ValueError("Unknown optimizer: %r" % (config["class_name"],))
    return cls.from_config(config.get("config") or {})


def get(identifier):
    return deserialize(identifier)
