"""Model layer: a Keras-shaped ``Sequential`` whose state is ONE flat fp32 buffer.

PyTorch replaces Keras at the framework layer.  The reference ships models around as
``{'model': model.to_json(), 'weights': model.get_weights()}`` (``distkeras/utils.py:80-86``) and
does all parameter-server algebra on the ragged list ``get_weights()`` returns
(``distkeras/workers.py:329-341``).  Here a model is

* a JSON-serialisable *spec* (list of layer configs -- the ``to_json`` analogue), and
* a single flat, contiguous fp32 parameter tensor with per-layer views (the ``get_weights``
  analogue); every commit / pull / optimizer kernel works on that flat buffer.

Two executors consume the same spec: the autograd path in this file (CPU oracle, any device) and
the native sm_100a engine in ``distkeras_b200.parallel.engine``.

Layouts (shared by both executors): activations are channels-last (NHWC); ``Dense`` kernels are
stored ``[out, in]``, ``Conv2D`` kernels ``[cout, kh, kw, cin]``.  ``get_weights`` / ``set_weights``
convert to / from the Keras conventions (``[in, out]`` and ``[kh, kw, cin, cout]``).
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# layer specs
# --------------------------------------------------------------------------------------------


class Layer:
    """Base layer spec: pure configuration, no tensors."""

    class_name = "Layer"

    def __init__(self, input_shape: Optional[Sequence[int]] = None, input_dim: Optional[int] = None,
                 name: Optional[str] = None):
        if input_dim is not None and input_shape is None:
            input_shape = (int(input_dim),)
        self.input_shape = tuple(int(s) for s in input_shape) if input_shape is not None else None
        self.name = name

    # -- spec ---------------------------------------------------------------------------------
    def get_config(self) -> dict:
        cfg = {}
        if self.input_shape is not None:
            cfg["input_shape"] = list(self.input_shape)
        if self.name is not None:
            cfg["name"] = self.name
        return cfg

    # -- shapes / parameters ------------------------------------------------------------------
    def output_shape(self, in_shape: Tuple[int, ...]) -> Tuple[int, ...]:
        return in_shape

    def param_shapes(self, in_shape: Tuple[int, ...]) -> List[Tuple[str, Tuple[int, ...], bool]]:
        """[(name, shape, trainable)] in flat-buffer order."""
        return []

    def init_params(self, in_shape, views: Dict[str, torch.Tensor], gen: torch.Generator) -> None:
        pass

    # -- autograd executor --------------------------------------------------------------------
    def forward(self, x: torch.Tensor, p: Dict[str, torch.Tensor], training: bool, ctx: dict) -> torch.Tensor:
        return x

    # -- Keras weight-layout conversion ---------------------------------------------------------
    def to_keras(self, name: str, t: torch.Tensor) -> torch.Tensor:
        return t

    def from_keras(self, name: str, t: torch.Tensor) -> torch.Tensor:
        return t


def _apply_activation(x: torch.Tensor, name: Optional[str]) -> torch.Tensor:
    if name is None or name == "linear":
        return x
    if name == "relu":
        return F.relu(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "softmax":
        return F.softmax(x, dim=-1)
    if name == "elu":
        return F.elu(x)
    raise ValueError(f"unknown activation {name!r}")


def _glorot_uniform_(t: torch.Tensor, fan_in: int, fan_out: int, gen: torch.Generator) -> None:
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    t.copy_((torch.rand(t.shape, generator=gen) * 2.0 - 1.0) * limit)


class Dense(Layer):
    class_name = "Dense"

    def __init__(self, units: int, activation: Optional[str] = None, use_bias: bool = True, **kw):
        super().__init__(**kw)
        self.units = int(units)
        self.activation = activation
        self.use_bias = bool(use_bias)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(units=self.units, activation=self.activation, use_bias=self.use_bias)
        return cfg

    def output_shape(self, in_shape):
        return tuple(in_shape[:-1]) + (self.units,)

    def param_shapes(self, in_shape):
        shapes = [("kernel", (self.units, int(in_shape[-1])), True)]
        if self.use_bias:
            shapes.append(("bias", (self.units,), True))
        return shapes

    def init_params(self, in_shape, views, gen):
        _glorot_uniform_(views["kernel"], int(in_shape[-1]), self.units, gen)
        if self.use_bias:
            views["bias"].zero_()

    def forward(self, x, p, training, ctx):
        y = F.linear(x, p["kernel"], p.get("bias"))
        if self.activation == "softmax" and ctx.get("logits_tail") is self:
            return y  # fused with the loss: keep logits
        return _apply_activation(y, self.activation)

    def to_keras(self, name, t):
        return t.t() if name == "kernel" else t

    def from_keras(self, name, t):
        return t.t() if name == "kernel" else t


class Activation(Layer):
    class_name = "Activation"

    def __init__(self, activation: str, **kw):
        super().__init__(**kw)
        self.activation = activation

    def get_config(self):
        cfg = super().get_config()
        cfg.update(activation=self.activation)
        return cfg

    def forward(self, x, p, training, ctx):
        if self.activation == "softmax" and ctx.get("logits_tail") is self:
            return x
        return _apply_activation(x, self.activation)


class Dropout(Layer):
    class_name = "Dropout"

    def __init__(self, rate: float, **kw):
        super().__init__(**kw)
        self.rate = float(rate)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(rate=self.rate)
        return cfg

    def forward(self, x, p, training, ctx):
        return F.dropout(x, self.rate, training=training)


class Flatten(Layer):
    class_name = "Flatten"

    def output_shape(self, in_shape):
        return (int(np.prod(in_shape)),)

    def forward(self, x, p, training, ctx):
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    class_name = "Reshape"

    def __init__(self, target_shape: Sequence[int], **kw):
        super().__init__(**kw)
        self.target_shape = tuple(int(s) for s in target_shape)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(target_shape=list(self.target_shape))
        return cfg

    def output_shape(self, in_shape):
        return self.target_shape

    def forward(self, x, p, training, ctx):
        return x.reshape((x.shape[0],) + self.target_shape)


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _conv_out(size: int, k: int, stride: int, pad: int) -> int:
    return (size + 2 * pad - k) // stride + 1


class Conv2D(Layer):
    """2-D convolution on NHWC activations (Keras ``Convolution2D`` in ``examples/mnist.py:150-155``)."""

    class_name = "Conv2D"

    def __init__(self, filters: int, kernel_size=3, strides=1, padding: str = "valid",
                 activation: Optional[str] = None, use_bias: bool = True, **kw):
        super().__init__(**kw)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        if self.strides[0] != self.strides[1]:
            raise ValueError("only square strides are supported")
        self.padding = padding
        self.activation = activation
        self.use_bias = bool(use_bias)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(filters=self.filters, kernel_size=list(self.kernel_size), strides=list(self.strides),
                   padding=self.padding, activation=self.activation, use_bias=self.use_bias)
        return cfg

    def pad_amount(self) -> int:
        if self.padding == "valid":
            return 0
        if self.padding == "same":
            if self.kernel_size[0] % 2 == 0:
                raise ValueError("'same' padding needs an odd kernel")
            return self.kernel_size[0] // 2
        raise ValueError(f"unknown padding {self.padding!r}")

    def output_shape(self, in_shape):
        h, w, _ = in_shape
        p, s = self.pad_amount(), self.strides[0]
        return (_conv_out(h, self.kernel_size[0], s, p), _conv_out(w, self.kernel_size[1], s, p), self.filters)

    def param_shapes(self, in_shape):
        cin = int(in_shape[-1])
        shapes = [("kernel", (self.filters, self.kernel_size[0], self.kernel_size[1], cin), True)]
        if self.use_bias:
            shapes.append(("bias", (self.filters,), True))
        return shapes

    def init_params(self, in_shape, views, gen):
        cin = int(in_shape[-1])
        rf = self.kernel_size[0] * self.kernel_size[1]
        _glorot_uniform_(views["kernel"], cin * rf, self.filters * rf, gen)
        if self.use_bias:
            views["bias"].zero_()

    def forward(self, x, p, training, ctx):
        w = p["kernel"].permute(0, 3, 1, 2)  # [cout, cin, kh, kw]
        y = F.conv2d(x.permute(0, 3, 1, 2), w, p.get("bias"), stride=self.strides[0], padding=self.pad_amount())
        return _apply_activation(y.permute(0, 2, 3, 1), self.activation)

    def to_keras(self, name, t):
        return t.permute(1, 2, 3, 0) if name == "kernel" else t

    def from_keras(self, name, t):
        return t.permute(3, 0, 1, 2) if name == "kernel" else t


Convolution2D = Conv2D


class MaxPooling2D(Layer):
    class_name = "MaxPooling2D"

    def __init__(self, pool_size=2, strides=None, **kw):
        super().__init__(**kw)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size

    def get_config(self):
        cfg = super().get_config()
        cfg.update(pool_size=list(self.pool_size), strides=list(self.strides))
        return cfg

    def output_shape(self, in_shape):
        h, w, c = in_shape
        return ((h - self.pool_size[0]) // self.strides[0] + 1, (w - self.pool_size[1]) // self.strides[1] + 1, c)

    def forward(self, x, p, training, ctx):
        y = F.max_pool2d(x.permute(0, 3, 1, 2), self.pool_size, self.strides)
        return y.permute(0, 2, 3, 1)


class GlobalAveragePooling2D(Layer):
    class_name = "GlobalAveragePooling2D"

    def output_shape(self, in_shape):
        return (int(in_shape[-1]),)

    def forward(self, x, p, training, ctx):
        return x.mean(dim=(1, 2))


class BatchNormalization(Layer):
    """Batch norm over the channel (last) axis.  As in Keras, the moving statistics are part of
    ``get_weights()`` and therefore of the center variable / committed residual."""

    class_name = "BatchNormalization"

    def __init__(self, momentum: float = 0.99, epsilon: float = 1e-3, **kw):
        super().__init__(**kw)
        self.momentum = float(momentum)
        self.epsilon = float(epsilon)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(momentum=self.momentum, epsilon=self.epsilon)
        return cfg

    def param_shapes(self, in_shape):
        c = int(in_shape[-1])
        return [("gamma", (c,), True), ("beta", (c,), True), ("moving_mean", (c,), False),
                ("moving_variance", (c,), False)]

    def init_params(self, in_shape, views, gen):
        views["gamma"].fill_(1.0)
        views["beta"].zero_()
        views["moving_mean"].zero_()
        views["moving_variance"].fill_(1.0)

    def forward(self, x, p, training, ctx):
        dims = tuple(range(x.dim() - 1))
        if training:
            mean = x.mean(dim=dims)
            var = x.var(dim=dims, unbiased=False)
            # moving statistics live in the flat buffer too; the in-place update is deferred until
            # after backward (views share the buffer's autograd version counter)
            ctx.setdefault("deferred", []).append((p["moving_mean"], mean.detach(), self.momentum))
            ctx.setdefault("deferred", []).append((p["moving_variance"], var.detach(), self.momentum))
        else:
            mean, var = p["moving_mean"], p["moving_variance"]
        return (x - mean) * torch.rsqrt(var + self.epsilon) * p["gamma"] + p["beta"]


class ResidualBlock(Layer):
    """ResNet basic block: conv3x3-BN-ReLU-conv3x3-BN (+ 1x1 projection when the shape changes),
    add, ReLU.  Not in the reference (BASELINE config 5 names ResNet-18); defined here."""

    class_name = "ResidualBlock"

    def __init__(self, filters: int, strides: int = 1, **kw):
        super().__init__(**kw)
        self.filters = int(filters)
        self.strides = int(strides)
        self.conv1 = Conv2D(filters, 3, strides, "same", use_bias=False)
        self.bn1 = BatchNormalization()
        self.conv2 = Conv2D(filters, 3, 1, "same", use_bias=False)
        self.bn2 = BatchNormalization()
        self.proj = Conv2D(filters, 1, strides, "valid", use_bias=False)
        self.bnp = BatchNormalization()

    def get_config(self):
        cfg = super().get_config()
        cfg.update(filters=self.filters, strides=self.strides)
        return cfg

    def _needs_proj(self, in_shape) -> bool:
        return self.strides != 1 or int(in_shape[-1]) != self.filters

    def _sub(self, in_shape):
        mid = self.conv1.output_shape(in_shape)
        subs = [("conv1", self.conv1, in_shape), ("bn1", self.bn1, mid), ("conv2", self.conv2, mid),
                ("bn2", self.bn2, mid)]
        if self._needs_proj(in_shape):
            subs += [("proj", self.proj, in_shape), ("bnp", self.bnp, mid)]
        return subs

    def output_shape(self, in_shape):
        return self.conv1.output_shape(in_shape)

    def param_shapes(self, in_shape):
        out = []
        for prefix, layer, shp in self._sub(in_shape):
            out += [(f"{prefix}.{n}", s, t) for n, s, t in layer.param_shapes(shp)]
        return out

    def init_params(self, in_shape, views, gen):
        for prefix, layer, shp in self._sub(in_shape):
            layer.init_params(shp, {k[len(prefix) + 1:]: v for k, v in views.items() if k.startswith(prefix + ".")}, gen)

    def forward(self, x, p, training, ctx):
        def sub(prefix):
            return {k[len(prefix) + 1:]: v for k, v in p.items() if k.startswith(prefix + ".")}

        y = self.conv1.forward(x, sub("conv1"), training, ctx)
        y = F.relu(self.bn1.forward(y, sub("bn1"), training, ctx))
        y = self.conv2.forward(y, sub("conv2"), training, ctx)
        y = self.bn2.forward(y, sub("bn2"), training, ctx)
        if "proj.kernel" in p:
            x = self.bnp.forward(self.proj.forward(x, sub("proj"), training, ctx), sub("bnp"), training, ctx)
        return F.relu(x + y)

    def to_keras(self, name, t):
        return t.permute(1, 2, 3, 0) if name.endswith("kernel") and t.dim() == 4 else t

    def from_keras(self, name, t):
        return t.permute(3, 0, 1, 2) if name.endswith("kernel") and t.dim() == 4 else t


LAYER_CLASSES = {c.class_name: c for c in (Dense, Activation, Dropout, Flatten, Reshape, Conv2D, MaxPooling2D,
                                           GlobalAveragePooling2D, BatchNormalization, ResidualBlock)}
LAYER_CLASSES["Convolution2D"] = Conv2D


# --------------------------------------------------------------------------------------------
# losses / metrics (autograd executor)
# --------------------------------------------------------------------------------------------
LOSS_ALIASES = {
    "categorical_crossentropy": "categorical_crossentropy",
    "sparse_categorical_crossentropy": "categorical_crossentropy",
    "binary_crossentropy": "binary_crossentropy",
    "mse": "mse",
    "mean_squared_error": "mse",
    # the remaining Keras-1 objectives (autograd executors; the native engine lowers cross-entropy and MSE)
    "mae": "mae", "mean_absolute_error": "mae",
    "mape": "mape", "mean_absolute_percentage_error": "mape",
    "msle": "msle", "mean_squared_logarithmic_error": "msle",
    "hinge": "hinge", "squared_hinge": "squared_hinge",
    "kld": "kld", "kullback_leibler_divergence": "kld",
    "poisson": "poisson", "cosine_proximity": "cosine_proximity",
}


def _labels_to_index(y: torch.Tensor) -> Optional[torch.Tensor]:
    if y.dim() == 1 or (y.dim() == 2 and y.shape[1] == 1 and not y.dtype.is_floating_point):
        return y.reshape(-1).long()
    return None


def compute_loss(loss: str, out: torch.Tensor, y: torch.Tensor, from_logits: bool) -> torch.Tensor:
    kind = LOSS_ALIASES.get(loss)
    if kind is None:
        raise ValueError(f"unsupported loss {loss!r}")
    if kind == "categorical_crossentropy":
        logp = F.log_softmax(out, dim=-1) if from_logits else torch.log(out.clamp_min(1e-7))
        idx = _labels_to_index(y)
        if idx is not None:
            return F.nll_loss(logp, idx)
        return -(y.to(logp.dtype) * logp).sum(dim=-1).mean()
    if from_logits:  # ``out`` are the pre-softmax scores: every other objective is defined on the probabilities
        out = torch.softmax(out, dim=-1)
    if kind == "binary_crossentropy":
        p = out.clamp(1e-7, 1 - 1e-7)
        t = y.to(p.dtype).reshape(p.shape)
        return -(t * torch.log(p) + (1 - t) * torch.log(1 - p)).mean()
    t = y.to(out.dtype).reshape(out.shape)
    if kind == "mse":
        return ((out - t) ** 2).mean()
    if kind == "mae":
        return (out - t).abs().mean()
    if kind == "mape":
        return 100.0 * ((t - out).abs() / t.abs().clamp_min(1e-7)).mean()
    if kind == "msle":
        return ((torch.log1p(out.clamp_min(1e-7)) - torch.log1p(t.clamp_min(1e-7))) ** 2).mean()
    if kind == "hinge":
        return torch.clamp(1.0 - t * out, min=0.0).mean()
    if kind == "squared_hinge":
        return (torch.clamp(1.0 - t * out, min=0.0) ** 2).mean()
    if kind == "kld":
        tc, oc = t.clamp(1e-7, 1.0), out.clamp(1e-7, 1.0)
        return (tc * torch.log(tc / oc)).sum(dim=-1).mean()
    if kind == "poisson":
        return (out - t * torch.log(out + 1e-7)).mean()
    if kind == "cosine_proximity":
        return -(F.normalize(t, dim=-1) * F.normalize(out, dim=-1)).sum(dim=-1).mean()
    raise ValueError(f"unsupported loss {loss!r}")


def compute_accuracy(out: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    idx = _labels_to_index(y)
    if idx is None:
        if y.dim() == 2 and y.shape[1] > 1:
            idx = y.argmax(dim=-1)
        else:
            return ((out.reshape(-1) > 0.5) == (y.reshape(-1) > 0.5)).float().mean()
    return (out.argmax(dim=-1) == idx).float().mean()


# --------------------------------------------------------------------------------------------
# Sequential
# --------------------------------------------------------------------------------------------
class ParamSegment:
    __slots__ = ("layer_index", "name", "offset", "shape", "size", "trainable")

    def __init__(self, layer_index, name, offset, shape, trainable):
        self.layer_index = layer_index
        self.name = name
        self.offset = offset
        self.shape = tuple(shape)
        self.size = int(np.prod(shape))
        self.trainable = trainable


class Sequential:
    """Linear stack of layers with one flat fp32 parameter buffer."""

    def __init__(self, layers: Optional[Sequence[Layer]] = None, name: str = "sequential", seed: Optional[int] = None):
        self.name = name
        self.layers: List[Layer] = []
        self.seed = seed
        self.flat: Optional[torch.Tensor] = None
        self.segments: List[ParamSegment] = []
        self.shapes: List[Tuple[int, ...]] = []  # shapes[i] = input shape of layer i; [-1] = output
        self.loss: Optional[str] = None
        self.optimizer = None
        self.metrics: List[str] = []
        self._replica = None
        for layer in layers or []:
            self.add(layer)

    # -- construction -----------------------------------------------------------------------
    def add(self, layer: Layer) -> None:
        self.layers.append(layer)
        self.flat = None  # rebuilt lazily

    @property
    def input_shape(self) -> Tuple[int, ...]:
        if not self.layers or self.layers[0].input_shape is None:
            raise ValueError("the first layer needs input_shape / input_dim")
        return self.layers[0].input_shape

    def build(self, device: Optional[torch.device] = None) -> "Sequential":
        if self.flat is not None:
            return self
        shape = self.input_shape
        self.shapes = [shape]
        self.segments = []
        offset = 0
        for li, layer in enumerate(self.layers):
            for name, shp, trainable in layer.param_shapes(shape):
                offset = (offset + 7) // 8 * 8  # every segment 32-byte aligned (16 B in the bf16 shadow: TMA)
                seg = ParamSegment(li, name, offset, shp, trainable)
                self.segments.append(seg)
                offset += seg.size
            shape = layer.output_shape(shape)
            self.shapes.append(shape)
        total = (offset + 7) // 8 * 8
        self.flat = torch.zeros(total, dtype=torch.float32)
        gen = torch.Generator()
        gen.manual_seed(self.seed if self.seed is not None else torch.seed() % (2 ** 31))
        for li, layer in enumerate(self.layers):
            layer.init_params(self.shapes[li], self._layer_views(self.flat, li), gen)
        if device is not None:
            self.flat = self.flat.to(device)
        return self

    @property
    def output_shape(self) -> Tuple[int, ...]:
        self.build()
        return self.shapes[-1]

    @property
    def num_params(self) -> int:
        self.build()
        return int(self.flat.numel())

    def count_params(self) -> int:
        self.build()
        return sum(s.size for s in self.segments)

    def _layer_views(self, flat: torch.Tensor, layer_index: int) -> Dict[str, torch.Tensor]:
        return {s.name: flat[s.offset:s.offset + s.size].view(s.shape)
                for s in self.segments if s.layer_index == layer_index}

    def trainable_mask(self) -> torch.Tensor:
        """1.0 where the flat element is trainable (BN statistics are not)."""
        self.build()
        m = torch.zeros_like(self.flat)
        for s in self.segments:
            if s.trainable:
                m[s.offset:s.offset + s.size] = 1.0
        return m

    def logits_tail(self) -> Optional[Layer]:
        """The trailing softmax (fused with categorical cross-entropy), if any."""
        if not self.layers:
            return None
        last = self.layers[-1]
        if getattr(last, "activation", None) == "softmax":
            return last
        return None

    # -- Keras-shaped API -------------------------------------------------------------------
    def get_config(self) -> dict:
        return {"class_name": "Sequential", "name": self.name,
                "layers": [{"class_name": l.class_name, "config": l.get_config()} for l in self.layers]}

    def to_json(self) -> str:
        return json.dumps(self.get_config())

    def get_weights(self) -> List[np.ndarray]:
        self.build()
        out = []
        for s in self.segments:
            t = self.flat[s.offset:s.offset + s.size].view(s.shape).detach().cpu()
            out.append(self.layers[s.layer_index].to_keras(s.name, t).contiguous().numpy().copy())
        return out

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        self.build()
        if len(weights) != len(self.segments):
            raise ValueError(f"expected {len(self.segments)} arrays, got {len(weights)}")
        with torch.no_grad():
            for s, w in zip(self.segments, weights):
                t = self.layers[s.layer_index].from_keras(s.name, torch.as_tensor(np.asarray(w), dtype=torch.float32))
                self.flat[s.offset:s.offset + s.size].copy_(t.reshape(-1).to(self.flat.device))

    def get_flat_weights(self) -> torch.Tensor:
        self.build()
        return self.flat

    def set_flat_weights(self, flat: torch.Tensor) -> None:
        self.build()
        with torch.no_grad():
            self.flat.copy_(flat.reshape(-1).to(self.flat.device, torch.float32))

    def copy(self) -> "Sequential":
        m = model_from_json(self.to_json())
        m.seed = self.seed
        m.build()
        m.set_flat_weights(self.get_flat_weights().detach().cpu())
        m.loss, m.optimizer, m.metrics = self.loss, self.optimizer, list(self.metrics)
        return m

    def summary(self) -> str:
        self.build()
        lines = [f"Model: {self.name}"]
        for li, layer in enumerate(self.layers):
            n = sum(s.size for s in self.segments if s.layer_index == li)
            lines.append(f"  {li:2d} {layer.class_name:<24s} out={self.shapes[li + 1]} params={n}")
        lines.append(f"Total params: {self.count_params()}")
        text = "\n".join(lines)
        print(text)
        return text

    # -- autograd executor ------------------------------------------------------------------
    def forward(self, x: torch.Tensor, flat: Optional[torch.Tensor] = None, training: bool = False,
                logits: bool = False, ctx: Optional[dict] = None) -> torch.Tensor:
        """Run the stack with parameters taken from ``flat`` (defaults to the model's buffer).

        Pass ``ctx`` to receive deferred state updates (BatchNorm moving statistics) and apply
        them with :func:`apply_deferred` after ``backward``; otherwise they are applied on return.
        """
        self.build()
        flat = self.flat if flat is None else flat
        own_ctx = ctx is None
        ctx = {} if ctx is None else ctx
        ctx["logits_tail"] = self.logits_tail() if logits else None
        x = prepare_input(x, self.input_shape, flat.device)
        for li, layer in enumerate(self.layers):
            x = layer.forward(x, self._layer_views(flat, li), training, ctx)
        if own_ctx:
            apply_deferred(ctx)
        return x

    def compile(self, loss: str = "categorical_crossentropy", optimizer="sgd", metrics: Sequence[str] = ("accuracy",),
                loss_weights=None) -> None:
        self.loss = loss
        self.optimizer = optimizer
        self.metrics = list(metrics or [])
        self._replica = None

    def _get_replica(self):
        if self._replica is None:
            from ..parallel.replica import TorchReplica

            if self.loss is None:
                raise RuntimeError("call compile() first")
            self._replica = TorchReplica(self, self.optimizer, self.loss, device=self.get_flat_weights().device,
                                         share_model_buffer=True)
        return self._replica

    def train_on_batch(self, x, y) -> List[float]:
        loss, acc = self._get_replica().train_on_batch(x, y)
        return [float(loss), float(acc)]

    def predict(self, x, batch_size: int = 8192) -> np.ndarray:
        self.build()
        x = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
        outs = []
        with torch.no_grad():
            for i in range(0, x.shape[0], batch_size):
                outs.append(self.forward(x[i:i + batch_size], training=False).float().cpu())
        return torch.cat(outs, dim=0).numpy()

    def evaluate(self, x, y, batch_size: int = 8192) -> List[float]:
        self.build()
        x = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
        y = torch.as_tensor(np.asarray(y)) if not isinstance(y, torch.Tensor) else y
        tot_l, tot_a, n = 0.0, 0.0, 0
        with torch.no_grad():
            for i in range(0, x.shape[0], batch_size):
                out = self.forward(x[i:i + batch_size], training=False, logits=True)
                yy = y[i:i + batch_size].to(out.device)
                b = out.shape[0]
                tot_l += float(compute_loss(self.loss or "categorical_crossentropy", out, yy,
                                            self.logits_tail() is not None)) * b
                tot_a += float(compute_accuracy(out, yy)) * b
                n += b
        return [tot_l / max(n, 1), tot_a / max(n, 1)]

    def fit(self, x, y, batch_size: int = 32, epochs: int = 1, shuffle: bool = True, seed: Optional[int] = None,
            verbose: int = 0) -> dict:
        """Keras-style single-process training loop over ``train_on_batch`` (partial last batch dropped, like
        the reference's workers).  Returns ``{"loss": [...], "accuracy": [...]}`` with one entry per epoch."""
        x = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
        y = torch.as_tensor(np.asarray(y)) if not isinstance(y, torch.Tensor) else y
        n = x.shape[0]
        g = torch.Generator()
        if seed is not None:
            g.manual_seed(seed)
        hist = {"loss": [], "accuracy": []}
        for epoch in range(int(epochs)):
            order = torch.randperm(n, generator=g) if shuffle else torch.arange(n)
            tot_l = tot_a = 0.0
            steps = n // batch_size
            for i in range(steps):
                idx = order[i * batch_size:(i + 1) * batch_size]
                l, a = self.train_on_batch(x[idx], y[idx])
                tot_l += l
                tot_a += a
            hist["loss"].append(tot_l / max(steps, 1))
            hist["accuracy"].append(tot_a / max(steps, 1))
            if verbose:
                print(f"epoch {epoch + 1}/{epochs}: loss {hist['loss'][-1]:.4f} accuracy {hist['accuracy'][-1]:.4f}")
        return hist

    # -- persistence (README TODO of the reference: "Save/Load Keras model") ------------------------
    def save_weights(self, path: str) -> None:
        torch.save({"flat": self.get_flat_weights().detach().cpu()}, path)

    def load_weights(self, path: str) -> None:
        flat = torch.load(path, weights_only=True)["flat"]
        if flat.numel() != self.num_params:
            raise ValueError(f"weight file holds {flat.numel()} parameters, the model has {self.num_params}")
        self.set_flat_weights(flat)

    def save(self, path: str) -> None:
        """Architecture (JSON) + weights in one file; see :func:`load_model`."""
        torch.save({"model": self.to_json(), "flat": self.get_flat_weights().detach().cpu(), "loss": self.loss,
                    "optimizer": self.optimizer if isinstance(self.optimizer, (str, dict, type(None))) else None}, path)

    def to(self, device) -> "Sequential":
        self.build()
        self.flat = self.flat.to(device)
        self._replica = None
        return self


def apply_deferred(ctx: dict) -> None:
    """Apply deferred exponential-moving-average updates collected during a training forward."""
    with torch.no_grad():
        for target, value, momentum in ctx.pop("deferred", []):
            target.mul_(momentum).add_(value * (1 - momentum))


def prepare_input(x, input_shape: Tuple[int, ...], device) -> torch.Tensor:
    """uint8 / float rows -> float32 tensor of shape [B, *input_shape] on ``device``."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    x = x.to(device)
    if x.dtype != torch.float32:
        x = x.float()
    if tuple(x.shape[1:]) != tuple(input_shape):
        x = x.reshape((x.shape[0],) + tuple(input_shape))
    return x


def model_from_config(cfg: dict):
    if cfg.get("class_name") == "Model":  # functional graph (several inputs / outputs)
        from .functional import Model

        return Model.from_config(cfg)
    layers = []
    for lc in cfg["layers"]:
        cls = LAYER_CLASSES[lc["class_name"]]
        layers.append(cls(**lc["config"]))
    return Sequential(layers, name=cfg.get("name", "sequential"))


def model_from_json(text: str):
    return model_from_config(json.loads(text))


def load_model(path: str) -> Sequential:
    """Inverse of :meth:`Sequential.save` (re-compiles with the stored loss / optimizer when present)."""
    d = torch.load(path, weights_only=True)  # str / dict / tensor payload only: no arbitrary unpickling
    m = model_from_json(d["model"])
    m.build()
    m.set_flat_weights(d["flat"])
    if d.get("loss"):
        m.compile(d["loss"], d.get("optimizer") or "sgd")
    return m
