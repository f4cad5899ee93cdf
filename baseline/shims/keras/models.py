"""``keras.models`` stand-in: ``Sequential`` on torch modules + autograd + ``torch.optim``.

Weight layout follows Keras (``Dense.kernel`` is ``[in, out]``, ``Conv2D.kernel`` is ``[kh, kw, in, out]``,
images are channels-last), so ``get_weights()`` / ``set_weights()`` round-trip with real Keras weight
lists.  ``get_weights()`` returns a ``WeightList`` whose ``__array__`` yields a 1-D *object* array of the
per-layer arrays: that is what ``np.asarray(model.get_weights())`` meant on the numpy the reference was
written against (ragged list -> object array); numpy >= 1.24 refuses to guess it.
"""
import json

import numpy as np

from . import backend as K
from . import optimizers as _optimizers
from .layers import LAYERS, Layer


class WeightList(list):
    """A plain list of per-layer ndarrays that converts to an object ndarray on request."""

    def __array__(self, dtype=None, copy=None):
        out = np.empty(len(self), dtype=object)
        for i, w in enumerate(self):
            out[i] = w
        return out


def _activation(name):
    import torch
    import torch.nn.functional as F

    if name in (None, "linear"):
        return lambda t: t
    if name == "relu":
        return F.relu
    if name == "softmax":
        return lambda t: F.softmax(t, dim=-1)
    if name == "sigmoid":
        return torch.sigmoid
    if name == "tanh":
        return torch.tanh
    raise ValueError("unsupported activation %r" % (name,))


class Sequential(object):
    def __init__(self, layers=None, name=None):
        self.layers = []
        self.name = name
        self._built = False
        self._optimizer = None
        for layer in layers or []:
            self.add(layer)

    # -- construction ----------------------------------------------------------------------------
    def add(self, layer):
        assert isinstance(layer, Layer)
        self.layers.append(layer)
        self._built = False

    def _build(self):
        """Create the torch parameters (Glorot-uniform kernels, zero biases -- the Keras defaults)."""
        if self._built:
            return
        import torch

        dev = K.device()
        shape = self.layers[0].input_shape
        if shape is None:
            raise ValueError("the first layer needs input_shape / input_dim")
        self._params = []      # torch parameters in Keras get_weights() order
        self._plan = []        # (kind, layer, param indices)
        gen = torch.Generator().manual_seed(1337)

        def glorot(*dims, fan_in, fan_out):
            lim = float(np.sqrt(6.0 / (fan_in + fan_out)))
            t = (torch.rand(*dims, generator=gen) * 2 - 1) * lim
            return torch.nn.Parameter(t.to(dev))

        for layer in self.layers:
            cn = layer.class_name
            idx = []
            if cn == "Dense":
                fan_in = int(shape[-1])
                idx.append(len(self._params)); self._params.append(glorot(fan_in, layer.units, fan_in=fan_in, fan_out=layer.units))
                if layer.use_bias:
                    idx.append(len(self._params)); self._params.append(torch.nn.Parameter(torch.zeros(layer.units, device=dev)))
                shape = tuple(shape[:-1]) + (layer.units,)
            elif cn == "Conv2D":
                h, w, c = shape
                kh, kw = layer.kernel_size
                idx.append(len(self._params))
                self._params.append(glorot(kh, kw, c, layer.filters, fan_in=kh * kw * c, fan_out=kh * kw * layer.filters))
                if layer.use_bias:
                    idx.append(len(self._params)); self._params.append(torch.nn.Parameter(torch.zeros(layer.filters, device=dev)))
                sh, sw = layer.strides
                if layer.padding == "same":
                    oh, ow = -(-h // sh), -(-w // sw)
                else:
                    oh, ow = (h - kh) // sh + 1, (w - kw) // sw + 1
                shape = (oh, ow, layer.filters)
            elif cn == "MaxPooling2D":
                h, w, c = shape
                shape = ((h - layer.pool_size[0]) // layer.strides[0] + 1, (w - layer.pool_size[1]) // layer.strides[1] + 1, c)
            elif cn == "Flatten":
                shape = (int(np.prod(shape)),)
            elif cn == "Reshape":
                shape = layer.target_shape
            self._plan.append((cn, layer, idx))
        self.output_shape = (None,) + tuple(shape)
        self._built = True

    # -- forward -----------------------------------------------------------------------------------
    def _forward(self, x, training):
        import torch.nn.functional as F

        P = self._params
        last = len(self._plan) - 1
        logits = None
        for li, (cn, layer, idx) in enumerate(self._plan):
            if cn == "Dense":
                x = x.matmul(P[idx[0]])
                if layer.use_bias:
                    x = x + P[idx[1]]
                if li == last and layer.activation == "softmax":
                    logits = x
                x = _activation(layer.activation)(x)
            elif cn == "Activation":
                if li == last and layer.activation == "softmax":
                    logits = x
                x = _activation(layer.activation)(x)
            elif cn == "Dropout":
                x = F.dropout(x, layer.rate, training)
            elif cn == "Flatten":
                x = x.reshape(x.shape[0], -1)
            elif cn == "Reshape":
                x = x.reshape((x.shape[0],) + tuple(layer.target_shape))
            elif cn == "Conv2D":
                w = P[idx[0]].permute(3, 2, 0, 1)                      # [kh, kw, in, out] -> [out, in, kh, kw]
                pad = 0
                if layer.padding == "same":
                    pad = (layer.kernel_size[0] // 2, layer.kernel_size[1] // 2)
                x = F.conv2d(x.permute(0, 3, 1, 2), w, P[idx[1]] if layer.use_bias else None, stride=layer.strides, padding=pad)
                x = _activation(layer.activation)(x.permute(0, 2, 3, 1))
            elif cn == "MaxPooling2D":
                x = F.max_pool2d(x.permute(0, 3, 1, 2), layer.pool_size, layer.strides).permute(0, 2, 3, 1)
            else:
                raise ValueError("unsupported layer %s" % cn)
        return x, logits

    # -- Keras API ---------------------------------------------------------------------------------
    def compile(self, optimizer, loss, metrics=None, loss_weights=None, **kwargs):
        self._build()
        self.loss = loss
        self.metrics = list(metrics or [])
        self.loss_weights = loss_weights
        self.optimizer = _optimizers.get(optimizer)
        self._optimizer = self.optimizer.build(self._params)

    def _tensor(self, a):
        import torch

        if isinstance(a, (list, tuple)) and len(a) == 1:
            a = a[0]
        t = torch.as_tensor(np.asarray(a))
        if t.dtype != torch.float32:
            t = t.float()
        return t.to(K.device(), non_blocking=True)

    def train_on_batch(self, x, y, **kwargs):
        import torch

        self._build()
        if self._optimizer is None:
            raise RuntimeError("compile() the model before training")
        xt, yt = self._tensor(x), self._tensor(y)
        out, logits = self._forward(xt, True)
        loss = self._loss(out, logits, yt)
        self._optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self._optimizer.step()
        res = [float(loss.item())]
        for m in self.metrics:
            if m in ("accuracy", "acc"):
                with torch.no_grad():
                    if yt.dim() == 2 and yt.shape[1] > 1:
                        acc = (out.argmax(dim=1) == yt.argmax(dim=1)).float().mean()
                    else:
                        acc = ((out.reshape(-1) > 0.5).float() == yt.reshape(-1)).float().mean()
                res.append(float(acc.item()))
        return res if len(res) > 1 else res[0]

    def _loss(self, out, logits, y):
        import torch
        import torch.nn.functional as F

        name = self.loss
        if name == "categorical_crossentropy":
            if logits is not None:
                return -(y * F.log_softmax(logits, dim=-1)).sum(dim=-1).mean()
            p = out / out.sum(dim=-1, keepdim=True)
            return -(y * torch.log(p.clamp(K.epsilon(), 1.0 - K.epsilon()))).sum(dim=-1).mean()
        if name in ("mse", "mean_squared_error"):
            return ((out - y.reshape(out.shape)) ** 2).mean()
        if name == "binary_crossentropy":
            return F.binary_cross_entropy(out.clamp(K.epsilon(), 1.0 - K.epsilon()), y.reshape(out.shape))
        raise ValueError("unsupported loss %r" % (name,))

    def predict(self, x, batch_size=32, **kwargs):
        import torch

        self._build()
        with torch.no_grad():
            out, _ = self._forward(self._tensor(x), False)
        return out.cpu().numpy()

    def get_weights(self):
        self._build()
        return WeightList(p.detach().cpu().numpy() for p in self._params)

    def set_weights(self, weights):
        import torch

        self._build()
        weights = list(weights)
        if len(weights) != len(self._params):
            raise ValueError("expected %d weight arrays, got %d" % (len(self._params), len(weights)))
        with torch.no_grad():
            for p, w in zip(self._params, weights):
                p.copy_(torch.as_tensor(np.asarray(w, dtype=np.float32)).reshape(p.shape))

    def count_params(self):
        self._build()
        return int(sum(p.numel() for p in self._params))

    def summary(self):
        self._build()
        print("Sequential (torch shim): %d layers, %d parameters" % (len(self.layers), self.count_params()))

    def get_config(self):
        return [{"class_name": l.class_name, "config": l._base_config()} for l in self.layers]

    def to_json(self, **kwargs):
        return json.dumps({"class_name": "Sequential", "config": self.get_config(), "keras_version": "2.0.8+torchshim",
                           "backend": "torch"})


def model_from_json(json_string, custom_objects=None):
    spec = json.loads(json_string)
    if spec.get("class_name") != "Sequential":
        raise ValueError("only Sequential models are supported by the shim")
    cfg = spec["config"]
    if isinstance(cfg, dict):  # Keras >= 2.2.3 nests the layer list
        cfg = cfg["layers"]
    model = Sequential()
    for l in cfg:
        model.add(LAYERS[l["class_name"]].from_config(l["config"]))
    return model
