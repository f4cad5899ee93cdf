"""Functional models: several inputs, several outputs, one flat parameter buffer.

The reference compiles whatever Keras model it is given with a *list* of feature columns and a *list* of
label columns -- one per model input / output -- plus ``loss_weights`` and ``metrics``
(``distkeras/workers.py:65-66, 75-76, 117-118, 140-148``).  :class:`Model` is the small functional API that
makes that meaningful here::

    a, b = Input((30,), name="a"), Input((8,), name="b")
    h = Dense(64, activation="relu")(Concatenate()([a, b]))
    cls = Dense(2, activation="softmax", name="cls")(h)
    reg = Dense(1, name="reg")(h)
    model = Model([a, b], [cls, reg])

It keeps the contract every trainer / parameter server relies on -- ``to_json`` / ``model_from_json``, ONE flat
fp32 buffer (``get_flat_weights``), Keras-ordered ``get_weights`` -- so all PS algorithms work on it unchanged.
Graphs run on the autograd executor (``TorchReplica``); the native sm_100a planner lowers ``Sequential`` stacks.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import core
from .core import Layer, ParamSegment, apply_deferred, prepare_input


class SymbolicTensor:
    """Output of a layer call (or a model input) while the graph is being described."""

    def __init__(self, layer: Optional[Layer], parents: Sequence["SymbolicTensor"], shape: Tuple[int, ...], name: str):
        self.layer, self.parents, self.shape, self.name = layer, list(parents), tuple(shape), name


def Input(shape: Sequence[int], name: Optional[str] = None) -> SymbolicTensor:
    return SymbolicTensor(None, [], tuple(int(s) for s in shape), name or "input")


def _call_layer(self: Layer, inputs):
    """``layer(tensor)`` / ``layer([tensors])`` -> symbolic output (installed as ``Layer.__call__``)."""
    ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
    if not all(isinstance(t, SymbolicTensor) for t in ins):
        raise TypeError("layers are called on Input(...) / layer outputs when building a functional Model")
    shape = self.merge_shape([t.shape for t in ins]) if isinstance(self, Merge) else self.output_shape(ins[0].shape)
    return SymbolicTensor(self, ins, shape, self.name or self.class_name.lower())


Layer.__call__ = _call_layer  # type: ignore[assignment]


class Merge(Layer):
    """Base of the layers that take several inputs."""

    def merge_shape(self, shapes: List[Tuple[int, ...]]) -> Tuple[int, ...]:
        raise NotImplementedError

    def forward_many(self, xs: List[torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError


class Concatenate(Merge):
    class_name = "Concatenate"

    def __init__(self, axis: int = -1, **kw):
        super().__init__(**kw)
        self.axis = int(axis)

    def get_config(self):
        cfg = super().get_config()
        cfg.update(axis=self.axis)
        return cfg

    def merge_shape(self, shapes):
        ax = self.axis if self.axis >= 0 else len(shapes[0]) + self.axis
        out = list(shapes[0])
        out[ax] = sum(s[ax] for s in shapes)
        return tuple(out)

    def forward_many(self, xs):
        return torch.cat(xs, dim=self.axis if self.axis < 0 else self.axis + 1)


class Add(Merge):
    class_name = "Add"

    def merge_shape(self, shapes):
        return tuple(shapes[0])

    def forward_many(self, xs):
        out = xs[0]
        for x in xs[1:]:
            out = out + x
        return out


core.LAYER_CLASSES.update({"Concatenate": Concatenate, "Add": Add})


class Model:
    """Directed acyclic graph of layers with the same flat-buffer contract as ``Sequential``."""

    def __init__(self, inputs: Union[SymbolicTensor, Sequence[SymbolicTensor]],
                 outputs: Union[SymbolicTensor, Sequence[SymbolicTensor]], name: str = "model", seed: Optional[int] = None):
        self.name, self.seed = name, seed
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        # topological order of the layer nodes reachable from the outputs
        self.nodes: List[SymbolicTensor] = []
        seen = set()

        def visit(t: SymbolicTensor):
            if id(t) in seen:
                return
            seen.add(id(t))
            for p in t.parents:
                visit(p)
            if t.layer is not None:
                self.nodes.append(t)
            elif not any(t is i for i in self.inputs):
                raise ValueError(f"graph input {t.name!r} is not listed in `inputs`")

        for o in self.outputs:
            visit(o)
        self.layers: List[Layer] = [n.layer for n in self.nodes]
        self.flat: Optional[torch.Tensor] = None
        self.segments: List[ParamSegment] = []
        self.loss = self.optimizer = self.loss_weights = None
        self.metrics: List[str] = []
        self._replica = None

    # -- shapes / parameters ------------------------------------------------------------------
    @property
    def input_shapes(self) -> List[Tuple[int, ...]]:
        return [t.shape for t in self.inputs]

    @property
    def output_shapes(self) -> List[Tuple[int, ...]]:
        return [t.shape for t in self.outputs]

    @property
    def input_shape(self):
        return self.inputs[0].shape if len(self.inputs) == 1 else self.input_shapes

    @property
    def output_shape(self):
        return self.outputs[0].shape if len(self.outputs) == 1 else self.output_shapes

    @property
    def num_inputs(self) -> int:
        return len(self.inputs)

    @property
    def num_outputs(self) -> int:
        return len(self.outputs)

    def _in_shape(self, node: SymbolicTensor) -> Tuple[int, ...]:
        return node.parents[0].shape

    def build(self, device: Optional[torch.device] = None) -> "Model":
        if self.flat is not None:
            return self
        self.segments = []
        offset = 0
        for li, node in enumerate(self.nodes):
            for pname, shp, trainable in node.layer.param_shapes(self._in_shape(node)):
                offset = (offset + 7) // 8 * 8
                seg = ParamSegment(li, pname, offset, shp, trainable)
                self.segments.append(seg)
                offset += seg.size
        self.flat = torch.zeros((offset + 7) // 8 * 8, dtype=torch.float32)
        gen = torch.Generator()
        gen.manual_seed(self.seed if self.seed is not None else torch.seed() % (2 ** 31))
        for li, node in enumerate(self.nodes):
            node.layer.init_params(self._in_shape(node), self._layer_views(self.flat, li), gen)
        if device is not None:
            self.flat = self.flat.to(device)
        return self

    @property
    def num_params(self) -> int:
        self.build()
        return int(self.flat.numel())

    def count_params(self) -> int:
        self.build()
        return sum(s.size for s in self.segments)

    def _layer_views(self, flat: torch.Tensor, layer_index: int) -> Dict[str, torch.Tensor]:
        return {s.name: flat[s.offset:s.offset + s.size].view(s.shape) for s in self.segments if s.layer_index == layer_index}

    def trainable_mask(self) -> torch.Tensor:
        self.build()
        m = torch.zeros_like(self.flat)
        for s in self.segments:
            if s.trainable:
                m[s.offset:s.offset + s.size] = 1.0
        return m

    def logits_tails(self) -> List[Optional[Layer]]:
        """Per output: the trailing softmax layer (fused with categorical cross-entropy), if any."""
        return [o.layer if getattr(o.layer, "activation", None) == "softmax" else None for o in self.outputs]

    def logits_tail(self) -> Optional[Layer]:
        return self.logits_tails()[0] if len(self.outputs) == 1 else None

    # -- Keras-shaped API -------------------------------------------------------------------
    def get_config(self) -> dict:
        ids = {id(t): f"in{i}" for i, t in enumerate(self.inputs)}
        ids.update({id(n): f"n{i}" for i, n in enumerate(self.nodes)})
        return {"class_name": "Model", "name": self.name,
                "inputs": [{"id": ids[id(t)], "shape": list(t.shape), "name": t.name} for t in self.inputs],
                "nodes": [{"id": ids[id(n)], "class_name": n.layer.class_name, "config": n.layer.get_config(),
                           "inbound": [ids[id(p)] for p in n.parents]} for n in self.nodes],
                "outputs": [ids[id(o)] for o in self.outputs]}

    def to_json(self) -> str:
        return json.dumps(self.get_config())

    @classmethod
    def from_config(cls, cfg: dict) -> "Model":
        tensors: Dict[str, SymbolicTensor] = {}
        inputs = []
        for spec in cfg["inputs"]:
            t = Input(spec["shape"], spec.get("name"))
            tensors[spec["id"]] = t
            inputs.append(t)
        for spec in cfg["nodes"]:
            layer = core.LAYER_CLASSES[spec["class_name"]](**spec["config"])
            ins = [tensors[i] for i in spec["inbound"]]
            tensors[spec["id"]] = layer(ins if isinstance(layer, Merge) else ins[0])
        return cls(inputs, [tensors[i] for i in cfg["outputs"]], name=cfg.get("name", "model"))

    def get_weights(self) -> List[np.ndarray]:
        self.build()
        return [self.layers[s.layer_index].to_keras(s.name, self.flat[s.offset:s.offset + s.size].view(s.shape).detach().cpu())
                .contiguous().numpy().copy() for s in self.segments]

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

    def copy(self) -> "Model":
        m = Model.from_config(self.get_config())
        m.seed = self.seed
        m.build()
        m.set_flat_weights(self.get_flat_weights().detach().cpu())
        m.loss, m.optimizer, m.metrics, m.loss_weights = self.loss, self.optimizer, list(self.metrics), self.loss_weights
        return m

    def to(self, device) -> "Model":
        self.build()
        self.flat = self.flat.to(device)
        self._replica = None
        return self

    def summary(self) -> str:
        self.build()
        lines = [f"Model: {self.name} ({len(self.inputs)} inputs, {len(self.outputs)} outputs)"]
        for li, n in enumerate(self.nodes):
            lines.append(f"  {li:2d} {n.layer.class_name:<20s} out={n.shape} params={sum(s.size for s in self.segments if s.layer_index == li)}")
        lines.append(f"Total params: {self.count_params()}")
        print("\n".join(lines))
        return "\n".join(lines)

    # -- autograd executor ------------------------------------------------------------------
    def forward(self, x, flat: Optional[torch.Tensor] = None, training: bool = False, logits: bool = False,
                ctx: Optional[dict] = None):
        """``x``: one tensor (single-input models) or a list, one per input.  Returns one tensor or a list."""
        self.build()
        flat = self.flat if flat is None else flat
        own_ctx = ctx is None
        ctx = {} if ctx is None else ctx
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        if len(xs) != len(self.inputs):
            raise ValueError(f"the model has {len(self.inputs)} inputs, got {len(xs)} arrays")
        vals = {id(t): prepare_input(v, t.shape, flat.device) for t, v in zip(self.inputs, xs)}
        tails = {id(l) for l in self.logits_tails() if l is not None} if logits else set()
        for li, n in enumerate(self.nodes):
            ctx["logits_tail"] = n.layer if id(n.layer) in tails else None
            ins = [vals[id(p)] for p in n.parents]
            if isinstance(n.layer, Merge):
                vals[id(n)] = n.layer.forward_many(ins)
            else:
                vals[id(n)] = n.layer.forward(ins[0], self._layer_views(flat, li), training, ctx)
        ctx["logits_tail"] = None
        if own_ctx:
            apply_deferred(ctx)
        outs = [vals[id(o)] for o in self.outputs]
        return outs[0] if len(outs) == 1 else outs

    def compile(self, loss="categorical_crossentropy", optimizer="sgd", metrics: Sequence[str] = ("accuracy",),
                loss_weights=None) -> None:
        self.loss, self.optimizer, self.metrics, self.loss_weights = loss, optimizer, list(metrics or []), loss_weights
        self._replica = None

    def _get_replica(self):
        if self._replica is None:
            from ..parallel.replica import TorchReplica

            if self.loss is None:
                raise RuntimeError("call compile() first")
            self._replica = TorchReplica(self, self.optimizer, self.loss, device=self.get_flat_weights().device,
                                         share_model_buffer=True, loss_weights=self.loss_weights, metrics=self.metrics)
        return self._replica

    def train_on_batch(self, x, y) -> List[float]:
        return [float(v) for v in self._get_replica().train_on_batch(x, y)]

    def predict(self, x, batch_size: int = 8192):
        self.build()
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        xs = [torch.as_tensor(np.asarray(a)) if not isinstance(a, torch.Tensor) else a for a in xs]
        chunks: List[List[torch.Tensor]] = [[] for _ in self.outputs]
        with torch.no_grad():
            for i in range(0, xs[0].shape[0], batch_size):
                out = self.forward([a[i:i + batch_size] for a in xs], training=False)
                for k, o in enumerate(out if isinstance(out, list) else [out]):
                    chunks[k].append(o.float().cpu())
        res = [torch.cat(c, dim=0).numpy() for c in chunks]
        return res[0] if len(res) == 1 else res

    def save_weights(self, path: str) -> None:
        torch.save({"flat": self.get_flat_weights().detach().cpu()}, path)

    def load_weights(self, path: str) -> None:
        self.set_flat_weights(torch.load(path, weights_only=True)["flat"])

    def save(self, path: str) -> None:
        torch.save({"model": self.to_json(), "flat": self.get_flat_weights().detach().cpu(), "loss": self.loss
                    if isinstance(self.loss, (str, type(None))) else None, "optimizer": None}, path)
