"""Native sm_100a executor: planner + replica.

``NativeReplica`` lowers a :class:`~distkeras_b200.models.core.Sequential` (Dense / Conv2D /
MaxPooling2D / Flatten / Dropout / Activation stacks) with its loss and worker optimizer into flat
op lists for the C++ executor (``csrc/engine.cu``).  One training step is

    input stage (cast + MinMax affine)  ->  tcgen05 GEMM per layer (bias / ReLU / dropout fused)
    -> fused softmax-xent (loss, accuracy, dZ)  ->  per layer: bias colsum, wgrad GEMM, dgrad GEMM
    (dReLU / dropout mask fused)  ->  fused flat optimizer (+ bf16 shadow)

All GEMMs are the hand-written tcgen05/TMEM/TMA kernel (``csrc/gemm_tcgen05.cu``); backward uses
its MN-major operand mode, so no tensor is ever transposed in memory: dgrad reads the bf16 weight
shadow ``[out, in]`` as an MN-major B operand and wgrad reads ``dZ [B, out]`` and the layer input
``[B, in]`` as MN-major A / B operands.  The reference executes the same math through
``keras_model.train_on_batch`` (``distkeras/workers.py:199-202, 327-342``).

The replica owns flat buffers ``W`` (fp32 master = the ``get_weights()`` analogue), ``G``, ``Wb``
(bf16 shadow), optimizer state and ``W1`` (last pulled center), which is exactly what the
parameter-server kernels operate on.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _native as N
from ..models.core import (Activation, BatchNormalization, Conv2D, Dense, Dropout, Flatten, GlobalAveragePooling2D,
                           MaxPooling2D, ResidualBlock, Sequential)
from ..ops.flat_optim import FlatOptimizer
from .replica import Replica

SLOT_X, SLOT_Y, SLOT_XB = 0, 1, 2


def _r8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


class UnsupportedByNativeEngine(Exception):
    """The model uses a layer / loss the native engine does not lower (autograd path is used)."""


class _Block:
    """One lowered layer group (a node of the native program)."""

    def __init__(self, kind: str):
        self.kind = kind  # dense | conv | pool | flatten | bn | gap | res
        self.act: Optional[str] = None
        self.drop_p = 0.0
        self.use_bias = True


def _conv_block(layer_index, layer: Conv2D, in_shape, out_shape) -> _Block:
    b = _Block("conv")
    b.layer_index = layer_index
    b.seg_prefix = ""
    b.in_shape, b.out_shape = tuple(in_shape), tuple(out_shape)
    b.kh, b.kw = layer.kernel_size
    b.stride, b.pad = layer.strides[0], layer.pad_amount()
    b.k_in = b.kh * b.kw * int(in_shape[-1])
    b.n_out = layer.filters
    b.act = layer.activation or "linear"
    b.use_bias = layer.use_bias
    if b.n_out % 8 != 0:
        raise UnsupportedByNativeEngine("Conv2D filters must be a multiple of 8")
    return b


def _bn_block(layer_index, layer: BatchNormalization, shape, prefix="") -> _Block:
    b = _Block("bn")
    b.layer_index, b.seg_prefix = layer_index, prefix
    b.shape = tuple(shape)
    b.channels = int(shape[-1])
    b.eps, b.momentum = layer.epsilon, layer.momentum
    b.act = "linear"
    if b.channels % 8 != 0:
        raise UnsupportedByNativeEngine("BatchNormalization channels must be a multiple of 8")
    return b


def _group_layers(model: Sequential) -> List[_Block]:
    """Fold Activation / Dropout layers into the preceding block; expand residual blocks."""
    blocks: List[_Block] = []
    if not isinstance(model, Sequential):
        raise UnsupportedByNativeEngine("functional (multi-input / multi-output) model")
    model.build()
    for li, layer in enumerate(model.layers):
        in_shape, out_shape = model.shapes[li], model.shapes[li + 1]
        if isinstance(layer, Dense):
            if len(in_shape) != 1:
                raise UnsupportedByNativeEngine("Dense on non-flat input")
            b = _Block("dense")
            b.layer_index, b.k_in, b.n_out, b.seg_prefix = li, int(in_shape[0]), layer.units, ""
            b.act = layer.activation or "linear"
            b.use_bias = layer.use_bias
            blocks.append(b)
        elif isinstance(layer, Conv2D):
            blocks.append(_conv_block(li, layer, in_shape, out_shape))
        elif isinstance(layer, MaxPooling2D):
            if layer.pool_size[0] != layer.pool_size[1] or layer.strides != layer.pool_size:
                raise UnsupportedByNativeEngine("only square non-overlapping max-pooling")
            b = _Block("pool")
            b.in_shape, b.out_shape, b.k = tuple(in_shape), tuple(out_shape), layer.pool_size[0]
            blocks.append(b)
        elif isinstance(layer, Flatten):
            blocks.append(_Block("flatten"))
        elif isinstance(layer, BatchNormalization):
            if len(in_shape) != 3:
                raise UnsupportedByNativeEngine("BatchNormalization on non-image input")
            blocks.append(_bn_block(li, layer, in_shape))
        elif isinstance(layer, GlobalAveragePooling2D):
            b = _Block("gap")
            b.in_shape = tuple(in_shape)
            blocks.append(b)
        elif isinstance(layer, ResidualBlock):
            b = _Block("res")
            b.layer_index = li
            b.in_shape, b.out_shape = tuple(in_shape), tuple(out_shape)
            mid = layer.conv1.output_shape(in_shape)
            b.conv1 = _conv_block(li, layer.conv1, in_shape, mid)
            b.conv1.seg_prefix = "conv1."
            b.bn1 = _bn_block(li, layer.bn1, mid, "bn1.")
            b.bn1.act = "relu"
            b.conv2 = _conv_block(li, layer.conv2, mid, mid)
            b.conv2.seg_prefix = "conv2."
            b.bn2 = _bn_block(li, layer.bn2, mid, "bn2.")
            b.proj = None
            if layer._needs_proj(in_shape):
                b.proj = _conv_block(li, layer.proj, in_shape, mid)
                b.proj.seg_prefix = "proj."
                b.bnp = _bn_block(li, layer.bnp, mid, "bnp.")
            blocks.append(b)
        elif isinstance(layer, Activation):
            if not blocks or blocks[-1].kind not in ("dense", "conv", "bn") or blocks[-1].act not in (None, "linear"):
                raise UnsupportedByNativeEngine("free-standing Activation")
            blocks[-1].act = layer.activation
        elif isinstance(layer, Dropout):
            if not blocks or blocks[-1].kind != "dense" or blocks[-1].drop_p:
                raise UnsupportedByNativeEngine("Dropout must follow a Dense block")
            blocks[-1].drop_p = layer.rate
        else:
            raise UnsupportedByNativeEngine(f"layer {layer.class_name}")
    for b in blocks[:-1]:
        if b.kind in ("dense", "conv", "bn") and b.act not in ("relu", "linear"):
            raise UnsupportedByNativeEngine(f"hidden activation {b.act!r}")
    last = blocks[-1]
    if last.kind != "dense" or last.act not in ("softmax", "linear", "sigmoid"):
        raise UnsupportedByNativeEngine("the model must end in a Dense softmax / linear / sigmoid head")
    if last.drop_p:
        raise UnsupportedByNativeEngine("dropout on the output layer")
    if blocks[0].kind not in ("dense", "conv"):
        raise UnsupportedByNativeEngine("the model must start with a Dense or Conv2D layer")
    return blocks


def _relu_mask_fusable(prev, rows: int, cols: int) -> bool:
    """True when ``prev`` (the block whose output gradient is being produced) ends in a plain ReLU
    whose saved output has exactly the gradient's layout, so the consumer can apply ``out > 0``."""
    if prev is None or prev.kind not in ("conv", "bn") or prev.act != "relu" or getattr(prev, "drop_p", 0) > 0:
        return False
    rec = getattr(prev, "out_rec", None)
    return rec is not None and rec["rows"] == rows and rec["cols"] == cols and rec["ld"] == cols


class NativeReplica(Replica):
    """A model replica executed by the native sm_100a engine on one GPU."""

    def __init__(self, model: Sequential, optimizer, loss: str, batch_size: int, device_index: int = 0,
                 in_dtype: str = "u8", input_affine: Tuple[float, float] = (1.0, 0.0), hist_slots: int = 1024,
                 dense_labels: bool = False, seed: int = 1234, training: bool = True,
                 pull_center_ptr: int = 0, region_steps: int = 1, comm_spec: Optional[dict] = None,
                 compact: Optional[bool] = None):
        self.lib = N.lib()
        model.build()
        self.model = model
        self.loss = loss
        self.B = int(batch_size)
        self.device = torch.device("cuda", device_index)
        self.device_index = device_index
        torch.cuda.set_device(self.device)
        self.in_dtype = {"u8": N.IN_U8, "f32": N.IN_F32, "bf16": N.IN_BF16}[in_dtype]
        self.in_torch_dtype = {"u8": torch.uint8, "f32": torch.float32, "bf16": torch.bfloat16}[in_dtype]
        self.scale, self.shift = float(input_affine[0]), float(input_affine[1])
        self.hist_slots = int(hist_slots)
        self.dense_labels = bool(dense_labels)
        self.seed = int(seed)
        self.blocks = _group_layers(model)
        self.region_steps = max(1, int(region_steps))
        self.comm_spec = dict(comm_spec) if comm_spec else None
        self._compact_request = compact
        head = self.blocks[-1]
        self.num_classes = head.n_out
        if loss in ("categorical_crossentropy", "sparse_categorical_crossentropy"):
            if head.act != "softmax":
                raise UnsupportedByNativeEngine("cross-entropy needs a softmax head")
            self.loss_kind = "xent"
        elif loss in ("mse", "mean_squared_error") and head.act == "linear":
            self.loss_kind = "mse"
            self.dense_labels = True
        else:
            raise UnsupportedByNativeEngine(f"loss {loss!r} with head {head.act!r}")

        dev = self.device
        P = model.num_params
        self.P = P
        self.W = model.get_flat_weights().detach().to(dev, torch.float32).clone().contiguous()
        self.Wb = torch.zeros(P, dtype=torch.bfloat16, device=dev)
        self.hist = torch.zeros(self.hist_slots, 2, dtype=torch.float32, device=dev)
        self.training = training
        self.pull_center_ptr = int(pull_center_ptr) if training else 0
        # compact (small-batch) program: region-level input staging, narrow forward tiles, and ONE fused
        # kernel for every weight gradient + bias gradient + optimizer update (+ the PS exchange)
        self.compact = self._compact_eligible()
        # device step counter: in the classic program the per-step input stage increments it before the
        # step's kernels read it; in the compact program it holds the 1-based index of the step being
        # executed and the fused backward-update kernel advances it when the step is complete
        self.step_base = 1 if self.compact else 0
        self.step_counter = torch.full((1,), self.step_base, dtype=torch.int32, device=dev)
        self._done_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.W1 = None
        if training:
            self.G = None if self.compact else torch.zeros(P, dtype=torch.float32, device=dev)
            self.opt = FlatOptimizer(optimizer, P, dev)
            if self.pull_center_ptr or self.comm_spec:
                self.W1 = self.W.clone()
        self._keep: list = []  # buffers referenced only by raw pointers
        # DK_TRACE=1: every GEMM / fused-update op of the training lists stamps the SM clock of its first CTA
        # at fixed points into a row of this buffer (tools/kernel_timeline.py)
        self._trace = torch.zeros(64, 16, dtype=torch.int64, device=dev) if os.environ.get("DK_TRACE") == "1" else None
        self._trace_names: list = []
        self.engine = self.lib.dk_engine_create()
        # L_step: training forward; L_bwd: loss + backward + optimizer; L_fwd: inference forward;
        # L_step_pull: training forward whose first GEMM pulls its weights from the PS (fused pull)
        self.L_step = self.lib.dk_engine_new_list(self.engine) if training else -1
        self.L_bwd = self.lib.dk_engine_new_list(self.engine) if training else -1
        self.L_fwd = self.lib.dk_engine_new_list(self.engine)
        self.L_bwd_comm = -1      # compact program: backward list whose fused update also exchanges with the PS
        self.L_in_region = self.L_in_step = -1
        self.L_step_pull = -1
        self.pull_segment = None
        self._input_feats = int(np.prod(model.input_shape))
        self._lower()
        self.refresh_shadow()
        self.iteration = 0
        # eager single-step staging (train_on_batch / predict convenience API)
        self._x_stage = torch.zeros(self.B, self._input_feats, dtype=self.in_torch_dtype, device=dev)
        if self.dense_labels:
            self._y_stage = torch.zeros(self.B, self.num_classes, dtype=torch.float32, device=dev)
        else:
            self._y_stage = torch.zeros(self.B, dtype=torch.int32, device=dev)

    def _compact_eligible(self) -> bool:
        """The compact program covers what the reference actually trains at its batch sizes: stacks of
        Dense (+ReLU / Dropout) layers ending in a softmax head with <= 16 classes, cross-entropy."""
        if self._compact_request is False or os.environ.get("DK_COMPACT", "1") == "0" or not self.training:
            return False
        if self.pull_center_ptr or self.loss_kind != "xent" or os.environ.get("DK_FUSED_HEAD", "1") == "0":
            return False
        limit = int(os.environ.get("DK_COMPACT_MAX_BATCH", "256"))
        if self.B > limit and self._compact_request is not True:
            return False
        blocks = self.blocks
        if any(b.kind != "dense" for b in blocks) or not (1 <= len(blocks) <= N.BWD_MAX_LAYERS):
            return False
        # hidden widths: multiples of 4 (fp32 rows stay 16-byte aligned); widths that are not multiples of 8 (the Higgs
        # MLP's 500) read their weights through the 8-padded bf16 shadow, which the fused update kernel refreshes itself
        for b in blocks[:-1]:
            if b.n_out % 4 != 0:
                return False
        head = blocks[-1]
        return (head.n_out <= 16 and head.act == "softmax" and head.k_in % 4 == 0
                and head.n_out * (_r8(head.k_in) + 4) * 4 <= 48 * 1024)

    # ------------------------------------------------------------------------------------------
    # op-list helpers
    # ------------------------------------------------------------------------------------------
    def _buf(self, *shape, dtype=torch.bfloat16) -> torch.Tensor:
        t = torch.zeros(*shape, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def _add(self, lst: int, kind: int, iargs, fargs=()) -> None:
        ia = (C.c_int64 * len(iargs))(*[int(v) for v in iargs])
        fa = (C.c_double * max(1, len(fargs)))(*[float(v) for v in fargs]) if fargs else (C.c_double * 1)(0.0)
        r = self.lib.dk_engine_add_op(self.engine, lst, kind, ia, len(iargs), fa, len(fargs))
        if r < 0:
            raise RuntimeError(f"dk_engine_add_op({kind}) failed: {r}")

    def _trace_row(self, name: str) -> int:
        """Device address of the next trace row (0 when tracing is off)."""
        if self._trace is None or len(self._trace_names) >= self._trace.shape[0]:
            return 0
        self._trace_names.append(name)
        return self._trace.data_ptr() + 128 * (len(self._trace_names) - 1)

    def _gemm(self, lst: int, A: int, lda: int, Bp: int, ldb: int, M: int, Nn: int, K: int, flags: int,
              ep: N.GemmEpilogue, bn: int = 0, splits: int = 0) -> None:
        ep.trace = self._trace_row(f"list{lst} gemm M={M} N={Nn} K={K} flags={flags} bn={bn}") or None
        r = self.lib.dk_engine_add_gemm(self.engine, lst, C.c_void_p(A), lda, C.c_void_p(Bp), ldb, M, Nn, K,
                                        flags, bn, splits, C.byref(ep))
        if r < 0:
            raise RuntimeError(f"dk_engine_add_gemm(M={M}, N={Nn}, K={K}, flags={flags}) failed: {r}")

    def _seg(self, layer_index: int, name: str):
        for s in self.model.segments:
            if s.layer_index == layer_index and s.name == name:
                return s
        return None

    # ------------------------------------------------------------------------------------------
    # lowering
    # ------------------------------------------------------------------------------------------
    def _lower(self) -> None:
        """Lower the block list into the forward lists (training / inference / fused-pull) and the
        backward list.  Every block contributes a forward emission and a backward closure."""
        B = self.B
        F = self._input_feats
        in_shape = tuple(self.model.input_shape)
        if self.compact:
            # the whole region's mini-batches are staged (cast + MinMax affine) by ONE launch; step j reads
            # rows [j B, (j + 1) B) through slot SLOT_XB (tensor maps re-encoded when the list is enqueued)
            x0 = self._buf(self.region_steps * B, _r8(F))
            self._xb_region = x0
            self.L_in_region = self.lib.dk_engine_new_list(self.engine)
            self.L_in_step = self.lib.dk_engine_new_list(self.engine)
            for lst, rows in ((self.L_in_region, self.region_steps * B), (self.L_in_step, B)):
                self._add(lst, N.OP_INPUT, [-(SLOT_X + 1), self.in_dtype, rows, F, x0.data_ptr(), _r8(F), 0, 0, 0, 0, F],
                          [self.scale, self.shift])
        else:
            x0 = self._buf(B, _r8(F))
        cur = dict(t=x0, rows=B, cols=F, ld=_r8(F), nhwc=in_shape if len(in_shape) == 3 else None,
                   slot=SLOT_XB if self.compact else None)
        first = self.blocks[0]
        # First convolution with 1 / 3 input channels (MNIST, CIFAR): the input stage writes the image with its
        # channels zero-padded to 32, so this layer too runs on the TMA-im2col kernels (forward + weight gradient)
        # and no column matrix exists anywhere in the network.
        self._cpad = 0
        if (first.kind == "conv" and os.environ.get("DK_IMPLICIT_CONV", "auto") != "0" and first.in_shape[-1] < 32
                and first.kh == first.kw and first.n_out % 8 == 0 and first.n_out <= 128 and len(self.blocks) > 1
                and (os.environ.get("DK_PAD_INPUT_CHANNELS", "auto") == "1"      # "1": always; "auto": not for 1-channel
                     or (os.environ.get("DK_PAD_INPUT_CHANNELS", "auto") == "auto"   # images (32x the bytes: the explicit
                         and first.in_shape[-1] >= 3))):                              # K = 9 path measured faster there)
            H0, W0, C0 = first.in_shape
            self._cpad = 32
            x0 = self._buf(B * H0 * W0, 32)
            cur = dict(t=x0, rows=B * H0 * W0, cols=32, ld=32, nhwc=(H0, W0, 32), slot=None, cpad=32)
        x0f = None
        if self.pull_center_ptr and first.kind == "dense" and first.k_in % 8 == 0 and F % 8 == 0:
            self.L_step_pull = self.lib.dk_engine_new_list(self.engine)
            x0f = self._buf(B, F, dtype=torch.float32)
        self._x0f = x0f
        self._lists = [l for l in (self.L_step, self.L_fwd, self.L_step_pull) if l >= 0]
        self._train_lists = (self.L_step, self.L_step_pull)
        self._pad_refresh: list = []
        # BatchNorm reduction scratch: zeroed once per step, sliced per BN (forward + backward sums)
        nbn = sum(4 * c for c in self._bn_channels())
        self._bn_scratch = self._buf(max(nbn, 1), dtype=torch.float32)
        self._bn_used = 0
        for lst in ([] if self.compact else self._lists):
            if nbn:
                self._add(lst, N.OP_MEMSET, [self._bn_scratch.data_ptr(), 0, nbn * 4])
            if self._cpad:   # rows = pixels, 3 (or 1) valid channels of 32
                H0, W0, C0 = first.in_shape
                self._add(lst, N.OP_INPUT,
                          [-(SLOT_X + 1), self.in_dtype, B * H0 * W0, C0, x0.data_ptr(), 32, 0, 0,
                           self.step_counter.data_ptr() if lst in self._train_lists else 0, 0, C0], [self.scale, self.shift])
                continue
            self._add(lst, N.OP_INPUT,
                      [-(SLOT_X + 1), self.in_dtype, B, F, x0.data_ptr(), cur["ld"], 0, 0,
                       self.step_counter.data_ptr() if lst in self._train_lists else 0,
                       x0f.data_ptr() if lst == self.L_step_pull else 0, F],
                      [self.scale, self.shift])
        nblocks = len(self.blocks)
        self._forked = set()
        self._side_streams = int(os.environ.get("DK_SIDE_STREAMS", "2"))
        backward = []  # (block, closure(grad, premasked, need_dx, prev_block) -> (grad_in, premasked_in))
        for bi, b in enumerate(self.blocks):
            is_last = bi == nblocks - 1
            if b.kind in ("dense", "conv"):
                cur, bw = self._emit_matmul(b, cur, bi, is_last)
            elif b.kind == "pool":
                cur, bw = self._emit_pool(b, cur)
            elif b.kind == "flatten":
                cur, bw = self._emit_flatten(b, cur)
            elif b.kind == "bn":
                cur, bw = self._emit_bn(b, cur)
            elif b.kind == "gap":
                cur, bw = self._emit_gap(b, cur)
            elif b.kind == "res":
                cur, bw = self._emit_res(b, cur, bi)
            else:  # pragma: no cover
                raise UnsupportedByNativeEngine(b.kind)
            b.out = cur
            backward.append((b, bw))
        self.logits = cur["t"]
        Cn = self.num_classes
        ldl = cur["ld"]
        self.probs = self._buf(B, Cn, dtype=torch.float32)
        self._zero_labels = self._buf(B, dtype=torch.int32)
        self._add(self.L_fwd, N.OP_XENT, [self.logits.data_ptr(), ldl, self._zero_labels.data_ptr(), 0, B, Cn, 0, 0,
                                          0, 0, self.probs.data_ptr(), 0, 0, 0])
        self._prepend_pad_refresh(self.L_fwd)
        if not self.training:
            return
        ldz = _r8(Cn)
        targets = [self.L_bwd]
        if self.compact and self.comm_spec:
            self.L_bwd_comm = self.lib.dk_engine_new_list(self.engine)
            targets.append(self.L_bwd_comm)
        for lst in targets:
            self._bwd_list = lst
            self._bwd_layers = []
            # ---- loss ----
            dz = self._buf(B, ldz)
            if self.loss_kind == "xent" and getattr(self, "head_fused", False):
                pass  # loss + dZ come out of the fused head op emitted by the last block's backward
            elif self.loss_kind == "xent":
                self._add(lst, N.OP_XENT, [self.logits.data_ptr(), ldl,
                                           0 if self.dense_labels else -(SLOT_Y + 1),
                                           -(SLOT_Y + 1) if self.dense_labels else 0,
                                           B, Cn, dz.data_ptr(), ldz, 0, 0, 0, self.hist.data_ptr(),
                                           self.step_counter.data_ptr(), self.hist_slots])
            else:
                self._add(lst, N.OP_ELOSS, [N.LOSS_MSE, self.logits.data_ptr(), -(SLOT_Y + 1), B, Cn, dz.data_ptr(), ldz,
                                            0, 0, self.hist.data_ptr(), self.step_counter.data_ptr(), self.hist_slots])
            if not self.compact:
                self._add(lst, N.OP_MEMSET, [self.G.data_ptr(), 0, self.P * 4])
            # ---- backward ----
            grad = dict(t=dz, rows=B, cols=Cn, ld=ldz)
            premasked = True  # dZ of the head is already w.r.t. the logits
            for bi in range(nblocks - 1, -1, -1):
                b, bw = backward[bi]
                prev = self.blocks[bi - 1] if bi > 0 else None
                grad, premasked = bw(grad, premasked, bi > 0, prev)
                if bi == 0:
                    break
            for sid in sorted(self._forked):
                self._add(lst, N.OP_JOIN, [sid])
            o = self.opt
            if self.compact:
                # ---- every weight gradient, bias gradient and the optimizer rule: ONE kernel ----
                self._emit_bwd_update(lst, with_comm=lst == self.L_bwd_comm)
            else:
                # ---- optimizer (one fused launch over the flat buffer, emits the bf16 shadow) ----
                self._add(lst, N.OP_OPTIM, [N.OPT_KINDS[o.kernel_kind], self.W.data_ptr(), self.G.data_ptr(), N.ptr(o.s0),
                                            N.ptr(o.s1), self.Wb.data_ptr(), self.P, int(o.nesterov),
                                            self.step_counter.data_ptr()],
                          [o.lr, o.p0, o.p1, o.eps, o.decay, 1.0])
            self._prepend_pad_refresh(lst)

    def _emit_bwd_update(self, lst: int, with_comm: bool) -> None:
        o = self.opt
        d = N.BwdUpdateDesc()
        layers = self._bwd_layers
        d.nlayers, d.batch = len(layers), self.B
        for i, l in enumerate(layers):
            ld = d.layer[i]
            ld.dz, ld.lddz, ld.x, ld.ldx, ld.x_slot = l["dz"], l["lddz"], l["x"], l["ldx"], l["x_slot"]
            ld.n_out, ld.k_in, ld.w_off, ld.b_off = l["n_out"], l["k_in"], l["w_off"], l["b_off"]
            ld.wb_pad, ld.ldwb_pad = l["wb_pad"], l["ldwb_pad"]
        d.w, d.s0, d.s1, d.wb = self.W.data_ptr(), N.ptr(o.s0), N.ptr(o.s1), self.Wb.data_ptr()
        d.w1 = N.ptr(self.W1)
        d.opt_kind = N.OPT_KINDS[o.kernel_kind]
        d.lr, d.p0, d.p1, d.eps, d.decay, d.nesterov = o.lr, o.p0, o.p1, o.eps, o.decay, int(o.nesterov)
        d.step, d.done_counter, d.step_inc = self.step_counter.data_ptr(), self._done_counter.data_ptr(), 1
        d.comm_mode, d.nshards, d.shard_per = N.COMM_NONE, 1, 1
        if with_comm:
            c = self.comm_spec
            d.comm_mode = c["mode"]
            d.comm_scale, d.alpha = float(c.get("scale", 1.0)), float(c.get("alpha", 0.0))
            d.scale_dev = c.get("scale_dev") or None
            shards = c["shards"]  # [(lo, hi, center_ptr)], equal-sized except the last
            if len(shards) > N.BWD_MAX_SHARDS:
                raise UnsupportedByNativeEngine("too many parameter-server shards for the fused exchange")
            d.nshards = len(shards)
            d.shard_per = max(1, shards[0][1] - shards[0][0])
            for i, (lo, hi, ptr) in enumerate(shards):
                if lo != i * d.shard_per:
                    raise UnsupportedByNativeEngine("parameter-server shards must be equal-sized and contiguous")
                d.shard_center[i] = ptr
            d.ctrl, d.worker, d.last_update = c["ctrl"], int(c["worker"]), c.get("last_update") or None
        d.trace = self._trace_row(f"list{lst} bwd_update comm={int(with_comm)}") or None
        r = self.lib.dk_engine_add_bwd_update(self.engine, lst, C.byref(d))
        if r < 0:
            raise RuntimeError(f"dk_engine_add_bwd_update failed: {r}")

    def _bn_channels(self) -> List[int]:
        out = []
        for b in self.blocks:
            if b.kind == "bn":
                out.append(b.channels)
            elif b.kind == "res":
                out += [b.bn1.channels, b.bn2.channels] + ([b.bnp.channels] if b.proj is not None else [])
        return out

    @staticmethod
    def _narrow_bn(M: int, Nn: int, floor: int) -> int:
        """Tile width for the compact program: with one or two M tiles the grid is N / bn CTAs, so the
        narrowest tile that keeps <= ~2 CTAs per SM spreads the weight stream over the most SMs."""
        m_tiles = (M + 127) // 128
        for bn in (16, 32, 64, 128):
            if bn >= floor and m_tiles * ((Nn + bn - 1) // bn) <= 296:
                return bn
        return 128

    def _plan_head_in_forward(self, b: "_Block", bi: int, rows: int, Nout: int):
        """Compact program: when block ``bi`` is the ReLU dense layer right below a softmax classifier the fused head
        kernel would handle, the head moves into the epilogue of THIS layer's forward GEMM (``DkGemmEpilogue.head_*``):
        one launch and one dependent-kernel latency less per step.  Returns the shared buffers (planned once) or None."""
        if getattr(self, "_fwd_head", None) is not None and self._fwd_head["bi"] == bi:
            return self._fwd_head
        if (not self.compact or not self.training or os.environ.get("DK_HEAD_IN_FWD", "1") == "0" or self.dense_labels
                or self.loss_kind != "xent" or bi + 2 != len(self.blocks) or b.kind != "dense" or b.act != "relu"
                or rows > 128 or Nout % 8 != 0 or (Nout + 15) // 16 > 32):
            return None
        nb = self.blocks[bi + 1]
        C_, Kh = nb.n_out, nb.k_in
        if (nb.kind != "dense" or nb.act == "relu" or C_ > 16 or Kh != Nout or Kh % 8 != 0 or C_ * (Kh + 4) * 4 > 48 * 1024
                or os.environ.get("DK_FUSED_HEAD", "1") == "0" or getattr(nb, "drop_p", 0) > 0):
            return None
        kseg = self._seg(nb.layer_index, nb.seg_prefix + "kernel")
        bseg = self._seg(nb.layer_index, nb.seg_prefix + "bias") if nb.use_bias else None
        ldz = _r8(C_)
        dev = self.device
        self._fwd_head = dict(
            bi=bi, w=self.Wb.data_ptr() + 2 * kseg.offset, ldw=Kh, C=C_,
            bias=(self.W.data_ptr() + 4 * bseg.offset) if bseg is not None else 0,
            acc=torch.zeros(2, 128, 16, dtype=torch.float32, device=dev), sync=torch.zeros(2, dtype=torch.int32, device=dev),
            dz=self._buf(rows, ldz), ldz=ldz, dh=self._buf(rows, Nout))
        return self._fwd_head

    @staticmethod
    def _pick_kch(K: int, a_rows: int, bn: int = 16, pool: int = 12 * 18432) -> int:
        """k-blocks per TMA request for a short-M forward GEMM (``DK_GEMM_KCH``): groups of ``kch`` whole 64-wide
        chunks go out as one 3-D request per operand, the K tail as 2-D requests; pick the ``kch`` with the fewest
        requests whose stage fits the kernel's shared-memory pool at least twice (0 = keep one k-block per request)."""
        full, total = K // 64, -(-K // 64)
        per = a_rows * 128 + bn * 128
        best, best_req = 0, 2 * total
        # requests up to ~16 KB cost the same TMA service time; bigger ones delay the first MMA (nothing of a request can
        # be consumed before all of it has landed)
        cap = max(2, 16384 // (max(a_rows, bn) * 128))
        for kch in range(2, min(15, full, cap, pool // (2 * per)) + 1):
            req = 2 * (full // kch) + 2 * (total - (full // kch) * kch)
            if req < best_req:
                best, best_req = kch, req
        return best

    def _bn_slice(self, floats: int) -> int:
        ptr = self._bn_scratch.data_ptr() + 4 * self._bn_used
        self._bn_used += floats
        return ptr

    # -- Dense / Conv2D -----------------------------------------------------------------------------
    def _emit_matmul(self, b: _Block, cur: dict, bi: int, is_last: bool):
        B, lists = self.B, self._lists
        wptr_f32, wb_ptr, g_ptr = self.W.data_ptr(), self.Wb.data_ptr(), (N.ptr(self.G) if self.training else 0)
        kseg = self._seg(b.layer_index, b.seg_prefix + "kernel")
        bseg = self._seg(b.layer_index, b.seg_prefix + "bias") if b.use_bias else None
        K, Nout = b.k_in, b.n_out
        cpad = cur.get("cpad", 0) if b.kind == "conv" else 0     # channels physically stored per input pixel
        if cpad:
            # weights [Nout, kh, kw, Cin] scattered into a zero-padded [Nout, kh, kw, cpad] shadow before every step
            taps, c_real = b.kh * b.kw, b.in_shape[-1]
            K = taps * cpad
            wpad = self._buf(Nout, K)
            wbp, wbld = wpad.data_ptr(), K
            self._pad_refresh.append((wpad.data_ptr(), cpad * 2, wb_ptr + 2 * kseg.offset, c_real * 2, c_real * 2, Nout * taps))
        elif K % 8 == 0:  # bf16 weight shadow [Nout, K] with a TMA-legal leading dimension
            wbp, wbld = wb_ptr + 2 * kseg.offset, K
        else:
            pad = self._buf(Nout, _r8(K))
            wbp, wbld = pad.data_ptr(), _r8(K)
            self._pad_refresh.append((pad.data_ptr(), _r8(K) * 2, wb_ptr + 2 * kseg.offset, K * 2, K * 2, Nout))
        inp = cur
        implicit = False
        cin_eff = c_real = 0
        if b.kind == "conv":
            H, Wd, Cin = b.in_shape
            c_real = Cin
            if cpad:
                Cin = cpad
            cin_eff = Cin
            OH, OW, _ = b.out_shape
            rows = B * OH * OW
            # implicit GEMM (DK_IMPLICIT_CONV=1): the forward and dgrad GEMMs gather their A operand from
            # the NHWC activation inside the kernel; the column matrix is then only needed by the wgrad
            # GEMM, so im2col moves off the critical path onto the wgrad branch of the backward list
            # default: layers the TMA-im2col kernel can feed (Cin a multiple of 64, or exactly 32); DK_IMPLICIT_CONV=1
            # also sends the other Cin % 8 == 0 layers to the thread-gather kernel, DK_IMPLICIT_CONV=0 disables both
            mode = os.environ.get("DK_IMPLICIT_CONV", "auto")
            implicit = (mode != "0" and (mode == "1" or Cin % 64 == 0 or Cin == 32) and Cin % 8 == 0 and Nout % 8 == 0
                        and cur["ld"] == Cin and wbld == K and b.kh == b.kw and not is_last)
            # weight gradient of an implicit layer: "tma" (default when the geometry allows: im2col operand produced
            # by TMA, no column matrix, bias gradient from the ones-tile MMA), "gather" (DK_IMPLICIT_WGRAD=1: the
            # cp.async gather kernel) or "explicit" (column matrix built on a side branch for a plain GEMM)
            wgrad_mode = "explicit"
            if implicit and os.environ.get("DK_IMPLICIT_WGRAD", "auto") == "1" and Nout <= 128:
                wgrad_mode = "gather"
            elif (implicit and os.environ.get("DK_IMPLICIT_WGRAD", "auto") != "0" and _r8(Nout) == Nout
                  and self.lib.dk_conv_wgrad_tma_supported(Cin, Nout, Nout, K)):
                wgrad_mode = "tma"
            implicit_wgrad = wgrad_mode != "explicit"
            need_col = (not implicit) or (self.training and not implicit_wgrad)
            col = self._buf(rows, _r8(K)) if need_col else None
            a_in = dict(t=col, rows=rows, cols=K, ld=_r8(K), nhwc=None)
            im2col_args = [cur["t"].data_ptr(), B, H, Wd, Cin, b.kh, b.kw, b.stride, b.pad, OH, OW,
                           col.data_ptr() if need_col else 0, _r8(K)]
            for lst in lists:
                if not implicit:
                    self._add(lst, N.OP_IM2COL, im2col_args)
                elif lst in self._train_lists and not implicit_wgrad:
                    # only the wgrad GEMM (backward list) reads the column matrix: build it on side branch 3
                    # while the forward chain continues; the backward list joins that branch before the wgrad
                    self._add(lst, N.OP_FORK, [3])
                    self.lib.dk_engine_set_build_stream(self.engine, 3)
                    self._add(lst, N.OP_IM2COL, im2col_args)
                    self.lib.dk_engine_set_build_stream(self.engine, 0)
                    self._forked.add(3)
        else:
            rows = cur["rows"]
            a_in = cur
        # classifier head (<= 16 classes, softmax cross-entropy): in the training lists the forward
        # GEMM, the loss and the dgrad GEMM collapse into ONE kernel (dk_dense_softmax_head)
        # (a head input width that is not a multiple of 8 -- the Higgs MLP's 500 -- runs on the 8-padded width: the pad
        # columns of the activation buffer and of the padded weight shadow are zeros, so they add nothing)
        Kh = wbld
        head_fused = (is_last and self.training and b.kind == "dense" and self.loss_kind == "xent" and Nout <= 16
                      and Kh == _r8(K) and a_in["ld"] == Kh and Nout * (Kh + 4) * 4 <= 48 * 1024
                      and b.act != "relu" and os.environ.get("DK_FUSED_HEAD", "1") != "0")
        self.head_fused = head_fused if is_last else getattr(self, "head_fused", False)
        if is_last and getattr(self, "_fwd_head", None) is not None and not head_fused:
            raise RuntimeError("planner bug: head moved into the forward GEMM but the classifier is not fusable")
        if is_last:
            ldl = _r8(Nout) if self.loss_kind == "xent" else Nout  # fp32 rows 16-byte aligned for TMA
            out = self._buf(rows, ldl, dtype=torch.float32)
            rec = dict(t=out, rows=rows, cols=Nout, ld=ldl, nhwc=None)
        else:
            out = self._buf(rows, _r8(Nout))
            rec = dict(t=out, rows=rows, cols=Nout, ld=_r8(Nout), nhwc=b.out_shape if b.kind == "conv" else None)
        for lst in lists:
            if head_fused and lst in self._train_lists:
                continue  # logits are produced (and consumed) inside the fused head kernel
            ep = N.GemmEpilogue()
            ep.bias = (wptr_f32 + 4 * bseg.offset) if bseg is not None else None
            ep.act = 1 if b.act == "relu" else 0
            ep.d, ep.ldd, ep.d_fp32 = out.data_ptr(), rec["ld"], 1 if is_last else 0
            ep.alpha = 1.0
            if lst in self._train_lists and b.drop_p > 0:
                ep.drop_p = b.drop_p
                ep.drop_seed = (self.seed * 7919 + bi * 104729) & 0xFFFFFFFF
                ep.step = self.step_counter.data_ptr()
            if lst == self.L_step_pull and bi == 0:
                # pull fused into the first GEMM: B operand = this layer's block of the center
                # variable in the PS GPU's HBM; the kernel refreshes W / W1 / Wb on the way
                off = kseg.offset
                self.pull_segment = (off, kseg.size)
                r = self.lib.dk_engine_add_gemm_pull(
                    self.engine, lst, C.c_void_p(self._x0f.data_ptr()), K, C.c_void_p(self.pull_center_ptr + 4 * off),
                    K, rows, Nout, K, C.c_void_p(self.W.data_ptr() + 4 * off),
                    C.c_void_p(self.W1.data_ptr() + 4 * off), C.c_void_p(self.Wb.data_ptr() + 2 * off), C.byref(ep))
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_gemm_pull failed: {r}")
                continue
            if implicit:
                r = self.lib.dk_engine_add_conv_gemm(self.engine, lst, C.c_void_p(cur["t"].data_ptr()), H, Wd, Cin, OH, OW,
                                                     b.kh, b.kw, b.stride, b.pad, 1, C.c_void_p(wbp), wbld, rows, Nout, K,
                                                     C.byref(ep))
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_conv_gemm(fwd) failed: {r}")
                continue
            bn = self._narrow_bn(rows, Nout, 16) if self.compact else 0
            fl = N.GEMM_SHORT_A if (self.compact and rows < 128) else 0   # one short M tile: short TMA box
            # the CTAs of a forward GEMM all read the same activation tile: clusters of up to 8 neighbours load 1/8 of
            # each k-block and multicast it, so an SM ingests its weights plus an eighth of the activations
            if fl and bn == 16 and os.environ.get("DK_GEMM_MCAST", "0") == "1" and self.lib.dk_gemm_mcast_cluster(rows) > 1:
                fl |= N.GEMM_MCAST_A
            elif fl and bn == 16 and os.environ.get("DK_SPLIT_M", "1") != "0":
                # the K loop of these GEMMs is bound by the bytes ONE SM can ingest (~50 GB/s): split the mini-batch
                # over 2 or 4 CTAs along M while the grid still fits the SMs (rows per CTA in bits 8..15 of the flags)
                n_tiles = (Nout + 15) // 16
                a_rows = _r8(rows)
                for split in (4, 2):
                    tr = rows // split
                    if rows % split == 0 and tr % 8 == 0 and n_tiles * split <= min(148, 64 if lst in self._train_lists and bi + 2 == len(self.blocks) else 148):
                        fl |= tr << 8
                        a_rows = tr
                        break
            if fl and bn == 16 and not (fl & N.GEMM_MCAST_A) and os.environ.get("DK_GEMM_KCH", "1") != "0":
                # a TMA request costs ~190 cycles whatever its size: load several 64-wide k-blocks per request (3-D view)
                fl |= self._pick_kch(K, a_rows if fl >> 8 else _r8(rows)) << 16
            if lst in self._train_lists and bn == 16 and a_in.get("slot") is None:
                fh = self._plan_head_in_forward(b, bi, rows, Nout)
                if fh is not None:
                    (ep.head_w, ep.head_ldw, ep.head_bias, ep.head_c) = fh["w"], fh["ldw"], fh["bias"], fh["C"]
                    ep.head_labels, ep.head_label_slot = 0, SLOT_Y
                    ep.head_acc, ep.head_sync = fh["acc"].data_ptr(), fh["sync"].data_ptr()
                    ep.head_dz, ep.head_ldz = fh["dz"].data_ptr(), fh["ldz"]
                    ep.head_dh, ep.head_lddh = fh["dh"].data_ptr(), Nout
                    ep.head_alpha = 1.0 / (1.0 - b.drop_p) if b.drop_p > 0 else 1.0
                    ep.head_hist, ep.head_step, ep.head_hist_slots = (self.hist.data_ptr(), self.step_counter.data_ptr(),
                                                                      self.hist_slots)
            if a_in.get("slot") is not None:
                ep.trace = self._trace_row(f"list{lst} gemm(slot) M={rows} N={Nout} K={K} bn={bn}") or None
                r = self.lib.dk_engine_add_gemm_slot(self.engine, lst, a_in["slot"], a_in["ld"], C.c_void_p(wbp), wbld,
                                                     rows, Nout, K, fl, bn, 1, C.byref(ep))
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_gemm_slot(M={rows}, N={Nout}, K={K}) failed: {r}")
                continue
            self._gemm(lst, a_in["t"].data_ptr(), a_in["ld"], wbp, wbld, rows, Nout, K, fl, ep, bn=bn,
                       splits=1 if bn else 0)
        b.out_rec = rec

        def backward(grad, premasked, need_dx, prev):
            lst = self._bwd_list
            head_din = None
            fh = getattr(self, "_fwd_head", None) if (head_fused and is_last) else None
            if fh is not None:
                # loss, dZ and dH were produced by the epilogue of the previous layer's forward GEMM
                grad = dict(t=fh["dz"], rows=rows, cols=Nout, ld=fh["ldz"])
                head_din = fh["dh"] if need_dx else None
                fuse_mask = True
            elif head_fused:
                fuse_mask = prev is not None and prev.kind == "dense" and prev.act == "relu"
                if prev is not None and prev.kind == "dense" and prev.drop_p > 0 and not fuse_mask:
                    raise UnsupportedByNativeEngine("dropout after a non-ReLU dense layer")
                alpha = 1.0 / (1.0 - prev.drop_p) if (fuse_mask and prev.drop_p > 0) else 1.0
                head_din = self._buf(rows, Kh) if need_dx else None
                self._add(lst, N.OP_HEAD,
                          [-(a_in["slot"] + 1) if a_in.get("slot") is not None else a_in["t"].data_ptr(), a_in["ld"], wbp,
                           wbld, (wptr_f32 + 4 * bseg.offset) if bseg is not None else 0,
                           0 if self.dense_labels else -(SLOT_Y + 1), -(SLOT_Y + 1) if self.dense_labels else 0,
                           rows, Nout, Kh, grad["t"].data_ptr(), grad["ld"], head_din.data_ptr() if need_dx else 0, Kh,
                           1 if fuse_mask else 0, self.hist.data_ptr(), self.step_counter.data_ptr(), self.hist_slots],
                          [alpha])
            if b.act == "relu" and not premasked:
                self._add(lst, N.OP_RELU_MASK, [grad["t"].data_ptr(), rec["t"].data_ptr(), rows * grad["ld"]])
            # parameter gradients only feed the optimizer: they run on the engine's side stream
            # (a parallel branch of the captured graph) while the dgrad chain continues
            # (graph branches).  DK_SIDE_STREAMS=2 (default): wgrads on branch 1, bias column sums on
            # branch 2, and the first layer's wgrad -- nothing is left to overlap it with -- on the main
            # stream; DK_SIDE_STREAMS=1: one branch carries both.
            if self.compact:
                # weight / bias gradient + optimizer of this layer are tiles of the fused update kernel
                padded = wbld != K
                self._bwd_layers.append(dict(
                    dz=grad["t"].data_ptr(), lddz=grad["ld"],
                    x=None if a_in.get("slot") is not None else a_in["t"].data_ptr(), ldx=a_in["ld"],
                    x_slot=a_in["slot"] if a_in.get("slot") is not None else -1, n_out=Nout, k_in=K,
                    w_off=kseg.offset, b_off=bseg.offset if bseg is not None else -1,
                    wb_pad=wbp if padded else None, ldwb_pad=wbld if padded else 0))
                if not need_dx:
                    return None, True
                if head_din is not None:
                    return dict(t=head_din, rows=rows, cols=K, ld=head_din.shape[1]), fuse_mask
                din = self._buf(rows, _r8(K))
                ep = N.GemmEpilogue()
                ep.d, ep.ldd, ep.alpha = din.data_ptr(), _r8(K), 1.0
                fuse_mask = prev is not None and prev.kind == "dense" and prev.act == "relu"
                if fuse_mask:
                    ep.mask, ep.ld_mask = prev.out_rec["t"].data_ptr(), prev.out_rec["ld"]
                    if prev.drop_p > 0:
                        ep.alpha = 1.0 / (1.0 - prev.drop_p)
                elif prev is not None and prev.kind == "dense" and prev.drop_p > 0:
                    raise UnsupportedByNativeEngine("dropout after a non-ReLU dense layer")
                self._gemm(lst, grad["t"].data_ptr(), grad["ld"], wbp, wbld, rows, K, Nout,
                           N.GEMM_B_MN | (N.GEMM_SHORT_A if rows < 128 else 0), ep, bn=self._narrow_bn(rows, K, 64), splits=1)
                return dict(t=din, rows=rows, cols=K, ld=_r8(K)), fuse_mask
            two = self._side_streams >= 2
            s_bias = 2 if two else (1 if need_dx else 0)
            s_wgrad = (1 if need_dx else 0) if two else 1
            if implicit and not implicit_wgrad:  # the column matrix was built on branch 3 during the forward pass
                self._add(lst, N.OP_JOIN, [3])
            for sid in {s_bias, s_wgrad} - {0}:
                self._add(lst, N.OP_FORK, [sid])
                self._forked.add(sid)
            self.lib.dk_engine_set_build_stream(self.engine, s_bias)
            if bseg is not None and not (implicit and wgrad_mode == "tma"):   # the TMA wgrad kernel also sums dZ
                self._add(lst, N.OP_COLSUM, [grad["t"].data_ptr(), rows, Nout, grad["ld"], g_ptr + 4 * bseg.offset],
                          [1.0])
            self.lib.dk_engine_set_build_stream(self.engine, s_wgrad)
            # wgrad: dW[Nout, K] = dZ^T[Nout, rows] * In[rows, K]  (both operands MN-major views)
            if implicit and wgrad_mode == "tma":
                if grad["ld"] != Nout:
                    raise UnsupportedByNativeEngine("implicit wgrad needs an unpadded output gradient")
                Hh, Ww, _ = b.in_shape
                Ci = cin_eff
                Oh, Ow, _ = b.out_shape
                dw_ptr = g_ptr + 4 * kseg.offset
                if cpad:   # gradient w.r.t. the channel-padded weights: scratch block, real channels copied out below
                    dwpad = self._buf(Nout, K, dtype=torch.float32)
                    dw_ptr = dwpad.data_ptr()
                    self._add(lst, N.OP_MEMSET, [dw_ptr, 0, Nout * K * 4])
                r = self.lib.dk_engine_add_conv_wgrad_tma(
                    self.engine, lst, C.c_void_p(inp["t"].data_ptr()), B, Hh, Ww, Ci, Oh, Ow, b.kh, b.kw, b.stride, b.pad,
                    C.c_void_p(grad["t"].data_ptr()), grad["ld"], C.c_void_p(dw_ptr), K, Nout,
                    C.c_void_p(g_ptr + 4 * bseg.offset) if bseg is not None else None)
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_conv_wgrad_tma failed: {r}")
                if cpad:
                    self._add(lst, N.OP_MEMCPY2D, [g_ptr + 4 * kseg.offset, c_real * 4, dw_ptr, cpad * 4, c_real * 4,
                                                   Nout * b.kh * b.kw])
            elif implicit and implicit_wgrad:
                if grad["ld"] != Nout:
                    raise UnsupportedByNativeEngine("implicit wgrad needs an unpadded output gradient")
                Hh, Ww, Ci = b.in_shape
                Oh, Ow, _ = b.out_shape
                r = self.lib.dk_engine_add_conv_wgrad(self.engine, lst, C.c_void_p(inp["t"].data_ptr()), Hh, Ww, Ci, Oh, Ow,
                                                      b.kh, b.kw, b.stride, b.pad, C.c_void_p(grad["t"].data_ptr()),
                                                      grad["ld"], C.c_void_p(g_ptr + 4 * kseg.offset), K, Nout, rows)
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_conv_wgrad failed: {r}")
            else:
                ep = N.GemmEpilogue()
                ep.d, ep.ldd, ep.d_fp32, ep.alpha = g_ptr + 4 * kseg.offset, K, 1, 1.0
                self._gemm(lst, grad["t"].data_ptr(), grad["ld"], a_in["t"].data_ptr(), a_in["ld"], Nout, K, rows,
                           N.GEMM_A_MN | N.GEMM_B_MN, ep)
            self.lib.dk_engine_set_build_stream(self.engine, 0)
            if not need_dx:
                return None, True
            if head_din is not None:  # the fused head already produced the (masked) input gradient
                return dict(t=head_din, rows=rows, cols=K, ld=head_din.shape[1]), fuse_mask
            if implicit and grad["ld"] == Nout:
                H, Wd, Cin = b.in_shape
                OH, OW, _ = b.out_shape
                # implicit dgrad: dX[(b, ih, iw), c] = sum_{kh', kw', co} dZ[b, (ih - off + kh') / s, ..., co] * Wd[c, (kh', kw', co)]
                # (Wd = the weights flipped / transposed once per step); the producer's dReLU mask is fused
                Kd = b.kh * b.kw * Nout
                wd = self._buf(Cin, Kd)
                self._add(lst, N.OP_WFLIP, [wbp, wbld, wd.data_ptr(), Kd, Nout, Cin, b.kh, b.kw])
                dx = self._buf(B * H * Wd, Cin)
                fuse_relu = _relu_mask_fusable(prev, B * H * Wd, Cin)
                ep = N.GemmEpilogue()
                ep.d, ep.ldd, ep.alpha = dx.data_ptr(), Cin, 1.0
                if fuse_relu:
                    ep.mask, ep.ld_mask = prev.out_rec["t"].data_ptr(), prev.out_rec["ld"]
                r = self.lib.dk_engine_add_conv_gemm(self.engine, lst, C.c_void_p(grad["t"].data_ptr()), OH, OW, Nout, H, Wd,
                                                     b.kh, b.kw, 1, b.kh - 1 - b.pad, b.stride, C.c_void_p(wd.data_ptr()), Kd,
                                                     B * H * Wd, Cin, Kd, C.byref(ep))
                if r < 0:
                    raise RuntimeError(f"dk_engine_add_conv_gemm(dgrad) failed: {r}")
                return dict(t=dx, rows=B * H * Wd, cols=Cin, ld=Cin), fuse_relu
            # dgrad: dIn[rows, K] = dZ[rows, Nout] * W[Nout, K]   (W read as an MN-major B operand)
            din = self._buf(rows, _r8(K))
            ep = N.GemmEpilogue()
            ep.d, ep.ldd, ep.alpha = din.data_ptr(), _r8(K), 1.0
            fuse_mask = (b.kind == "dense" and prev is not None and prev.kind == "dense" and prev.act == "relu")
            if fuse_mask:
                ep.mask, ep.ld_mask = prev.out_rec["t"].data_ptr(), prev.out_rec["ld"]
                if prev.drop_p > 0:
                    ep.alpha = 1.0 / (1.0 - prev.drop_p)
            elif b.kind == "dense" and prev is not None and prev.kind == "dense" and prev.drop_p > 0:
                raise UnsupportedByNativeEngine("dropout after a non-ReLU dense layer")
            self._gemm(lst, grad["t"].data_ptr(), grad["ld"], wbp, wbld, rows, K, Nout, N.GEMM_B_MN, ep)
            if b.kind == "conv":
                H, Wd, Cin = b.in_shape
                OH, OW, _ = b.out_shape
                dx = self._buf(B * H * Wd, Cin)
                # dReLU of the producing conv / BN block is applied while the gradient image is assembled
                fuse_relu = _relu_mask_fusable(prev, B * H * Wd, Cin)
                self._add(lst, N.OP_COL2IM, [din.data_ptr(), _r8(K), B, H, Wd, Cin, b.kh, b.kw, b.stride, b.pad, OH, OW,
                                             dx.data_ptr(), prev.out_rec["t"].data_ptr() if fuse_relu else 0])
                return dict(t=dx, rows=B * H * Wd, cols=Cin, ld=Cin), fuse_relu
            return dict(t=din, rows=rows, cols=K, ld=_r8(K)), fuse_mask

        return rec, backward

    # -- pooling / reshapes ---------------------------------------------------------------------------
    def _emit_pool(self, b: _Block, cur: dict):
        B = self.B
        H, Wd, Cc = b.in_shape
        OH, OW, _ = b.out_shape
        out = self._buf(B * OH * OW, Cc)
        rec = dict(t=out, rows=B * OH * OW, cols=Cc, ld=Cc, nhwc=b.out_shape)
        for lst in self._lists:
            self._add(lst, N.OP_MAXPOOL_FWD, [cur["t"].data_ptr(), B, H, Wd, Cc, b.k, b.k, out.data_ptr()])
        inp = cur

        def backward(grad, premasked, need_dx, prev):
            dx = self._buf(B * H * Wd, Cc)
            # the pool's input is the previous block's post-ReLU output: its dReLU mask is (x > 0),
            # which the pooling backward already has in registers -> no separate mask pass
            fuse_relu = _relu_mask_fusable(prev, B * H * Wd, Cc)
            self._add(self._bwd_list, N.OP_MAXPOOL_BWD, [inp["t"].data_ptr(), out.data_ptr(), grad["t"].data_ptr(), B, H,
                                                     Wd, Cc, b.k, b.k, dx.data_ptr(), 1 if fuse_relu else 0])
            return dict(t=dx, rows=B * H * Wd, cols=Cc, ld=Cc), fuse_relu

        return rec, backward

    def _emit_flatten(self, b: _Block, cur: dict):
        if cur["ld"] != cur["cols"]:
            raise UnsupportedByNativeEngine("flatten of a padded activation")
        feat = cur["rows"] * cur["cols"] // self.B
        if feat % 8 != 0:
            raise UnsupportedByNativeEngine("flattened feature count must be a multiple of 8")
        rec = dict(t=cur["t"], rows=self.B, cols=feat, ld=feat, nhwc=None)
        inp = cur

        def backward(grad, premasked, need_dx, prev):
            return dict(t=grad["t"], rows=inp["rows"], cols=inp["cols"], ld=inp["ld"]), premasked

        return rec, backward

    def _emit_gap(self, b: _Block, cur: dict):
        B = self.B
        H, Wd, Cc = b.in_shape
        P = H * Wd
        out = self._buf(B, Cc)
        rec = dict(t=out, rows=B, cols=Cc, ld=Cc, nhwc=None)
        for lst in self._lists:
            self._add(lst, N.OP_GAP_FWD, [cur["t"].data_ptr(), B, P, Cc, out.data_ptr()])

        def backward(grad, premasked, need_dx, prev):
            dx = self._buf(B * P, Cc)
            self._add(self._bwd_list, N.OP_GAP_BWD, [grad["t"].data_ptr(), B, P, Cc, dx.data_ptr()])
            return dict(t=dx, rows=B * P, cols=Cc, ld=Cc), False

        return rec, backward

    # -- BatchNormalization (+ ReLU) ------------------------------------------------------------------
    def _emit_bn(self, b: _Block, cur: dict):
        if cur["ld"] != cur["cols"] or cur["cols"] != b.channels:
            raise UnsupportedByNativeEngine("BatchNormalization on a padded activation")
        rows, Cc = cur["rows"], b.channels
        wf, g_ptr = self.W.data_ptr(), (N.ptr(self.G) if self.training else 0)
        seg = {n: self._seg(b.layer_index, b.seg_prefix + n) for n in ("gamma", "beta", "moving_mean", "moving_variance")}
        gamma, beta = wf + 4 * seg["gamma"].offset, wf + 4 * seg["beta"].offset
        mm, mv = wf + 4 * seg["moving_mean"].offset, wf + 4 * seg["moving_variance"].offset
        out = self._buf(rows, Cc)
        rec = dict(t=out, rows=rows, cols=Cc, ld=Cc, nhwc=cur["nhwc"])
        saved_mean, saved_invstd = self._buf(Cc, dtype=torch.float32), self._buf(Cc, dtype=torch.float32)
        relu = 1 if b.act == "relu" else 0
        fsum, bsum = self._bn_slice(2 * Cc), self._bn_slice(2 * Cc)
        for lst in self._lists:
            if lst in self._train_lists:
                self._add(lst, N.OP_BN_FWD, [cur["t"].data_ptr(), rows, Cc, fsum, saved_mean.data_ptr(),
                                             saved_invstd.data_ptr(), mm, mv, gamma, beta, relu, out.data_ptr()],
                          [b.eps, b.momentum])
            else:
                self._add(lst, N.OP_BN_INF, [cur["t"].data_ptr(), rows, Cc, mm, mv, gamma, beta, relu, out.data_ptr()],
                          [b.eps])
        inp = cur
        b.out_rec = rec

        def backward(grad, premasked, need_dx, prev):
            dx = self._buf(rows, Cc)
            mask = out.data_ptr() if (relu and not premasked) else 0
            self._add(self._bwd_list, N.OP_BN_BWD, [grad["t"].data_ptr(), inp["t"].data_ptr(), mask, rows, Cc,
                                                saved_mean.data_ptr(), saved_invstd.data_ptr(), gamma, bsum,
                                                g_ptr + 4 * seg["gamma"].offset, g_ptr + 4 * seg["beta"].offset,
                                                dx.data_ptr()])
            return dict(t=dx, rows=rows, cols=Cc, ld=Cc), True

        return rec, backward

    # -- residual block: conv-BN-ReLU-conv-BN (+ projection) , add, ReLU ---------------------------------
    def _emit_res(self, b: _Block, cur: dict, bi: int):
        x = cur
        y1, bw_c1 = self._emit_matmul(b.conv1, x, bi, False)
        a1, bw_b1 = self._emit_bn(b.bn1, y1)
        y2, bw_c2 = self._emit_matmul(b.conv2, a1, bi, False)
        z2, bw_b2 = self._emit_bn(b.bn2, y2)
        if b.proj is not None:
            yp, bw_cp = self._emit_matmul(b.proj, x, bi, False)
            sc, bw_bp = self._emit_bn(b.bnp, yp)
        else:
            sc = x
            if sc["ld"] != sc["cols"]:
                raise UnsupportedByNativeEngine("identity shortcut on a padded activation")
        rows, Cc = z2["rows"], z2["cols"]
        out = self._buf(rows, Cc)
        rec = dict(t=out, rows=rows, cols=Cc, ld=Cc, nhwc=b.out_shape)
        for lst in self._lists:
            self._add(lst, N.OP_ADD, [out.data_ptr(), z2["t"].data_ptr(), sc["t"].data_ptr(), rows * Cc, 1])
        b.out_rec = rec

        def backward(grad, premasked, need_dx, prev):
            lst = self._bwd_list
            if not premasked:
                self._add(lst, N.OP_RELU_MASK, [grad["t"].data_ptr(), out.data_ptr(), rows * Cc])
            g2, _ = bw_b2(grad, True, True, None)
            ga1, _ = bw_c2(g2, True, True, None)
            gy1, _ = bw_b1(ga1, False, True, None)       # ReLU of bn1 masked inside the BN backward
            gx1, _ = bw_c1(gy1, True, need_dx, None)
            if b.proj is not None:
                gp, _ = bw_bp(grad, True, True, None)
                gx2, _ = bw_cp(gp, True, need_dx, None)
            else:
                gx2 = grad
            if not need_dx:
                return None, True
            n = gx1["rows"] * gx1["ld"]
            dx = self._buf(gx1["rows"], gx1["ld"])
            self._add(lst, N.OP_ADD, [dx.data_ptr(), gx1["t"].data_ptr(), gx2["t"].data_ptr(), n, 0])
            return dict(t=dx, rows=gx1["rows"], cols=gx1["cols"], ld=gx1["ld"]), False

        return rec, backward

    def pull_rest_ranges(self):
        """Flat ranges NOT covered by the fused-pull GEMM (pulled by the plain pull kernel)."""
        if self.pull_segment is None:
            return [(0, self.P)]
        off, size = self.pull_segment
        size8 = (size + 7) // 8 * 8
        return [(lo, hi) for lo, hi in ((0, off), (off + size8, self.P)) if hi > lo]

    def _prepend_pad_refresh(self, lst: int) -> None:
        # pad refresh ops are appended; order inside a list only matters relative to the GEMMs that
        # read the padded shadows, so they live in a dedicated list run before `lst`.
        if not self._pad_refresh:
            return
        if not hasattr(self, "L_pad"):
            self.L_pad = self.lib.dk_engine_new_list(self.engine)
            for dst, dpitch, src, spitch, width, height in self._pad_refresh:
                self._add(self.L_pad, N.OP_MEMCPY2D, [dst, dpitch, src, spitch, width, height])

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    def _run(self, lst: int) -> None:
        r = self.lib.dk_engine_run(self.engine, lst, C.c_void_p(N.current_stream()))
        if r != 0:
            raise RuntimeError(f"dk_engine_run(list {lst}) failed: {r}")

    def refresh_shadow(self) -> None:
        """Recompute the bf16 shadow from the fp32 master (after an external write to ``W``)."""
        N.check(self.lib.dk_cast_bf16(self.W.data_ptr(), self.Wb.data_ptr(), self.P, C.c_void_p(N.current_stream())),
                "dk_cast_bf16")
        self.refresh_pads()

    def refresh_pads(self) -> None:
        """Re-derive the 8-padded weight shadows from the flat bf16 shadow (after a pull or any other external write)."""
        if hasattr(self, "L_pad"):
            self._run(self.L_pad)

    weights_changed = refresh_shadow

    def enqueue_region_input(self, x_ptr: int) -> None:
        """Compact program: stage (cast + affine) the mini-batches of a whole region -- ``region_steps * B``
        rows starting at ``x_ptr`` -- with one launch; steps then run with ``staged_row = j * B``."""
        self.lib.dk_engine_set_slot(self.engine, SLOT_X, C.c_void_p(x_ptr))
        self._run(self.L_in_region)

    def enqueue_step(self, x_ptr: int, y_ptr: int, fused_pull: bool = False, comm: bool = False,
                     staged_row: Optional[int] = None) -> None:
        """Enqueue one training step reading its batch from device pointers (graph-capturable).
        With ``fused_pull`` the first layer's weights are pulled from the parameter server inside
        its forward GEMM (the caller pulls every other segment beforehand).  Compact program: with
        ``staged_row`` the batch was already staged by :meth:`enqueue_region_input`; ``comm`` makes the
        fused backward-update kernel also perform the window-boundary exchange with the PS."""
        self.lib.dk_engine_set_slot(self.engine, SLOT_Y, C.c_void_p(y_ptr))
        if self.compact:
            xr = self._xb_region
            if staged_row is None:
                self.lib.dk_engine_set_slot(self.engine, SLOT_X, C.c_void_p(x_ptr))
                self._run(self.L_in_step)
                staged_row = 0
            self.lib.dk_engine_set_slot(self.engine, SLOT_XB, C.c_void_p(xr.data_ptr() + staged_row * xr.shape[1] * 2))
            # (no pad refresh here: the fused update kernel keeps the padded weight shadows current itself; whoever
            # writes the flat shadow from outside -- a pull, set_flat -- calls refresh_pads())
            self._run(self.L_step)
            if comm and self.L_bwd_comm < 0:
                raise RuntimeError("the replica was planned without a comm_spec")
            self._run(self.L_bwd_comm if comm else self.L_bwd)
            return
        self.lib.dk_engine_set_slot(self.engine, SLOT_X, C.c_void_p(x_ptr))
        if hasattr(self, "L_pad"):
            self._run(self.L_pad)
        if fused_pull:
            if self.L_step_pull < 0:
                raise RuntimeError("the replica was not planned with pull_center_ptr")
            self._run(self.L_step_pull)
        else:
            self._run(self.L_step)
        self._run(self.L_bwd)

    def enqueue_forward(self, x_ptr: int) -> None:
        self.lib.dk_engine_set_slot(self.engine, SLOT_X, C.c_void_p(x_ptr))
        if self.compact:
            self._run(self.L_in_step)
            self.lib.dk_engine_set_slot(self.engine, SLOT_XB, C.c_void_p(self._xb_region.data_ptr()))
        if hasattr(self, "L_pad"):
            self._run(self.L_pad)
        self._run(self.L_fwd)

    def step_kernel_count(self) -> int:
        return (self.lib.dk_engine_list_kernels(self.engine, self.L_step)
                + self.lib.dk_engine_list_kernels(self.engine, self.L_bwd))

    def launches(self) -> int:
        return int(self.lib.dk_engine_launches(self.engine))

    # -- eager convenience API (Replica interface) ------------------------------------------------
    def _stage_inputs(self, x, y=None) -> None:
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        if x.shape[0] != self.B:
            raise ValueError(f"native replica was planned for batch {self.B}, got {x.shape[0]}")
        self._x_stage.copy_(x.reshape(self.B, -1).to(self.in_torch_dtype), non_blocking=True)
        if y is not None:
            y = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))
            if self.dense_labels:
                self._y_stage.copy_(y.reshape(self.B, -1).to(torch.float32), non_blocking=True)
            else:
                if y.dim() == 2 and y.shape[1] > 1:
                    y = y.argmax(dim=1)
                self._y_stage.copy_(y.reshape(-1).to(torch.int32), non_blocking=True)

    def train_on_batch(self, x, y):
        self._stage_inputs(x, y)
        slot = (int(self.step_counter.item()) - self.step_base) % self.hist_slots
        self.hist[slot].zero_()
        self.enqueue_step(self._x_stage.data_ptr(), self._y_stage.data_ptr())
        self.iteration += 1
        rec = self.hist[slot].tolist()
        return float(rec[0]), float(rec[1])

    @torch.no_grad()
    def predict(self, x) -> torch.Tensor:
        self._stage_inputs(x)
        self.enqueue_forward(self._x_stage.data_ptr())
        return self.probs.clone()

    def set_flat(self, flat: torch.Tensor) -> None:
        with torch.no_grad():
            self.W.copy_(flat.to(self.W.device, torch.float32))
        self.refresh_shadow()

    def ensure_snapshot(self) -> torch.Tensor:
        if self.W1 is None:
            self.W1 = self.W.clone()
        return self.W1

    def set_learning_rate(self, lr: float) -> None:
        raise NotImplementedError("re-plan the replica to change the learning rate")

    def close(self) -> None:
        if self.engine:
            self.lib.dk_engine_destroy(self.engine)
            self.engine = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def try_native_replica(model: Sequential, optimizer, loss: str, batch_size: int, device_index: int,
                       **kw) -> Optional[NativeReplica]:
    """``NativeReplica`` if the model is lowerable, else ``None`` (caller uses the autograd path)."""
    try:
        return NativeReplica(model, optimizer, loss, batch_size, device_index, **kw)
    except UnsupportedByNativeEngine:
        return None


def native_predict(model: Sequential, x: torch.Tensor, batch_size: int, device, label_index=None):
    """Batched inference through the native engine (reference op K14); ``None`` if not lowerable.

    ``label_index=(activation_threshold, default_index)`` also runs the LabelIndexTransformer rule
    (``distkeras/transformers.py:302-350``) on every batch of probabilities while they are still in HBM (one
    ``dk_label_index`` launch per batch) and returns ``(probs, indices)``."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    n = int(x.shape[0])
    if n == 0:
        empty = torch.empty(0, model.output_shape[-1])
        return empty if label_index is None else (empty, torch.empty(0, dtype=torch.int64))
    bs = min(batch_size, max(8, _r8(n)))
    in_dtype = "u8" if x.dtype == torch.uint8 else "f32"
    try:
        rep = NativeReplica(model, "sgd", "categorical_crossentropy", bs, idx, in_dtype=in_dtype, training=False)
    except UnsupportedByNativeEngine:
        return None
    outs, idxs = [], []
    flat = x.reshape(n, -1)
    for i in range(0, n, bs):
        chunk = flat[i:i + bs]
        m = chunk.shape[0]
        if m < bs:
            pad = torch.zeros(bs - m, flat.shape[1], dtype=flat.dtype)
            chunk = torch.cat([chunk, pad], dim=0)
        probs = rep.predict(chunk)
        if label_index is not None:
            out_idx = torch.empty(bs, dtype=torch.int32, device=probs.device)
            N.check(rep.lib.dk_label_index(probs.data_ptr(), bs, probs.shape[1], float(label_index[0]),
                                           int(label_index[1]), out_idx.data_ptr(), None, None,
                                           C.c_void_p(N.current_stream())), "dk_label_index")
            idxs.append(out_idx[:m].cpu().to(torch.int64))
        outs.append(probs[:m].cpu())
    rep.close()
    if label_index is not None:
        return torch.cat(outs, dim=0), torch.cat(idxs, dim=0)
    return torch.cat(outs, dim=0)
