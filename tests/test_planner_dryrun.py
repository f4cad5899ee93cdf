"""The native engine's planner (``parallel/engine.py``) normally only runs on a GPU box.  This dry run lowers
every zoo model on CPU -- real host-side op lists in the C++ engine, a recording proxy in place of the calls that
need a CUDA driver (tensor-map encoding) -- under every planner switch, so a Python-level mistake in a lowering
branch cannot hide until the GPU tests run."""
import collections
from unittest import mock

import pytest
import torch

from distkeras_b200 import _native as N
from distkeras_b200.models import ZOO
from distkeras_b200.parallel import engine as eng

GPU_ONLY = ("dk_engine_add_gemm", "dk_engine_add_gemm_pull", "dk_engine_add_conv_gemm", "dk_engine_add_conv_wgrad",
            "dk_engine_add_gemm_slot", "dk_engine_add_bwd_update", "dk_engine_add_conv_wgrad_tma")
OP_NAMES = {v: k for k, v in vars(N).items() if k.startswith("OP_") and isinstance(v, int)}


class RecordingLib:
    """Delegates host-only engine calls to the real library and records the GPU-only ones."""

    def __init__(self, real):
        self._real = real
        self.calls = collections.Counter()
        self.ops = collections.Counter()

    def __getattr__(self, name):
        real = getattr(self._real, name)
        if name in GPU_ONLY:
            def fake(*args):
                self.calls[name] += 1
                return 0
            return fake
        if name == "dk_engine_add_op":
            def add_op(h, lst, kind, *rest):
                self.ops[OP_NAMES.get(int(kind), int(kind))] += 1
                return real(h, lst, kind, *rest)
            return add_op
        return real


def _lower(model_name, batch, monkeypatch, native_lib, **env):
    for k in ("DK_IMPLICIT_CONV", "DK_IMPLICIT_WGRAD", "DK_SIDE_STREAMS", "DK_FUSED_HEAD", "DK_COMPACT", "DK_COMPACT_MAX_BATCH", "DK_HEAD_IN_FWD"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    lib = RecordingLib(native_lib)
    cpu = torch.device("cpu")
    with mock.patch.object(eng.N, "lib", lambda: lib), mock.patch.object(eng.torch, "device", lambda *a, **k: cpu), \
            mock.patch.object(eng.torch.cuda, "set_device", lambda *a, **k: None), \
            mock.patch.object(eng.NativeReplica, "refresh_shadow", lambda self: None):
        rep = eng.NativeReplica(ZOO[model_name](seed=0), "adam", "categorical_crossentropy", batch, 0, in_dtype="u8",
                                input_affine=(1 / 255.0, 0.0))
        sizes = {n: lib.dk_engine_list_size(rep.engine, getattr(rep, n)) for n in ("L_step", "L_bwd", "L_fwd")}
        rep.close()
    return lib, sizes


def test_compact_lowering_of_small_batch_mlp(monkeypatch, native_lib):
    """Batch <= 256 dense stacks: region-level input stage, slot-fed first GEMM, ONE fused backward-update op."""
    lib, sizes = _lower("mnist_mlp", 64, monkeypatch, native_lib)
    assert lib.calls["dk_engine_add_bwd_update"] == 1 and lib.calls["dk_engine_add_gemm_slot"] == 2  # L_step + L_fwd
    assert lib.ops["OP_OPTIM"] == 0 and lib.ops["OP_COLSUM"] == 0 and lib.ops["OP_MEMSET"] == 0
    # the classifier head rides in the epilogue of the second forward GEMM: no head op at all
    assert lib.ops["OP_INPUT"] == 2 and lib.ops["OP_HEAD"] == 0
    # L_step: fwd1 (slot-fed), fwd2 + head | L_bwd: dgrad2, fused update | L_fwd: fwd1, fwd2, head GEMM, softmax
    assert lib.calls["dk_engine_add_gemm"] == 4 and sizes["L_bwd"] == 0 and sizes["L_fwd"] == 1
    lib, sizes = _lower("mnist_mlp", 64, monkeypatch, native_lib, DK_HEAD_IN_FWD="0")
    assert lib.ops["OP_HEAD"] == 1 and sizes["L_bwd"] == 1


def test_compact_lowering_of_the_higgs_mlp(monkeypatch, native_lib):
    """500-wide layers (not a multiple of 8): compact program with padded weight shadows, head kernel on the padded width."""
    lib, sizes = _lower("higgs_mlp", 64, monkeypatch, native_lib)
    assert lib.calls["dk_engine_add_bwd_update"] == 1 and lib.ops["OP_HEAD"] == 1 and lib.ops["OP_OPTIM"] == 0
    # L_step: fwd1 (slot), fwd2, fwd3 | L_bwd: head, dgrad3, dgrad2, update | L_fwd: 4 GEMMs + softmax; 4 padded shadows
    assert lib.calls["dk_engine_add_gemm"] == 7 and lib.calls["dk_engine_add_gemm_slot"] == 2 and lib.ops["OP_MEMCPY2D"] == 4


@pytest.mark.parametrize("model_name,batch", [("mnist_mlp", 512), ("higgs_mlp", 512), ("mnist_convnet", 32),
                                              ("cifar10_cnn", 32), ("resnet18", 4)])
def test_default_lowering(model_name, batch, monkeypatch, native_lib):
    lib, sizes = _lower(model_name, batch, monkeypatch, native_lib)
    assert sizes["L_step"] > 0 and sizes["L_bwd"] > sizes["L_step"] // 2 and sizes["L_fwd"] > 0
    assert lib.calls["dk_engine_add_gemm"] > 0
    # implicit (TMA-im2col) convolution is the default for layers with 32 / 64k input channels
    assert (lib.calls["dk_engine_add_conv_gemm"] > 0) == (model_name in ("mnist_convnet", "cifar10_cnn", "resnet18"))
    assert lib.ops["OP_OPTIM"] == 1 and lib.ops["OP_FORK"] == lib.ops["OP_FORK"]  # one optimizer launch per step
    if model_name in ("mnist_mlp", "cifar10_cnn", "higgs_mlp", "mnist_convnet"):
        # fused head in training (500 / 225-wide inputs run on the 8-padded width), softmax kernel in inference
        assert lib.ops["OP_HEAD"] == 1 and lib.ops["OP_XENT"] == 1
    if "cnn" in model_name or "convnet" in model_name:
        assert lib.ops["OP_RELU_MASK"] == 0  # every conv dReLU is fused into the dgrad epilogue / col2im / pool backward


@pytest.mark.parametrize("env", [dict(DK_FUSED_HEAD="0"), dict(DK_SIDE_STREAMS="1"), dict(DK_IMPLICIT_CONV="1"),
                                 dict(DK_IMPLICIT_CONV="1", DK_IMPLICIT_WGRAD="1"), dict(DK_IMPLICIT_CONV="0")],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
@pytest.mark.parametrize("model_name,batch", [("mnist_mlp", 512), ("cifar10_cnn", 32), ("resnet18", 4)])
def test_lowering_under_every_switch(model_name, batch, env, monkeypatch, native_lib):
    lib, sizes = _lower(model_name, batch, monkeypatch, native_lib, **env)
    assert sizes["L_step"] > 0 and sizes["L_bwd"] > 0
    conv_model = model_name != "mnist_mlp"
    if env.get("DK_FUSED_HEAD") == "0":
        assert lib.ops["OP_HEAD"] == 0 and lib.ops["OP_XENT"] == 2
    if env.get("DK_IMPLICIT_CONV") == "0":
        assert lib.calls["dk_engine_add_conv_gemm"] == 0
    if env.get("DK_IMPLICIT_CONV") == "1" and conv_model:
        assert lib.calls["dk_engine_add_conv_gemm"] > 0 and lib.ops["OP_WFLIP"] > 0
        if env.get("DK_IMPLICIT_WGRAD") == "1":
            assert lib.calls["dk_engine_add_conv_wgrad"] > 0
        else:  # default: the TMA-im2col weight gradient wherever the geometry allows it
            assert lib.calls["dk_engine_add_conv_wgrad"] == 0 and lib.calls["dk_engine_add_conv_wgrad_tma"] > 0
    if env.get("DK_SIDE_STREAMS") == "1" and not conv_model:  # conv models also join the im2col branch per layer
        assert lib.ops["OP_JOIN"] == 1
