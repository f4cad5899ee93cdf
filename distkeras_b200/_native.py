"""ctypes binding of the native sm_100a runtime (``csrc/`` -> ``lib/libdistkeras_b200.so``).

The library is built in-tree by ``build_native.py`` (``__graft_entry__.build``).  On a machine
with a GPU the native path is mandatory: :func:`lib` raises if the shared object is missing
instead of silently falling back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdistkeras_b200.so")

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()

i32, i64, f32, f64, vp = C.c_int, C.c_long, C.c_float, C.c_double, C.c_void_p
u32 = C.c_uint


class GemmEpilogue(C.Structure):
    """Mirror of ``DkGemmEpilogue`` (csrc/gemm.h)."""

    _fields_ = [
        ("bias", vp), ("bias_along_m", i32), ("act", i32), ("mask", vp), ("ld_mask", i32),
        ("d", vp), ("ldd", i32), ("d_fp32", i32), ("accumulate", i32), ("dt", vp), ("lddt", i32),
        ("alpha", f32), ("drop_p", f32), ("drop_seed", u32), ("step", vp), ("tma_store", i32), ("tma_mask", i32),
        ("trace", vp),
        # classifier head fused into the producing GEMM (csrc/gemm.h)
        ("head_w", vp), ("head_ldw", i32), ("head_bias", vp), ("head_c", i32), ("head_labels", vp),
        ("head_label_slot", i32), ("head_acc", vp), ("head_sync", vp), ("head_dz", vp), ("head_ldz", i32),
        ("head_dh", vp), ("head_lddh", i32), ("head_alpha", f32), ("head_hist", vp), ("head_step", vp),
        ("head_hist_slots", i32),
    ]


BWD_MAX_LAYERS, BWD_MAX_SHARDS = 6, 16
COMM_NONE, COMM_EXCHANGE, COMM_ELASTIC = 0, 1, 2


class BwdLayerDesc(C.Structure):
    """Mirror of ``DkBwdLayerDesc`` (csrc/dense_fused.h)."""

    _fields_ = [("dz", vp), ("lddz", i64), ("x", vp), ("ldx", i64), ("x_slot", i32), ("n_out", i32), ("k_in", i32),
                ("w_off", i64), ("b_off", i64), ("wb_pad", vp), ("ldwb_pad", i64)]


class BwdUpdateDesc(C.Structure):
    """Mirror of ``DkBwdUpdateDesc`` (csrc/dense_fused.h)."""

    _fields_ = [("nlayers", i32), ("batch", i32), ("layer", BwdLayerDesc * BWD_MAX_LAYERS),
                ("w", vp), ("s0", vp), ("s1", vp), ("w1", vp), ("wb", vp),
                ("opt_kind", i32), ("lr", f32), ("p0", f32), ("p1", f32), ("eps", f32), ("decay", f32), ("nesterov", i32),
                ("step", vp), ("done_counter", vp), ("step_inc", i32),
                ("comm_mode", i32), ("comm_scale", f32), ("scale_dev", vp), ("alpha", f32),
                ("nshards", i32), ("shard_per", i64), ("shard_center", vp * BWD_MAX_SHARDS),
                ("ctrl", vp), ("worker", i32), ("last_update", vp), ("trace", vp)]


# op kinds (csrc/engine.h)
OP_INPUT, OP_GEMM, OP_XENT, OP_ROWSUM, OP_TRANSPOSE, OP_OPTIM, OP_IM2COL, OP_COL2IM = range(8)
OP_MAXPOOL_FWD, OP_MAXPOOL_BWD, OP_RELU_MASK, OP_ADD, OP_MEMSET = 8, 9, 10, 11, 12
OP_PS_COMMIT, OP_PS_PULL, OP_PS_EXCHANGE, OP_PS_ELASTIC, OP_PS_DAMPED, OP_PS_TICKET = 13, 14, 15, 16, 17, 18
OP_LOCK_ACQUIRE, OP_LOCK_RELEASE, OP_EAMSGD_PRE, OP_EAMSGD_POST, OP_CAST, OP_ELOSS = 19, 20, 21, 22, 23, 24
OP_MEMCPY, OP_LABEL_INDEX, OP_COLSUM, OP_MEMCPY2D, OP_FORK, OP_JOIN, OP_GEMM_PULL = 25, 26, 27, 28, 29, 30, 31
OP_BN_FWD, OP_BN_INF, OP_BN_BWD, OP_GAP_FWD, OP_GAP_BWD, OP_HEAD = 32, 33, 34, 35, 36, 37
OP_CONV_GEMM, OP_WFLIP, OP_CONV_WGRAD, OP_BWD_UPDATE, OP_CONV_WGRAD_TMA = 38, 39, 40, 41, 42
GEMM_TF32, GEMM_A_MN, GEMM_B_MN, GEMM_PERSISTENT, GEMM_PAIR, GEMM_SHORT_A, GEMM_MCAST_A = 1, 2, 4, 8, 16, 32, 64

OPT_KINDS = {"sgd": 0, "momentum": 1, "adagrad": 2, "rmsprop": 3, "adam": 4, "adadelta": 5, "adamax": 6, "nadam": 7}
IN_U8, IN_F32, IN_BF16 = 0, 1, 2
LOSS_XENT, LOSS_MSE, LOSS_BCE = 0, 1, 2

# control block words (csrc/ps.h)
CTRL_NUM_UPDATES, CTRL_LOCK_NEXT, CTRL_LOCK_SERVING, CTRL_STOP, CTRL_SHARD_NEXT, CTRL_WORKERS_DONE = 0, 1, 2, 3, 4, 5
CTRL_BARRIER = 6
CTRL_HEARTBEAT, CTRL_DONE_FLAGS, CTRL_STALENESS_HIST, CTRL_WORDS = 16, 56, 96, 128
CTRL_MAX_WORKERS = 40

_SIGNATURES = {
    # gemm
    "dk_tmap_encode_2d": (i32, [vp, vp, i32, i64, i64, i64, i32]),
    "dk_gemm_pick_bn": (i32, [i32]),
    "dk_gemm_mcast_cluster": (i32, [i32]),
    "dk_gemm_mcast_box_rows": (i32, [i32]),
    "dk_gemm_pick_bn2": (i32, [i32, i32]),
    "dk_gemm_pick_bn_splitk": (i32, [i32, i32, i32]),
    "dk_gemm_pick_splits_pair": (i32, [i32, i32, i32, i32]),
    "dk_gemm_tn": (i32, [vp, i64, vp, i64, C.POINTER(GemmEpilogue), i32, i32, i32, i32, i32, vp]),
    "dk_gemm_tn_ex": (i32, [vp, i64, vp, i64, C.POINTER(GemmEpilogue), i32, i32, i32, i32, i32, i32, vp]),
    "dk_gemm_pick_splits": (i32, [i32, i32, i32, i32, i32]),
    "dk_conv_gemm": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, C.POINTER(GemmEpilogue), i32, i32,
                           i32, vp]),
    "dk_conv_weight_flip": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "dk_conv_tma": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, C.POINTER(GemmEpilogue), i32, i32, vp]),
    "dk_conv_tma_supported": (i32, [i32, i32, i32, i32, i32]),
    "dk_conv_wgrad_tma": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, i32, vp, vp]),
    "dk_conv_wgrad_tma_supported": (i32, [i32, i32, i64, i64]),
    "dk_conv_pick_bn": (i32, [i32]),
    "dk_conv_gather_mode": (i32, [i32]),
    "dk_conv_wgrad": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, i32, i32, i32, vp]),
    "dk_engine_add_conv_wgrad": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, i32, i32]),
    # ps
    "dk_ps_commit": (i32, [vp, vp, vp, i64, f32, vp, vp, i32, u32, vp]),
    "dk_ps_pull": (i32, [vp, vp, vp, vp, i64, vp, vp, vp]),
    "dk_ps_exchange": (i32, [vp, vp, vp, vp, i64, f32, vp, vp, i32, u32, vp, vp]),
    "dk_ps_elastic": (i32, [vp, vp, vp, i64, f32, vp, i32, u32, vp]),
    "dk_ps_damped_exchange": (i32, [vp, vp, vp, vp, i64, f32, f32, vp, i32, u32, vp]),
    "dk_ps_ticket": (i32, [vp, vp, vp, vp]),
    "dk_ps_fetch_add": (i32, [vp, u32, vp, vp]),
    "dk_ps_barrier": (i32, [vp, i32, vp, vp, i32, vp]),
    "dk_ps_easgd_read": (i32, [vp, vp, vp, vp, i64, f32, vp]),
    "dk_ps_easgd_add": (i32, [vp, vp, i64, vp, i32, vp]),
    "dk_ps_lock_acquire": (i32, [vp, vp, vp]),
    "dk_ps_lock_release": (i32, [vp, vp, vp]),
    "dk_ps_average": (i32, [C.POINTER(vp), i32, i64, i64, vp]),
    "dk_ps_copy": (i32, [vp, vp, i64, vp]),
    # optim / loss / nn
    "dk_optim_step": (i32, [i32, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, f32, vp]),
    "dk_eamsgd_pre": (i32, [vp, vp, vp, vp, i64, f32, vp]),
    "dk_eamsgd_post": (i32, [vp, vp, vp, vp, i64, f32, vp]),
    "dk_cast_bf16": (i32, [vp, vp, i64, vp]),
    "dk_softmax_xent": (i32, [vp, i32, vp, vp, i32, i32, vp, i32, vp, i32, vp, vp, vp, i32, vp]),
    "dk_dense_softmax_head": (i32, [vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, vp, i32, vp, i32, f32, i32, vp, vp, i32,
                                    vp]),
    "dk_elementwise_loss": (i32, [i32, vp, vp, i32, i32, vp, i32, vp, i32, vp, vp, i32, vp]),
    "dk_input_stage": (i32, [vp, i32, i32, i32, f32, f32, vp, i32, vp, i32, vp, vp, i32, vp]),
    "dk_gemm_pull": (i32, [vp, i64, vp, i64, C.POINTER(GemmEpilogue), i32, i32, i32, vp, vp, vp, vp]),
    "dk_transpose_bf16": (i32, [vp, i32, i32, i32, vp, i32, vp]),
    "dk_rowsum_bf16": (i32, [vp, i32, i32, i32, vp, f32, vp]),
    "dk_colsum_bf16": (i32, [vp, i32, i32, i32, vp, f32, vp]),
    "dk_im2col": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "dk_col2im": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "dk_maxpool_fwd": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "dk_maxpool_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "dk_maxpool_bwd_ex": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "dk_col2im_ex": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "dk_relu_mask_bf16": (i32, [vp, vp, i64, vp]),
    "dk_bn_forward": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp]),
    "dk_bn_inference": (i32, [vp, i64, i32, vp, vp, vp, vp, f32, i32, vp, vp]),
    "dk_bn_backward": (i32, [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dk_gap_fwd": (i32, [vp, i32, i32, i32, vp, vp]),
    "dk_gap_bwd": (i32, [vp, i32, i32, i32, vp, vp]),
    "dk_add_bf16": (i32, [vp, vp, vp, i64, i32, vp]),
    "dk_label_index": (i32, [vp, i32, i32, f32, i32, vp, vp, vp, vp]),
    # fabric
    "dk_device_count": (i32, []),
    "dk_set_device": (i32, [i32]),
    "dk_fabric_alloc": (i32, [i64, C.POINTER(vp)]),
    "dk_fabric_free": (i32, [vp]),
    "dk_ipc_export": (i32, [vp, vp]),
    "dk_ipc_open": (i32, [vp, C.POINTER(vp)]),
    "dk_ipc_close": (i32, [vp]),
    "dk_can_access_peer": (i32, [i32, i32]),
    "dk_enable_peer_access": (i32, [i32, i32]),
    "dk_host_alloc_pinned": (i32, [i64, C.POINTER(vp)]),
    "dk_host_free_pinned": (i32, [vp]),
    "dk_host_register": (i32, [vp, i64]),
    "dk_host_unregister": (i32, [vp]),
    "dk_memcpy_async": (i32, [vp, vp, i64, i32, vp]),
    "dk_memset_async": (i32, [vp, i32, i64, vp]),
    "dk_stream_sync": (i32, [vp]),
    "dk_device_sync": (i32, []),
    "dk_build_info": (C.c_char_p, []),
    # engine
    "dk_engine_create": (vp, []),
    "dk_engine_destroy": (None, [vp]),
    "dk_engine_new_list": (i32, [vp]),
    "dk_engine_set_build_stream": (i32, [vp, i32]),
    "dk_engine_clear_list": (i32, [vp, i32]),
    "dk_engine_set_slot": (i32, [vp, i32, vp]),
    "dk_engine_add_op": (i32, [vp, i32, i32, C.POINTER(C.c_int64), i32, C.POINTER(f64), i32]),
    "dk_engine_add_conv_gemm": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, i32, i32, i32,
                                      C.POINTER(GemmEpilogue)]),
    "dk_engine_add_gemm": (i32, [vp, i32, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, C.POINTER(GemmEpilogue)]),
    "dk_engine_add_gemm_pull": (i32, [vp, i32, vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, C.POINTER(GemmEpilogue)]),
    "dk_engine_add_conv_wgrad_tma": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, i64, i32, vp]),
    "dk_engine_add_gemm_slot": (i32, [vp, i32, i32, i64, vp, i64, i32, i32, i32, i32, i32, i32, C.POINTER(GemmEpilogue)]),
    "dk_engine_add_bwd_update": (i32, [vp, i32, C.POINTER(BwdUpdateDesc)]),
    "dk_bwd_update": (i32, [C.POINTER(BwdUpdateDesc), vp]),
    "dk_bwd_update_desc_bytes": (i64, []),
    "dk_engine_run": (i32, [vp, i32, vp]),
    "dk_engine_list_size": (i32, [vp, i32]),
    "dk_engine_list_kernels": (i32, [vp, i32]),
    "dk_engine_launches": (i64, [vp]),
}


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    """Load (once) and return the native library; raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"native runtime {LIB_PATH} is missing: run `python build_native.py` "
                "(the sm_100a kernels are mandatory on the GPU path; there is no PyTorch fallback)")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.dk_bwd_update_desc_bytes() != C.sizeof(BwdUpdateDesc):
            raise RuntimeError("ctypes mirror of DkBwdUpdateDesc is out of sync with csrc/dense_fused.h")
        _lib = handle
    return _lib


def check(code: int, what: str = "native call") -> None:
    if code != 0:
        raise RuntimeError(f"{what} failed with code {code}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, ``None`` -> 0."""
    return 0 if t is None else int(t.data_ptr())


def current_stream() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


class CudaBuffer:
    """Expose a raw device allocation to torch through ``__cuda_array_interface__``."""

    def __init__(self, address: int, nbytes: int, typestr: str, shape, owner=None):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(address), False), "version": 2,
            "strides": None,
        }
        self._owner = owner
        self.nbytes = nbytes


def as_tensor(address: int, shape, dtype, device_index: int, owner=None):
    """Wrap raw device memory (e.g. a peer-mapped IPC pointer) as a torch tensor view."""
    import numpy as np
    import torch

    np_dtype = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.bfloat16: None,
                torch.int64: "<i8", torch.uint32 if hasattr(torch, "uint32") else None: "<u4"}.get(dtype)
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    if np_dtype is None:  # bf16: view through int16
        nbytes = int(np.prod(shape)) * 2
        buf = CudaBuffer(address, nbytes, "<i2", shape, owner)
        return torch.as_tensor(buf, device=torch.device("cuda", device_index)).view(torch.bfloat16)
    nbytes = int(np.prod(shape)) * int(np_dtype[-1])
    buf = CudaBuffer(address, nbytes, np_dtype, shape, owner)
    return torch.as_tensor(buf, device=torch.device("cuda", device_index))
