"""Python entry to the tcgen05 GEMM (``csrc/gemm_tcgen05.cu``) for tests and micro-benchmarks."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _native as N


def gemm_tn(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
            bias_along_m: bool = False, relu: bool = False, mask: Optional[torch.Tensor] = None, out_fp32: bool = False,
            alpha: float = 1.0, bn: int = 0, tf32: bool = False, splits: int = 1, persistent: bool = False,
            pair: bool = False) -> torch.Tensor:
    """``D[M, N] = epilogue(A B^T)``.

    ``a`` is ``[M, K]`` (K-major) or, with ``a_mn``, the MN-major storage ``[K, M]``; same for ``b``
    (``[N, K]`` / ``[K, N]``).  bf16 inputs (fp32 with ``tf32``), fp32 accumulation in TMEM.
    """
    assert a.is_cuda and b.is_cuda and a.is_contiguous() and b.is_contiguous()
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    Nn = b.shape[1] if b_mn else b.shape[0]
    alloc = torch.zeros if splits != 1 else torch.empty
    out = alloc(M, Nn, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=a.device)
    ep = N.GemmEpilogue()
    ep.bias = N.ptr(bias) or None
    ep.bias_along_m = int(bias_along_m)
    ep.act = int(relu)
    if mask is not None:
        ep.mask, ep.ld_mask = mask.data_ptr(), mask.stride(0)
    ep.d, ep.ldd, ep.d_fp32, ep.alpha = out.data_ptr(), Nn, int(out_fp32), alpha
    flags = (N.GEMM_TF32 if tf32 else 0) | (N.GEMM_A_MN if a_mn else 0) | (N.GEMM_B_MN if b_mn else 0)
    if pair:  # cta_group::2: two CTAs share one 256 x bn tile
        flags |= N.GEMM_PAIR
        bn = bn or (256 if Nn > 128 else 128)
    if persistent:
        flags |= N.GEMM_PERSISTENT
        bn = bn or (256 if Nn > 128 else (128 if Nn > 64 else 64))
    if splits == 0:
        splits = N.lib().dk_gemm_pick_splits(M, Nn, K, bn or N.lib().dk_gemm_pick_bn(Nn), int(tf32))
    N.check(N.lib().dk_gemm_tn_ex(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), C.byref(ep), M, Nn, K, flags, bn,
                                  splits, C.c_void_p(N.current_stream())), "dk_gemm_tn_ex")
    return out
