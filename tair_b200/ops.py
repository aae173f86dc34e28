"""Thin torch-tensor front end over the C ABI (``include/tair_b200.h``).

torch supplies device memory and the current stream; all arithmetic happens in
the sm_100a kernels of ``libtair_b200.so``.  Every function raises if handed a
CPU tensor — there is deliberately no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SILU, Epilogue, TairError  # noqa: F401

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TairError(f"{name}: expected a CUDA tensor (tair_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TairError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: torch.Tensor, name: str) -> tuple[torch.Tensor, int]:
    """2-D view with unit inner stride; returns (tensor, row stride in elements)."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise TairError(f"{name}: expected a 2-D tensor with contiguous rows, got {tuple(t.shape)} / {t.stride()}")
    return t, t.stride(0)


def _make_epilogue(out, M, n_out, bias, residual, rowgroup, rows_per_group, act) -> Epilogue:
    e = Epilogue()
    _cuda(out, "out")
    if out.dtype not in (BF16, torch.float32):
        raise TairError("out must be bf16 or fp32")
    o2, ldc = _rows(out, "out")
    if o2.shape[0] != M or o2.shape[1] != n_out:
        raise TairError(f"out has shape {tuple(o2.shape)}, expected {(M, n_out)}")
    e.out, e.ldc = o2.data_ptr(), ldc
    e.out_fp32 = 1 if out.dtype == torch.float32 else 0
    e.act = act
    if bias is not None:
        _cuda(bias, "bias", torch.float32)
        if not bias.is_contiguous():
            raise TairError("bias must be contiguous")
        e.bias = bias.data_ptr()
    if residual is not None:
        _cuda(residual, "residual", BF16)
        r2, ldr = _rows(residual, "residual")
        if tuple(r2.shape) != (M, n_out):
            raise TairError(f"residual has shape {tuple(r2.shape)}, expected {(M, n_out)}")
        e.residual, e.ldr = r2.data_ptr(), ldr
    if rowgroup is not None:
        _cuda(rowgroup, "rowgroup", torch.float32)
        g2, ldg = _rows(rowgroup, "rowgroup")
        if rows_per_group <= 0 or g2.shape[0] * rows_per_group < M or g2.shape[1] < n_out:
            raise TairError("rowgroup shape does not cover the output")
        e.rowgroup, e.ldg, e.rows_per_group = g2.data_ptr(), ldg, rows_per_group
    return e


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, residual=None, rowgroup=None, rows_per_group=0,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, out_dtype=BF16) -> torch.Tensor:
    """``epilogue(a @ w.T)``; a [M,K] bf16, w [N,K] bf16 (nn.Linear layout)."""
    _cuda(a, "a", BF16), _cuda(w, "w", BF16)
    a2, lda = _rows(a, "a")
    w2, ldw = _rows(w, "w")
    M, K = a2.shape
    N = w2.shape[0]
    if w2.shape[1] != K:
        raise TairError(f"gemm: K mismatch {K} vs {w2.shape[1]}")
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype)
    e = _make_epilogue(out, M, n_out, bias, residual, rowgroup, rows_per_group, act)
    rc = _lib.lib().tair_gemm_bf16(a2.data_ptr(), lda, w2.data_ptr(), ldw, M, N, K, C.byref(e), _stream())
    _lib.check(rc, "tair_gemm_bf16")
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, stride: int = 1, bias=None, residual=None, rowgroup=None,
            rows_per_group=0, act: int = ACT_NONE, out: Optional[torch.Tensor] = None,
            out_dtype=BF16) -> torch.Tensor:
    """3x3 / pad 1 convolution on channels-last bf16: x [B,H,W,Cin], w [Cout, 9*Cin] -> [B,Ho,Wo,Cout]."""
    _cuda(x, "x", BF16), _cuda(w, "w", BF16)
    if x.dim() != 4 or not x.is_contiguous():
        raise TairError("conv3x3: x must be a contiguous [B,H,W,C] tensor")
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    if w.dim() != 2 or w.shape[1] != 9 * Cin or not w.is_contiguous():
        raise TairError(f"conv3x3: w must be contiguous [Cout, 9*Cin], got {tuple(w.shape)}")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    M = B * Ho * Wo
    if out is None:
        out = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=out_dtype)
    o2 = out.view(M, -1) if out.dim() == 4 else out
    r2 = residual.view(M, -1) if (residual is not None and residual.dim() == 4) else residual
    e = _make_epilogue(o2, M, Cout, bias, r2, rowgroup, rows_per_group, act)
    rc = _lib.lib().tair_conv3x3_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, stride, C.byref(e), _stream())
    _lib.check(rc, "tair_conv3x3_bf16")
    return out


def launch_count() -> int:
    return int(_lib.lib().tair_launch_count())


def reset_launch_count() -> None:
    _lib.lib().tair_launch_count_reset()


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, B: int, H: int, Lq: int, Lk: int,
              head_dim: int = 64, scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v per (batch, head).

    q [B*Lq, >=H*D], k/v [B*Lk, >=H*D] are 2-D bf16 views with unit inner stride (column slices of a fused
    projection buffer are fine); head h lives in columns [h*D, (h+1)*D).  Returns [B*Lq, H*D] bf16.
    """
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _cuda(t, n, BF16)
    q2, ldq = _rows(q, "q")
    k2, ldk = _rows(k, "k")
    v2, ldv = _rows(v, "v")
    if q2.shape[0] != B * Lq or k2.shape[0] != B * Lk or v2.shape[0] != B * Lk:
        raise TairError("attention: row counts do not match B*Lq / B*Lk")
    if out is None:
        out = torch.empty((B * Lq, H * head_dim), device=q.device, dtype=BF16)
    o2, ldo = _rows(out, "out")
    if scale is None:
        scale = head_dim ** -0.5
    rc = _lib.lib().tair_attention_bf16(q2.data_ptr(), ldq, k2.data_ptr(), ldk, v2.data_ptr(), ldv, o2.data_ptr(), ldo,
                                        B, H, Lq, Lk, head_dim, float(scale), _stream())
    _lib.check(rc, "tair_attention_bf16")
    return out
