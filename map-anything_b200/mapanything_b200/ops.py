"""Tensor-level wrappers over the C-ABI kernels.

PyTorch is used for device memory and streams only: every function here hands raw device pointers and
the current CUDA stream to `libmapanything_b200.so`.  `LAUNCHES` counts kernel launches issued through
this module (bench.py reports it as `gpu_launches`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import MA_ACT_GELU, MA_ACT_NONE, MA_ACT_RELU, MA_BF16, MA_F32, GemmEpilogue, check  # noqa: F401

LAUNCHES = 0


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return MA_BF16
    if t.dtype == torch.float32:
        return MA_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: Optional[torch.Tensor], n: int, name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n or not t.is_cuda:
        raise ValueError(f"{name} must be a contiguous CUDA float32 tensor with {n} elements")
    return t


def gemm(
    x: torch.Tensor,
    w: torch.Tensor,
    out: torch.Tensor,
    *,
    bias: Optional[torch.Tensor] = None,
    act: int = MA_ACT_NONE,
    colscale: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    residual_row_mod: int = 0,
    out_relu: Optional[torch.Tensor] = None,
    rows_per_group_in: int = 0,
    rows_per_group_out: int = 0,
    row_offset_out: int = 0,
    block_n: int = 0,
) -> torch.Tensor:
    """out = epilogue(x @ w.T); x [M,K] bf16, w [N,K] bf16 (row strides may exceed K). See ma_gemm_bf16."""
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("gemm operands must be bfloat16")
    if x.dim() != 2 or w.dim() != 2 or x.stride(1) != 1 or w.stride(1) != 1 or out.stride(-1) != 1:
        raise ValueError("gemm operands must be 2-D with contiguous rows")
    M, K = x.shape
    N, Kw = w.shape
    if K != Kw:
        raise ValueError(f"K mismatch: x {tuple(x.shape)} vs w {tuple(w.shape)}")
    ep = GemmEpilogue()
    ep.out = out.data_ptr()
    ep.ldo = out.stride(-2)
    ep.out_dtype = _dt(out)
    ep.act = act
    ep.bias = _ptr(_f32c(bias, N, "bias"))
    ep.colscale = _ptr(_f32c(colscale, N, "colscale"))
    if residual is not None:
        ep.residual = residual.data_ptr()
        ep.ldr = residual.stride(-2)
        ep.residual_dtype = _dt(residual)
    ep.residual_row_mod = residual_row_mod
    if out_relu is not None:
        if out_relu.dtype != torch.bfloat16:
            raise TypeError("out_relu must be bfloat16")
        ep.out_relu = out_relu.data_ptr()
        ep.ldo_relu = out_relu.stride(-2)
    ep.rows_per_group_in = rows_per_group_in
    ep.rows_per_group_out = rows_per_group_out
    ep.row_offset_out = row_offset_out
    lib = _lib.load()
    check(
        lib.ma_gemm_bf16(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), M, N, K, C.byref(ep), block_n, _stream()),
        "ma_gemm_bf16",
    )
    _count()
    return out


def attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    out: torch.Tensor,
    *,
    num_heads: int,
    num_seqs: int,
    q_len: int,
    kv_len: int,
    q_seq_stride: Optional[int] = None,
    kv_seq_stride: Optional[int] = None,
    scale: Optional[float] = None,
) -> torch.Tensor:
    """softmax(q k^T * scale) v per (sequence, head), head_dim 64.

    q/k/v/out are 2-D token-major bf16 views whose rows may be strided column slices of a wider matrix
    (e.g. qkv[:, :D], qkv[:, D:2D], qkv[:, 2D:]); no head permutation is materialised. See ma_attention_fwd."""
    for t in (q, k, v, out):
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("attention operands must be 2-D bf16 with contiguous rows")
    if q_seq_stride is None:
        q_seq_stride = q_len
    if kv_seq_stride is None:
        kv_seq_stride = kv_len
    if scale is None:
        scale = 64 ** -0.5
    lib = _lib.load()
    # Column offsets are folded into the base pointers; the tensor map width is the row stride, which
    # always covers [col0, col0 + heads*64) of the parent matrix when the view is a column slice of it.
    check(
        lib.ma_attention_fwd(
            q.data_ptr(), q.stride(0), q.shape[0], 0,
            k.data_ptr(), k.stride(0), k.shape[0], 0,
            v.data_ptr(), v.stride(0), 0,
            out.data_ptr(), out.stride(0), 0,
            num_seqs, num_heads, q_len, kv_len, q_seq_stride, kv_seq_stride, float(scale), _stream(),
        ),
        "ma_attention_fwd",
    )
    _count()
    return out
