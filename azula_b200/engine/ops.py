r"""Python face of the sm_100a kernels: torch tensors in, C-ABI calls out (``include/azb.h``).

Activation convention of the native backbone: NHWC bf16, i.e. tensors of shape
``(N, H, W, C)`` whose last dimension is contiguous; the pixel stride ``ld`` may exceed ``C``
so that a producer can write straight into a channel slice of a concatenation buffer.
Nothing here falls back to torch arithmetic: a missing library raises.
"""

from __future__ import annotations

import torch

from ctypes import c_float, c_int, c_int64, c_void_p
from dataclasses import dataclass
from torch import Tensor

from .. import _lib

_lib.register({
    "azb_conv_gemm_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64,
         c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p],
    ),
})


def _ld(t: Tensor) -> int:
    """Pixel stride (elements) of an NHWC view; checks the layout the kernels assume."""
    assert t.stride(-1) == 1, "channel dimension must be contiguous"
    if t.ndim == 4:
        n, h, w, _ = t.shape
        ld = t.stride(2)
        assert t.stride(1) == ld * w and (n == 1 or t.stride(0) == ld * w * h), "not a dense NHWC view"
        return ld
    assert t.ndim == 2
    return t.stride(0)


@dataclass
class PackedConv:
    r"""Weights of one conv / linear layer in kernel layout."""

    w: Tensor  # bf16 (c_out_rows, taps, k_per_tap)
    bias: Tensor | None  # fp32 (c_out,)
    c_in: int
    c_out: int
    taps: int

    @property
    def c_out_rows(self) -> int:
        return self.w.shape[0]

    @property
    def k_per_tap(self) -> int:
        return self.w.shape[2]


def pack_conv(weight: Tensor, bias: Tensor | None) -> PackedConv:
    r"""(C_out, C_in[, kh, kw]) fp32 -> bf16 (C_out_rows, taps, K_tap): tap-major K, both padded with
    zeros (K_tap to a multiple of 64, rows to the N tile) so that TMA boxes never straddle taps."""
    if weight.ndim == 3:  # Conv1d k=1
        weight = weight[..., 0]
    if weight.ndim == 2:
        weight = weight[:, :, None, None]
    c_out, c_in, kh, kw = weight.shape
    assert (kh, kw) in ((1, 1), (3, 3))
    taps = kh * kw
    k_pad = -(-c_in // 64) * 64
    tile = 128 if c_out >= 128 else 64 if c_out >= 64 else 32 if c_out >= 32 else 16
    rows = -(-c_out // tile) * tile
    w = torch.zeros(rows, taps, k_pad, dtype=torch.bfloat16, device=weight.device)
    w[:c_out, :, :c_in] = weight.permute(0, 2, 3, 1).reshape(c_out, taps, c_in).to(torch.bfloat16)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(w=w.contiguous(), bias=b, c_in=c_in, c_out=c_out, taps=taps)


def conv(x: Tensor, pc: PackedConv, out: Tensor | None = None, residual: Tensor | None = None,
         nchw_f32: bool = False) -> Tensor:
    r"""3x3 (pad 1) / 1x1 convolution or linear layer on tcgen05 (``azb_conv_gemm_bf16``).

    x: (N, H, W, C_in) or (rows, C_in) bf16.  Returns bf16 NHWC (or fp32 NCHW when ``nchw_f32``).
    """
    assert x.dtype == torch.bfloat16 and x.is_cuda
    if x.ndim == 2:
        n, h, w = 1, 1, x.shape[0]
    else:
        n, h, w, _ = x.shape
    assert x.shape[-1] == pc.c_in, (x.shape, pc.c_in)
    if out is None:
        if nchw_f32:
            out = torch.empty((n, pc.c_out, h, w), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((*x.shape[:-1], pc.c_out), dtype=torch.bfloat16, device=x.device)
    out_ld = 0 if nchw_f32 else _ld(out)
    _lib.check(
        _lib.lib().azb_conv_gemm_bf16(
            x.data_ptr(), n, h, w, pc.c_in, _ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps, pc.k_per_tap,
            _lib.ptr(pc.bias), _lib.ptr(residual), 0 if residual is None else _ld(residual), out.data_ptr(), out_ld,
            1 if nchw_f32 else 0, _lib.stream_ptr(x.device),
        ),
        "azb_conv_gemm_bf16",
    )
    return out
