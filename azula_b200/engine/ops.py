r"""Python face of the sm_100a kernels: torch tensors in, C-ABI calls out (``include/azb.h``).

Activation convention of the native backbone: NHWC bf16, i.e. tensors of shape
``(N, H, W, C)`` whose last dimension is contiguous; the pixel stride ``ld`` may exceed ``C``
so that a producer can write straight into a channel slice of a concatenation buffer.
Nothing here falls back to torch arithmetic: a missing library raises.
"""

from __future__ import annotations

import math
import os
import torch

import ctypes

from ctypes import POINTER, byref, c_float, c_int, c_int32, c_int64, c_void_p
from dataclasses import dataclass
from torch import Tensor

from .. import _lib

class AzbConv(ctypes.Structure):
    r"""``AzbConv`` of ``include/azb.h``: every option of the convolution / linear kernel (zero = not used)."""

    _fields_ = [
        ("act", c_void_p), ("n", c_int64), ("h", c_int64), ("w", c_int64), ("c_in", c_int64), ("act_ld", c_int64),
        ("wpack", c_void_p), ("c_out", c_int64), ("c_out_rows", c_int64), ("k_per_tap", c_int64),
        ("taps", c_int32), ("stride", c_int32), ("act_fn", c_int32), ("out_mode", c_int32), ("stat_gran", c_int32),
        ("res_up", c_int32),
        ("bias", c_void_p), ("gate", c_void_p), ("gate_ld", c_int64), ("gate_rows", c_int64),
        ("residual", c_void_p), ("res_ld", c_int64), ("out", c_void_p), ("out_ld", c_int64), ("colsum", c_void_p),
        ("act2", c_void_p), ("c_in2", c_int64), ("act2_ld", c_int64), ("k2", c_int64), ("gn_acc", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("in_coef", c_void_p), ("in_silu", c_int32), ("in_up", c_int32),
        ("in_norm", c_int32), ("in_eps", c_float), ("in_rowstat", c_void_p), ("in_mod", c_void_p), ("in_mod_ld", c_int64),
        ("rowstat", c_void_p), ("out_up", c_int32), ("reserved_", c_int32),
    ]


class AzbConvChoice(ctypes.Structure):
    r"""``AzbConvChoice`` of ``include/azb.h``: what the convolution launcher would do for a descriptor."""

    _fields_ = [("halo", c_int32), ("pair", c_int32), ("lean", c_int32), ("block_n", c_int32), ("splits", c_int32),
                ("tiles", c_int32), ("epi", c_int32)]


_lib.register({
    "azb_conv_bf16": (c_int, [POINTER(AzbConv), c_void_p]),
    "azb_conv_choice": (c_int, [POINTER(AzbConv), POINTER(AzbConvChoice)]),
    "azb_gn_coef_f32": (
        c_int,
        [c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_float, c_void_p,
         c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p],
    ),
    "azb_gn_pool_acc_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
         c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    ),
    "azb_conv_tuning": (c_int, [c_int, c_int]),
    "azb_debug_trace": (c_int, [c_void_p]),
    "azb_conv_tf32": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_void_p,
         c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p],
    ),
    "azb_gn_stats_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_int64,
                                 c_void_p]),
    "azb_gn_apply_f32": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
         c_void_p, c_int64, c_int, c_int, c_void_p],
    ),
    "azb_nchw_to_nhwc_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "azb_attention_qknorm_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p],
    ),
    "azb_attention_f16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p],
    ),
    "azb_zero_bytes": (c_int, [c_void_p, c_int64, c_void_p]),
    "azb_gn_apply_acc_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
         c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    ),
    "azb_conv_gemm_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64,
         c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p],
    ),
    "azb_conv_gemm_stats_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64,
         c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p],
    ),
    "azb_conv_skip_stats_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
         c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p],
    ),
    "azb_conv_colsum_rows": (c_int, [c_int64, c_int64, c_int64, POINTER(c_int64), POINTER(c_int64)]),
    "azb_gn_finalize_f32": (
        c_int,
        [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p,
         c_void_p],
    ),
    "azb_gn_stats_workspace": (c_int, [c_int64, c_int64, c_int64, c_int64, POINTER(c_int64)]),
    "azb_gn_stats_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "azb_gn_apply_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p],
    ),
    "azb_attention_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
         c_void_p],
    ),
    "azb_attention_mma_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
         c_void_p],
    ),
    "azb_im2col3x3_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "azb_timestep_features_f32": (c_int, [c_void_p, c_int, c_int64, c_int64, c_float, c_void_p, c_void_p]),
    "azb_linear_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "azb_add_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "azb_conv2d_bf16": (
        c_int,
        [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int64,
         c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p,
         c_int, c_void_p],
    ),
    "azb_rownorm_mod_bf16": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_float, c_void_p, c_int64, c_int64, c_void_p],
    ),
    "azb_segment_rmsnorm_bf16": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p]),
    "azb_qk_norm_rope_bf16": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_float, c_void_p, c_int64, c_void_p]),
    "azb_patchify_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "azb_unpatchify_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "azb_linear_gather_f32": (
        c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p],
    ),
})

ACT = {None: 0, "none": 0, "silu": 1, "relu": 2, "relu2": 3}
NORM = {"layer": 0, "rms": 1}


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
    tile = 128 if c_out >= 128 else 64 if c_out >= 64 else 32 if c_out >= 32 else 16  # row padding; the kernel picks the N tile
    rows = -(-c_out // tile) * tile
    w = torch.zeros(rows, taps, k_pad, dtype=torch.bfloat16, device=weight.device)
    w[:c_out, :, :c_in] = weight.permute(0, 2, 3, 1).reshape(c_out, taps, c_in).to(torch.bfloat16)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(w=w.contiguous(), bias=b, c_in=c_in, c_out=c_out, taps=taps)


def pack_conv_up(weight: Tensor, bias: Tensor | None) -> PackedConv:
    r"""Weights of ``conv3x3(upsample2x(z))`` as four 2 x 2 convolutions of ``z`` (one per output phase (dy, dx)):
    (C_out, C_in, 3, 3) fp32 -> bf16 (C_out_rows, 16, K_tap), entry ``(dy * 2 + dx) * 4 + a * 2 + b`` = the sum (in
    fp32) of the taps (kh, kw) that read half-resolution pixel (i - 1 + dy + a, j - 1 + dx + b) for output pixel
    (2 i + dy, 2 j + dx): rows dy = 0: {0}, {1, 2}; dy = 1: {0, 1}, {2}; columns alike.  ``AzbConv::in_up = 2``."""
    c_out, c_in, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    w = weight.detach().to(torch.float32)
    sel = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    taps = []
    for dy in (0, 1):
        for dx in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    taps.append(sum(w[:, :, i, j] for i in sel[dy][a] for j in sel[dx][b]))
    k_pad = -(-c_in // 64) * 64
    tile = 128 if c_out >= 128 else 64 if c_out >= 64 else 32 if c_out >= 32 else 16
    rows = -(-c_out // tile) * tile
    out = torch.zeros(rows, 16, k_pad, dtype=torch.bfloat16, device=weight.device)
    out[:c_out, :, :c_in] = torch.stack(taps, dim=1).to(torch.bfloat16)
    b_ = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(w=out.contiguous(), bias=b_, c_in=c_in, c_out=c_out, taps=16)


@dataclass
class PackedConvSkip:
    r"""conv3x3 and the 1x1 skip connection of a ResBlock as one GEMM: weights [rows][9 * k_per_tap + k2]."""

    w: Tensor
    bias: Tensor
    c_in: int
    c_in2: int
    c_out: int
    k_per_tap: int
    k2: int

    @property
    def c_out_rows(self) -> int:
        return self.w.shape[0]


def pack_conv_skip(conv: PackedConv, skip: PackedConv) -> PackedConvSkip:
    assert conv.taps == 9 and skip.taps == 1 and conv.c_out == skip.c_out and conv.c_out_rows == skip.c_out_rows
    rows = conv.c_out_rows
    w = torch.cat((conv.w.reshape(rows, -1), skip.w.reshape(rows, -1)), dim=1).contiguous()
    return PackedConvSkip(w=w, bias=(conv.bias + skip.bias).contiguous(), c_in=conv.c_in, c_in2=skip.c_in, c_out=conv.c_out,
                          k_per_tap=conv.k_per_tap, k2=skip.k_per_tap)


def conv_skip(x: Tensor, x2: Tensor, pc: PackedConvSkip, out: Tensor | None = None, colsum: Tensor | None = None) -> Tensor:
    r"""``conv3x3(x) + conv1x1(x2) + bias`` in one launch (``azb_conv_skip_stats_bf16``); both NHWC bf16."""
    n, h, w, _ = x.shape
    assert x2.shape[:3] == x.shape[:3] and x.shape[-1] == pc.c_in and x2.shape[-1] == pc.c_in2
    if out is None:
        out = torch.empty(n, h, w, pc.c_out, dtype=torch.bfloat16, device=x.device)
    _lib.check(
        _lib.lib().azb_conv_skip_stats_bf16(
            x.data_ptr(), n, h, w, pc.c_in, _ld(x), x2.data_ptr(), pc.c_in2, _ld(x2), pc.w.data_ptr(), pc.c_out,
            pc.c_out_rows, pc.k_per_tap, pc.k2, pc.bias.data_ptr(), out.data_ptr(), _ld(out), _lib.ptr(colsum),
            1 if colsum is None else pc.c_out // colsum.shape[1], _lib.stream_ptr(x.device),
        ),
        "azb_conv_skip_stats_bf16",
    )
    return out


def colsum_rows(n: int, h: int, w: int) -> tuple[int, bool]:
    r"""(rows of the column-sum buffer of an (n, h, w) convolution output, whether it can feed
    :func:`gn_finalize`)."""
    rows, ok = c_int64(0), c_int64(0)
    _lib.check(_lib.lib().azb_conv_colsum_rows(n, h, w, byref(rows), byref(ok)), "azb_conv_colsum_rows")
    return rows.value, bool(ok.value)


def gn_finalize(parts: list[tuple[Tensor, int]], n: int, h: int, w: int, stats: Tensor | None = None,
                groups: int = 32, eps: float = 1e-5) -> Tensor:
    r"""GroupNorm statistics (N, groups, 2) from the column sums of one or two convolutions
    (``azb_gn_finalize_f32``); ``parts`` = [(colsum, channels), ...] in channel order."""
    (a, ca), (b, cb) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
    if stats is None:
        stats = torch.empty(n, groups, 2, dtype=torch.float32, device=a.device)
    gran = lambda t, c: 1 if t is None else c // t.shape[1]  # noqa: E731  (channels per colsum entry)
    _lib.check(
        _lib.lib().azb_gn_finalize_f32(a.data_ptr(), ca, gran(a, ca), _lib.ptr(b), cb, gran(b, cb), n, h, w, groups, eps,
                                       stats.data_ptr(), _lib.stream_ptr(a.device)),
        "azb_gn_finalize_f32",
    )
    return stats


def conv(x: Tensor, pc: PackedConv, out: Tensor | None = None, residual: Tensor | None = None,
         nchw_f32: bool = False, colsum: Tensor | None = None) -> Tensor:
    r"""3x3 (pad 1) / 1x1 convolution or linear layer on tcgen05 (``azb_conv_gemm_bf16``).

    x: (N, H, W, C_in) or (rows, C_in) bf16.  Returns bf16 NHWC (or fp32 NCHW when ``nchw_f32``).
    With ``colsum`` (fp32 (rows, C_out, 2) or (rows, C_out / 8, 2), see :func:`colsum_rows`) the epilogue
    also emits the per-channel (per-8-channel-block) sums the consuming GroupNorm needs
    (``azb_conv_gemm_stats_bf16``).
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
    if colsum is not None:
        assert not nchw_f32 and colsum.dtype == torch.float32 and colsum.is_contiguous()
        _lib.check(
            _lib.lib().azb_conv_gemm_stats_bf16(
                x.data_ptr(), n, h, w, pc.c_in, _ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps,
                pc.k_per_tap, _lib.ptr(pc.bias), _lib.ptr(residual), 0 if residual is None else _ld(residual),
                out.data_ptr(), out_ld, colsum.data_ptr(), pc.c_out // colsum.shape[1], _lib.stream_ptr(x.device),
            ),
            "azb_conv_gemm_stats_bf16",
        )
        return out
    _lib.check(
        _lib.lib().azb_conv_gemm_bf16(
            x.data_ptr(), n, h, w, pc.c_in, _ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps, pc.k_per_tap,
            _lib.ptr(pc.bias), _lib.ptr(residual), 0 if residual is None else _ld(residual), out.data_ptr(), out_ld,
            1 if nchw_f32 else 0, _lib.stream_ptr(x.device),
        ),
        "azb_conv_gemm_bf16",
    )
    return out


GN_GROUPS = 32
GN_EPS = 1e-5


class GroupNormScratch:
    r"""Reusable workspace of the GroupNorm reduction (partials, per-image counters, statistics)."""

    def __init__(self, device, max_n: int = 64, max_partial: int = 1 << 22) -> None:
        self.partial = torch.empty(max_partial, dtype=torch.float32, device=device)
        self.counters = torch.zeros(max_n, dtype=torch.int32, device=device)

    def need(self, n: int, hw: int, c: int, groups: int) -> None:
        want = c_int64(0)
        _lib.check(_lib.lib().azb_gn_stats_workspace(n, hw, c, groups, byref(want)))
        if want.value > self.partial.numel() or n > self.counters.numel():
            raise RuntimeError("GroupNormScratch too small")


def gn_stats(x: Tensor, scratch: GroupNormScratch, stats: Tensor | None = None, groups: int = GN_GROUPS,
             eps: float = GN_EPS) -> Tensor:
    r"""(N, H, W, C) or (N, T, C) bf16 -> stats (N, groups, 2) = (mean, rstd) in fp32."""
    n, c = x.shape[0], x.shape[-1]
    hw = math.prod(x.shape[1:-1])
    scratch.need(n, hw, c, groups)
    if stats is None:
        stats = torch.empty(n, groups, 2, dtype=torch.float32, device=x.device)
    _lib.check(
        _lib.lib().azb_gn_stats_bf16(
            x.data_ptr(), _ld(x), n, hw, c, groups, eps, scratch.partial.data_ptr(), stats.data_ptr(),
            scratch.counters.data_ptr(), _lib.stream_ptr(x.device),
        ),
        "azb_gn_stats_bf16",
    )
    return stats


def gn_apply(x: Tensor, out: Tensor | None = None, stats: Tensor | None = None, gamma: Tensor | None = None,
             beta: Tensor | None = None, scale_shift: Tensor | None = None, ss_step: Tensor | None = None,
             ss_step_stride: int = 0, silu: bool = True, mode: int = 0, groups: int = GN_GROUPS) -> Tensor:
    r"""Normalise + modulate + activate + resample in one pass (``azb_gn_apply_bf16``).

    x (N, H, W, C) bf16.  ``scale_shift`` is fp32 (rows, 2C) with rows in {1, N} (or a per-step table
    indexed by the device counter ``ss_step``).  ``mode``: 0 same, 1 nearest x2, 2 average pool 2x2.
    """
    n, h, w, c = x.shape
    ho, wo = (h * 2, w * 2) if mode == 1 else (h // 2, w // 2) if mode == 2 else (h, w)
    if out is None:
        out = torch.empty(n, ho, wo, c, dtype=torch.bfloat16, device=x.device)
    ss_stride = 0
    if scale_shift is not None:
        assert scale_shift.dtype == torch.float32 and scale_shift.shape[-1] == 2 * c and scale_shift.is_contiguous()
        if ss_step is None and scale_shift.ndim == 2 and scale_shift.shape[0] == n and n > 1:
            ss_stride = 2 * c
    _lib.check(
        _lib.lib().azb_gn_apply_bf16(
            x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), n, h, w, c, groups, _lib.ptr(stats), _lib.ptr(gamma),
            _lib.ptr(beta), _lib.ptr(scale_shift), ss_stride, _lib.ptr(ss_step), ss_step_stride, int(silu), mode,
            _lib.stream_ptr(x.device),
        ),
        "azb_gn_apply_bf16",
    )
    return out


def attention(qkv: Tensor, heads: int, new_order: bool = False, out: Tensor | None = None, kernel: str = "auto") -> Tensor:
    r"""(N, T, 3C) bf16 -> (N, T, C) bf16, legacy (per-head q|k|v) or new (q|k|v blocks) channel order.
    ``kernel="mma"`` forces the warp-level kernel (``"auto"``: tcgen05 for head width 64)."""
    n, t, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    if out is None:
        out = torch.empty(n, t, c, dtype=torch.bfloat16, device=qkv.device)
    hs, kd, vd = (d, c, 2 * c) if new_order else (3 * d, d, 2 * d)
    fn = _lib.lib().azb_attention_mma_bf16 if kernel == "mma" else _lib.lib().azb_attention_bf16
    _lib.check(
        fn(qkv.data_ptr(), qkv.stride(1), out.data_ptr(), out.stride(1), n, t, heads, d, hs, kd, vd,
           _lib.stream_ptr(qkv.device)),
        "azb_attention_bf16",
    )
    return out


def im2col3x3(x: Tensor, k_pad: int = 64, out: Tensor | None = None) -> Tensor:
    r"""fp32 NCHW network input -> bf16 (N, H, W, k_pad) patches for the first 3x3 convolution."""
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(n, h, w, k_pad, dtype=torch.bfloat16, device=x.device)
    _lib.check(
        _lib.lib().azb_im2col3x3_f32(x.data_ptr(), out.data_ptr(), n, c, h, w, k_pad, _lib.stream_ptr(x.device)),
        "azb_im2col3x3_f32",
    )
    return out


def pack_first_conv(weight: Tensor, bias: Tensor, k_pad: int = 64) -> PackedConv:
    r"""(C_out, C, 3, 3) -> a 1-tap GEMM weight over the im2col patches (k = tap*C + c)."""
    c_out, c = weight.shape[:2]
    flat = torch.zeros(c_out, k_pad, dtype=torch.float32, device=weight.device)
    flat[:, : 9 * c] = weight.permute(0, 2, 3, 1).reshape(c_out, 9 * c)
    return pack_conv(flat, bias)


def timestep_features(t: Tensor, dim: int, out: Tensor | None = None, max_period: float = 10000.0) -> Tensor:
    rows = t.numel()
    assert t.dtype in (torch.int64, torch.float32) and t.is_contiguous()
    if out is None:
        out = torch.empty(rows, dim, dtype=torch.float32, device=t.device)
    _lib.check(
        _lib.lib().azb_timestep_features_f32(
            t.data_ptr(), _lib.DTYPE_CODE[t.dtype], rows, dim, max_period, out.data_ptr(), _lib.stream_ptr(t.device)
        ),
        "azb_timestep_features_f32",
    )
    return out


def linear_f32(x: Tensor, weight: Tensor, bias: Tensor | None, silu_in: bool = False, out: Tensor | None = None) -> Tensor:
    m, k = x.shape
    nn_ = weight.shape[0]
    assert x.dtype == torch.float32 and weight.dtype == torch.float32 and x.is_contiguous() and weight.is_contiguous()
    if out is None:
        out = torch.empty(m, nn_, dtype=torch.float32, device=x.device)
    _lib.check(
        _lib.lib().azb_linear_f32(
            x.data_ptr(), weight.data_ptr(), _lib.ptr(bias), out.data_ptr(), m, nn_, k, int(silu_in),
            _lib.stream_ptr(x.device),
        ),
        "azb_linear_f32",
    )
    return out


def add_rows(y: Tensor, table: Tensor, idx: Tensor) -> Tensor:
    _lib.check(
        _lib.lib().azb_add_rows_f32(
            y.data_ptr(), table.data_ptr(), idx.data_ptr(), y.shape[0], y.shape[1], _lib.stream_ptr(y.device)
        ),
        "azb_add_rows_f32",
    )
    return y


# ------------------------------------------------------------------ in-repo backbones (nn.unet / nn.dit)


def token_grid(rows: int) -> tuple[int, int, int]:
    r"""(n, h, w) under which a (rows, C) token matrix is handed to the convolution entry so that an M tile is
    128 consecutive rows."""
    w = math.gcd(rows, 16)
    return 1, rows // w, w


def conv2d(x: Tensor, pc: PackedConv, out: Tensor | None = None, stride: int = 1, act: str | None = None,
           gate: Tensor | None = None, gate_rows: int = 0, residual: Tensor | None = None, nchw_f32: bool = False) -> Tensor:
    r"""``out = residual + gate[sample] * act(conv(x) + bias)`` (``azb_conv2d_bf16``).

    x: (N, H, W, C_in) bf16 NHWC, or (rows, C_in) tokens.  ``gate``: fp32 (C_out,) shared or (samples, >= C_out)
    rows (a slice of a wider matrix is fine: its row stride is used); ``gate_rows`` = output pixels (tokens) per
    sample, default one image.
    """
    assert x.dtype == torch.bfloat16 and x.is_cuda
    tokens = x.ndim == 2
    n, h, w = token_grid(x.shape[0]) if tokens else x.shape[:3]
    ho, wo = -(-h // stride), -(-w // stride)
    if out is None:
        if nchw_f32:
            out = torch.empty((n, pc.c_out, ho, wo), dtype=torch.float32, device=x.device)
        elif tokens:
            out = torch.empty((x.shape[0], pc.c_out), dtype=torch.bfloat16, device=x.device)
        else:
            out = torch.empty((n, ho, wo, pc.c_out), dtype=torch.bfloat16, device=x.device)
    gate_ld = 0
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.stride(-1) == 1
        if gate.ndim == 2 and gate.shape[0] > 1:
            gate_ld = gate.stride(0)
        gate_rows = gate_rows or ho * wo
    _lib.check(
        _lib.lib().azb_conv2d_bf16(
            x.data_ptr(), n, h, w, pc.c_in, _ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps, pc.k_per_tap,
            stride, _lib.ptr(pc.bias), ACT[act], _lib.ptr(gate), gate_ld, gate_rows, _lib.ptr(residual),
            0 if residual is None else _ld(residual), out.data_ptr(), 0 if nchw_f32 else _ld(out),
            1 if nchw_f32 else 0, None, 1, _lib.stream_ptr(x.device),
        ),
        "azb_conv2d_bf16",
    )
    return out


def rownorm_mod(x: Tensor, kind: str = "layer", mod: Tensor | None = None, rows_per_sample: int = 0,
                out: Tensor | None = None, eps: float = 1e-5) -> Tensor:
    r"""``(1 + a) * norm(x) + b`` over the last dimension of a bf16 NHWC / token tensor
    (``azb_rownorm_mod_bf16``); ``mod``: fp32 (>= 2C,) or (samples, >= 2C) = [a | b | ...]."""
    c = x.shape[-1]
    rows = x.numel() // c if x.is_contiguous() else math.prod(x.shape[:-1])
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    mod_ld = 0
    if mod is not None:
        assert mod.dtype == torch.float32 and mod.stride(-1) == 1 and mod.shape[-1] >= 2 * c
        if mod.ndim == 2 and mod.shape[0] > 1:
            mod_ld = mod.stride(0)
            rows_per_sample = rows_per_sample or rows // mod.shape[0]
        else:
            rows_per_sample = rows
    _lib.check(
        _lib.lib().azb_rownorm_mod_bf16(
            x.data_ptr(), _ld(x) if x.ndim in (2, 4) else x.stride(-2), out.data_ptr(),
            _ld(out) if out.ndim in (2, 4) else out.stride(-2), rows, c, NORM[kind], eps, _lib.ptr(mod), mod_ld,
            rows_per_sample, _lib.stream_ptr(x.device),
        ),
        "azb_rownorm_mod_bf16",
    )
    return out


def segment_rmsnorm_(x: Tensor, segs: int, d: int, eps: float = 1e-5) -> Tensor:
    r"""In-place RMS normalisation of the first ``segs`` d-wide channel segments of every row of a
    (rows, ld) bf16 matrix (``azb_segment_rmsnorm_bf16``)."""
    assert x.ndim == 2 and x.dtype == torch.bfloat16
    _lib.check(
        _lib.lib().azb_segment_rmsnorm_bf16(x.data_ptr(), x.stride(0), x.shape[0], segs, d, eps, _lib.stream_ptr(x.device)),
        "azb_segment_rmsnorm_bf16",
    )
    return x


def patchify(x: Tensor, p: int, q: int, k_pad: int, out: Tensor | None = None) -> Tensor:
    r"""fp32 NCHW -> bf16 tokens (N * H/p * W/q, k_pad) (``azb_patchify_f32``)."""
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and h % p == 0 and w % q == 0
    hp, wp = h // p, w // q
    if out is None:
        out = torch.empty(n * hp * wp, k_pad, dtype=torch.bfloat16, device=x.device)
    _lib.check(
        _lib.lib().azb_patchify_f32(x.data_ptr(), out.data_ptr(), n, c, hp, wp, p, q, k_pad, _lib.stream_ptr(x.device)),
        "azb_patchify_f32",
    )
    return out


def unpatchify(yt: Tensor, n: int, c: int, hp: int, wp: int, p: int, q: int, out: Tensor | None = None) -> Tensor:
    r"""Channel-major fp32 GEMM output (c*p*q, tokens) -> fp32 NCHW (``azb_unpatchify_f32``)."""
    assert yt.dtype == torch.float32 and yt.is_contiguous()
    if out is None:
        out = torch.empty(n, c, hp * p, wp * q, dtype=torch.float32, device=yt.device)
    _lib.check(
        _lib.lib().azb_unpatchify_f32(yt.data_ptr(), out.data_ptr(), n, c, hp, wp, p, q, _lib.stream_ptr(yt.device)),
        "azb_unpatchify_f32",
    )
    return out


def linear_gather(x: Tensor, xoff: Tensor | None, weight: Tensor, bias: Tensor | None, silu_in: bool = False,
                  out: Tensor | None = None) -> Tensor:
    r"""``y[:, j] = b[j] + act(x[:, xoff[j] : xoff[j] + K]) @ w[j]`` in fp32 (``azb_linear_gather_f32``)."""
    m = x.shape[0]
    nn_, k = weight.shape
    assert x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.is_contiguous() and x.stride(1) == 1
    if out is None:
        out = torch.empty(m, nn_, dtype=torch.float32, device=x.device)
    _lib.check(
        _lib.lib().azb_linear_gather_f32(
            x.data_ptr(), x.stride(0), _lib.ptr(xoff), weight.data_ptr(), _lib.ptr(bias), out.data_ptr(), m, nn_, k,
            int(silu_in), _lib.stream_ptr(x.device),
        ),
        "azb_linear_gather_f32",
    )
    return out


def conv_desc(x: Tensor, pc, out: Tensor, *, grid: tuple[int, int, int] | None = None, stride: int = 1, act: int = 0,
              gate: int | None = None, gate_ld: int = 0, gate_rows: int = 0, residual: Tensor | None = None,
              nchw_f32: bool = False, x2: Tensor | None = None, gn_acc: Tensor | None = None, gran: int = 8,
              workspace: Tensor | None = None, in_coef: Tensor | None = None, in_silu: bool = False,
              res_up: bool = False, in_up: bool = False, in_norm: int = 0, in_eps: float = 1e-5,
              in_rowstat: Tensor | None = None, in_mod: int | None = None, in_mod_ld: int = 0,
              rowstat: Tensor | None = None, out_up: bool = False) -> AzbConv:
    r"""Fills an :class:`AzbConv` for ``azb_conv_bf16``; ``pc`` is a :class:`PackedConv` or, with ``x2``, a
    :class:`PackedConvSkip``.  ``gn_acc``: int64 (N, C_out / gran, 4) exact GroupNorm accumulators (zeroed by the
    caller).  ``in_coef``: fp32 (N, C_in, 2) from :func:`gn_coef` -- the convolution then reads
    ``act(a x + b)`` instead of ``x`` (only where :func:`conv_choice` reports ``halo``)."""
    n, h, w = grid if grid is not None else x.shape[:3]
    d = AzbConv()
    d.act, d.n, d.h, d.w, d.c_in, d.act_ld = x.data_ptr(), n, h, w, pc.c_in, x.stride(-2)
    d.wpack, d.c_out, d.c_out_rows, d.k_per_tap = pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.k_per_tap
    d.taps, d.stride, d.act_fn, d.out_mode, d.stat_gran = (9 if x2 is not None else pc.taps), stride, act, int(nchw_f32), gran
    d.bias = _lib.ptr(pc.bias)
    d.gate, d.gate_ld, d.gate_rows = gate, gate_ld, gate_rows
    d.residual, d.res_ld = _lib.ptr(residual), (0 if residual is None else residual.stride(-2))
    d.res_up = int(res_up)  # residual given at half resolution, upsampled (nearest) while it is added
    d.out, d.out_ld = out.data_ptr(), (0 if nchw_f32 else out.stride(-2))
    if x2 is not None:
        d.act2, d.c_in2, d.act2_ld, d.k2 = x2.data_ptr(), pc.c_in2, x2.stride(-2), pc.k2
    if gn_acc is not None:
        assert gn_acc.dtype == torch.int64 and gn_acc.is_contiguous() and gn_acc.numel() == n * (pc.c_out // gran) * 4
        d.gn_acc = gn_acc.data_ptr()
    if workspace is not None:  # zero-initialised uint8 scratch for split-K (flags stay zero between launches)
        d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if in_coef is not None:
        assert in_coef.dtype == torch.float32 and in_coef.is_contiguous() and in_coef.numel() == n * pc.c_in * 2
        d.in_coef, d.in_silu = in_coef.data_ptr(), int(in_silu)
    # x is (n, h / 2, w / 2, c), `grid` = the upsampled extents: 1 = read through a nearest 2x upsampling tensor map,
    # 2 = phase-decomposed (weights from pack_conv_up, 2.25 x fewer FLOPs)
    d.in_up = (2 if getattr(pc, "taps", 9) == 16 else 1) if in_up else 0
    # per-pixel LayerNorm (1) / RMSNorm (2) + modulation of the input, statistics from the producer's `rowstat`;
    # `in_mod` is a raw device address of [a(C_in) | b(C_in)] fp32 rows
    if in_norm:
        assert in_rowstat is not None and in_rowstat.dtype == torch.float32 and in_rowstat.is_contiguous()
        assert in_rowstat.numel() == n * h * w * (pc.c_in // 64) * 2 and in_mod
        d.in_norm, d.in_eps, d.in_rowstat, d.in_mod, d.in_mod_ld = in_norm, in_eps, in_rowstat.data_ptr(), in_mod, in_mod_ld
    if rowstat is not None:
        ho, wo = -(-h // stride), -(-w // stride)
        assert rowstat.dtype == torch.float32 and rowstat.is_contiguous() and rowstat.numel() == n * ho * wo * (pc.c_out // 64) * 2
        d.rowstat = rowstat.data_ptr()
    d.out_up = int(out_up)  # `out` is (n, 2 h, 2 w, c): the result is stored through a nearest 2x upsampling
    return d


def conv_choice(d: AzbConv) -> AzbConvChoice:
    r"""``azb_conv_choice``: the launcher's decision for a descriptor (halo tiles, CTA pairs, N tile, split-K)."""
    c = AzbConvChoice()
    _lib.check(_lib.lib().azb_conv_choice(byref(d), byref(c)), "azb_conv_choice")
    return c


def gn_coef(n: int, h: int, w: int, parts: list[tuple[Tensor, int]], gamma: Tensor, beta: Tensor,
            scale_shift: Tensor | None = None, silu: bool = True, gran: int = 8, groups: int = GN_GROUPS,
            eps: float = GN_EPS, out: Tensor | None = None) -> Tensor:
    r"""``azb_gn_coef_f32``: fp32 (N, C, 2) coefficients {A, B} of the GroupNorm (+ scale / shift, + SiLU) transform
    of an (N, H, W, C) tensor whose producers accumulated ``parts`` = [(acc, channels), ...]."""
    (a, ca), (b, cb) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
    c = ca + cb
    if out is None:
        out = torch.empty(n, c, 2, dtype=torch.float32, device=a.device)
    ss_stride = 0
    if scale_shift is not None and scale_shift.ndim == 2 and scale_shift.shape[0] == n and n > 1:
        ss_stride = scale_shift.stride(0)
    _lib.check(
        _lib.lib().azb_gn_coef_f32(n, h, w, c, groups, a.data_ptr(), ca, _lib.ptr(b), cb, gran, eps, gamma.data_ptr(),
                                   beta.data_ptr(), _lib.ptr(scale_shift), ss_stride, int(silu), out.data_ptr(),
                                   _lib.stream_ptr(a.device)),
        "azb_gn_coef_f32",
    )
    return out


def conv_acc(x: Tensor, pc, out: Tensor | None = None, residual: Tensor | None = None, x2: Tensor | None = None,
             gran: int = 8, workspace: Tensor | None = None, in_coef: Tensor | None = None,
             in_silu: bool = False, res_up: bool = False, in_up: bool = False) -> tuple[Tensor, Tensor]:
    r"""Convolution that also returns the exact GroupNorm accumulators of its output: (out, int64 (N, C_out / gran, 4))."""
    n, h, w, _ = x.shape
    if in_up:
        h, w = 2 * h, 2 * w
    if out is None:
        out = torch.empty(n, h, w, pc.c_out, dtype=torch.bfloat16, device=x.device)
    acc = torch.zeros(n, pc.c_out // gran, 4, dtype=torch.int64, device=x.device)
    d = conv_desc(x, pc, out, grid=(n, h, w), residual=residual, x2=x2, gn_acc=acc, gran=gran, workspace=workspace,
                  in_coef=in_coef, in_silu=in_silu, res_up=res_up, in_up=in_up)
    _lib.check(_lib.lib().azb_conv_bf16(byref(d), _lib.stream_ptr(x.device)), "azb_conv_bf16")
    return out, acc


KNOB_PAIR, KNOB_PREFETCH, KNOB_SPLITK, KNOB_GN_WAVE, KNOB_BLOCKN, KNOB_LEAN, KNOB_HALO = 0, 1, 2, 3, 4, 5, 6
KNOB_HALO_SA, KNOB_HALO_SB, KNOB_HALO_SPREAD, KNOB_PDL, KNOB_ROWEPI = 7, 8, 9, 10, 11


def conv_tuning(knob: int, value: int) -> None:
    r"""``azb_conv_tuning``: overrides an automatic choice of the convolution launcher (-1 restores it)."""
    _lib.check(_lib.lib().azb_conv_tuning(knob, value), "azb_conv_tuning")


if os.environ.get("AZB_PDL", "") == "0":  # A/B switch: plain stream-ordered launches
    try:
        conv_tuning(KNOB_PDL, 0)
    except Exception:  # library not built yet: the first real call reports it
        pass

if os.environ.get("AZB_PAIR", "") in ("0", "1"):  # A/B switch: CTA pairs never / whenever possible
    try:
        conv_tuning(KNOB_PAIR, int(os.environ["AZB_PAIR"]))
    except Exception:
        pass

if os.environ.get("AZB_ROWEPI", "") == "0":  # A/B switch: the shared-memory-transpose epilogue instead of TMA stores
    try:
        conv_tuning(KNOB_ROWEPI, 0)
    except Exception:
        pass

SPLITK_WORKSPACE_BYTES = 16 << 20


def splitk_workspace(device) -> Tensor:
    r"""Zeroed scratch that lets ``azb_conv_bf16`` split the reduction of small-map / long-K layers over CTAs."""
    return torch.zeros(SPLITK_WORKSPACE_BYTES, dtype=torch.uint8, device=device)


def gn_apply_acc(x: Tensor, parts: list[tuple[Tensor, int]], gamma: Tensor, beta: Tensor, out: Tensor | None = None,
                 scale_shift: Tensor | None = None, silu: bool = True, mode: int = 0, gran: int = 8,
                 groups: int = GN_GROUPS, eps: float = GN_EPS) -> Tensor:
    r"""GroupNorm + modulation + SiLU + resampling with statistics from exact accumulators
    (``azb_gn_apply_acc_bf16``); ``parts`` = [(acc, channels), ...] of the one or two producers of ``x``."""
    n, h, w, c = x.shape
    ho, wo = (h * 2, w * 2) if mode == 1 else (h // 2, w // 2) if mode == 2 else (h, w)
    if out is None:
        out = torch.empty(n, ho, wo, c, dtype=torch.bfloat16, device=x.device)
    (a, ca), (b, cb) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
    ss_stride = 0
    if scale_shift is not None and scale_shift.ndim == 2 and scale_shift.shape[0] == n and n > 1:
        ss_stride = scale_shift.stride(0)
    _lib.check(
        _lib.lib().azb_gn_apply_acc_bf16(
            x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), n, h, w, c, groups, a.data_ptr(), ca, _lib.ptr(b), cb, gran,
            eps, gamma.data_ptr(), beta.data_ptr(), _lib.ptr(scale_shift), ss_stride, int(silu), mode,
            _lib.stream_ptr(x.device),
        ),
        "azb_gn_apply_acc_bf16",
    )
    return out


# ------------------------------------------------------------------------- reference-numerics (TF32) mode


def pack_conv_f32(weight: Tensor, bias: Tensor | None, c_in_pad: int | None = None) -> PackedConv:
    r"""(C_out, C_in[, kh, kw]) -> fp32 (C_out_rows, taps, K_tap) with K_tap = C_in rounded up to 32 (one 128-byte row of
    fp32): the weights of ``azb_conv_tf32``.  ``c_in_pad``: the (zero padded) channel count the activation is stored
    with (the 3-channel network input is padded to 4)."""
    if weight.ndim == 3:
        weight = weight[..., 0]
    if weight.ndim == 2:
        weight = weight[:, :, None, None]
    c_out, c_in, kh, kw = weight.shape
    assert (kh, kw) in ((1, 1), (3, 3))
    taps = kh * kw
    c_store = c_in if c_in_pad is None else c_in_pad
    k_pad = -(-c_store // 32) * 32
    tile = 128 if c_out >= 128 else 64 if c_out >= 64 else 32 if c_out >= 32 else 16
    rows = -(-c_out // tile) * tile
    w = torch.zeros(rows, taps, k_pad, dtype=torch.float32, device=weight.device)
    w[:c_out, :, :c_in] = weight.detach().permute(0, 2, 3, 1).reshape(c_out, taps, c_in).to(torch.float32)
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return PackedConv(w=w.contiguous(), bias=b, c_in=c_store, c_out=c_out, taps=taps)


def conv_tf32(x: Tensor, pc: PackedConv, out: Tensor | None = None, stride: int = 1, act: str | None = None,
              residual: Tensor | None = None, nchw: bool = False, out_f16: bool = False) -> Tensor:
    r"""``azb_conv_tf32``: x fp32 NHWC (or (rows, C) tokens) -> fp32 NHWC / fp16 NHWC / fp32 NCHW."""
    assert x.dtype == torch.float32 and x.is_cuda and pc.w.dtype == torch.float32
    tokens = x.ndim == 2
    n, h, w = token_grid(x.shape[0]) if tokens else x.shape[:3]
    ho, wo = -(-h // stride), -(-w // stride)
    if out is None:
        if nchw:
            out = torch.empty((n, pc.c_out, ho, wo), dtype=torch.float32, device=x.device)
        else:
            shape = (x.shape[0], pc.c_out) if tokens else (n, ho, wo, pc.c_out)
            out = torch.empty(shape, dtype=torch.float16 if out_f16 else torch.float32, device=x.device)
    _lib.check(
        _lib.lib().azb_conv_tf32(
            x.data_ptr(), n, h, w, pc.c_in, _ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps, pc.k_per_tap, stride,
            _lib.ptr(pc.bias), ACT[act], _lib.ptr(residual), 0 if residual is None else _ld(residual), out.data_ptr(),
            0 if nchw else _ld(out), int(nchw), int(out_f16), _lib.stream_ptr(x.device),
        ),
        "azb_conv_tf32",
    )
    return out


def gn_stats_workspace_f32(n: int, groups: int, device) -> Tensor:
    r"""Zeroed scratch for :func:`gn_stats_f32` (arrival counters + fp64 partial sums of up to 16 CTAs per slot)."""
    slots = n * groups
    return torch.zeros(-(-slots * 4 // 256) * 256 + slots * 16 * 16, dtype=torch.uint8, device=device)


def gn_stats_f32(x: Tensor, groups: int = GN_GROUPS, eps: float = GN_EPS, stats: Tensor | None = None,
                 workspace: Tensor | None = None) -> Tensor:
    r"""``azb_gn_stats_f32``: (N, groups, 2) = {mean, rstd} of an fp32 NHWC tensor."""
    n, h, w, c = x.shape
    if stats is None:
        stats = torch.empty(n, groups, 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().azb_gn_stats_f32(x.data_ptr(), _ld(x), n, h * w, c, groups, eps, stats.data_ptr(), _lib.ptr(workspace),
                                           0 if workspace is None else workspace.numel(), _lib.stream_ptr(x.device)),
               "azb_gn_stats_f32")
    return stats


def gn_apply_f32(x: Tensor, stats: Tensor | None = None, gamma: Tensor | None = None, beta: Tensor | None = None,
                 scale_shift: Tensor | None = None, silu: bool = False, mode: int = 0, out: Tensor | None = None,
                 groups: int = GN_GROUPS) -> Tensor:
    r"""``azb_gn_apply_f32``: normalise (+ scale / shift) (+ SiLU) (+ nearest 2x upsampling / 2 x 2 pooling), fp32 NHWC."""
    n, h, w, c = x.shape
    ho, wo = (2 * h, 2 * w) if mode == 1 else (h // 2, w // 2) if mode == 2 else (h, w)
    if out is None:
        out = torch.empty(n, ho, wo, c, dtype=torch.float32, device=x.device)
    ss_stride = 0
    if scale_shift is not None and scale_shift.ndim == 2 and scale_shift.shape[0] > 1:
        ss_stride = scale_shift.stride(0)
    _lib.check(
        _lib.lib().azb_gn_apply_f32(x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), n, h, w, c, groups, _lib.ptr(stats),
                                    _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(scale_shift), ss_stride, int(silu), mode,
                                    _lib.stream_ptr(x.device)),
        "azb_gn_apply_f32",
    )
    return out


def attention_qknorm(qkv: Tensor, heads: int, eps: float = 1e-5, out: Tensor | None = None) -> Tensor:
    r"""``azb_attention_qknorm_bf16``: attention over a (N, T, 3C) bf16 projection in [q | k | v] channel order with the
    per-head RMS normalisation of q and k folded into the logits (T <= 256, head width 64)."""
    n, t, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    if out is None:
        out = torch.empty(n, t, c, dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.lib().azb_attention_qknorm_bf16(qkv.data_ptr(), qkv.stride(1), out.data_ptr(), out.stride(1), n, t, heads, d, d,
                                                    c, 2 * c, eps, _lib.stream_ptr(qkv.device)), "azb_attention_qknorm_bf16")
    return out


def attention_f16(qkv: Tensor, heads: int, new_order: bool = False, out: Tensor | None = None) -> Tensor:
    r"""``azb_attention_f16``: qkv fp16 (N, T, 3C) -> fp32 (N, T, C)."""
    assert qkv.dtype == torch.float16
    n, t, c3 = qkv.shape
    c = c3 // 3
    d = c // heads
    if out is None:
        out = torch.empty(n, t, c, dtype=torch.float32, device=qkv.device)
    hs, kd, vd = (d, c, 2 * c) if new_order else (3 * d, d, 2 * d)
    _lib.check(_lib.lib().azb_attention_f16(qkv.data_ptr(), qkv.stride(1), out.data_ptr(), out.stride(1), n, t, heads, d, hs, kd,
                                            vd, _lib.stream_ptr(qkv.device)), "azb_attention_f16")
    return out
