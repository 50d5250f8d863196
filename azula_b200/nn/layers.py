r"""Common layers (interface of ``azula/nn/layers.py``).

These are the plain-torch definitions: they serve CPU tensors, training and autograd, and fix the
parameter names (``state_dict`` keys) of the reference.  On a CUDA device, under ``torch.no_grad()``,
the backbones that use them (:mod:`azula_b200.nn.unet`, :mod:`azula_b200.nn.dit`) do not call them:
their forward is a launch plan over ``libazb.so`` (:mod:`azula_b200.engine.unet`,
:mod:`azula_b200.engine.dit`).
"""

from __future__ import annotations

__all__ = [
    "ConvNd",
    "LayerNorm",
    "Patchify",
    "RMSNorm",
    "ReLU2",
    "SineEncoding",
    "SwiGLU",
    "Unpatchify",
]

import math
import torch
import torch.nn as nn

from collections.abc import Sequence
from torch import Tensor

_CONV = {0: nn.Linear, 1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}


def ConvNd(in_channels: int, out_channels: int, spatial: int = 2, identity_init: bool = False, **kwargs) -> nn.Module:
    r"""Returns an N-dimensional convolution (``azula/nn/layers.py:25-68``); ``spatial=0`` is a
    linear layer.  With ``identity_init`` the first ``in_channels`` filters start as
    :math:`10^{-2} W + I` (a pseudo-identity through the kernel centre)."""
    if spatial not in _CONV:
        raise NotImplementedError()

    conv = _CONV[spatial](in_channels, out_channels, **kwargs)

    if identity_init:
        with torch.no_grad():
            head = conv.weight[:in_channels]
            centre = tuple(k // 2 for k in conv.weight.shape[2:])
            eye = torch.zeros_like(head)
            for i in range(min(in_channels, out_channels)):
                eye[(i, i, *centre)] = 1.0
            head.mul_(1e-2).add_(eye)

    return conv


class ReLU2(nn.Module):
    r""":math:`y = \max(x, 0)^2` (``azula/nn/layers.py:71-82``)."""

    def forward(self, x: Tensor) -> Tensor:
        return torch.relu(x).square()


class SwiGLU(nn.Module):
    r""":math:`y = x_1 \, x_2 \, \sigma(x_2)` over interleaved channel pairs, :math:`(*, 2C) \to (*, C)`
    (``azula/nn/layers.py:89-117``)."""

    def forward(self, x: Tensor) -> Tensor:
        pairs = x.unflatten(-1, (-1, 2))
        return pairs[..., 0] * nn.functional.silu(pairs[..., 1])


def _at_least_fp32(x: Tensor) -> Tensor:
    return x.to(torch.promote_types(x.dtype, torch.float32))


class LayerNorm(nn.Module):
    r"""Standardisation along ``dim`` without affine parameters, computed in at least float32 with
    the unbiased variance of :func:`torch.var_mean` (``azula/nn/layers.py:120-155``)."""

    def __init__(self, dim: int | Sequence[int], eps: float = 1e-5) -> None:
        super().__init__()

        self.dim = dim
        self.eps = eps

    def extra_repr(self) -> str:
        return f"dim={self.dim}"

    def forward(self, x: Tensor) -> Tensor:
        h = _at_least_fp32(x)
        var, mean = torch.var_mean(h, dim=self.dim, keepdim=True)
        return ((h - mean) * torch.rsqrt(var + self.eps)).to(x.dtype)


class RMSNorm(nn.Module):
    r""":math:`y = x / \sqrt{\mathbb{E}[x^2] + \epsilon}` along ``dim``, in at least float32
    (``azula/nn/layers.py:158-195``)."""

    def __init__(self, dim: int | Sequence[int], eps: float = 1e-5) -> None:
        super().__init__()

        self.dim = dim
        self.eps = eps

    def extra_repr(self) -> str:
        return f"dim={self.dim}"

    def forward(self, x: Tensor) -> Tensor:
        h = _at_least_fp32(x)
        return (h * torch.rsqrt(h.square().mean(dim=self.dim, keepdim=True) + self.eps)).to(x.dtype)


class Patchify(nn.Module):
    r"""Moves patches of shape ``patch_shape`` of the trailing spatial dimensions into the channel
    dimension: :math:`(*, Z, A a, B b) \to (*, Z a b, A, B)`, or :math:`(*, A, B, Z a b)` with
    ``channel_last`` (``azula/nn/layers.py:198-222``)."""

    def __init__(self, patch_shape: Sequence[int], channel_last: bool = False) -> None:
        super().__init__()

        self.patch_shape = tuple(patch_shape)
        self.channel_last = channel_last

    def extra_repr(self) -> str:
        return f"{self.patch_shape}, channel_last={self.channel_last}"

    def forward(self, x: Tensor) -> Tensor:
        nd = len(self.patch_shape)
        lead = x.ndim - nd - 1  # batch dimensions
        for i, p in enumerate(self.patch_shape):
            x = x.unflatten(lead + 1 + 2 * i, (-1, p))  # (*, Z, A, a, B, b, ...)
        grid = [lead + 1 + 2 * i for i in range(nd)]
        inner = [lead + 2 + 2 * i for i in range(nd)]
        batch = list(range(lead))
        if self.channel_last:
            x = x.permute(*batch, *grid, lead, *inner)
            return x.flatten(lead + nd)
        x = x.permute(*batch, lead, *inner, *grid)
        return x.flatten(lead, lead + nd)


class Unpatchify(nn.Module):
    r"""The inverse of :class:`Patchify` (``azula/nn/layers.py:225-246``)."""

    def __init__(self, patch_shape: Sequence[int], channel_last: bool = False) -> None:
        super().__init__()

        self.patch_shape = tuple(patch_shape)
        self.channel_last = channel_last

    def extra_repr(self) -> str:
        return f"{self.patch_shape}, channel_last={self.channel_last}"

    def forward(self, x: Tensor) -> Tensor:
        nd = len(self.patch_shape)
        lead = x.ndim - nd - 1
        batch = list(range(lead))
        if self.channel_last:  # (*, A, B, Z a b)
            x = x.unflatten(-1, (-1, *self.patch_shape))  # (*, A, B, Z, a, b)
            grid = [lead + i for i in range(nd)]
            inner = [lead + nd + 1 + i for i in range(nd)]
            z = lead + nd
        else:  # (*, Z a b, A, B)
            x = x.unflatten(lead, (-1, *self.patch_shape))  # (*, Z, a, b, A, B)
            inner = [lead + 1 + i for i in range(nd)]
            grid = [lead + 1 + nd + i for i in range(nd)]
            z = lead
        order = [*batch, z]
        for g, a in zip(grid, inner, strict=True):
            order += [g, a]
        x = x.permute(*order)  # (*, Z, A, a, B, b)
        for i in range(nd):
            x = x.flatten(lead + 1 + i, lead + 2 + i)
        return x


class SineEncoding(nn.Module):
    r"""Sinusoidal encoding :math:`e = [\sin(x \omega^{-i}), \cos(x \omega^{-i})]` with
    :math:`i` on a uniform grid of :math:`D/2` points in :math:`[0, 1]`
    (``azula/nn/layers.py:249-299``)."""

    def __init__(self, features: int, omega: float = 1e4) -> None:
        super().__init__()

        assert features % 2 == 0

        self.features = features
        self.omega = omega

    def forward(self, x: Tensor) -> Tensor:
        h = _at_least_fp32(x).unsqueeze(-1)
        freqs = torch.linspace(0, 1, self.features // 2, dtype=h.dtype, device=h.device)
        freqs = torch.exp(math.log(1 / self.omega) * freqs)
        return torch.cat((torch.sin(h * freqs), torch.cos(h * freqs)), dim=-1).to(x.dtype)
