r"""U-Net building blocks (interface of ``azula/nn/unet.py``).

Parameter names and shapes equal the reference's, so its checkpoints load with
``load_state_dict``.  Execution: on a CUDA device with autograd disabled, :meth:`UNet.forward` runs
the launch plan of :mod:`azula_b200.engine.unet` (NHWC bf16 activations, tcgen05 convolutions with
SiLU / gated-residual epilogues, fused Ada-Norm-Zero prologue); otherwise the plain torch
definition below, which is the reference's arithmetic.
"""

from __future__ import annotations

__all__ = ["UNet", "UNetBlock"]

import torch
import torch.nn as nn

from collections.abc import Sequence
from torch import Tensor

from .layers import ConvNd, LayerNorm, RMSNorm
from .utils import NativeCache, checkpoint


class _SplitMod(nn.Module):
    r"""``(..., 3C) -> (3, ..., C, 1 x spatial)``: the three Ada-Norm-Zero vectors, broadcastable
    over the spatial dimensions (the reference uses an einops ``Rearrange`` here)."""

    def __init__(self, spatial: int) -> None:
        super().__init__()

        self.spatial = spatial

    def forward(self, x: Tensor) -> Tensor:
        x = x.unflatten(-1, (3, -1)).movedim(-2, 0)
        return x.reshape(*x.shape, *(1,) * self.spatial)


class UNetBlock(nn.Module):
    r"""Modulated U-Net block (``azula/nn/unet.py:18-122``):

    .. math:: y = x + c \odot \mathrm{FFN}((1 + a) \odot \mathrm{norm}(x) + b)

    with :math:`(a, b, c)` an MLP of the modulation vector (or free parameters when
    ``mod_features = 0``) and FFN = conv, SiLU, conv.

    Arguments:
        channels: The number of channels :math:`C`.
        mod_features: The number of modulating features :math:`D`.
        norm: The kind of normalization: ``"group"``, ``"layer"`` or ``"rms"``.
        groups: The number of groups of the group normalization.
        ffn_factor: The channel factor in the FFN.
        spatial: The number of spatial dimensions :math:`N`.
        dropout: The dropout rate in :math:`[0, 1]`.
        checkpointing: Whether to use activation checkpointing or not.
        kwargs: Keyword arguments passed to :func:`azula_b200.nn.layers.ConvNd`.
    """

    def __init__(
        self,
        channels: int,
        mod_features: int = 0,
        norm: str = "layer",
        groups: int = 16,
        ffn_factor: int = 1,
        spatial: int = 2,
        dropout: float | None = None,
        checkpointing: bool = False,
        **kwargs,
    ) -> None:
        super().__init__()

        self.checkpointing = checkpointing
        self.channels = channels
        self.norm_kind = norm

        if norm == "layer":
            self.norm = LayerNorm(dim=-spatial - 1, eps=1e-5)
        elif norm == "rms":
            self.norm = RMSNorm(dim=-spatial - 1, eps=1e-5)
        elif norm == "group":
            self.norm = nn.GroupNorm(num_groups=min(groups, channels), num_channels=channels, affine=False, eps=1e-5)
        else:
            raise NotImplementedError()

        if mod_features > 0:
            self.ada_zero = nn.Sequential(
                nn.Linear(mod_features, mod_features),
                nn.SiLU(),
                nn.Linear(mod_features, 3 * channels),
                _SplitMod(spatial),
            )
            self.ada_zero[2].weight.data.mul_(1e-2)
        else:
            self.ada_zero = nn.Parameter(torch.randn(3, channels, *(1,) * spatial))
            self.ada_zero.data.mul_(1e-2)

        self.ffn = nn.Sequential(
            ConvNd(channels, ffn_factor * channels, spatial=spatial, **kwargs),
            nn.SiLU(),
            nn.Identity() if dropout is None else nn.Dropout(dropout),
            ConvNd(ffn_factor * channels, channels, spatial=spatial, **kwargs),
        )

    def _forward(self, x: Tensor, mod: Tensor | None = None) -> Tensor:
        a, b, c = self.ada_zero if torch.is_tensor(self.ada_zero) else self.ada_zero(mod)
        y = (a + 1) * self.norm(x) + b
        return x + c * self.ffn(y)

    def forward(self, x: Tensor, mod: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tensor, with shape :math:`(B, C, L_1, ..., L_N)`.
            mod: The modulation vector, with shape :math:`(D)` or :math:`(B, D)`.
        """
        if self.checkpointing:
            return checkpoint(self._forward, reentrant=not self.training)(x, mod)
        return self._forward(x, mod)


class UNet(NativeCache, nn.Module):
    r"""Modulated U-Net (``azula/nn/unet.py:125-259``).

    Arguments:
        in_channels: The number of input channels :math:`C_i`.
        out_channels: The number of output channels :math:`C_o`.
        cond_channels: The number of condition channels :math:`C_c`.
        hid_channels: The numbers of channels at each depth.
        hid_blocks: The numbers of hidden blocks at each depth.
        kernel_size: The kernel size of all convolutions.
        stride: The stride of the downsampling convolutions.
        spatial: The number of spatial dimensions :math:`N`.
        periodic: Whether the spatial dimensions are periodic or not.
        identity_init: Initialize down/upsampling convolutions as identity.
        kwargs: Keyword arguments passed to :class:`UNetBlock`.
    """

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        cond_channels: int = 0,
        hid_channels: Sequence[int] = (64, 128, 256),
        hid_blocks: Sequence[int] = (3, 3, 3),
        kernel_size: int | Sequence[int] = 3,
        stride: int | Sequence[int] = 2,
        spatial: int = 2,
        periodic: bool = False,
        identity_init: bool = False,
        **kwargs,
    ) -> None:
        super().__init__()

        assert len(hid_blocks) == len(hid_channels)

        kernel_size = [kernel_size] * spatial if isinstance(kernel_size, int) else list(kernel_size)
        stride = [stride] * spatial if isinstance(stride, int) else list(stride)

        conv = dict(  # noqa: C408
            kernel_size=tuple(kernel_size),
            padding=tuple(k // 2 for k in kernel_size),
            padding_mode="circular" if periodic else "zeros",
            spatial=spatial,
        )

        self.descent, self.ascent = nn.ModuleList(), nn.ModuleList()

        depth = len(hid_blocks)
        for i, blocks in enumerate(hid_blocks):
            down, up = nn.ModuleList(), nn.ModuleList()

            for _ in range(blocks):
                down.append(UNetBlock(hid_channels[i], **conv, **kwargs))
                up.append(UNetBlock(hid_channels[i], **conv, **kwargs))

            if i == 0:
                down.insert(0, ConvNd(in_channels + cond_channels, hid_channels[0], **conv))
                up.append(ConvNd(hid_channels[0], out_channels, **conv))
            else:
                down.insert(
                    0, ConvNd(hid_channels[i - 1], hid_channels[i], stride=stride, identity_init=identity_init, **conv)
                )
                up.append(nn.Upsample(scale_factor=tuple(stride), mode="nearest"))

            if i + 1 < depth:
                up.insert(
                    0, ConvNd(hid_channels[i] + hid_channels[i + 1], hid_channels[i], identity_init=identity_init, **conv)
                )

            self.descent.append(down)
            self.ascent.insert(0, up)

        self._native: dict = {}  # kernel-layout weights and launch plans (engine/unet.py)

    def forward(self, x: Tensor, mod: Tensor | None = None, cond: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tensor, with shape :math:`(B, C_i, L_1, ..., L_N)`.
            mod: The modulation vector, with shape :math:`(D)` or :math:`(B, D)`.
            cond: The condition tensor, with shape :math:`(B, C_c, L_1, ..., L_N)`.

        Returns:
            The output tensor, with shape :math:`(B, C_o, L_1, ..., L_N)`.
        """
        if cond is not None:
            x = torch.cat((x, cond), dim=1)

        if x.is_cuda and not torch.is_grad_enabled():
            from .. import engine
            from ..engine import unet as _engine

            if engine.native_enabled() and _engine.supports(self, x, mod):
                return _engine.forward(self, x, mod)

        return self._torch_forward(x, mod)

    def _torch_forward(self, x: Tensor, mod: Tensor | None) -> Tensor:
        skips: list[Tensor | None] = []

        for level in self.descent:
            skips.append(x if skips else None)  # the input of every level but the first is a skip
            for layer in level:
                x = layer(x, mod) if isinstance(layer, UNetBlock) else layer(x)

        if hasattr(self, "bottleneck"):
            x = self.bottleneck(x, mod)

        for level in self.ascent:
            for layer in level:
                x = layer(x, mod) if isinstance(layer, UNetBlock) else layer(x)

            y = skips.pop()
            if y is None:
                continue
            for d in range(2, x.ndim):  # odd sizes: the upsampled tensor may be one too long
                if x.shape[d] > y.shape[d]:
                    x = torch.narrow(x, d, 0, y.shape[d])
            x = torch.cat((y, x), dim=1)

        return x
