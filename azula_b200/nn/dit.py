r"""Diffusion Transformer (DiT) building blocks (interface of ``azula/nn/dit.py``).

Parameter names and shapes equal the reference's.  On a CUDA device with autograd disabled,
:meth:`DiT.forward` runs the launch plan of :mod:`azula_b200.engine.dit` (bf16 tokens, tcgen05
GEMMs with SiLU / gated-residual epilogues, fused RMS-norm + modulation, flash attention);
otherwise the plain torch definition below.

References:
    | Scalable Diffusion Models with Transformers (Peebles et al., 2022)
    | https://arxiv.org/abs/2212.09748
"""

from __future__ import annotations

__all__ = ["DiT", "DiTBlock"]

import torch
import torch.nn as nn

from torch import Tensor
from typing import Literal

from .attention import MultiheadSelfAttention
from .layers import ReLU2, RMSNorm, SineEncoding, SwiGLU
from .utils import NativeCache, checkpoint


class _SplitTokenMod(nn.Module):
    r"""``(..., 3C) -> (3, ..., 1, C)``: Ada-Norm-Zero vectors broadcastable over tokens."""

    def forward(self, x: Tensor) -> Tensor:
        return x.unflatten(-1, (3, -1)).movedim(-2, 0).unsqueeze(-2)


class _JoinPositions(nn.Module):
    r"""``(..., P, C) -> (..., P C)``."""

    def forward(self, x: Tensor) -> Tensor:
        return x.flatten(-2)


class DiTBlock(nn.Module):
    r"""Modulated DiT block (``azula/nn/dit.py:24-132``):

    .. math:: y' = (1 + a) \odot \mathrm{RMSNorm}(x) + b \qquad
        y = x + c \odot \mathrm{FFN}(y' + \mathrm{MSA}(y'))

    Arguments:
        channels: The number of channels :math:`C`.
        mod_features: The number of modulating features :math:`D`.
        ffn_factor: The channel factor in the FFN.
        ffn_activation: The activation function in the FFN.
        dropout: The dropout rate in :math:`[0, 1]`.
        checkpointing: Whether to use activation checkpointing or not.
        kwargs: Keyword arguments passed to :class:`MultiheadSelfAttention`.
    """

    def __init__(
        self,
        channels: int,
        mod_features: int = 0,
        ffn_factor: int = 4,
        ffn_activation: Literal["relu", "relu2", "silu", "swiglu"] = "silu",
        dropout: float | None = None,
        checkpointing: bool = False,
        **kwargs,
    ) -> None:
        super().__init__()

        self.checkpointing = checkpointing
        self.channels = channels
        self.ffn_activation = ffn_activation

        if hasattr(nn, "RMSNorm"):
            self.norm = nn.RMSNorm(channels, elementwise_affine=False, eps=1e-5)
        else:
            self.norm = RMSNorm(dim=-1, eps=1e-5)

        if mod_features > 0:
            self.ada_zero = nn.Sequential(
                nn.Linear(mod_features, mod_features),
                nn.SiLU(),
                nn.Linear(mod_features, 3 * channels),
                _SplitTokenMod(),
            )
            self.ada_zero[2].weight.data.mul_(1e-2)
        else:
            self.ada_zero = nn.Parameter(torch.randn(3, channels))
            self.ada_zero.data.mul_(1e-2)

        self.msa = MultiheadSelfAttention(channels, **kwargs)

        activations = {"relu": nn.ReLU, "relu2": ReLU2, "silu": nn.SiLU, "swiglu": SwiGLU}
        if ffn_activation not in activations:
            raise NotImplementedError(f"Unknown activation '{ffn_activation}'.")
        shrink = 2 if ffn_activation == "swiglu" else 1

        self.ffn = nn.Sequential(
            nn.Linear(channels, ffn_factor * channels),
            activations[ffn_activation](),
            nn.Identity() if dropout is None else nn.Dropout(dropout),
            nn.Linear(ffn_factor * channels // shrink, channels),
        )

    def _forward(self, x: Tensor, mod: Tensor | None = None, pos: Tensor | None = None, mask: Tensor | None = None) -> Tensor:
        a, b, c = self.ada_zero if torch.is_tensor(self.ada_zero) else self.ada_zero(mod)
        y = (a + 1) * self.norm(x) + b
        y = y + self.msa(y, pos, mask)
        return x + c * self.ffn(y)

    def forward(self, x: Tensor, mod: Tensor | None = None, pos: Tensor | None = None, mask: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tokens :math:`x`, with shape :math:`(*, L, C)`.
            mod: The modulation vector, with shape :math:`(D)` or :math:`(*, D)`.
            pos: The postition coordinates, with shape :math:`(*, L, N)`.
            mask: The attention mask, with shape :math:`(*, L, L)`.
        """
        if self.checkpointing:
            return checkpoint(self._forward, reentrant=not self.training)(x, mod, pos, mask)
        return self._forward(x, mod, pos, mask)


class DiT(NativeCache, nn.Module):
    r"""Modulated DiT-like network over tokens (``azula/nn/dit.py:135-218``).

    Arguments:
        in_channels: The number of input channels :math:`C_i`.
        out_channels: The number of output channels :math:`C_o`.
        cond_channels: The number of condition channels :math:`C_c`.
        mod_features: The number of modulating features :math:`D`.
        pos_channels: The number of positional channels :math:`P`.
        hid_channels: The numbers of hidden token channels :math:`C_h`.
        hid_blocks: The number of hidden transformer blocks.
        kwargs: Keyword arguments passed to :class:`DiTBlock`.
    """

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        cond_channels: int = 0,
        mod_features: int = 0,
        pos_channels: int = 1,
        hid_channels: int = 1024,
        hid_blocks: int = 3,
        **kwargs,
    ) -> None:
        super().__init__()

        self.in_proj = nn.Linear(in_channels + cond_channels, hid_channels)
        self.out_proj = nn.Linear(hid_channels, out_channels)

        self.pos_embedding = nn.Sequential(
            SineEncoding(hid_channels, omega=1e2),
            _JoinPositions(),
            nn.Linear(pos_channels * hid_channels, hid_channels, bias=False),
        )
        self.pos_embedding[2].weight.data.mul_(1e-2)

        self.blocks = nn.ModuleList([
            DiTBlock(channels=hid_channels, pos_channels=pos_channels, mod_features=mod_features, **kwargs)
            for _ in range(hid_blocks)
        ])

        self._native: dict = {}  # kernel-layout weights and launch plans (engine/dit.py)

    def forward(self, x: Tensor, mod: Tensor | None = None, pos: Tensor | None = None, cond: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tensor, with shape :math:`(*, L, C_i)`.
            mod: The modulation vector, with shape :math:`(D)` or :math:`(*, D)`.
            pos: The position tensor, with shape :math:`(*, L, P)`. If :py:`None`, the sequence indices.
            cond: The condition tensor, with shape :math:`(*, L, C_c)`.

        Returns:
            The output tensor, with shape :math:`(*, L, C_o)`.
        """
        if cond is not None:
            x = torch.cat((x, cond), dim=-1)

        if x.is_cuda and not torch.is_grad_enabled():
            from .. import engine
            from ..engine import dit as _engine

            where = "arange" if pos is None else pos
            if engine.native_enabled() and _engine.supports(self, x, mod, where):
                return _engine.forward(self, x, mod, where)

        if pos is None:
            pos = torch.arange(x.shape[-2], dtype=x.dtype, device=x.device)[..., None]

        x = self.in_proj(x)
        x = x + self.pos_embedding(pos)

        for block in self.blocks:
            x = block(x, mod, pos=pos)

        return self.out_proj(x)
