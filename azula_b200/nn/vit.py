r"""Vision Transformer (ViT) building blocks (interface of ``azula/nn/vit.py``).

References:
    | An Image is Worth 16x16 Words: Transformers for Image Recognition at Scale (Dosovitskiy et al., 2021)
    | https://arxiv.org/abs/2010.11929
"""

from __future__ import annotations

__all__ = ["ViT"]

import math
import torch

from collections.abc import Sequence
from torch import Tensor

from .dit import DiT
from .layers import Patchify, Unpatchify


class ViT(DiT):
    r"""Modulated ViT-like network: patchify, :class:`azula_b200.nn.dit.DiT` over the patch tokens
    with grid positions, unpatchify (``azula/nn/vit.py:22-108``).

    Arguments:
        in_channels: The number of input channels :math:`C_i`.
        out_channels: The number of output channels :math:`C_o`.
        cond_channels: The number of condition channels :math:`C_c`.
        mod_features: The number of modulating features :math:`D`.
        hid_channels: The numbers of hidden token channels.
        hid_blocks: The number of hidden transformer blocks.
        spatial: The number of spatial dimensions :math:`N`.
        patch_size: The patch size or shape.
        unpatch_size: The unpatch size or shape.
        kwargs: Keyword arguments passed to :class:`azula_b200.nn.dit.DiTBlock`.
    """

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        cond_channels: int = 0,
        mod_features: int = 0,
        hid_channels: int = 1024,
        hid_blocks: int = 3,
        spatial: int = 2,
        patch_size: int | Sequence[int] = 1,
        unpatch_size: int | Sequence[int] | None = None,
        **kwargs,
    ) -> None:
        patch_size = [patch_size] * spatial if isinstance(patch_size, int) else list(patch_size)

        if unpatch_size is None:
            unpatch_size = patch_size
        elif isinstance(unpatch_size, int):
            unpatch_size = [unpatch_size] * spatial

        assert len(patch_size) == len(unpatch_size) == spatial

        super().__init__(
            in_channels=math.prod(patch_size) * in_channels,
            out_channels=math.prod(unpatch_size) * out_channels,
            cond_channels=math.prod(patch_size) * cond_channels,
            mod_features=mod_features,
            pos_channels=spatial,
            hid_channels=hid_channels,
            hid_blocks=hid_blocks,
            **kwargs,
        )

        self.patch = Patchify(patch_size, channel_last=True)
        self.unpatch = Unpatchify(unpatch_size, channel_last=True)
        self.spatial = spatial

    def forward(self, x: Tensor, mod: Tensor | None = None, cond: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tensor, with shape :math:`(B, C_i, L_1, ..., L_N)`.
            mod: The modulation vector, with shape :math:`(D)` or :math:`(B, D)`.
            cond: The condition tensor, with :math:`(B, C_c, L_1, ..., L_N)`.

        Returns:
            The output tensor, with shape :math:`(B, C_o, L_1, ..., L_N)`.
        """
        if x.is_cuda and not torch.is_grad_enabled():
            from .. import engine
            from ..engine import dit as _engine

            if engine.native_enabled() and _engine.supports_image(self, x, mod, cond):
                return _engine.forward_image(self, x, mod, cond)

        x = self.patch(x)
        if cond is not None:
            cond = self.patch(cond)

        grid = x.shape[1:-1]
        pos = torch.cartesian_prod(*(torch.arange(size, dtype=x.dtype, device=x.device) for size in grid))
        pos = pos.reshape(-1, len(grid))

        y = super().forward(x.flatten(1, -2), mod, pos=pos, cond=None if cond is None else cond.flatten(1, -2))

        return self.unpatch(y.unflatten(-2, grid))
