r"""Attention layers (interface of ``azula/nn/attention.py``)."""

from __future__ import annotations

__all__ = ["MultiheadSelfAttention"]

import math
import torch
import torch.nn as nn

from torch import BoolTensor, Tensor

from .layers import RMSNorm


class MultiheadSelfAttention(nn.Module):
    r"""Multi-head self-attention (``azula/nn/attention.py:17-121``).

    Arguments:
        channels: The number of channels :math:`H \times C`.
        pos_channels: The number of positional channels :math:`P` (RoPE only).
        attention_heads: The number of attention heads :math:`H`.
        qkv_bias: Whether the query-key-value projection has a bias.
        qk_norm: Whether queries and keys are RMS-normalized per head.
        rope: Whether to use rotary positional embedding (RoPE).
        dropout: The dropout rate in :math:`[0, 1]`.
    """

    def __init__(
        self,
        channels: int,
        pos_channels: int = 1,
        attention_heads: int = 1,
        qkv_bias: bool = True,
        qk_norm: bool = True,
        rope: bool = False,
        dropout: float | None = None,
    ) -> None:
        super().__init__()

        assert channels % attention_heads == 0

        self.qkv_proj = nn.Linear(channels, 3 * channels, bias=qkv_bias)
        self.y_proj = nn.Linear(channels, channels, bias=False)

        if not qk_norm:
            self.qk_norm = nn.Identity()
        elif hasattr(nn, "RMSNorm"):
            self.qk_norm = nn.RMSNorm(channels // attention_heads, elementwise_affine=False, eps=1e-5)
        else:
            self.qk_norm = RMSNorm(dim=-1, eps=1e-5)

        if rope:
            magnitude = torch.exp(math.log(1e-1) * torch.rand(channels // 2, 1))
            direction = torch.randn(channels // 2, pos_channels)
            direction = direction / torch.linalg.norm(direction, dim=-1, keepdim=True)

            self.theta_proj = nn.Linear(pos_channels, channels // 2, bias=False)
            self.theta_proj.weight.data.copy_(magnitude * direction)
        else:
            self.theta_proj = None

        self.heads = attention_heads
        self.dropout = 0.0 if dropout is None else dropout

    def _split(self, t: Tensor) -> Tensor:
        r""":math:`(*, L, H C) \to (*, H, L, C)`."""
        return t.unflatten(-1, (self.heads, -1)).movedim(-2, -3)

    def forward(self, x: Tensor, pos: Tensor | None = None, mask: BoolTensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input tokens :math:`x`, with shape :math:`(*, L, H \times C)`.
            pos: Optional position vectors :math:`p`, with shape :math:`(*, L, P)`.
            mask: Optional attention mask, with shape :math:`(L, L)`.

        Returns:
            The ouput tokens :math:`y`, with shape :math:`(*, L, H \times C)`.
        """
        q, k, v = (self._split(t) for t in self.qkv_proj(x).chunk(3, dim=-1))
        q, k = self.qk_norm(q), self.qk_norm(k)

        if self.theta_proj is not None:
            q, k = apply_rope(q, k, self._split(self.theta_proj(pos)))

        y = nn.functional.scaled_dot_product_attention(
            query=q, key=k, value=v, attn_mask=mask, dropout_p=self.dropout if self.training else 0.0
        )

        return self.y_proj(y.movedim(-3, -2).flatten(-2))


def apply_rope(q: Tensor, k: Tensor, theta: Tensor) -> tuple[Tensor, Tensor]:
    r"""Rotates consecutive channel pairs of :math:`q` and :math:`k` by the angles
    :math:`\theta` (``azula/nn/attention.py:124-156``), in at least float32."""
    dtype = torch.promote_types(q.dtype, k.dtype)
    wide = torch.promote_types(dtype, torch.float32)
    rot = torch.polar(torch.ones_like(theta, dtype=wide), theta.to(wide))

    def turn(t: Tensor) -> Tensor:
        z = torch.view_as_complex(t.to(wide).unflatten(-1, (-1, 2)).contiguous())
        return torch.view_as_real(z * rot).flatten(-2).to(dtype)

    return turn(q), turn(k)
