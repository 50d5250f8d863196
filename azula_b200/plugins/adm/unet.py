r"""The ADM (guided-diffusion) U-Net as a flat *layout* plus two executors.

The reference vendors openai/guided-diffusion's ``UNetModel`` as a tree of ``nn.Module``
classes (``azula/plugins/adm/_src/unet.py:387-634``).  Here the network is described once by a
:class:`Layout` -- an ordered list of units (stem convolution, residual units, attention units)
derived from the card configuration -- and the module only *owns the parameters*, registered
under the checkpoint's own names (``input_blocks.N.M...``, ``middle_block...``,
``output_blocks...``, ``time_embed...``, ``out...``) so that a guided-diffusion ``state_dict``
loads unchanged.  Two executors interpret the layout:

* CUDA tensors under ``torch.no_grad``: :mod:`azula_b200.engine.adm` -- NHWC bf16 activations,
  tcgen05 implicit-GEMM convolutions, fused GroupNorm/SiLU/scale-shift/resample passes and
  flash-style attention, all through the C ABI of ``libazb.so`` (no torch arithmetic);
* everything else (CPU tensors, autograd): :func:`forward_torch`, the same maths in plain
  fp32 torch ops, which doubles as the differentiable path guidance methods need.
"""

from __future__ import annotations

__all__ = ["Unit", "Layout", "UNetModel", "forward_torch", "timestep_embedding"]

import math
import torch
import torch.nn as nn
import torch.nn.functional as F

from collections.abc import Sequence
from dataclasses import dataclass, field
from torch import Tensor

GROUPS = 32  # normalization() of the reference: GroupNorm(32, C), _src/nn.py:80-87
SAME, UP, DOWN = 0, 1, 2  # resampling codes shared with azb_gn_apply_bf16 (include/azb.h)


@dataclass(frozen=True)
class Unit:
    r"""One executable unit of the network."""

    kind: str  # "stem" | "res" | "attn"
    path: str  # prefix of its parameters in the state_dict
    cin: int
    cout: int
    resample: int = SAME  # res only
    heads: int = 0  # attn only


@dataclass
class Layout:
    r"""What ``UNetModel.__init__`` of the reference builds (``_src/unet.py:468-603``), as data."""

    in_channels: int
    out_channels: int
    model_channels: int
    num_classes: int | None
    new_attention_order: bool
    scale_shift: bool
    dropout: float
    encoder: list[list[Unit]] = field(default_factory=list)  # input_blocks; every output is a skip
    middle: list[Unit] = field(default_factory=list)
    decoder: list[list[Unit]] = field(default_factory=list)  # output_blocks; each pops one skip

    @property
    def embed_dim(self) -> int:
        return 4 * self.model_channels

    @property
    def final_channels(self) -> int:
        return self.decoder[-1][-1].cout

    def units(self):
        for block in (*self.encoder, self.middle, *self.decoder):
            yield from block

    def res_units(self) -> list[Unit]:
        return [u for u in self.units() if u.kind == "res"]


def make_layout(
    in_channels: int,
    model_channels: int,
    out_channels: int,
    num_res_blocks: int,
    attention_resolutions,
    dropout: float = 0.0,
    channel_mult: Sequence[int] = (1, 2, 4, 8),
    num_classes: int | None = None,
    num_heads: int = 1,
    num_head_channels: int = -1,
    num_heads_upsample: int = -1,
    use_scale_shift_norm: bool = False,
    resblock_updown: bool = False,
    use_new_attention_order: bool = False,
) -> Layout:
    r"""Derives the unit list from the constructor arguments of the reference
    (``_src/unet.py:420-438``); ``attention_resolutions`` are downsampling RATES."""
    if not resblock_updown:
        raise NotImplementedError("only resblock_updown=True networks are supported (all ADM cards use it)")
    if num_heads_upsample == -1:
        num_heads_upsample = num_heads

    def heads_of(ch: int, n: int) -> int:
        return n if num_head_channels == -1 else ch // num_head_channels

    lay = Layout(
        in_channels=in_channels,
        out_channels=out_channels,
        model_channels=model_channels,
        num_classes=num_classes,
        new_attention_order=use_new_attention_order,
        scale_shift=use_scale_shift_norm,
        dropout=dropout,
    )
    rates = set(attention_resolutions)
    width = int(channel_mult[0] * model_channels)
    lay.encoder.append([Unit("stem", "input_blocks.0.0", in_channels, width)])
    pending = [width]  # channels of the skips, in production order
    rate = 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            n = len(lay.encoder)
            block = [Unit("res", f"input_blocks.{n}.0", width, int(mult * model_channels))]
            width = block[0].cout
            if rate in rates:
                block.append(Unit("attn", f"input_blocks.{n}.1", width, width, heads=heads_of(width, num_heads)))
            lay.encoder.append(block)
            pending.append(width)
        if level + 1 < len(channel_mult):
            n = len(lay.encoder)
            lay.encoder.append([Unit("res", f"input_blocks.{n}.0", width, width, resample=DOWN)])
            pending.append(width)
            rate *= 2
    lay.middle = [
        Unit("res", "middle_block.0", width, width),
        Unit("attn", "middle_block.1", width, width, heads=heads_of(width, num_heads)),
        Unit("res", "middle_block.2", width, width),
    ]
    for level in reversed(range(len(channel_mult))):
        for i in range(num_res_blocks + 1):
            n = len(lay.decoder)
            cout = int(model_channels * channel_mult[level])
            block = [Unit("res", f"output_blocks.{n}.0", width + pending.pop(), cout)]
            width = cout
            if rate in rates:
                block.append(
                    Unit("attn", f"output_blocks.{n}.{len(block)}", width, width, heads=heads_of(width, num_heads_upsample))
                )
            if level and i == num_res_blocks:
                block.append(Unit("res", f"output_blocks.{n}.{len(block)}", width, width, resample=UP))
                rate //= 2
            lay.decoder.append(block)
    return lay


# ------------------------------------------------------------------------------- parameters


def _at(**children: nn.Module) -> nn.Module:
    r"""A bare container whose children sit at the given (numeric) names, e.g. ``_at(_0=..., _2=...)``."""
    box = nn.Module()
    for name, child in children.items():
        box.add_module(name.lstrip("_"), child)
    return box


def _res_parameters(u: Unit, embed_dim: int, scale_shift: bool) -> nn.Module:
    box = nn.Module()
    box.in_layers = _at(_0=nn.GroupNorm(GROUPS, u.cin), _2=nn.Conv2d(u.cin, u.cout, 3, padding=1))
    box.emb_layers = _at(_1=nn.Linear(embed_dim, 2 * u.cout if scale_shift else u.cout))
    box.out_layers = _at(_0=nn.GroupNorm(GROUPS, u.cout), _3=nn.Conv2d(u.cout, u.cout, 3, padding=1))
    if u.cin != u.cout:
        box.skip_connection = nn.Conv2d(u.cin, u.cout, 1)
    return box


def _attn_parameters(u: Unit) -> nn.Module:
    box = nn.Module()
    box.norm = nn.GroupNorm(GROUPS, u.cin)
    box.qkv = nn.Conv1d(u.cin, 3 * u.cin, 1)
    box.proj_out = nn.Conv1d(u.cin, u.cin, 1)
    return box


class UNetModel(nn.Module):
    r"""ADM U-Net: ``forward(x, timesteps, y=None)`` with the reference's signature
    (``_src/unet.py:605-634``) and ``state_dict`` naming.

    Arguments are those of the reference constructor (``_src/unet.py:420-438``).
    """

    def __init__(
        self,
        image_size: int,
        in_channels: int,
        model_channels: int,
        out_channels: int,
        num_res_blocks: int,
        attention_resolutions,
        dropout: float = 0,
        channel_mult: Sequence[int] = (1, 2, 4, 8),
        conv_resample: bool = True,
        dims: int = 2,
        num_classes: int | None = None,
        use_checkpoint: bool = False,
        num_heads: int = 1,
        num_head_channels: int = -1,
        num_heads_upsample: int = -1,
        use_scale_shift_norm: bool = False,
        resblock_updown: bool = False,
        use_new_attention_order: bool = False,
    ) -> None:
        super().__init__()
        if dims != 2:
            raise NotImplementedError("only 2-d data is supported")

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_classes = num_classes
        self.layout = lay = make_layout(
            in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout,
            tuple(channel_mult), num_classes, num_heads, num_head_channels, num_heads_upsample,
            use_scale_shift_norm, resblock_updown, use_new_attention_order,
        )

        self.time_embed = _at(_0=nn.Linear(model_channels, lay.embed_dim), _2=nn.Linear(lay.embed_dim, lay.embed_dim))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, lay.embed_dim)
        self.input_blocks, self.middle_block, self.output_blocks = nn.Module(), nn.Module(), nn.Module()
        self.out = _at(_0=nn.GroupNorm(GROUPS, lay.final_channels),
                       _2=nn.Conv2d(lay.final_channels, out_channels, 3, padding=1))
        for u in lay.units():
            if u.kind == "stem":
                params = nn.Conv2d(u.cin, u.cout, 3, padding=1)
            elif u.kind == "res":
                params = _res_parameters(u, lay.embed_dim, lay.scale_shift)
            else:
                params = _attn_parameters(u)
            self._plant(u.path, params)
        # zero_module() sites of the reference (_src/unet.py:207,285,602)
        with torch.no_grad():
            for u in lay.units():
                if u.kind == "res":
                    last = self.get_submodule(u.path + ".out_layers.3")
                elif u.kind == "attn":
                    last = self.get_submodule(u.path + ".proj_out")
                else:
                    continue
                nn.init.zeros_(last.weight), nn.init.zeros_(last.bias)
            nn.init.zeros_(self.out.get_submodule("2").weight), nn.init.zeros_(self.out.get_submodule("2").bias)

        self._native = {}  # (device, signature) -> engine plan, see azula_b200.engine.adm
        # "bf16" (default): the fast native path, bf16 activations and tensor-core operands, stated bf16 tolerance;
        # "tf32": the reference-numerics path (engine/adm_tf32.py): fp32 activations, TF32 operands, fp32 everything else
        self.precision = "bf16"

    def _plant(self, path: str, module: nn.Module) -> None:
        node = self
        *parents, leaf = path.split(".")
        for name in parents:
            if not hasattr(node, name):
                node.add_module(name, nn.Module())
            node = getattr(node, name)
        node.add_module(leaf, module)

    def forward(self, x: Tensor, timesteps: Tensor, y: Tensor | None = None) -> Tensor:
        r"""
        Arguments:
            x: The input :math:`(N, C, H, W)`.
            timesteps: The (discrete) time steps, shape :math:`(N)` or :math:`(1)`.
            y: Class labels :math:`(N)` for class-conditional networks.

        Returns:
            The output :math:`(N, C', H, W)`.
        """
        if (y is not None) != (self.num_classes is not None):
            raise ValueError("must specify y if and only if the model is class-conditional")
        if x.is_cuda and not torch.is_grad_enabled() and not (self.training and self.layout.dropout > 0):
            from ... import engine as _engine
            from ...engine import adm as engine

            if _engine.native_enabled():
                if getattr(self, "precision", "bf16") == "tf32":  # reference numerics: fp32 activations, TF32 contractions
                    from ...engine import adm_tf32

                    return adm_tf32.forward(self, x, timesteps, y)
                return engine.forward(self, x, timesteps, y)
        state = dict(self.named_parameters())
        return forward_torch(self.layout, state, x, timesteps, y, dropout=self.layout.dropout if self.training else 0.0)

    def _apply(self, fn, *args, **kwargs):
        self._native.clear()  # packed weights and plans belong to the old device / dtype
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_native"] = {}  # launch plans hold device pointers: never copied or pickled
        return state


# ----------------------------------------------------------------- plain torch executor (fp32)


def timestep_embedding(timesteps: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    r"""``[cos | sin]`` features of ``_src/nn.py:90-108``."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat((torch.cos(args), torch.sin(args)), dim=-1)
    return F.pad(emb, (0, 1)) if dim % 2 else emb


def _norm(p, path: str, x: Tensor) -> Tensor:
    return F.group_norm(x.float(), GROUPS, p[path + ".weight"], p[path + ".bias"], 1e-5).to(x.dtype)


def _res_torch(p, u: Unit, x: Tensor, emb: Tensor, scale_shift: bool, dropout: float) -> Tensor:
    h = F.silu(_norm(p, u.path + ".in_layers.0", x))
    if u.resample == UP:
        h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
    elif u.resample == DOWN:
        h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    h = F.conv2d(h, p[u.path + ".in_layers.2.weight"], p[u.path + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), p[u.path + ".emb_layers.1.weight"], p[u.path + ".emb_layers.1.bias"]).to(h.dtype)
    e = e[:, :, None, None]
    if scale_shift:
        scale, shift = e.chunk(2, dim=1)
        h = F.silu(_norm(p, u.path + ".out_layers.0", h) * (1 + scale) + shift)
    else:
        h = F.silu(_norm(p, u.path + ".out_layers.0", h + e))
    if dropout > 0:
        h = F.dropout(h, dropout)
    h = F.conv2d(h, p[u.path + ".out_layers.3.weight"], p[u.path + ".out_layers.3.bias"], padding=1)
    if u.cin != u.cout:
        x = F.conv2d(x, p[u.path + ".skip_connection.weight"], p[u.path + ".skip_connection.bias"])
    return x + h


def _attn_torch(p, u: Unit, x: Tensor, new_order: bool) -> Tensor:
    n, c, hh, ww = x.shape
    seq = x.reshape(n, c, hh * ww)
    qkv = F.conv1d(_norm(p, u.path + ".norm", seq), p[u.path + ".qkv.weight"], p[u.path + ".qkv.bias"])
    d = c // u.heads
    if new_order:  # q | k | v blocks, each split into heads (_src/unet.py:361-381)
        q, k, v = (part.reshape(n * u.heads, d, -1) for part in qkv.chunk(3, dim=1))
    else:  # per head q | k | v (_src/unet.py:328-345)
        q, k, v = qkv.reshape(n * u.heads, 3 * d, -1).split(d, dim=1)
    scale = d**-0.25
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1).to(weight.dtype)
    a = torch.einsum("bts,bcs->bct", weight, v).reshape(n, c, -1)
    a = F.conv1d(a, p[u.path + ".proj_out.weight"], p[u.path + ".proj_out.bias"])
    return (seq + a).reshape(n, c, hh, ww)


def forward_torch(lay: Layout, p: dict[str, Tensor], x: Tensor, timesteps: Tensor, y: Tensor | None = None,
                  dropout: float = 0.0) -> Tensor:
    r"""The network in plain torch ops on NCHW tensors (any device, differentiable)."""
    emb = timestep_embedding(timesteps, lay.model_channels)
    emb = F.linear(emb, p["time_embed.0.weight"], p["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), p["time_embed.2.weight"], p["time_embed.2.bias"])
    if lay.num_classes is not None:
        emb = emb + p["label_emb.weight"][y]

    def run(block: list[Unit], h: Tensor) -> Tensor:
        for u in block:
            if u.kind == "res":
                h = _res_torch(p, u, h, emb, lay.scale_shift, dropout)
            elif u.kind == "attn":
                h = _attn_torch(p, u, h, lay.new_attention_order)
            else:
                h = F.conv2d(h, p[u.path + ".weight"], p[u.path + ".bias"], padding=1)
        return h

    skips = []
    h = x.to(p["time_embed.0.weight"].dtype)
    for block in lay.encoder:
        h = run(block, h)
        skips.append(h)
    h = run(lay.middle, h)
    for block in lay.decoder:
        h = run(block, torch.cat((h, skips.pop()), dim=1))
    h = F.silu(_norm(p, "out.0", h.to(x.dtype)))
    return F.conv2d(h, p["out.2.weight"], p["out.2.bias"], padding=1)
