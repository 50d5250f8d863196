"""Oracle (test infrastructure): functional fp32 restatement of the ADM UNet forward.

The reference vendors openai/guided-diffusion's ``UNetModel`` as nn.Module classes
(azula/plugins/adm/_src/unet.py).  This file restates its *inference forward* as pure
functions over a guided-diffusion ``state_dict`` and a small block table derived from the
card config, so that (i) it shares no structure with the reference classes and (ii) the
CUDA engine's per-block outputs can be checked against ``forward(..., taps=...)``.

Citations are ``file:line`` in the reference checkout.  Never imported by product code.
"""

from __future__ import annotations

import math
import torch
import torch.nn.functional as F

from dataclasses import dataclass, field
from torch import Tensor

GN_GROUPS = 32  # azula/plugins/adm/_src/nn.py:87
GN_EPS = 1e-5  # torch.nn.GroupNorm default


@dataclass
class Res:
    """One residual block: prefix in the state_dict, channels and resampling mode."""

    key: str
    cin: int
    cout: int
    mode: str = "same"  # "same" | "up" | "down"


@dataclass
class Attn:
    key: str
    ch: int
    heads: int


@dataclass
class Table:
    """Block table of one UNet: what the constructor at _src/unet.py:424-603 builds."""

    model_ch: int
    in_ch: int
    out_ch: int
    down: list[list] = field(default_factory=list)  # per input block: list of Res/Attn/("conv", key)
    mid: list = field(default_factory=list)
    up: list[list] = field(default_factory=list)
    num_classes: int | None = None
    new_attention_order: bool = False
    scale_shift: bool = True


def block_table(
    image_size: int = 64,
    image_channels: int = 3,
    learn_var: bool = True,
    num_channels: int = 128,
    channel_mult=(1, 2, 3, 4),
    num_res_blocks: int = 2,
    attention_resolutions=(32, 16, 8),
    num_heads: int = 1,
    num_head_channels: int = -1,
    num_classes: int | None = None,
    resblock_updown: bool = False,
    use_scale_shift_norm: bool = False,
    use_new_attention_order: bool = False,
    **_ignored,
) -> Table:
    """Derives the block table from a card config.

    Mirrors the loop structure of _src/unet.py:468-603 and the argument mapping of
    azula/plugins/adm/__init__.py:164-202 (attention at downsample rates image_size // r).
    Only ``resblock_updown=True`` style resampling and 2-d data are restated (all cards
    reachable from the scope table use them, cards.yaml:1-98).
    """
    assert resblock_updown, "oracle restates the resblock_updown variant only"
    rates = {image_size // r for r in attention_resolutions}
    out_ch = 2 * image_channels if learn_var else image_channels

    def heads(ch):
        return num_heads if num_head_channels == -1 else ch // num_head_channels

    tab = Table(
        model_ch=num_channels,
        in_ch=image_channels,
        out_ch=out_ch,
        num_classes=num_classes,
        new_attention_order=use_new_attention_order,
        scale_shift=use_scale_shift_norm,
    )
    ch = int(channel_mult[0] * num_channels)
    tab.down.append([("conv", "input_blocks.0.0")])
    skips = [ch]
    ds, idx = 1, 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            co = int(mult * num_channels)
            blk = [Res(f"input_blocks.{idx}.0", ch, co)]
            ch = co
            if ds in rates:
                blk.append(Attn(f"input_blocks.{idx}.1", ch, heads(ch)))
            tab.down.append(blk)
            skips.append(ch)
            idx += 1
        if level != len(channel_mult) - 1:
            tab.down.append([Res(f"input_blocks.{idx}.0", ch, ch, "down")])
            skips.append(ch)
            idx += 1
            ds *= 2
    tab.mid = [
        Res("middle_block.0", ch, ch),
        Attn("middle_block.1", ch, heads(ch)),
        Res("middle_block.2", ch, ch),
    ]
    idx = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = skips.pop()
            co = int(num_channels * mult)
            blk = [Res(f"output_blocks.{idx}.0", ch + ich, co)]
            ch = co
            if ds in rates:
                blk.append(Attn(f"output_blocks.{idx}.{len(blk)}", ch, heads(ch)))
            if level and i == num_res_blocks:
                blk.append(Res(f"output_blocks.{idx}.{len(blk)}", ch, ch, "up"))
                ds //= 2
            tab.up.append(blk)
            idx += 1
    return tab


# --------------------------------------------------------------------------------- pieces


def time_features(timesteps: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """[cos | sin] sinusoidal features. Follows _src/nn.py:90-108."""
    half = dim // 2
    freqs = torch.exp(
        -math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half
    ).to(timesteps.device)
    ang = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    if dim % 2:
        emb = F.pad(emb, (0, 1))
    return emb


def _gn(sd, key, x):
    return F.group_norm(x, GN_GROUPS, sd[key + ".weight"], sd[key + ".bias"], GN_EPS)


def _conv(sd, key, x, pad):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=pad)


def res_block(sd, b: Res, x: Tensor, emb: Tensor, scale_shift: bool = True) -> Tensor:
    """Follows _src/unet.py:227-247 (and the up/down branch :228-233)."""
    h = F.silu(_gn(sd, b.key + ".in_layers.0", x))
    if b.mode == "up":
        h = F.interpolate(h, scale_factor=2, mode="nearest")  # :106
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif b.mode == "down":
        h = F.avg_pool2d(h, 2, 2)  # :133
        x = F.avg_pool2d(x, 2, 2)
    h = _conv(sd, b.key + ".in_layers.2", h, 1)
    e = F.linear(F.silu(emb), sd[b.key + ".emb_layers.1.weight"], sd[b.key + ".emb_layers.1.bias"])
    e = e[..., None, None].to(h.dtype)
    if scale_shift:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn(sd, b.key + ".out_layers.0", h) * (1 + scale) + shift
        h = F.silu(h)
    else:
        h = F.silu(_gn(sd, b.key + ".out_layers.0", h + e))
    h = _conv(sd, b.key + ".out_layers.3", h, 1)
    if b.cin != b.cout:
        w = sd[b.key + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[b.key + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def attention(sd, b: Attn, x: Tensor, new_order: bool = False) -> Tensor:
    """Follows _src/unet.py:290-296 with QKVAttentionLegacy :328-345 / QKVAttention :361-381."""
    n, c, hh, ww = x.shape
    t = hh * ww
    xf = x.reshape(n, c, t)
    y = F.group_norm(xf, GN_GROUPS, sd[b.key + ".norm.weight"], sd[b.key + ".norm.bias"], GN_EPS)
    qkv = F.conv1d(y, sd[b.key + ".qkv.weight"], sd[b.key + ".qkv.bias"])
    d = c // b.heads
    if new_order:
        q, k, v = (z.reshape(n * b.heads, d, t) for z in qkv.chunk(3, dim=1))
    else:
        q, k, v = qkv.reshape(n * b.heads, 3 * d, t).split(d, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    w = torch.einsum("bct,bcs->bts", q * s, k * s)
    w = torch.softmax(w.float(), dim=-1).to(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(n, c, t)
    a = F.conv1d(a, sd[b.key + ".proj_out.weight"], sd[b.key + ".proj_out.bias"])
    return (xf + a).reshape(n, c, hh, ww)


def forward(
    sd: dict[str, Tensor],
    tab: Table,
    x: Tensor,
    timesteps: Tensor,
    y: Tensor | None = None,
    taps: dict | None = None,
) -> Tensor:
    """UNet forward. Follows _src/unet.py:605-634.

    ``taps`` (optional dict) receives every block output keyed by its state_dict prefix.
    """
    emb = time_features(timesteps, tab.model_ch).to(sd["time_embed.0.weight"].dtype)  # (float64 state: the tie-breaker run)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if tab.num_classes is not None:
        emb = emb + sd["label_emb.weight"][y]

    def run(blk, h):
        for b in blk:
            if isinstance(b, Res):
                h = res_block(sd, b, h, emb, tab.scale_shift)
            elif isinstance(b, Attn):
                h = attention(sd, b, h, tab.new_attention_order)
            else:
                h = _conv(sd, b[1], h, 1)
            if taps is not None:
                taps[b.key if not isinstance(b, tuple) else b[1]] = h
        return h

    hs = []
    h = x
    for blk in tab.down:
        h = run(blk, h)
        hs.append(h)
    h = run(tab.mid, h)
    for blk in tab.up:
        h = run(blk, torch.cat([h, hs.pop()], dim=1))
    h = F.silu(_gn(sd, "out.0", h))
    return _conv(sd, "out.2", h, 1)


def seeded_state(reference_state: dict[str, Tensor], seed: int = 1234) -> dict[str, Tensor]:
    """Overwrites EVERY tensor of a state_dict from a seeded CPU generator.

    A default-initialised ADM outputs exactly 0 (zero_module, _src/unet.py:207,285,602) and
    ``skip_init`` leaves memory uninitialised, so fixtures must not rely on module init.
    Matrices/convs ~ N(0, 1/sqrt(fan_in)) scaled by 0.7, biases ~ N(0, 0.02), norm gains 1+N(0,0.1).
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(reference_state):
        v = reference_state[k]
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        r = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if v.ndim >= 2:
            fan_in = v[0].numel()
            r = r * (0.7 / math.sqrt(fan_in))
        elif k.endswith("weight"):
            r = 1 + 0.1 * r
        else:
            r = 0.02 * r
        out[k] = r.to(v.dtype)
    return out


def state_shapes(tab: Table) -> dict[str, tuple[int, ...]]:
    """Names and shapes of the guided-diffusion ``state_dict`` for a block table.

    Matches what the constructor registers at _src/unet.py:458-603 (ResBlock :177-219,
    AttentionBlock :274-285); lets tests build seeded weights without the reference.
    """
    shp: dict[str, tuple[int, ...]] = {}
    emb = 4 * tab.model_ch

    def lin(key, i, o):
        shp[key + ".weight"], shp[key + ".bias"] = (o, i), (o,)

    def conv(key, i, o, k):
        shp[key + ".weight"], shp[key + ".bias"] = (o, i, k, k), (o,)

    def norm(key, c):
        shp[key + ".weight"], shp[key + ".bias"] = (c,), (c,)

    lin("time_embed.0", tab.model_ch, emb)
    lin("time_embed.2", emb, emb)
    if tab.num_classes is not None:
        shp["label_emb.weight"] = (tab.num_classes, emb)
    for blk in [*tab.down, tab.mid, *tab.up]:
        for b in blk:
            if isinstance(b, Res):
                norm(b.key + ".in_layers.0", b.cin)
                conv(b.key + ".in_layers.2", b.cin, b.cout, 3)
                lin(b.key + ".emb_layers.1", emb, 2 * b.cout if tab.scale_shift else b.cout)
                norm(b.key + ".out_layers.0", b.cout)
                conv(b.key + ".out_layers.3", b.cout, b.cout, 3)
                if b.cin != b.cout:
                    conv(b.key + ".skip_connection", b.cin, b.cout, 1)
            elif isinstance(b, Attn):
                norm(b.key + ".norm", b.ch)
                shp[b.key + ".qkv.weight"], shp[b.key + ".qkv.bias"] = (3 * b.ch, b.ch, 1), (3 * b.ch,)
                shp[b.key + ".proj_out.weight"], shp[b.key + ".proj_out.bias"] = (b.ch, b.ch, 1), (b.ch,)
            else:
                conv(b[1], tab.in_ch, int(tab.down[1][0].cin), 3)
    norm("out.0", tab.up[-1][-1].cout)
    conv("out.2", tab.up[-1][-1].cout, tab.out_ch, 3)
    return shp
