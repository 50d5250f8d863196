"""Oracle (test infrastructure): functional fp32 restatement of the in-repo backbones.

``azula.nn.unet.UNet`` and ``azula.nn.vit.ViT`` / ``azula.nn.dit.DiT`` of the reference are
nn.Module classes; here their *inference forward* is restated as pure functions over a reference
``state_dict`` (same keys) and the constructor arguments, so the CUDA engine and the host mirror
can be checked against something that shares no code with either.  Pinned against outputs of the
unmodified reference by ``oracle/gen_golden_nn.py`` -> ``tests/golden/nn_{unet,vit,dit,samplers}.npz``
(``tests/test_nn_cpu.py``).

Citations are ``file:line`` in the reference checkout.  Never imported by product code.
"""

from __future__ import annotations

import math
import torch
import torch.nn.functional as F

from torch import Tensor

EPS = 1e-5  # azula/nn/unet.py:52-60, azula/nn/dit.py:53, azula/nn/attention.py:53


# ------------------------------------------------------------------------------------------ shared
def ada_zero(sd: dict, key: str, mod: Tensor | None, trailing: int) -> tuple[Tensor, Tensor, Tensor]:
    """(a, b, c) of one block: Linear -> SiLU -> Linear -> split in 3 (azula/nn/unet.py:64-75,98-101;
    azula/nn/dit.py:57-67,96-99); `trailing` singleton dimensions are appended for broadcasting."""
    if key + ".ada_zero" in sd:  # mod_features = 0: free parameters
        abc = sd[key + ".ada_zero"]
        return abc[0], abc[1], abc[2]
    h = F.linear(mod, sd[key + ".ada_zero.0.weight"], sd[key + ".ada_zero.0.bias"])
    h = F.linear(F.silu(h), sd[key + ".ada_zero.2.weight"], sd[key + ".ada_zero.2.bias"])
    h = h.unflatten(-1, (3, -1)).movedim(-2, 0)
    h = h.reshape(*h.shape, *(1,) * trailing)
    return h[0], h[1], h[2]


def channel_layer_norm(x: Tensor, dim: int) -> Tensor:
    """azula/nn/layers.py:152-155: (x - mean) * rsqrt(var + eps) with torch.var_mean (unbiased)."""
    v, m = torch.var_mean(x, dim=dim, keepdim=True)
    return (x - m) * torch.rsqrt(v + EPS)


def rms_norm(x: Tensor, dim: int = -1) -> Tensor:
    """azula/nn/layers.py:192-195 == torch.nn.RMSNorm(elementwise_affine=False, eps=1e-5)."""
    return x * torch.rsqrt(x.square().mean(dim=dim, keepdim=True) + EPS)


# -------------------------------------------------------------------------------------------- UNet
def unet_block(sd: dict, key: str, x: Tensor, mod: Tensor | None, norm: str = "layer", groups: int = 16) -> Tensor:
    """UNetBlock._forward (azula/nn/unet.py:97-107): y = x + c * ffn((a + 1) * norm(x) + b)."""
    a, b, c = ada_zero(sd, key, mod, trailing=2)
    if norm == "layer":
        y = channel_layer_norm(x, dim=1)
    elif norm == "rms":
        y = rms_norm(x, dim=1)
    else:
        y = F.group_norm(x, min(groups, x.shape[1]), eps=EPS)  # azula/nn/unet.py:54-60 (affine=False)
    y = (a + 1) * y + b
    y = F.conv2d(y, sd[key + ".ffn.0.weight"], sd[key + ".ffn.0.bias"], padding=1)
    y = F.conv2d(F.silu(y), sd[key + ".ffn.3.weight"], sd[key + ".ffn.3.bias"], padding=1)
    return x + c * y


def unet_forward(sd: dict, x: Tensor, mod: Tensor | None, hid_blocks=(3, 3, 3), norm: str = "layer",
                 groups: int = 16) -> Tensor:
    """UNet.forward (azula/nn/unet.py:207-259) for spatial=2, kernel 3, stride 2, zero padding.

    Module order inside a level follows the constructor (:170-205): descent level i =
    [conv (stride 2 when i > 0), blocks...]; ascent level i = [conv (when i + 1 < depth), blocks...,
    Upsample (i > 0) | out conv (i = 0)], stored deepest level first.
    """
    depth = len(hid_blocks)
    memory = []
    for i in range(depth):
        memory.append(x if memory else None)  # :227-231
        key = f"descent.{i}"
        x = F.conv2d(x, sd[f"{key}.0.weight"], sd[f"{key}.0.bias"], stride=2 if i > 0 else 1, padding=1)
        for j in range(hid_blocks[i]):
            x = unet_block(sd, f"{key}.{j + 1}", x, mod, norm, groups)
    for pos in range(depth):
        i = depth - 1 - pos
        key = f"ascent.{pos}"
        j = 0
        if i + 1 < depth:
            x = F.conv2d(x, sd[f"{key}.0.weight"], sd[f"{key}.0.bias"], padding=1)
            j = 1
        for _ in range(hid_blocks[i]):
            x = unet_block(sd, f"{key}.{j}", x, mod, norm, groups)
            j += 1
        if i > 0:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")  # :194
        else:
            x = F.conv2d(x, sd[f"{key}.{j}.weight"], sd[f"{key}.{j}.bias"], padding=1)  # :197
        y = memory.pop()  # :249-257
        if y is None:
            continue
        x = x[..., : y.shape[2], : y.shape[3]]
        x = torch.cat((y, x), dim=1)
    return x


# ------------------------------------------------------------------------------------------ DiT / ViT
def sine_encoding(x: Tensor, features: int, omega: float) -> Tensor:
    """azula/nn/layers.py:283-299."""
    x = x.unsqueeze(-1)
    freqs = torch.exp(math.log(1 / omega) * torch.linspace(0, 1, features // 2, dtype=x.dtype, device=x.device))
    return torch.cat((torch.sin(x * freqs), torch.cos(x * freqs)), dim=-1)


def rope(t: Tensor, theta: Tensor) -> Tensor:
    """Rotation of consecutive channel pairs by theta (azula/nn/attention.py:135-156)."""
    re, im = t[..., 0::2], t[..., 1::2]
    cos, sin = torch.cos(theta), torch.sin(theta)
    return torch.stack((re * cos - im * sin, re * sin + im * cos), dim=-1).flatten(-2)


def self_attention(sd: dict, key: str, x: Tensor, heads: int, qk_norm: bool = True, pos: Tensor | None = None) -> Tensor:
    """MultiheadSelfAttention.forward without mask (azula/nn/attention.py:89-108):
    qkv split as "(n H C)", per-head RMS-normalised q and k, rotary embedding when the module has a
    ``theta_proj`` (:97-100: theta = theta_proj(pos) split per head), softmax(q k^T / sqrt(C)) v."""
    B, L, D = x.shape
    qkv = F.linear(x, sd[key + ".qkv_proj.weight"], sd.get(key + ".qkv_proj.bias"))
    q, k, v = qkv.reshape(B, L, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
    if qk_norm:
        q, k = rms_norm(q), rms_norm(k)
    if key + ".theta_proj.weight" in sd:
        theta = F.linear(pos, sd[key + ".theta_proj.weight"])  # (L, D / 2)
        theta = theta.reshape(L, heads, D // heads // 2).permute(1, 0, 2)  # (H, L, C / 2)
        q, k = rope(q, theta), rope(k, theta)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(D // heads), dim=-1)
    y = (att @ v).transpose(1, 2).reshape(B, L, D)
    return F.linear(y, sd[key + ".y_proj.weight"])


_ACT = {"silu": F.silu, "relu": F.relu, "relu2": lambda t: F.relu(t).square()}


def dit_block(sd: dict, key: str, x: Tensor, mod: Tensor | None, heads: int, qk_norm: bool = True,
              activation: str = "silu", pos: Tensor | None = None) -> Tensor:
    """DiTBlock._forward (azula/nn/dit.py:89-107)."""
    a, b, c = ada_zero(sd, key, mod, trailing=0)
    if a.ndim == 2:  # (B, C) -> (B, 1, C)   ("... (n C) -> n ... 1 C", :62)
        a, b, c = a[:, None], b[:, None], c[:, None]
    y = (a + 1) * rms_norm(x) + b
    y = y + self_attention(sd, key + ".msa", y, heads, qk_norm, pos)
    y = F.linear(y, sd[key + ".ffn.0.weight"], sd[key + ".ffn.0.bias"])
    y = F.linear(_ACT[activation](y), sd[key + ".ffn.3.weight"], sd[key + ".ffn.3.bias"])
    return x + c * y


def dit_forward(sd: dict, x: Tensor, mod: Tensor | None, pos: Tensor, hid_blocks: int, heads: int,
                qk_norm: bool = True, activation: str = "silu") -> Tensor:
    """DiT.forward (azula/nn/dit.py:180-218): tokens (B, L, C_i), positions (L, P)."""
    x = F.linear(x, sd["in_proj.weight"], sd["in_proj.bias"])
    hid = x.shape[-1]
    emb = sine_encoding(pos, hid, omega=1e2).flatten(-2)  # :153-157
    x = x + F.linear(emb, sd["pos_embedding.2.weight"])
    for i in range(hid_blocks):
        x = dit_block(sd, f"blocks.{i}", x, mod, heads, qk_norm, activation, pos)
    return F.linear(x, sd["out_proj.weight"], sd["out_proj.bias"])


def vit_forward(sd: dict, x: Tensor, mod: Tensor | None, patch: int, hid_blocks: int, heads: int,
                qk_norm: bool = True, activation: str = "silu", cond: Tensor | None = None) -> Tensor:
    """ViT.forward (azula/nn/vit.py:79-108): patchify channel-last "(Z a b)" (the condition image separately,
    concatenated per token, :97-100 with azula/nn/dit.py:200-201), grid positions, DiT, unpatchify."""
    B, C, H, W = x.shape
    hp, wp = H // patch, W // patch
    tokens = lambda t: t.reshape(B, t.shape[1], hp, patch, wp, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, hp * wp, -1)  # noqa: E731
    tok = tokens(x) if cond is None else torch.cat((tokens(x), tokens(cond)), dim=-1)
    ii, jj = torch.meshgrid(torch.arange(hp, dtype=x.dtype, device=x.device),
                            torch.arange(wp, dtype=x.dtype, device=x.device), indexing="ij")
    pos = torch.stack((ii.flatten(), jj.flatten()), dim=-1)  # cartesian_prod, :99-101
    y = dit_forward(sd, tok, mod, pos, hid_blocks, heads, qk_norm, activation)
    co = y.shape[-1] // (patch * patch)
    return y.reshape(B, hp, wp, co, patch, patch).permute(0, 3, 1, 4, 2, 5).reshape(B, co, H, W)
