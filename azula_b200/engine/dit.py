r"""Native sm_100a executor of :class:`azula_b200.nn.dit.DiT` / :class:`azula_b200.nn.vit.ViT`.

Replaces ``DiT.forward`` / ``ViT.forward`` of the reference (``azula/nn/dit.py:180-218``,
``azula/nn/vit.py:79-108``, blocks ``azula/nn/dit.py:89-107``, attention ``azula/nn/attention.py:
101-121``) by a launch plan over bf16 token matrices ``(B L, C)``:

    azb_patchify_f32            pixels -> tokens (ViT)                                  nn/vit.py:97
    azb_conv2d_bf16 (taps = 1)  in_proj + bias + positional embedding (as the residual) nn/dit.py:206-212
    per block
      azb_rownorm_mod_bf16      y = (1 + a) * RMSNorm(x) + b                            nn/dit.py:102-103
      azb_conv2d_bf16           qkv = y W^T + bias                                      nn/attention.py:101
      azb_segment_rmsnorm_bf16  per-head RMS norm of q and k, in place                  nn/attention.py:103
      azb_attention_bf16        softmax(q k^T / sqrt(d)) v, logits stay on chip         nn/attention.py:110-116
      azb_conv2d_bf16           y2 = y + attn W_y^T            (residual epilogue)      nn/attention.py:118-119, nn/dit.py:104
      azb_conv2d_bf16           h = act(y2 W_1^T + b_1)        (activation epilogue)    nn/dit.py:105
      azb_conv2d_bf16           out = x + c * (h W_2^T + b_2)  (gate + residual)        nn/dit.py:105-106
    azb_conv2d_bf16 (fp32 out)  out_proj, then azb_unpatchify_f32 (ViT)                 nn/dit.py:216, nn/vit.py:105

The positional embedding depends only on the token grid: it is evaluated once per plan with the
module's own torch layers and kept as a bf16 ``(B L, C)`` residual.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from torch import Tensor

from .. import _lib
from . import ops
from .plan import LaunchPlan, ModulationBank, fingerprint, note_use

_MAX_PLANS = 2
_ACTS = {"silu": 1, "relu": 2, "relu2": 3}
FOLD_QK_NORM = True  # A/B switch: q / k RMS normalisation inside the short-sequence attention kernel


def structure_ok(model) -> bool:
    cached = model._native.get("structure_ok")
    if cached is not None:
        return cached
    ok = True
    hid = model.in_proj.out_features
    ok &= hid % 8 == 0 and hid <= 2048
    feats = set()
    for b in model.blocks:
        msa = b.msa
        d = hid // msa.heads
        ok &= b.ffn_activation in _ACTS
        ok &= d in (16, 32, 64, 128, 256)
        ok &= msa.theta_proj is None or msa.theta_proj.out_features * 2 == hid
        ok &= isinstance(msa.qk_norm, (nn.Identity, nn.RMSNorm)) if hasattr(nn, "RMSNorm") else isinstance(msa.qk_norm, nn.Identity)
        if hasattr(nn, "RMSNorm") and isinstance(msa.qk_norm, nn.RMSNorm):
            ok &= msa.qk_norm.weight is None
        ok &= b.ffn[0].out_features % 8 == 0
        ok &= hasattr(nn, "RMSNorm") and isinstance(b.norm, nn.RMSNorm) and b.norm.weight is None
        feats.add(0 if torch.is_tensor(b.ada_zero) else b.ada_zero[0].in_features)
    ok &= len(feats) <= 1
    model._native["structure_ok"] = bool(ok)
    return bool(ok)


def _common_ok(model, mod: Tensor | None, batch: int) -> bool:
    if not structure_ok(model):
        return False
    if any(isinstance(m, nn.Dropout) and m.training and m.p > 0 for m in model.modules()):
        return False
    if any(b.msa.training and b.msa.dropout > 0 for b in model.blocks):
        return False
    if mod is not None and (mod.ndim not in (1, 2) or (mod.ndim == 2 and mod.shape[0] not in (1, batch))):
        return False
    return True


def supports(model, x: Tensor, mod: Tensor | None, pos) -> bool:
    r"""Token interface: native when ``pos`` is the canonical sequence index (``pos="arange"``) or ONE set of position
    vectors (L, P) shared by the batch (per-sample positions take the torch path)."""
    if not (x.ndim == 3 and x.is_floating_point() and x.numel() > 0 and _common_ok(model, mod, x.shape[0])):
        return False
    if isinstance(pos, str):
        return True
    return torch.is_tensor(pos) and pos.ndim == 2 and pos.shape[0] == x.shape[1] and pos.is_floating_point()


def supports_image(model, x: Tensor, mod: Tensor | None, cond: Tensor | None) -> bool:
    if x.ndim != 4 or model.spatial != 2 or not x.is_floating_point() or x.numel() == 0:
        return False
    if cond is not None and (cond.ndim != 4 or cond.shape[0] != x.shape[0] or cond.shape[2:] != x.shape[2:]):
        return False
    p, q = model.patch.patch_shape
    if x.shape[2] % p or x.shape[3] % q or model.unpatch.patch_shape != model.patch.patch_shape:
        return False
    return _common_ok(model, mod, x.shape[0])


class Packed:
    r"""Kernel-layout copy of the parameters on one device."""

    def __init__(self, model, device: torch.device) -> None:
        self.fingerprint = fingerprint(model)
        self.device = device
        pk = lambda lin: ops.pack_conv(lin.weight.detach().to(device), None if lin.bias is None else lin.bias.detach().to(device))  # noqa: E731
        w_in = model.in_proj.weight.detach().to(device)
        k8 = -(-w_in.shape[1] // 8) * 8  # the GEMM reads channels in 16-byte vectors: zero-pad the input width
        w_in = torch.nn.functional.pad(w_in, (0, k8 - w_in.shape[1]))
        self.in_proj = ops.pack_conv(w_in, None if model.in_proj.bias is None else model.in_proj.bias.detach().to(device))
        self.out_proj = pk(model.out_proj)
        self.k_pad = self.in_proj.k_per_tap
        self.block = []
        for b in model.blocks:
            self.block.append({
                "qkv": pk(b.msa.qkv_proj), "y": pk(b.msa.y_proj), "ffn1": pk(b.ffn[0]), "ffn2": pk(b.ffn[3]),
                "heads": b.msa.heads, "qk_norm": not isinstance(b.msa.qk_norm, nn.Identity), "rope": b.msa.theta_proj is not None,
                "qk_eps": getattr(b.msa.qk_norm, "eps", None) or 1e-5, "eps": b.norm.eps or 1e-5,
                "act": _ACTS[b.ffn_activation],
            })
        self.bank = ModulationBank(list(model.blocks), device)


class Plan(LaunchPlan):
    r"""The launch list of one (batch, tokens, modulation rows, positions) signature."""

    def __init__(self, model, packed: Packed, batch: int, tokens: int, rows: int, pos: Tensor, device: torch.device) -> None:
        super().__init__(device)
        self.packed = packed
        arena = self.arena
        B, L = batch, tokens
        R = B * L
        hid = packed.in_proj.c_out
        self.grid = ops.token_grid(R)
        self.hid_buf, self.abc = packed.bank.buffers(rows, device)
        self.mod_ld = self.abc.stride(0) if (rows == B and B > 1) else 0

        with torch.no_grad():
            posf = pos.to(device=device, dtype=torch.float32)
            emb = model.pos_embedding(posf).to(torch.bfloat16)  # (L, C)
            # rotary embedding: the angles depend only on the positions -> {cos, sin} tables (L, C / 2) per block
            self.rot = []
            for b in model.blocks:
                if b.msa.theta_proj is None:
                    self.rot.append(None)
                else:
                    theta = b.msa.theta_proj(posf.to(b.msa.theta_proj.weight.dtype)).to(torch.float32)  # (L, C / 2), head-major
                    self.rot.append(torch.stack((torch.cos(theta), torch.sin(theta)), dim=-1).contiguous())
        self.posemb = emb.expand(B, L, hid).reshape(R, hid).contiguous()

        self.tok = arena.pin(arena.take(R, packed.k_pad))
        self.tok.zero_()
        x = arena.take(R, hid)
        self.conv(self.tok, packed.in_proj, x, grid=self.grid, residual=self.posemb)
        for (b, w), rot in zip(zip(model.blocks, packed.block, strict=True), self.rot, strict=True):
            off = packed.bank.offset[id(b)]
            abc = self.abc.data_ptr() + 4 * off
            y = arena.take(R, hid)
            self.rownorm(x, y, 1, abc, self.mod_ld, L, eps=w["eps"])
            qkv = arena.take(R, 3 * hid)
            self.conv(y, w["qkv"], qkv, grid=self.grid)
            heads, d = w["heads"], hid // w["heads"]
            # short sequences, no rotation: the q / k RMS normalisation is folded into the attention kernel's logits
            fold = w["qk_norm"] and rot is None and L <= 256 and d == 64 and FOLD_QK_NORM
            if fold:
                pass
            elif rot is not None:  # RMS norm (if any) + rotary embedding of q and k in one in-place pass
                self.keep += [qkv, rot]
                self._emit("qk_norm", 0.0, 2.0 * 2 * R * 2 * hid, self.lib.azb_qk_norm_rope_bf16, qkv.data_ptr(), 3 * hid, R,
                           heads, d, int(w["qk_norm"]), w["qk_eps"], rot.data_ptr(), L, desc=f"{R}x{2 * heads}x{d} +rope")
            elif w["qk_norm"]:
                self.keep.append(qkv)
                self._emit("qk_norm", 0.0, 2.0 * 2 * R * 2 * hid, self.lib.azb_segment_rmsnorm_bf16, qkv.data_ptr(), 3 * hid,
                           R, 2 * heads, d, w["qk_eps"], desc=f"{R}x{2 * heads}x{d}")
            att = arena.take(R, hid)
            self.keep += [qkv, att]
            if fold:
                self._emit("attention", 4.0 * B * heads * L * L * d, 2.0 * R * 4 * hid, self.lib.azb_attention_qknorm_bf16,
                           qkv.data_ptr(), 3 * hid, att.data_ptr(), hid, B, L, heads, d, d, hid, 2 * hid, w["qk_eps"],
                           desc=f"{B}x{heads}x{L}x{d} +qk-norm")
            else:
                self._emit("attention", 4.0 * B * heads * L * L * d, 2.0 * R * 4 * hid, self.lib.azb_attention_bf16,
                           qkv.data_ptr(), 3 * hid, att.data_ptr(), hid, B, L, heads, d, d, hid, 2 * hid,
                           desc=f"{B}x{heads}x{L}x{d}")
            arena.give(qkv)
            y2 = arena.take(R, hid)
            self.conv(att, w["y"], y2, grid=self.grid, residual=y)
            arena.give(att)
            arena.give(y)
            h = arena.take(R, w["ffn1"].c_out)
            self.conv(y2, w["ffn1"], h, grid=self.grid, act=w["act"])
            arena.give(y2)
            out = arena.take(R, hid)
            self.conv(h, w["ffn2"], out, grid=self.grid, gate=abc + 4 * 2 * hid, gate_ld=self.mod_ld, gate_rows=L, residual=x)
            arena.give(h)
            arena.give(x)
            x = out
        self.final = arena.pin(x)
        self.yt = torch.empty(packed.out_proj.c_out, R, dtype=torch.float32, device=device)
        self.scratch_bytes = arena.bytes

    @property
    def launches(self) -> int:
        return len(self.ops) + 3 + (2 if self.packed.bank.hidden else 0)

    def run_core(self, mod: Tensor | None) -> Tensor:
        r"""tokens in ``self.tok`` -> channel-major fp32 output ``self.yt`` (C_o, B L)."""
        lib, pk = self.lib, self.packed
        s = _lib.stream_ptr(self.device)
        if pk.bank.hidden:
            pk.bank.run(lib, mod, self.hid_buf, self.abc, s)
        self.replay()
        oc, f = pk.out_proj, self.final
        n, h, w = self.grid
        _lib.check(lib.azb_conv2d_bf16(f.data_ptr(), n, h, w, oc.c_in, f.stride(0), oc.w.data_ptr(), oc.c_out,
                                       oc.c_out_rows, 1, oc.k_per_tap, 1, _lib.ptr(oc.bias), 0, None, 0, 0, None, 0,
                                       self.yt.data_ptr(), 0, 1, None, 1, s), "azb_conv2d_bf16")
        return self.yt


def _prepare(model, device, mod: Tensor | None):
    cache = model._native
    packed = cache.get("packed")
    if packed is None or packed.fingerprint != fingerprint(model) or packed.device != device:
        ok = cache.get("structure_ok")
        cache.clear()
        cache["structure_ok"] = ok
        packed = cache["packed"] = Packed(model, device)
    rows = 1
    if packed.bank.hidden:
        if mod is None:
            raise ValueError("this network needs a modulation vector `mod`")
        mod = mod.to(device=device, dtype=torch.float32).reshape(-1, packed.bank.features).contiguous()
        rows = mod.shape[0]
    return cache, packed, mod, rows


def _plan(model, cache, packed, key, batch, tokens, rows, pos_fn, device) -> Plan:
    plan = cache.get(key)
    if plan is None:
        plans = [k for k in cache if isinstance(k, tuple)]
        while len(plans) >= _MAX_PLANS:
            del cache[plans.pop(0)]
        plan = cache[key] = Plan(model, packed, batch, tokens, rows, pos_fn(), device)
    note_use(model, packed, plan)
    return plan


def forward(model, x: Tensor, mod: Tensor | None, pos) -> Tensor:
    r"""``DiT.forward``: (B, L, C_i) -> (B, L, C_o), same dtype.  ``pos``: ``"arange"`` (sequence indices) or a
    (L, P) tensor of positions shared by the batch (its plan is keyed on the tensor's address and version and keeps
    the tensor alive: positional embedding and rotary tables are evaluated once per plan)."""
    device = x.device
    B, L, cin = x.shape
    with torch.cuda.device(device):
        cache, packed, mod, rows = _prepare(model, device, mod)
        if isinstance(pos, str):
            key, pos_fn = ("tokens", B, L, rows), (lambda: torch.arange(L, dtype=torch.float32, device=device)[:, None])
        else:
            key, pos_fn = ("tokens", B, L, rows, pos.data_ptr(), pos._version, tuple(pos.shape)), (lambda: pos)
        plan = _plan(model, cache, packed, key, B, L, rows, pos_fn, device)
        if not isinstance(pos, str):
            plan.pos_ref = pos
        plan.tok[:, :cin].copy_(x.reshape(B * L, cin))
        yt = plan.run_core(mod)
        out = yt.t().reshape(B, L, -1).to(x.dtype)
    return out


def forward_image(model, x: Tensor, mod: Tensor | None, cond: Tensor | None) -> Tensor:
    r"""``ViT.forward``: (B, C_i, H, W) -> (B, C_o, H, W), same dtype."""
    device = x.device
    if cond is not None:
        # patchify(x) ++ patchify(cond) per token (nn/vit.py:97-100) == patchify of the channel concatenation: the
        # token's channel index is (z p + a) q + b with the pixel channel z slowest
        x = torch.cat((x, cond.to(x.dtype)), dim=1)
    B, c, H, W = x.shape
    p, q = model.patch.patch_shape
    hp, wp = H // p, W // q
    with torch.cuda.device(device):
        cache, packed, mod, rows = _prepare(model, device, mod)

        def grid_positions():
            ii = torch.arange(hp, dtype=torch.float32, device=device)
            jj = torch.arange(wp, dtype=torch.float32, device=device)
            return torch.cartesian_prod(ii, jj).reshape(-1, 2)

        plan = _plan(model, cache, packed, ("image", B, hp, wp, rows), B, hp * wp, rows, grid_positions, device)
        xin = x.to(torch.float32).contiguous()
        s = _lib.stream_ptr(device)
        _lib.check(plan.lib.azb_patchify_f32(xin.data_ptr(), plan.tok.data_ptr(), B, c, hp, wp, p, q, packed.k_pad, s),
                   "azb_patchify_f32")
        yt = plan.run_core(mod)
        co = packed.out_proj.c_out // (p * q)
        out = torch.empty(B, co, H, W, dtype=torch.float32, device=device)
        _lib.check(plan.lib.azb_unpatchify_f32(yt.data_ptr(), out.data_ptr(), B, co, hp, wp, p, q, s), "azb_unpatchify_f32")
    return out.to(x.dtype)
