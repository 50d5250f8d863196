r"""The REFERENCE-NUMERICS executor of the ADM U-Net: fp32 activations, TF32 tensor-core contractions.

The reference runs the backbone in the module's dtype -- fp32 -- and, under PyTorch's default flags, cuDNN executes
its convolutions on the tensor cores in TF32 (10-bit operand mantissa, fp32 accumulation); everything else (GroupNorm,
SiLU, residuals, the embedding MLPs) is fp32 arithmetic (``azula/plugins/adm/__init__.py:118-123``,
``azula/denoise.py:314-320``).  ``model.precision = "tf32"`` (:func:`azula_b200.engine.set_precision`) makes the native
path do exactly that instead of the bf16 fast path:

    activations   fp32 NHWC in HBM (pixel stride ``ld`` >= C: decoder concatenations are channel slices, no ``cat``)
    conv / linear ``azb_conv_tf32``: TMA tiles of fp32, ``tcgen05.mma.kind::tf32``, fp32 accumulators in tensor
                  memory, fp32 bias / residual in the epilogue, nothing rounded on the way out
    GroupNorm     ``azb_gn_stats_f32`` (mean / rstd per image and group) + ``azb_gn_apply_f32`` (affine, scale / shift,
                  SiLU, 2x up / down-sampling), fp32 in and out
    attention     q, k, v leave the projection's epilogue as fp16 (10-bit mantissa, as TF32) for ``azb_attention_f16``
                  (fp32 softmax and accumulation, fp32 result)
    embeddings    the fp32 kernels of the bf16 plan (``azb_timestep_features_f32``, ``azb_linear_f32``)

It is a flat launch list over statically allocated buffers like the bf16 plan, so it is captured into the sampler's
CUDA graph the same way; it trades the bf16 plan's fusions (GroupNorm on halo tiles, phase-decomposed upsampling, CTA
pairs) for the reference's numerics: ~3x slower than bf16, still faster than eager.
"""

from __future__ import annotations

import math
import torch

from torch import Tensor

from .. import _lib
from . import ops
from .plan import fingerprint, note_use

_MAX_PLANS = 2


class PackedTF32:
    r"""fp32 kernel-layout copy of a model's parameters on one device."""

    def __init__(self, model, device: torch.device) -> None:
        lay = model.layout
        if not lay.scale_shift:
            raise NotImplementedError("native ADM path needs use_scale_shift_norm=True (all ADM cards use it)")
        p = {k: v.detach().to(device) for k, v in model.named_parameters()}
        f32 = lambda key: p[key].to(torch.float32).contiguous()  # noqa: E731
        self.fingerprint = fingerprint(model)
        self.device = device
        self.c_in_pad = -(-lay.in_channels // 4) * 4
        self.time0 = (f32("time_embed.0.weight"), f32("time_embed.0.bias"))
        self.time2 = (f32("time_embed.2.weight"), f32("time_embed.2.bias"))
        self.labels = f32("label_emb.weight") if lay.num_classes is not None else None
        pk = lambda w, b, **kw: ops.pack_conv_f32(p[w], p[b], **kw)  # noqa: E731
        self.unit: dict[str, dict] = {}
        emb_w, emb_b, offset = [], [], 0
        for u in lay.units():
            if u.kind == "stem":
                self.unit[u.path] = {"conv": pk(u.path + ".weight", u.path + ".bias", c_in_pad=self.c_in_pad)}
            elif u.kind == "res":
                if (u.cin // 32) % 4 or (u.cout // 32) % 4:
                    raise NotImplementedError("the TF32 mode needs GroupNorm groups of a multiple of 4 channels")
                self.unit[u.path] = {
                    "gn1": (f32(u.path + ".in_layers.0.weight"), f32(u.path + ".in_layers.0.bias")),
                    "conv1": pk(u.path + ".in_layers.2.weight", u.path + ".in_layers.2.bias"),
                    "gn2": (f32(u.path + ".out_layers.0.weight"), f32(u.path + ".out_layers.0.bias")),
                    "conv2": pk(u.path + ".out_layers.3.weight", u.path + ".out_layers.3.bias"),
                    "skip": (pk(u.path + ".skip_connection.weight", u.path + ".skip_connection.bias") if u.cin != u.cout else None),
                    "emb_offset": offset,
                }
                emb_w.append(f32(u.path + ".emb_layers.1.weight"))
                emb_b.append(f32(u.path + ".emb_layers.1.bias"))
                offset += 2 * u.cout
            else:
                if u.cin // u.heads not in (32, 64, 128, 256):
                    raise NotImplementedError(f"attention head width {u.cin // u.heads} not in (32, 64, 128, 256)")
                self.unit[u.path] = {
                    "gn": (f32(u.path + ".norm.weight"), f32(u.path + ".norm.bias")),
                    "qkv": pk(u.path + ".qkv.weight", u.path + ".qkv.bias"),
                    "proj": pk(u.path + ".proj_out.weight", u.path + ".proj_out.bias"),
                }
        self.emb_total = offset
        self.emb_w, self.emb_b = torch.cat(emb_w).contiguous(), torch.cat(emb_b).contiguous()
        self.out_gn = (f32("out.0.weight"), f32("out.0.bias"))
        self.out_conv = pk("out.2.weight", "out.2.bias")


class _Arena:
    def __init__(self, device) -> None:
        self.device, self.idle, self.owner, self.bytes = device, [], {}, 0

    def take(self, *shape: int, dtype=torch.float32) -> Tensor:
        need = math.prod(shape) * (2 if dtype == torch.float16 else 4)
        fit = [t for t in self.idle if t.numel() >= need]
        if fit:
            flat = min(fit, key=Tensor.numel)
            self.idle = [t for t in self.idle if t is not flat]
        else:
            flat = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.bytes += need
        view = flat[:need].view(dtype).view(shape)
        self.owner[id(view)] = flat
        return view

    def give(self, view: Tensor) -> None:
        self.idle.append(self.owner.pop(id(view)))


class PlanTF32:
    r"""The launch list of one (batch, height, width, embedding rows) signature."""

    def __init__(self, model, packed: PackedTF32, n: int, h: int, w: int, rows: int, device: torch.device) -> None:
        self.lay = lay = model.layout
        self.packed = packed
        self.n, self.h, self.w, self.rows, self.device = n, h, w, rows, device
        self.lib = _lib.lib()
        self.ops: list[tuple] = []
        self.meta: list[tuple] = []
        self.keep: list = []
        arena = self.arena = _Arena(device)
        f32 = dict(dtype=torch.float32, device=device)
        D = lay.embed_dim
        self.feat = torch.empty(rows, lay.model_channels, **f32)
        self.emb1 = torch.empty(rows, D, **f32)
        self.emb = torch.empty(rows, D, **f32)
        self.emb_all = torch.empty(rows, packed.emb_total, **f32)
        self.x_nhwc = torch.empty(n, h, w, packed.c_in_pad, **f32)
        self.stats_ws = ops.gn_stats_workspace_f32(n, ops.GN_GROUPS, device)

        L = len(lay.encoder)
        sizes, hh, ww = [], h, w
        for block in lay.encoder:
            if block[0].kind == "res" and block[0].resample == 2:
                if hh % 2 or ww % 2:
                    raise ValueError(f"spatial size {(h, w)} is not divisible by the network's downsampling")
                hh, ww = hh // 2, ww // 2
            sizes.append((hh, ww, block[-1].cout))
        self.cat = []
        carried = lay.middle[-1].cout
        for j, block in enumerate(lay.decoder):
            sh, sw, sc = sizes[L - 1 - j]
            buf = arena.take(n, sh, sw, carried + sc)
            arena.owner.pop(id(buf))  # concat buffers live for the whole forward
            self.cat.append((buf, carried))
            carried = block[-1].cout
        self.final = arena.take(n, h, w, lay.final_channels)
        arena.owner.pop(id(self.final))

        cur = None
        for i, block in enumerate(lay.encoder):
            buf, left = self.cat[L - 1 - i]
            cur = self._block(block, cur, buf[..., left:])
        cur = self._block(lay.middle, cur, self.cat[0][0][..., : self.cat[0][1]])
        for j, block in enumerate(lay.decoder):
            dest = self.cat[j + 1][0][..., : self.cat[j + 1][1]] if j + 1 < L else self.final
            cur = self._block(block, self.cat[j][0], dest)
        st = self._stats(self.final)
        self._apply(self.final, self.final, st, packed.out_gn, None, True, 0)
        self.scratch_bytes = arena.bytes

    # ------------------------------------------------------------------------ emitters
    def _emit(self, kind: str, flops: float, nbytes: float, fn, *args, desc: str = "") -> None:
        self.ops.append((fn, args))
        self.meta.append((kind, flops, nbytes, desc))

    def _conv(self, x: Tensor, pc, out: Tensor, residual: Tensor | None = None, out_f16: bool = False) -> None:
        n, h, w = x.shape[:3]
        self.keep += [x, out, pc.w, pc.bias] + ([residual] if residual is not None else [])
        flops = 2.0 * n * h * w * pc.c_out * pc.taps * x.shape[-1]
        nbytes = 4.0 * (n * h * w * (x.shape[-1] + pc.c_out * (2 if residual is not None else 1)) + pc.c_out * pc.taps * x.shape[-1])
        self._emit("conv3x3" if pc.taps == 9 else "conv1x1", flops, nbytes, self.lib.azb_conv_tf32, x.data_ptr(), n, h, w, pc.c_in,
                   ops._ld(x), pc.w.data_ptr(), pc.c_out, pc.c_out_rows, pc.taps, pc.k_per_tap, 1, _lib.ptr(pc.bias), 0,
                   _lib.ptr(residual), 0 if residual is None else ops._ld(residual), out.data_ptr(), ops._ld(out), 0, int(out_f16),
                   desc=f"{n}x{h}x{w} {x.shape[-1]}->{pc.c_out}" + (" +res" if residual is not None else ""))

    def _stats(self, x: Tensor) -> Tensor:
        n, h, w, c = x.shape
        st = torch.empty(n, ops.GN_GROUPS, 2, dtype=torch.float32, device=self.device)
        self.keep += [x, st]
        self._emit("gn_stats", 0.0, 4.0 * n * h * w * c, self.lib.azb_gn_stats_f32, x.data_ptr(), ops._ld(x), n, h * w, c,
                   ops.GN_GROUPS, ops.GN_EPS, st.data_ptr(), self.stats_ws.data_ptr(), self.stats_ws.numel(), desc=f"{n}x{h}x{w}x{c}")
        return st

    def _apply(self, x: Tensor, out: Tensor, stats: Tensor | None, affine, emb_offset: int | None, silu: bool, mode: int) -> None:
        n, h, w, c = x.shape
        gamma, beta = affine if affine is not None else (None, None)
        ss_ptr, ss_stride = None, 0
        if emb_offset is not None:
            ss_ptr = self.emb_all.data_ptr() + 4 * emb_offset
            ss_stride = self.packed.emb_total if (self.rows == n and n > 1) else 0
        self.keep += [x, out] + ([gamma, beta] if gamma is not None else [])
        px_out = n * h * w * (4 if mode == 1 else 1) // (4 if mode == 2 else 1)
        self._emit("gn_apply", 0.0, 4.0 * c * (n * h * w + px_out), self.lib.azb_gn_apply_f32, x.data_ptr(), ops._ld(x),
                   out.data_ptr(), ops._ld(out), n, h, w, c, ops.GN_GROUPS, _lib.ptr(stats), _lib.ptr(gamma), _lib.ptr(beta),
                   ss_ptr, ss_stride, int(silu), mode, desc=f"{n}x{h}x{w}x{c} mode{mode}")

    # ------------------------------------------------------------------------ structure
    def _block(self, block, x: Tensor | None, dest: Tensor) -> Tensor:
        arena, temp = self.arena, None
        for k, u in enumerate(block):
            last = k + 1 == len(block)
            if u.kind == "stem":
                self._conv(self.x_nhwc, self.packed.unit[u.path]["conv"], dest)
                return dest
            n, h, w, _ = x.shape
            ho, wo = (2 * h, 2 * w) if u.resample == 1 else (h // 2, w // 2) if u.resample == 2 else (h, w)
            out = dest if last else arena.take(n, ho, wo, u.cout)
            (self._res if u.kind == "res" else self._attn)(u, x, out)
            if temp is not None:
                arena.give(temp)
            temp = None if last else out
            x = out
        return dest

    def _res(self, u, x: Tensor, out: Tensor) -> None:
        r"""``ResBlock._forward`` (``_src/unet.py:227-247``) with ``use_scale_shift_norm``."""
        w_, arena = self.packed.unit[u.path], self.arena
        n, ho, wo, _ = out.shape
        st1 = self._stats(x)
        h1 = arena.take(n, ho, wo, u.cin)
        self._apply(x, h1, st1, w_["gn1"], None, True, u.resample)  # SiLU(GN(x)), then up / down
        xr = x
        if u.resample:
            xr = arena.take(n, ho, wo, u.cin)
            self._apply(x, xr, None, None, None, False, u.resample)  # x_upd on the raw input
        h2 = arena.take(n, ho, wo, u.cout)
        self._conv(h1, w_["conv1"], h2)
        arena.give(h1)
        st2 = self._stats(h2)
        self._apply(h2, h2, st2, w_["gn2"], w_["emb_offset"], True, 0)  # SiLU(GN(h) (1 + scale) + shift), in place
        sk = xr
        if w_["skip"] is not None:
            sk = arena.take(n, ho, wo, u.cout)
            self._conv(xr, w_["skip"], sk)
        self._conv(h2, w_["conv2"], out, residual=sk)
        arena.give(h2)
        if sk is not xr:
            arena.give(sk)
        if xr is not x:
            arena.give(xr)

    def _attn(self, u, x: Tensor, out: Tensor) -> None:
        r"""``AttentionBlock._forward`` (``_src/unet.py:290-296``)."""
        w_, arena = self.packed.unit[u.path], self.arena
        n, h, w, c = x.shape
        st = self._stats(x)
        y = arena.take(n, h, w, c)
        self._apply(x, y, st, w_["gn"], None, False, 0)
        qkv = arena.take(n, h, w, 3 * c, dtype=torch.float16)
        self._conv(y, w_["qkv"], qkv, out_f16=True)
        arena.give(y)
        a = arena.take(n, h, w, c)
        d = c // u.heads
        hs, kd, vd = (d, c, 2 * c) if self.lay.new_attention_order else (3 * d, d, 2 * d)
        self.keep += [qkv, a]
        t = h * w
        self._emit("attention", 4.0 * n * u.heads * t * t * d, 2.0 * n * t * 3 * c + 4.0 * n * t * c, self.lib.azb_attention_f16,
                   qkv.data_ptr(), 3 * c, a.data_ptr(), c, n, t, u.heads, d, hs, kd, vd, desc=f"{n}x{u.heads}x{t}x{d}")
        arena.give(qkv)
        self._conv(a, w_["proj"], out, residual=x)
        arena.give(a)

    # ------------------------------------------------------------------------ running
    @property
    def launches(self) -> int:
        return len(self.ops) + 7

    def profile(self, detail: list | None = None) -> dict[str, dict]:
        s = _lib.stream_ptr(self.device)
        events = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.ops) + 1)]
        events[0].record()
        for i, (fn, args) in enumerate(self.ops):
            _lib.check(fn(*args, s), fn.__name__)
            events[i + 1].record()
        torch.cuda.synchronize(self.device)
        table: dict[str, dict] = {}
        for i, (kind, flops, nbytes, desc) in enumerate(self.meta):
            row = table.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            ms = events[i].elapsed_time(events[i + 1])
            row["launches"] += 1
            row["ms"] += ms
            row["flops"] += flops
            row["bytes"] += nbytes
            if detail is not None:
                detail.append((kind, desc, ms, flops, nbytes))
        return table

    def run(self, x: Tensor, timesteps: Tensor, y: Tensor | None, out: Tensor) -> Tensor:
        lib, pk, lay = self.lib, self.packed, self.lay
        s = _lib.stream_ptr(self.device)
        chk = _lib.check
        n, c, h, w = x.shape
        chk(lib.azb_nchw_to_nhwc_f32(x.data_ptr(), self.x_nhwc.data_ptr(), n, c, h, w, pk.c_in_pad, s), "azb_nchw_to_nhwc_f32")
        rows, D = self.rows, lay.embed_dim
        chk(lib.azb_timestep_features_f32(timesteps.data_ptr(), _lib.DTYPE_CODE[timesteps.dtype], rows, lay.model_channels,
                                          10000.0, self.feat.data_ptr(), s), "azb_timestep_features_f32")
        chk(lib.azb_linear_f32(self.feat.data_ptr(), pk.time0[0].data_ptr(), pk.time0[1].data_ptr(), self.emb1.data_ptr(), rows, D,
                               lay.model_channels, 0, s), "azb_linear_f32")
        chk(lib.azb_linear_f32(self.emb1.data_ptr(), pk.time2[0].data_ptr(), pk.time2[1].data_ptr(), self.emb.data_ptr(), rows, D, D,
                               1, s), "azb_linear_f32")
        if y is not None:
            chk(lib.azb_add_rows_f32(self.emb.data_ptr(), pk.labels.data_ptr(), y.data_ptr(), rows, D, s), "azb_add_rows_f32")
        chk(lib.azb_linear_f32(self.emb.data_ptr(), pk.emb_w.data_ptr(), pk.emb_b.data_ptr(), self.emb_all.data_ptr(), rows,
                               pk.emb_total, D, 1, s), "azb_linear_f32")
        for fn, args in self.ops:
            rc = fn(*args, s)
            if rc:
                chk(rc, fn.__name__)
        oc, f = pk.out_conv, self.final
        chk(lib.azb_conv_tf32(f.data_ptr(), n, h, w, oc.c_in, ops._ld(f), oc.w.data_ptr(), oc.c_out, oc.c_out_rows, oc.taps,
                              oc.k_per_tap, 1, _lib.ptr(oc.bias), 0, None, 0, out.data_ptr(), 0, 1, 0, s), "azb_conv_tf32")
        return out


def forward(model, x: Tensor, timesteps: Tensor, y: Tensor | None = None) -> Tensor:
    r"""``UNetModel.forward`` in the reference-numerics mode: (N, C, H, W) -> (N, C', H, W), same dtype."""
    device = x.device
    lay = model.layout
    if x.ndim != 4 or x.shape[1] != lay.in_channels:
        raise ValueError(f"expected an input of shape (N, {lay.in_channels}, H, W), got {tuple(x.shape)}")
    n, _, h, w = x.shape
    with torch.cuda.device(device):
        cache = model._native
        packed = cache.get("packed_tf32")
        if packed is None or packed.fingerprint != fingerprint(model) or packed.device != device:
            for k in [k for k in cache if k == "packed_tf32" or (isinstance(k, tuple) and k[0] == "tf32")]:
                del cache[k]
            packed = cache["packed_tf32"] = PackedTF32(model, device)
        timesteps = timesteps.reshape(-1)
        if timesteps.numel() not in (1, n):
            raise ValueError(f"timesteps must have 1 or {n} elements, got {timesteps.numel()}")
        if y is not None and timesteps.numel() != n:
            timesteps = timesteps.expand(n)
        tdtype = torch.float32 if timesteps.is_floating_point() else torch.int64
        timesteps = timesteps.to(device=device, dtype=tdtype).contiguous()
        rows = timesteps.numel()
        if y is not None:
            y = y.reshape(-1).to(device=device, dtype=torch.int64).contiguous()
            if y.numel() != n:
                raise ValueError(f"y must have {n} elements")
        key = ("tf32", n, h, w, rows)
        plan = cache.get(key)
        if plan is None:
            plans = [k for k in cache if isinstance(k, tuple) and k[0] == "tf32"]
            while len(plans) >= _MAX_PLANS:
                del cache[plans.pop(0)]
            plan = cache[key] = PlanTF32(model, packed, n, h, w, rows, device)
        note_use(model, packed, plan)
        xin = x.to(torch.float32).contiguous()
        out = torch.empty((n, lay.out_channels, h, w), dtype=torch.float32, device=device)
        plan.run(xin, timesteps, y, out)
    return out.to(x.dtype)
