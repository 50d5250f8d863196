r"""Native sm_100a executor of the ADM U-Net layout (:mod:`azula_b200.plugins.adm.unet`).

Replaces ``UNetModel.forward`` of the reference (``azula/plugins/adm/_src/unet.py:605-634``
and the blocks it calls, ``:227-247,290-296,328-345``; ~10^3 ATen/cuDNN launches per forward,
fp32 NCHW, every GroupNorm / SiLU / scale-shift / residual / resample a separate HBM round
trip) by a *plan*: a flat list of C-ABI launches (``include/azb.h``) over statically allocated
NHWC bf16 buffers, built once per input signature.

Data layout in HBM
    * activations: NHWC bf16, pixel stride ``ld`` >= C, so producers write straight into channel
      slices of the decoder's concatenation buffers -- ``torch.cat`` (``_src/unet.py:631``)
      disappears: encoder block *i* stores its output in the right half of the buffer that
      decoder block *L-1-i* normalises and convolves, the previous decoder block in the left half;
    * weights: bf16 ``[C_out][tap][C_in]`` (K-major, zero padded to the 64-element TMA box),
      packed once per parameter version; biases, GroupNorm affine and embeddings fp32;
    * the 42 per-block ``emb_layers`` linears are concatenated into one fp32 matrix, evaluated
      by one launch per forward; blocks read their ``[scale | shift]`` slice in place.

Per residual unit: stats -> apply(+SiLU, +resample) -> conv3x3 (tcgen05) -> stats ->
apply(+scale/shift +SiLU, in place) -> conv3x3 (tcgen05, + fused skip residual); attention unit:
stats -> apply -> qkv GEMM -> flash attention -> proj GEMM (+ fused residual).  No allocation,
synchronisation or host read happens while a plan runs, so it is CUDA-graph capturable.
"""

from __future__ import annotations

import math
import torch

from ctypes import byref, c_int64
from torch import Tensor

from .. import _lib
from . import ops
from .plan import note_use

_MAX_PLANS = 2
FUSE_NORM = True  # GroupNorm + SiLU ride on the halo tiles of the convolution that consumes them (A/B switch)
UP_PHASES = True  # upsampling blocks: conv1(up(.)) as four 2 x 2 convolutions of the half-resolution tensor
RESAMPLE_SHORTCUTS = True  # up blocks: residual read through the upsampling; down blocks: both branches in one pass
FUSE_NORM_SKIP = True  # ... also when the ResBlock's 1x1 skip operand is part of the GEMM


class Packed:
    r"""Kernel-layout copy of a model's parameters on one device."""

    def __init__(self, model, device: torch.device) -> None:
        lay = model.layout
        if not lay.scale_shift:
            raise NotImplementedError("native ADM path needs use_scale_shift_norm=True (all ADM cards use it)")
        p = {k: v.detach().to(device) for k, v in model.named_parameters()}
        f32 = lambda key: p[key].to(torch.float32).contiguous()  # noqa: E731
        self.fingerprint = fingerprint(model)
        self.k_pad = -(-9 * lay.in_channels // 64) * 64
        self.time0 = (f32("time_embed.0.weight"), f32("time_embed.0.bias"))
        self.time2 = (f32("time_embed.2.weight"), f32("time_embed.2.bias"))
        self.labels = f32("label_emb.weight") if lay.num_classes is not None else None
        self.unit: dict[str, dict] = {}
        emb_w, emb_b, offset = [], [], 0
        for u in lay.units():
            if u.kind == "stem":
                w = ops.pack_first_conv(p[u.path + ".weight"].float(), p[u.path + ".bias"], self.k_pad)
                self.unit[u.path] = {"conv": w}
            elif u.kind == "res":
                conv2 = ops.pack_conv(p[u.path + ".out_layers.3.weight"], p[u.path + ".out_layers.3.bias"])
                skip = (ops.pack_conv(p[u.path + ".skip_connection.weight"], p[u.path + ".skip_connection.bias"])
                        if u.cin != u.cout else None)
                if skip is not None and skip.taps == 1:
                    conv2, skip = ops.pack_conv_skip(conv2, skip), None  # one GEMM: K = [9 taps | skip channels]
                self.unit[u.path] = {
                    "gn1": (f32(u.path + ".in_layers.0.weight"), f32(u.path + ".in_layers.0.bias")),
                    "conv1": ops.pack_conv(p[u.path + ".in_layers.2.weight"], p[u.path + ".in_layers.2.bias"]),
                    # upsampling block: conv1(up(.)) as four 2 x 2 convolutions of the half-resolution tensor
                    "conv1_up": (ops.pack_conv_up(p[u.path + ".in_layers.2.weight"], p[u.path + ".in_layers.2.bias"])
                                 if u.resample == 1 and UP_PHASES else None),
                    "gn2": (f32(u.path + ".out_layers.0.weight"), f32(u.path + ".out_layers.0.bias")),
                    "conv2": conv2,
                    "skip": skip,
                    "emb_offset": offset,
                }
                emb_w.append(f32(u.path + ".emb_layers.1.weight"))
                emb_b.append(f32(u.path + ".emb_layers.1.bias"))
                offset += 2 * u.cout
            else:
                if u.cin // u.heads not in (16, 32, 64, 128, 192, 256):
                    raise NotImplementedError(f"attention head width {u.cin // u.heads} not in (16, 32, 64, 128, 192, 256)")
                self.unit[u.path] = {
                    "gn": (f32(u.path + ".norm.weight"), f32(u.path + ".norm.bias")),
                    "qkv": ops.pack_conv(p[u.path + ".qkv.weight"], p[u.path + ".qkv.bias"]),
                    "proj": ops.pack_conv(p[u.path + ".proj_out.weight"], p[u.path + ".proj_out.bias"]),
                }
        self.emb_total = offset
        self.emb_w, self.emb_b = torch.cat(emb_w).contiguous(), torch.cat(emb_b).contiguous()
        self.out_gn = (f32("out.0.weight"), f32("out.0.bias"))
        self.out_conv = ops.pack_conv(p["out.2.weight"], p["out.2.bias"])


def fingerprint(model) -> tuple:
    return tuple((q.data_ptr(), q._version) for q in model.parameters())


class _Arena:
    r"""Plan-build-time allocator of bf16 scratch with reuse (one stream => sequential lifetimes)."""

    def __init__(self, device) -> None:
        self.device = device
        self.idle: list[Tensor] = []
        self.owner: dict[int, Tensor] = {}
        self.bytes = 0

    def take(self, *shape: int) -> Tensor:
        need = math.prod(shape)
        fit = [t for t in self.idle if t.numel() >= need]
        if fit:
            flat = min(fit, key=Tensor.numel)
            self.idle = [t for t in self.idle if t is not flat]
        else:
            flat = torch.empty(need, dtype=torch.bfloat16, device=self.device)
            self.bytes += 2 * need
        view = flat[:need].view(shape)
        self.owner[id(view)] = flat
        return view

    def give(self, view: Tensor) -> None:
        self.idle.append(self.owner.pop(id(view)))


class Plan:
    r"""The launch list of one (batch, height, width, embedding rows) signature."""

    def __init__(self, model, packed: Packed, n: int, h: int, w: int, rows: int, device: torch.device) -> None:
        self.lay = lay = model.layout
        self.packed = packed
        self.n, self.h, self.w, self.rows, self.device = n, h, w, rows, device
        self.lib = _lib.lib()
        self.ops: list[tuple] = []
        self.meta: list[tuple] = []
        # every width of the network is a multiple of model_channels: with model_channels % 256 == 0 all GroupNorm
        # groups (C / 32 channels) are multiples of 8 channels and the epilogues emit one sum per 8-channel block
        self.stat_gran = 8 if lay.model_channels % 256 == 0 else 1
        self.acc_of: dict[tuple, tuple] = {}  # activation view -> (exact GroupNorm accumulators of its producer, channels)
        self.cat_parts: dict[tuple, list] = {}   # concatenation buffer -> [left view, right view]
        self.keep: list[Tensor] = []  # everything the launch list points into
        arena = self.arena = _Arena(device)
        f32 = dict(dtype=torch.float32, device=device)

        # ---- embedding path (fp32): features -> MLP -> (+ label rows) -> all emb_layers at once
        D = lay.embed_dim
        self.feat = torch.empty(rows, lay.model_channels, **f32)
        self.emb1 = torch.empty(rows, D, **f32)
        self.emb = torch.empty(rows, D, **f32)
        self.emb_all = torch.empty(rows, packed.emb_total, **f32)

        # ---- GroupNorm: one pool of exact accumulators (int64 {sum, sumsq} x {hi, lo} per image and channel block),
        # one slice per convolution output, cleared by ONE memset at the start of a forward
        widths = [u.cout for u in lay.units() if u.kind != "attn"] + [u.cout for u in lay.units() if u.kind == "res"] + [
            u.cin for u in lay.units() if u.kind == "attn"]
        self.acc_pool = torch.zeros(max(4 * n * sum(-(-c // self.stat_gran) for c in widths), 4), dtype=torch.int64,
                                    device=device)
        self.acc_used = 0
        self._emit("zero", 0.0, 8.0 * self.acc_pool.numel(), self.lib.azb_zero_bytes, self.acc_pool.data_ptr(),
                   8 * self.acc_pool.numel())
        self.splitk_ws = ops.splitk_workspace(device)
        # ---- GroupNorm fallback workspace (feature maps too small for the fused sums)
        self.gn_partial_need = 0
        self.counters = torch.zeros(max(n, 1), dtype=torch.int32, device=device)

        # ---- concatenation buffers of the decoder
        L = len(lay.encoder)
        assert len(lay.decoder) == L
        # spatial size of each encoder output
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
            assert block[0].cin == carried + sc, (block[0], carried, sc)
            buf = arena.take(n, sh, sw, carried + sc)
            self.cat.append((buf, carried))
            self.cat_parts[self._key(buf)] = [buf[..., :carried], buf[..., carried:]]
            carried = block[-1].cout
        for buf, _ in self.cat:  # concat buffers live for the whole forward
            arena.owner.pop(id(buf))
        self.patches = arena.take(n, h, w, packed.k_pad)
        arena.owner.pop(id(self.patches))
        self.final = arena.take(n, h, w, lay.final_channels)
        arena.owner.pop(id(self.final))

        def skip_slot(i: int) -> Tensor:
            buf, left = self.cat[L - 1 - i]
            return buf[..., left:]

        # ---- encoder
        cur = None
        for i, block in enumerate(lay.encoder):
            cur = self._block(block, cur, skip_slot(i))
        # ---- middle
        cur = self._block(lay.middle, cur, self.cat[0][0][..., : self.cat[0][1]])
        # ---- decoder
        for j, block in enumerate(lay.decoder):
            if j + 1 < L:
                nxt, left = self.cat[j + 1]
                dest = nxt[..., :left]
            else:
                dest = self.final
            cur = self._block(block, self.cat[j][0], dest)
        # ---- head: GroupNorm + SiLU in place; the last convolution is bound per call (output tensor)
        st = self._stats(self.final)
        self.out_coef = None
        # (the output tensor is bound per call: any valid pointer serves the query)
        if self._fusable(st, self.final, packed.out_conv, self.final, nchw_f32=True):
            self.out_coef = self._coef(self.final, st, packed.out_gn, None, True)
        else:
            self._apply(self.final, self.final, st, packed.out_gn, None, True, 0)

        self.gn_partial = torch.empty(max(self.gn_partial_need, 1), **f32)
        self._bind_partial()
        self.scratch_bytes = arena.bytes

    # ------------------------------------------------------------------------ plan building
    def _emit(self, kind: str, flops: float, nbytes: float, fn, *args, desc: str = "") -> None:
        r"""Queues one launch; ``flops`` / ``nbytes`` are its ALGORITHMIC work."""
        self.ops.append((fn, args))
        self.meta.append((kind, flops, nbytes, desc))

    @staticmethod
    def _key(t: Tensor) -> tuple:
        return (t.data_ptr(), tuple(t.shape))

    def _acc_for(self, out: Tensor, c_out: int) -> Tensor | None:
        r"""A zeroed-per-forward slice of the accumulator pool for the GroupNorm sums of ``out``, or ``None``
        when a 32-row slab of the convolution's M tiles would straddle two images (tiny feature maps)."""
        n, h, w = out.shape[:3]
        self.acc_of.pop(self._key(out), None)  # the buffer may be a recycled one
        if not ops.colsum_rows(n, h, w)[1]:
            return None
        count = n * (c_out // self.stat_gran) * 4
        if self.acc_used + count > self.acc_pool.numel():
            raise RuntimeError("GroupNorm accumulator pool exhausted")
        acc = self.acc_pool[self.acc_used : self.acc_used + count].view(n, c_out // self.stat_gran, 4)
        self.acc_used += count
        self.acc_of[self._key(out)] = (acc, c_out)
        return acc

    def _conv(self, x: Tensor, pc, out: Tensor, residual: Tensor | None = None, stats: bool = False,
              x2: Tensor | None = None, in_coef: Tensor | None = None, in_silu: bool = True, res_up: bool = False,
              in_up: bool = False) -> None:
        r"""Queues a convolution (``azb_conv_bf16``); with ``stats`` its epilogue also adds the exact sums from
        which the GroupNorm(s) consuming ``out`` derive their statistics (no read pass over ``out``, no reduction
        launch).  With ``x2`` the ResBlock's 1x1 skip connection is part of the same GEMM.  With ``in_coef``
        (:meth:`_coef`) the GroupNorm + SiLU of the INPUT is applied on the fly to the halo tiles."""
        n, h, w, _ = x.shape
        if in_up:  # x at half resolution, read through a nearest 2x upsampling
            h, w = 2 * h, 2 * w
        acc = self._acc_for(out, pc.c_out) if stats else None
        if not stats:
            self.acc_of.pop(self._key(out), None)
        d = ops.conv_desc(x, pc, out, grid=(n, h, w), residual=residual, x2=x2, gn_acc=acc, gran=self.stat_gran,
                          workspace=self.splitk_ws, in_coef=in_coef, in_silu=in_silu, res_up=res_up, in_up=in_up)
        self.keep += [d, x, out, pc.w] + [t for t in (residual, pc.bias, x2, in_coef) if t is not None]
        taps = 9 if x2 is not None else pc.taps
        k_extra = pc.c_in2 if x2 is not None else 0
        phased = getattr(pc, "taps", 9) == 16
        if phased:    # phase-decomposed upsampling convolution: 4 taps per output pixel EXECUTED; the algorithmic
            taps = 9  # work of the reference layer (what the rates are quoted on) stays 9 taps
        flops = 2.0 * n * h * w * pc.c_out * (taps * pc.c_in + k_extra)
        nbytes = 2.0 * (n * h * w * (pc.c_in + k_extra + pc.c_out * (2 if residual is not None else 1))
                        + pc.c_out * (taps * pc.c_in + k_extra))
        desc = f"{n}x{h}x{w} {pc.c_in}->{pc.c_out}" + (" gn+" if in_coef is not None else "") + (" up+" if in_up else "") + (
            (" +res(up)" if res_up else " +res") if residual is not None else "") + (
            f" +skip1x1({pc.c_in2})" if x2 is not None else "") + (" +stats" if acc is not None else "")
        if phased:
            desc += " (4 x 2x2 phases)"
        self._emit("conv3x3" if taps == 9 else "conv1x1", flops, nbytes, self.lib.azb_conv_bf16, byref(d), desc=desc)

    def _conv_skip(self, x: Tensor, x2: Tensor, pc: ops.PackedConvSkip, out: Tensor, in_coef: Tensor | None = None) -> None:
        self._conv(x, pc, out, stats=True, x2=x2, in_coef=in_coef)

    def _pool_dual(self, x: Tensor, out: Tensor, out_raw: Tensor, stats, affine) -> None:
        r"""Queues ``azb_gn_pool_acc_bf16``: both branches of a downsampling ResBlock from one read of ``x``."""
        n, h, w, c = x.shape
        gamma, beta = affine
        (a, ca), (b, cb) = stats[1][0], (stats[1][1] if len(stats[1]) > 1 else (None, 0))
        for t in (out, out_raw):
            self.acc_of.pop(self._key(t), None)
        self.keep += [x, out, out_raw, a, gamma, beta] + ([b] if b is not None else [])
        self._emit(
            "gn_apply", 0.0, 2.0 * c * (n * h * w + 2 * n * h * w // 4), self.lib.azb_gn_pool_acc_bf16, x.data_ptr(), ops._ld(x),
            out.data_ptr(), ops._ld(out), out_raw.data_ptr(), ops._ld(out_raw), n, h, w, c, ops.GN_GROUPS, a.data_ptr(), ca,
            _lib.ptr(b), cb, self.stat_gran, ops.GN_EPS, gamma.data_ptr(), beta.data_ptr(), None, 0, 1,
            desc=f"{n}x{h}x{w}x{c} mode2 dual",
        )

    def _fusable(self, stats, x: Tensor, pc, out: Tensor, residual: Tensor | None = None, x2: Tensor | None = None,
                 nchw_f32: bool = False, in_up: bool = False) -> bool:
        r"""Whether the GroupNorm (+ SiLU) in front of this convolution can ride on its halo tiles: statistics from
        exact accumulators and a halo kernel for the shape (``azb_conv_choice``)."""
        if not FUSE_NORM or stats is None or stats[0] != "acc" or (x2 is not None and not FUSE_NORM_SKIP):
            return False
        if self.stat_gran != 8 and not nchw_f32:
            # the convolution will also emit per-channel sums for the next GroupNorm (widths whose groups are not
            # multiples of 8 channels, e.g. the 192-channel imagenet_64x64 card): no halo kernel does that
            return False
        n, h, w, _ = x.shape
        grid = (n, 2 * h, 2 * w) if in_up else (n, h, w)
        d = ops.conv_desc(x, pc, out, grid=grid, residual=residual, x2=x2, gran=self.stat_gran, nchw_f32=nchw_f32,
                          in_up=in_up, in_coef=self._probe_coef(n, pc.c_in))  # (the transform itself is what asks for halo tiles)
        try:
            return bool(ops.conv_choice(d).halo)
        except _lib.AzbError:  # e.g. the phase-decomposed form on a map below 16 x 8 half-resolution pixels
            return False

    def _probe_coef(self, n: int, c: int) -> Tensor:
        r"""A correctly shaped coefficient buffer for ``azb_conv_choice`` queries (nothing is launched)."""
        t = getattr(self, "_probe", None)
        if t is None or t.numel() < n * c * 2:
            t = self._probe = torch.empty(n * c * 2, dtype=torch.float32, device=self.device)
        return t[: n * c * 2]

    def _coef(self, x: Tensor, stats, affine, emb_offset: int | None, silu: bool) -> Tensor:
        r"""Queues ``azb_gn_coef_f32`` for the GroupNorm over ``x``: fp32 (N, C, 2) transform coefficients."""
        n, h, w, c = x.shape
        gamma, beta = affine
        ss_ptr, ss_stride = None, 0
        if emb_offset is not None:
            ss_ptr = self.emb_all.data_ptr() + 4 * emb_offset
            ss_stride = self.packed.emb_total if (self.rows == n and n > 1) else 0
        (a, ca), (b, cb) = stats[1][0], (stats[1][1] if len(stats[1]) > 1 else (None, 0))
        coef = torch.empty(n, c, 2, dtype=torch.float32, device=self.device)
        self.keep += [coef, a, gamma, beta] + ([b] if b is not None else [])
        self._emit(
            "gn_coef", 0.0, 8.0 * n * c, self.lib.azb_gn_coef_f32, n, h, w, c, ops.GN_GROUPS, a.data_ptr(), ca,
            _lib.ptr(b), cb, self.stat_gran, ops.GN_EPS, gamma.data_ptr(), beta.data_ptr(), ss_ptr, ss_stride, int(silu),
            coef.data_ptr(), desc=f"{n}x{c}",
        )
        return coef

    def _stats(self, x: Tensor):
        r"""Where the GroupNorm over ``x`` finds its statistics: ``("acc", parts)`` -- the exact accumulators of
        the one or two producers of ``x`` -- or ``("stats", tensor)`` from the stand-alone reduction pass."""
        n, c = x.shape[0], x.shape[-1]
        hw = math.prod(x.shape[1:-1])
        parts = self.cat_parts.get(self._key(x), [x])
        sources = [self.acc_of.get(self._key(t)) for t in parts]
        if all(src is not None for src in sources) and (c // ops.GN_GROUPS) % self.stat_gran == 0:
            return ("acc", sources)
        want = c_int64(0)
        _lib.check(self.lib.azb_gn_stats_workspace(n, hw, c, ops.GN_GROUPS, byref(want)), "azb_gn_stats_workspace")
        self.gn_partial_need = max(self.gn_partial_need, want.value)
        stats = torch.empty(n, ops.GN_GROUPS, 2, dtype=torch.float32, device=self.device)
        self.keep += [x, stats]
        self._emit(
            "gn_stats", 0.0, 2.0 * n * hw * c,
            self.lib.azb_gn_stats_bf16, x.data_ptr(), ops._ld(x), n, hw, c, ops.GN_GROUPS, ops.GN_EPS, "partial",
            stats.data_ptr(), self.counters.data_ptr(),
        )
        return ("stats", stats)

    def _bind_partial(self) -> None:
        ptr = self.gn_partial.data_ptr()
        self.ops = [(fn, tuple(ptr if isinstance(a, str) else a for a in args)) for fn, args in self.ops]

    def _apply(self, x: Tensor, out: Tensor, stats, affine, emb_offset: int | None, silu: bool, mode: int) -> None:
        n, h, w, c = x.shape
        gamma, beta = affine if affine is not None else (None, None)
        self.acc_of.pop(self._key(out), None)
        ss_ptr, ss_stride = None, 0
        if emb_offset is not None:
            ss_ptr = self.emb_all.data_ptr() + 4 * emb_offset
            ss_stride = self.packed.emb_total if (self.rows == n and n > 1) else 0
        self.keep += [x, out] + ([gamma, beta] if gamma is not None else [])
        px_out = n * h * w * (4 if mode == 1 else 1) // (4 if mode == 2 else 1)
        desc = f"{n}x{h}x{w}x{c} mode{mode}"
        if stats is not None and stats[0] == "acc":
            (a, ca), (b, cb) = stats[1][0], (stats[1][1] if len(stats[1]) > 1 else (None, 0))
            self.keep += [a] + ([b] if b is not None else [])
            self._emit(
                "gn_apply", 0.0, 2.0 * c * (n * h * w + px_out),
                self.lib.azb_gn_apply_acc_bf16, x.data_ptr(), ops._ld(x), out.data_ptr(), ops._ld(out), n, h, w, c,
                ops.GN_GROUPS, a.data_ptr(), ca, _lib.ptr(b), cb, self.stat_gran, ops.GN_EPS, _lib.ptr(gamma),
                _lib.ptr(beta), ss_ptr, ss_stride, int(silu), mode, desc=desc,
            )
            return
        st = stats[1] if stats is not None else None
        if st is not None:
            self.keep.append(st)
        self._emit(
            "gn_apply", 0.0, 2.0 * c * (n * h * w + px_out),
            self.lib.azb_gn_apply_bf16, x.data_ptr(), ops._ld(x), out.data_ptr(), ops._ld(out), n, h, w, c,
            ops.GN_GROUPS, _lib.ptr(st), _lib.ptr(gamma), _lib.ptr(beta), ss_ptr, ss_stride, None, 0, int(silu), mode,
            desc=desc,
        )

    def _block(self, block, x: Tensor | None, dest: Tensor) -> Tensor:
        r"""Emits the units of one block; the last unit writes ``dest``."""
        arena = self.arena
        temp = None
        for k, u in enumerate(block):
            last = k + 1 == len(block)
            if u.kind == "stem":
                self._conv(self.patches, self.packed.unit[u.path]["conv"], dest, stats=True)
                return dest
            n, h, w, _ = x.shape
            ho, wo = (2 * h, 2 * w) if u.resample == 1 else (h // 2, w // 2) if u.resample == 2 else (h, w)
            out = dest if last else arena.take(n, ho, wo, u.cout)
            if u.kind == "res":
                self._res(u, x, out)
            else:
                self._attn(u, x, out)
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
        h2 = arena.take(n, ho, wo, u.cout)
        res_up = False
        if not u.resample and self._fusable(st1, x, w_["conv1"], h2):
            # SiLU(GN(x)) is applied to the halo tiles of conv1: no normalised copy of x in HBM
            self._conv(x, w_["conv1"], h2, stats=True, in_coef=self._coef(x, st1, w_["gn1"], None, True))
            xr = x
        else:
            h1, h1_done = arena.take(n, ho, wo, u.cin), False
            identity_skip = w_["skip"] is None and not isinstance(w_["conv2"], ops.PackedConvSkip)
            if u.resample == 1 and identity_skip and RESAMPLE_SHORTCUTS:
                # upsampling block: x_upd = up(x) is never stored, conv2's epilogue reads x through the 2x upsampling;
                # nor is up(SiLU(GN(x))) when conv1 can load its halo tiles through an upsampling tensor map
                xr, res_up = x, True
                if w_.get("conv1_up") is not None and self._fusable(st1, x, w_["conv1_up"], h2, in_up=True):
                    # ... as four 2 x 2 convolutions of x (one per output phase): 2.25 x fewer FLOPs
                    self._conv(x, w_["conv1_up"], h2, stats=True, in_coef=self._coef(x, st1, w_["gn1"], None, True), in_up=True)
                    h1_done = True
                elif self._fusable(st1, x, w_["conv1"], h2, in_up=True):
                    self._conv(x, w_["conv1"], h2, stats=True, in_coef=self._coef(x, st1, w_["gn1"], None, True), in_up=True)
                    h1_done = True
                else:
                    self._apply(x, h1, st1, w_["gn1"], None, True, 1)
            elif u.resample == 2 and st1[0] == "acc" and RESAMPLE_SHORTCUTS:
                # downsampling block: ONE pass over x writes pool(SiLU(GN(x))) and x_upd = pool(x)
                xr = arena.take(n, ho, wo, u.cin)
                self._pool_dual(x, h1, xr, st1, w_["gn1"])
            else:
                self._apply(x, h1, st1, w_["gn1"], None, True, u.resample)  # SiLU(GN(x)) then up / down
                if u.resample:
                    xr = arena.take(n, ho, wo, u.cin)
                    self._apply(x, xr, None, None, None, False, u.resample)  # x_upd on the raw input
                else:
                    xr = x
            if not h1_done:
                self._conv(h1, w_["conv1"], h2, stats=True)
            arena.give(h1)
        st2 = self._stats(h2)
        skip_fused = isinstance(w_["conv2"], ops.PackedConvSkip)
        sk = xr
        if not skip_fused and w_["skip"] is not None:
            sk = arena.take(n, ho, wo, u.cout)
            self._conv(xr, w_["skip"], sk)
        if self._fusable(st2, h2, w_["conv2"], out, residual=None if (skip_fused or res_up) else sk,
                         x2=xr if skip_fused else None):
            coef2 = self._coef(h2, st2, w_["gn2"], w_["emb_offset"], True)  # SiLU(GN(h) (1 + scale) + shift) on the fly
        else:
            coef2 = None
            self._apply(h2, h2, st2, w_["gn2"], w_["emb_offset"], True, 0)  # the same, in place
        if skip_fused:
            self._conv_skip(h2, xr, w_["conv2"], out, in_coef=coef2)  # skip_connection(x) + h, the 1x1 folded into the GEMM's K
        else:
            self._conv(h2, w_["conv2"], out, residual=sk, stats=True, in_coef=coef2, res_up=res_up)  # x + h
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
        qkv = arena.take(n, h, w, 3 * c)
        self._conv(y, w_["qkv"], qkv)
        arena.give(y)
        a = arena.take(n, h, w, c)
        d = c // u.heads
        hs, kd, vd = (d, c, 2 * c) if self.lay.new_attention_order else (3 * d, d, 2 * d)
        self.keep += [qkv, a]
        t = h * w
        self._emit("attention", 4.0 * n * u.heads * t * t * d, 2.0 * n * t * 4 * c, self.lib.azb_attention_bf16, qkv.data_ptr(), 3 * c, a.data_ptr(), c, n, h * w, u.heads, d, hs, kd, vd)
        arena.give(qkv)
        self._conv(a, w_["proj"], out, residual=x, stats=True)
        arena.give(a)

    # ------------------------------------------------------------------------------ running
    @property
    def launches(self) -> int:
        r"""Kernels launched by one :meth:`run` (without the label lookup; the accumulator memset is not a kernel)."""
        return sum(1 for kind, *_ in self.meta if kind != "zero") + 6

    def profile(self, detail: list | None = None) -> dict[str, dict]:
        r"""Times every queued launch with CUDA events on the current stream (buffers keep whatever
        the last :meth:`run` left in them); returns per kernel kind: launches, ms, flops, bytes.
        ``detail`` (a list) receives one (kind, description, ms, flops, bytes) tuple per launch."""
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
            row["launches"] += 1
            ms = events[i].elapsed_time(events[i + 1])
            row["ms"] += ms
            if detail is not None:
                detail.append((kind, desc, ms, flops, nbytes))
            row["flops"] += flops
            row["bytes"] += nbytes
        return table

    def run(self, x: Tensor, timesteps: Tensor, y: Tensor | None, out: Tensor) -> Tensor:
        lib, pk, lay = self.lib, self.packed, self.lay
        s = _lib.stream_ptr(self.device)
        chk = _lib.check
        n, c, h, w = x.shape
        chk(lib.azb_im2col3x3_f32(x.data_ptr(), self.patches.data_ptr(), n, c, h, w, pk.k_pad, s), "azb_im2col3x3_f32")
        rows, D = self.rows, lay.embed_dim
        chk(lib.azb_timestep_features_f32(timesteps.data_ptr(), _lib.DTYPE_CODE[timesteps.dtype], rows,
                                          lay.model_channels, 10000.0, self.feat.data_ptr(), s), "azb_timestep_features_f32")
        chk(lib.azb_linear_f32(self.feat.data_ptr(), pk.time0[0].data_ptr(), pk.time0[1].data_ptr(),
                               self.emb1.data_ptr(), rows, D, lay.model_channels, 0, s), "azb_linear_f32")
        chk(lib.azb_linear_f32(self.emb1.data_ptr(), pk.time2[0].data_ptr(), pk.time2[1].data_ptr(),
                               self.emb.data_ptr(), rows, D, D, 1, s), "azb_linear_f32")
        if y is not None:
            chk(lib.azb_add_rows_f32(self.emb.data_ptr(), pk.labels.data_ptr(), y.data_ptr(), rows, D, s), "azb_add_rows_f32")
        chk(lib.azb_linear_f32(self.emb.data_ptr(), pk.emb_w.data_ptr(), pk.emb_b.data_ptr(), self.emb_all.data_ptr(),
                               rows, pk.emb_total, D, 1, s), "azb_linear_f32")
        for fn, args in self.ops:
            rc = fn(*args, s)
            if rc:
                chk(rc, fn.__name__)
        d = ops.conv_desc(self.final, pk.out_conv, out, nchw_f32=True, in_coef=self.out_coef, in_silu=True)
        chk(lib.azb_conv_bf16(byref(d), s), "azb_conv_bf16")
        return out


def forward(model, x: Tensor, timesteps: Tensor, y: Tensor | None = None) -> Tensor:
    r"""``UNetModel.forward`` on a CUDA device: (N, C, H, W) any float dtype -> (N, C', H, W) same dtype."""
    device = x.device
    lay = model.layout
    if x.ndim != 4 or x.shape[1] != lay.in_channels:
        raise ValueError(f"expected an input of shape (N, {lay.in_channels}, H, W), got {tuple(x.shape)}")
    n, _, h, w = x.shape
    with torch.cuda.device(device):
        cache = model._native
        packed = cache.get("packed")
        if packed is None or packed.fingerprint != fingerprint(model) or packed.emb_w.device != device:
            for k in [k for k in cache if k == "packed" or (isinstance(k, tuple) and k[0] != "tf32")]:
                del cache[k]
            packed = cache["packed"] = Packed(model, device)

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

        key = (n, h, w, rows)
        plan = cache.get(key)
        if plan is None:
            plans = [k for k in cache if isinstance(k, tuple) and k[0] != "tf32"]
            while len(plans) >= _MAX_PLANS:
                del cache[plans.pop(0)]
            plan = cache[key] = Plan(model, packed, n, h, w, rows, device)
        note_use(model, packed, plan)

        xin = x.to(torch.float32).contiguous()
        out = torch.empty((n, lay.out_channels, h, w), dtype=torch.float32, device=device)
        plan.run(xin, timesteps, y, out)
    return out.to(x.dtype)
