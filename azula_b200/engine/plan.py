r"""Launch plans: a backbone forward as a flat, pre-bound list of C-ABI calls.

A plan is built once per input signature.  Building allocates every activation buffer (from an
:class:`Arena` that recycles dead scratch), packs nothing, launches nothing; running a plan is a
loop of ``ctypes`` calls with constant arguments on the current stream -- no allocation, no
synchronisation, no host read -- so a plan can be captured into a CUDA graph together with the
transition kernel (:mod:`azula_b200.engine.loop`).
"""

from __future__ import annotations

import contextlib
import math
import torch

from torch import Tensor

from .. import _lib


class Arena:
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

    def pin(self, view: Tensor) -> Tensor:
        r"""Marks a buffer as living for the whole forward (never recycled)."""
        self.owner.pop(id(view))
        return view


class LaunchPlan:
    r"""Queue of (entry point, arguments) pairs plus their ALGORITHMIC work for profiling."""

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self.lib = _lib.lib()
        self.ops: list[tuple] = []
        self.meta: list[tuple] = []
        self.keep: list = []  # everything the launch list points into
        self.arena = Arena(device)

    def _emit(self, kind: str, flops: float, nbytes: float, fn, *args, desc: str = "") -> None:
        r"""Queues one launch; ``flops`` / ``nbytes`` are its algorithmic work."""
        self.ops.append((fn, args))
        self.meta.append((kind, flops, nbytes, desc))

    def replay(self) -> None:
        s = _lib.stream_ptr(self.device)
        for fn, args in self.ops:
            rc = fn(*args, s)
            if rc:
                _lib.check(rc, fn.__name__)

    def profile(self, detail: list | None = None) -> dict[str, dict]:
        r"""Times every queued launch with CUDA events on the current stream (buffers keep whatever
        the last run left in them); returns per kernel kind: launches, ms, flops, bytes.
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

    # ------------------------------------------------------------------ shared emitters
    def conv(self, x: Tensor, pc, out: Tensor, *, grid: tuple[int, int, int] | None = None, stride: int = 1,
             act: int = 0, gate: int | None = None, gate_ld: int = 0, gate_rows: int = 0, residual: Tensor | None = None,
             nchw_f32: bool = False, kind: str | None = None) -> None:
        r"""Queues ``azb_conv2d_bf16``.  ``x`` is an NHWC view, or any row-major (rows, C) view with
        ``grid`` = the (n, h, w) it is presented as.  ``gate`` is a raw device address (fp32)."""
        from . import ops

        n, h, w = grid if grid is not None else x.shape[:3]
        ho, wo = -(-h // stride), -(-w // stride)
        x_ld = x.stride(-2)
        self.keep += [x, out, pc.w] + ([residual] if residual is not None else []) + ([pc.bias] if pc.bias is not None else [])
        flops = 2.0 * n * ho * wo * pc.c_out * pc.taps * pc.c_in
        nbytes = 2.0 * (n * h * w * pc.c_in + pc.c_out * pc.taps * pc.c_in) + n * ho * wo * pc.c_out * (
            4.0 if nchw_f32 else 2.0 * (2 if residual is not None else 1))
        desc = f"{n}x{h}x{w} {pc.c_in}->{pc.c_out}" + (f" s{stride}" if stride > 1 else "") + (
            " +act" if act else "") + (" +gate" if gate else "") + (" +res" if residual is not None else "")
        self._emit(
            kind or ("conv3x3" if pc.taps == 9 else "gemm"), flops, nbytes,
            self.lib.azb_conv2d_bf16, x.data_ptr(), n, h, w, pc.c_in, x_ld, pc.w.data_ptr(), pc.c_out, pc.c_out_rows,
            pc.taps, pc.k_per_tap, stride, _lib.ptr(pc.bias), act, gate, gate_ld, gate_rows, _lib.ptr(residual),
            0 if residual is None else residual.stride(-2), out.data_ptr(), 0 if nchw_f32 else out.stride(-2),
            1 if nchw_f32 else 0, None, 1, desc=desc,
        )
        _ = ops  # (entry points registered on import)

    def rownorm(self, x: Tensor, out: Tensor, kind: int, mod: int | None, mod_ld: int, rows_per_sample: int,
                eps: float = 1e-5) -> None:
        r"""Queues ``azb_rownorm_mod_bf16`` over all rows of ``x`` (last dimension = channels)."""
        c = x.shape[-1]
        rows = math.prod(x.shape[:-1])
        self.keep += [x, out]
        self._emit(
            "rownorm", 0.0, 4.0 * rows * c,
            self.lib.azb_rownorm_mod_bf16, x.data_ptr(), x.stride(-2), out.data_ptr(), out.stride(-2), rows, c, kind, eps,
            mod, mod_ld, max(rows_per_sample, 1), desc=f"{rows}x{c}",
        )


class ModulationBank:
    r"""All Ada-Norm-Zero MLPs of a network (``Linear(D, D) -> SiLU -> Linear(D, 3C)`` per block,
    ``azula/nn/unet.py:65-72`` / ``azula/nn/dit.py:57-64``) evaluated by TWO launches per forward:
    the first linears as one stacked fp32 GEMV, the second ones as one gathered GEMV whose output
    column j reads the hidden slice of its own block.  Blocks whose modulation is a free parameter
    (``mod_features = 0``) contribute constant rows."""

    def __init__(self, blocks: list, device: torch.device) -> None:
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.offset: dict[int, int] = {}  # id(block) -> first column of its [a | b | c] row
        self.width: dict[int, int] = {}
        w0, b0, w1, b1, xoff, const = [], [], [], [], [], []
        col = 0
        learned = [b for b in blocks if not torch.is_tensor(b.ada_zero)]
        free = [b for b in blocks if torch.is_tensor(b.ada_zero)]
        self.features = learned[0].ada_zero[0].in_features if learned else 0
        for j, b in enumerate(learned):
            lin0, lin1 = b.ada_zero[0], b.ada_zero[2]
            assert lin0.in_features == self.features
            w0.append(f32(lin0.weight)), b0.append(f32(lin0.bias))
            w1.append(f32(lin1.weight)), b1.append(f32(lin1.bias))
            xoff.append(torch.full((lin1.out_features,), j * self.features, dtype=torch.int32))
            self.offset[id(b)], self.width[id(b)] = col, lin1.out_features
            col += lin1.out_features
        self.learned_cols = col
        for b in free:
            row = f32(b.ada_zero).reshape(-1)
            const.append(row)
            self.offset[id(b)], self.width[id(b)] = col, row.numel()
            col += row.numel()
        self.cols = col
        self.hidden = len(learned) * self.features
        if learned:
            self.w0, self.b0 = torch.cat(w0).contiguous(), torch.cat(b0).contiguous()
            self.w1, self.b1 = torch.cat(w1).contiguous(), torch.cat(b1).contiguous()
            self.xoff = torch.cat(xoff).to(device)
        self.const = torch.cat(const).contiguous() if const else None

    def buffers(self, rows: int, device: torch.device) -> tuple[Tensor | None, Tensor]:
        r"""(hidden (rows, blocks * D), abc (rows, cols)) for one plan; constant rows pre-filled."""
        hid = torch.empty(rows, max(self.hidden, 1), dtype=torch.float32, device=device) if self.hidden else None
        abc = torch.zeros(rows, max(self.cols, 4), dtype=torch.float32, device=device)
        if self.const is not None:
            abc[:, self.learned_cols : self.cols] = self.const
        return hid, abc

    def run(self, lib, mod: Tensor, hid: Tensor, abc: Tensor, stream: int) -> None:
        r"""mod (rows, D) fp32 contiguous -> abc[:, :learned_cols]."""
        if not self.hidden:
            return
        rows = mod.shape[0]
        _lib.check(lib.azb_linear_f32(mod.data_ptr(), self.w0.data_ptr(), self.b0.data_ptr(), hid.data_ptr(), rows,
                                      self.hidden, self.features, 0, stream), "azb_linear_f32")
        _lib.check(lib.azb_linear_gather_f32(hid.data_ptr(), hid.stride(0), self.xoff.data_ptr(), self.w1.data_ptr(),
                                             self.b1.data_ptr(), abc.data_ptr(), rows, self.learned_cols, self.features,
                                             1, stream), "azb_linear_gather_f32")


def fingerprint(model) -> tuple:
    return tuple((q.data_ptr(), q._version) for q in model.parameters())


# ---- which packed weights / plans a piece of code used (the fused loop pins them for the lifetime of its graph)
_tracker: list | None = None


def note_use(model, packed, plan) -> None:
    r"""Called by the native forwards with the objects whose device memory their launches address."""
    if _tracker is not None and not any(p is plan for _, _, p in _tracker):
        _tracker.append((model, packed, plan))


@contextlib.contextmanager
def track_use():
    r"""Collects (model, packed weights, plan) triples of every native forward executed inside the context."""
    global _tracker
    previous, _tracker = _tracker, []
    try:
        yield _tracker
    finally:
        _tracker = previous
