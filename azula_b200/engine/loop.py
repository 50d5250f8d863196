r"""The fused sampling loop: backbone forward + ONE transition kernel per stage, graph-captured.

Replaces the Python hot loop of ``azula/sample.py:151-157`` (and the multi-step loops ``:510-537`` ff.) for every
sampler of ``azula/sample.py`` over preconditioned denoisers, optionally wrapped in classifier-free guidance
(``azula/guidance/cfg.py``).  Per stage the device executes ``F = backbone(x_in, time_in)`` followed by
``azb_step_ex_f32`` (which also emits the next pre-scaled ``x_in``) and ``azb_advance``; the stage index, Philox
state and time input live in device memory, so the very same captured graph serves every stage and the loop issues
no per-step host arithmetic and no ATen dispatch.

Lifetime and staleness (what a captured graph bakes in): raw device pointers of the backbone's parameters or of
their packed kernel-layout copies, of the launch plans' activation arenas, of the coefficient table.  A
:class:`FusedLoop` therefore (i) holds strong references to the packed weights and plans its capture used, so that
cache eviction or ``module.to(...)`` on the model cannot free memory the graph still reads, and (ii) is keyed on
a fingerprint ``(data_ptr, _version)`` of every parameter and buffer of the denoiser and on the schedule's
attributes: after ``load_state_dict`` / an optimiser step / a mutated schedule the sampler builds a fresh loop
instead of replaying the old weights.
"""

from __future__ import annotations

import ctypes
import math
import torch
import torch.nn as nn

from torch import Tensor

from .. import _lib
from ..nn.utils import get_module_dtype
from . import plan as _plan
from . import table as _table

_F_DTYPES = (torch.float32, torch.bfloat16, torch.float16)


def default_generator(device: torch.device | None = None) -> torch.Generator:
    torch.cuda.init()  # the generator tuple is empty until CUDA is initialised
    device = torch.device("cuda") if device is None else torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[index]


def _guided(den) -> bool:
    return bool(getattr(den, "guided_inner", False))


def supports(sampler, x: Tensor, kwargs: dict | None = None) -> bool:
    r"""Whether :class:`FusedLoop` can run this (sampler, input) pair."""
    from ..denoise import Preconditioned

    from . import native_enabled

    den = sampler.denoiser
    if _guided(den):
        if not den.fusable() or kwargs is None or not isinstance(kwargs.get("positive"), dict):
            return False
        if not isinstance(kwargs.get("negative", {}), dict) or x.ndim < 1:
            return False
        den = den.denoiser
    return (
        x.is_cuda
        and native_enabled()
        and x.dtype == torch.float32
        and x.numel() > 0
        and not x.requires_grad
        and isinstance(den, Preconditioned)
        and den.fusable()
        and get_module_dtype(den.backbone) in (None, *_F_DTYPES)
    )


def _freeze(value):
    if torch.is_tensor(value):
        return ("tensor", tuple(value.shape), value.dtype, value.device)
    if isinstance(value, dict):
        return ("dict", tuple(sorted((k, _freeze(v)) for k, v in value.items())))
    try:
        hash(value)
        return value
    except TypeError:
        return ("id", id(value))


def _tensors_of(obj) -> tuple:
    if isinstance(obj, nn.Module):
        return tuple((t.data_ptr(), t._version) for t in (*obj.parameters(), *obj.buffers()))
    return ()


def _describe(obj) -> tuple:
    r"""Hashable description of an object's public scalar attributes plus the (address, version) of the tensors it
    owns: changes whenever something a coefficient table or a captured graph froze may have changed."""
    items = []
    for k, v in sorted(vars(obj).items()):
        if k.startswith("_") or isinstance(v, nn.Module):
            continue
        if torch.is_tensor(v):
            items.append((k, v.data_ptr(), v._version))
        else:
            items.append((k, _freeze(v)))
    return (type(obj), tuple(items), _tensors_of(obj))


def signature(sampler, x: Tensor, kwargs: dict):
    r"""Cache key of a :class:`FusedLoop`: everything baked into its table, buffers and graph."""
    den = sampler.denoiser
    inner = _table.inner_denoiser(den)
    return (
        type(sampler), tuple(x.shape), x.device, sampler.start, sampler.stop, sampler.steps, sampler._signature(),
        sampler.dtype, sampler.device, sampler.shard, sampler.graph, sampler.unroll, type(den), _describe(inner),
        _describe(inner.schedule), get_module_dtype(inner.backbone),
        tuple(getattr(m, "precision", None) for m in inner.backbone.modules() if hasattr(m, "precision")),
        # the guidance strength lives in device memory: its VALUE is not part of the key
        tuple(sorted((k, ("scalar",) if (k == "guidance" and _guided(den) and not torch.is_tensor(v)) else _freeze(v))
                     for k, v in kwargs.items())),
    )


def _clone(v):
    if torch.is_tensor(v):
        return v.clone()
    if isinstance(v, dict):
        return {k: _clone(u) for k, u in v.items()}
    return v


def _refresh(static, new) -> None:
    if torch.is_tensor(static):
        static.copy_(new)
    elif isinstance(static, dict):
        for k in static:
            _refresh(static[k], new[k])


class FusedLoop:
    r"""State of one (sampler, input signature): coefficient table, static buffers, graph."""

    def __init__(self, sampler, x: Tensor, kwargs: dict, graph: bool | None, unroll: int | None) -> None:
        self.sampler = sampler
        self.outer = sampler.denoiser
        self.denoiser = _table.inner_denoiser(sampler.denoiser)  # strong references: ids in no key can be recycled
        self.schedule = self.denoiser.schedule
        self.device = x.device
        self.shape = tuple(x.shape)
        self.table = _table.build(sampler, x.device)
        if self.table is None:
            return
        tab = self.table
        self.stages = tab.steps

        self.select = getattr(self.denoiser, "output_select", lambda: None)()
        self.in_dtype = get_module_dtype(self.denoiser.backbone) or torch.float32

        # ---- classifier-free guidance: both branches in ONE forward of 2B samples when their keyword sets agree,
        # else two forwards of B; either way the combine happens inside the transition kernel
        self.guided = _guided(self.outer)
        self.kwargs = _clone(kwargs)
        self.batched = False
        self.guidance = None
        if self.guided:
            self.guidance = torch.ones((), dtype=torch.float32, device=self.device)
            pos, neg = self.kwargs["positive"], self.kwargs.get("negative", {})
            batchable = lambda v: torch.is_tensor(v) and v.ndim >= 1 and v.shape[0] == self.shape[0]  # noqa: E731
            wants = getattr(self.outer, "batched", None)
            same = set(pos) == set(neg) and all(
                (batchable(pos[k]) and batchable(neg[k]) and pos[k].shape == neg[k].shape and pos[k].dtype == neg[k].dtype)
                or (not torch.is_tensor(pos[k]) and not torch.is_tensor(neg[k]) and _freeze(pos[k]) == _freeze(neg[k]))
                for k in pos
            )
            self.batched = bool(same and wants is not False)
            if self.batched:
                self.pair_kwargs = {k: (torch.cat((pos[k], neg[k])) if torch.is_tensor(pos[k]) else pos[k]) for k in pos}
        copies = 2 if self.batched else 1
        self.shared = {k: v for k, v in self.kwargs.items() if k not in ("positive", "negative", "guidance")} if self.guided else self.kwargs
        if self.batched:  # shared per-sample tensors accompany both halves of the 2B batch
            self.shared = {k: (torch.cat((v, v)) if (torch.is_tensor(v) and v.ndim >= 1 and v.shape[0] == self.shape[0]) else v)
                           for k, v in self.shared.items()}

        # ---- static device state
        self.x = torch.empty_like(x, memory_format=torch.contiguous_format)
        self.x_alt = torch.empty_like(self.x) if tab.alt else None
        self.hist = torch.empty((tab.slots, *self.shape), dtype=torch.float32, device=self.device) if tab.slots else None
        self.x_in = torch.empty((copies * self.shape[0], *self.shape[1:]) if copies > 1 else self.shape,
                                dtype=self.in_dtype, device=self.device)
        self.step_idx = torch.zeros((), dtype=torch.int32, device=self.device)
        self.philox = torch.zeros(2, dtype=torch.int64, device=self.device)  # {offset, seed}
        self.time_in = tab.time[0].clone()

        self.rng_threads, self.offset_inc, self.rng_first = sampler._rng_layout(x.numel())

        if unroll is None:
            unroll_stages = self.stages if (x.numel() <= (1 << 16) and self.stages <= 4096) else tab.per_step
        else:
            unroll_stages = math.gcd(sampler.steps, max(1, int(unroll))) * tab.per_step
        self.unroll = max(1, math.gcd(self.stages, unroll_stages))

        self.pinned: list = []  # (model, packed weights, plan) triples the capture used: kept alive with the graph
        self.graph = None
        self.graph_error: Exception | None = None
        if graph is None or graph:
            try:
                self._capture()
            except Exception as e:  # capture is best effort; the eager fused loop stays exact
                torch.cuda.synchronize(self.device)
                self.graph, self.graph_error = None, e
                if graph:
                    raise

    # ------------------------------------------------------------------ one stage on the stream
    def _forward(self) -> tuple[Tensor, Tensor | None]:
        den = self.denoiser
        if not self.guided:
            return den.call_backbone(self.x_in, self.time_in, **self.kwargs), None
        if self.batched:
            return den.call_backbone(self.x_in, self.time_in, **self.pair_kwargs, **self.shared), None
        pos = den.call_backbone(self.x_in, self.time_in, **self.kwargs["positive"], **self.shared)
        neg = den.call_backbone(self.x_in, self.time_in, **self.kwargs.get("negative", {}), **self.shared)
        return pos, neg

    def _step(self) -> None:
        tab = self.table
        out, neg = self._forward()
        if out.dtype not in _F_DTYPES:
            out = out.to(torch.float32)
        out = out.contiguous()
        if neg is not None:
            neg = neg.to(out.dtype).contiguous()

        numel = self.x.numel()
        copies = 2 if self.batched else 1
        if out.numel() % copies or (self.select is None and out.numel() != copies * numel):
            raise RuntimeError(f"backbone output has {out.numel()} elements, expected {copies * numel}")
        if self.select is None:
            n_per, batch, stride = numel, 1, numel
        else:  # the mean is predicted by the first `select` channels of dim 1
            batch = self.shape[0]
            n_per = numel // batch
            stride = out.numel() // (copies * batch)
        f_neg = None
        if self.guided:
            f_neg = neg.data_ptr() if neg is not None else out.data_ptr() + batch * stride * out.element_size()
        self._last_out = (out, neg)  # eager mode: keep F alive until the kernel ran

        d = _lib.AzbStep()
        d.src[0], d.src[1] = self.x.data_ptr(), _lib.ptr(self.x_alt)
        d.dst[0], d.dst[1] = self.x.data_ptr(), _lib.ptr(self.x_alt)
        d.f, d.f_neg, d.guidance = out.data_ptr(), f_neg, _lib.ptr(self.guidance)
        d.eps, d.x_in_next, d.hist = None, self.x_in.data_ptr(), _lib.ptr(self.hist)
        d.table, d.step_idx, d.philox_state = tab.coef.data_ptr(), self.step_idx.data_ptr(), self.philox.data_ptr()
        d.f_batch_stride, d.n_per_sample, d.batch, d.hist_stride = stride, n_per, batch, numel
        d.offset_host, d.offset_inc, d.rng_threads, d.rng_elem_offset = 0, self.offset_inc, self.rng_threads, self.rng_first
        d.seed = 0
        d.f_dtype, d.in_dtype = _lib.DTYPE_CODE[out.dtype], _lib.DTYPE_CODE[self.in_dtype]
        d.row_floats, d.x_in_copies = _lib.ROW_COLS, copies
        d.noise_hint = -1 if tab.noiseless else 0
        lib = _lib.lib()
        stream = _lib.stream_ptr(self.device)
        _lib.check(lib.azb_step_ex_f32(ctypes.byref(d), stream), "azb_step_ex_f32")
        row = tab.time[0]
        _lib.check(
            lib.azb_advance(
                self.step_idx.data_ptr(), None, 0, tab.time.data_ptr(), self.time_in.data_ptr(), row.element_size(),
                max(1, row.numel()), self.stages, stream,
            ),
            "azb_advance",
        )

    def _reset(self, x: Tensor, kwargs: dict, seed: int, offset: int) -> None:
        self.x.copy_(x)
        x_in = (x * self.table.c_in0).to(self.in_dtype)  # (c_in * x_t).to(dtype), azula/denoise.py:317
        if self.batched:
            self.x_in[: self.shape[0]].copy_(x_in)
            self.x_in[self.shape[0] :].copy_(x_in)
        else:
            self.x_in.copy_(x_in.reshape(self.x_in.shape))
        self.step_idx.zero_()
        # seeds are 64-bit patterns; store them as the int64 with the same bits
        self.philox.copy_(torch.tensor([offset, seed - (1 << 64) if seed >= (1 << 63) else seed], dtype=torch.int64))
        self.time_in.copy_(self.table.time[0])
        _refresh(self.kwargs, kwargs)
        if self.guided:
            g = kwargs.get("guidance", 1.0)
            self.guidance.copy_(g.to(torch.float32).reshape(()) if torch.is_tensor(g) else torch.tensor(float(g)))
            if self.batched:
                pos, neg = kwargs["positive"], kwargs.get("negative", {})
                for k, v in self.pair_kwargs.items():
                    if torch.is_tensor(v):
                        v.copy_(torch.cat((pos[k], neg[k])))
                for k, v in self.shared.items():
                    if torch.is_tensor(v) and v is not self.kwargs.get(k):
                        v.copy_(torch.cat((kwargs[k], kwargs[k])))

    def _capture(self) -> None:
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with _plan.track_use() as used:
            with torch.cuda.stream(side):
                self.x.fill_(0.5)  # finite dummy state for the warm-up pass (lazy init of libraries)
                self.x_in.fill_(0.5)
                if self.x_alt is not None:
                    self.x_alt.fill_(0.5)
                if self.hist is not None:
                    self.hist.zero_()
                self.step_idx.zero_()
                self._step()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.step_idx.zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(self.unroll):
                    self._step()
        self.pinned = used
        self.graph = graph

    def valid(self) -> bool:
        r"""Whether the packed weights the graph reads still describe the live parameters (belt and braces: the
        sampler's key already contains the parameter fingerprint)."""
        return all(packed.fingerprint == _plan.fingerprint(model) for model, packed, _ in self.pinned)

    # ------------------------------------------------------------------------------ full loop
    @torch.no_grad()
    def run(self, x: Tensor, kwargs: dict, progress) -> Tensor:
        gen = default_generator(self.device)
        seed, offset = gen.initial_seed(), gen.get_offset()
        self._reset(x, kwargs, seed, offset)

        if self.graph is not None:
            for _ in progress(range(self.stages // self.unroll)):
                self.graph.replay()
        else:
            per = self.table.per_step
            for _ in progress(range(self.stages // per)):
                for _ in range(per):
                    self._step()

        # the reference draws one randn_like per noisy stage, whatever its amplitude (azula/sample.py:214,259)
        gen.set_offset(offset + self.table.draws * self.offset_inc)
        return self.x.clone().reshape(x.shape)
