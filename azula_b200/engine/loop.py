r"""The fused sampling loop: backbone forward + ONE transition kernel per step, graph-captured.

Replaces the Python hot loop of ``azula/sample.py:151-157`` for the samplers whose transition
is affine in :math:`(x_t, F, \varepsilon)` (DDPM, DDIM) over preconditioned denoisers.  Per
step the device executes ``F = backbone(x_in, time_in)`` followed by ``azb_step_f32`` (which
also emits the next pre-scaled ``x_in``) and ``azb_advance``; the step index, Philox offset
and time input live in device memory, so the very same captured graph serves every step and
the loop issues no per-step host arithmetic.
"""

from __future__ import annotations

import math
import torch

from torch import Tensor

from .. import _lib
from ..nn.utils import get_module_dtype
from . import table as _table

_F_DTYPES = (torch.float32, torch.bfloat16, torch.float16)


def default_generator(device: torch.device | None = None) -> torch.Generator:
    torch.cuda.init()  # the generator tuple is empty until CUDA is initialised
    device = torch.device("cuda") if device is None else torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[index]


def supports(sampler, x: Tensor) -> bool:
    r"""Whether :class:`FusedLoop` can run this (sampler, input) pair."""
    from ..denoise import Preconditioned

    from . import native_enabled

    den = sampler.denoiser
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
    try:
        hash(value)
        return value
    except TypeError:
        return ("id", id(value))


class FusedLoop:
    r"""State of one (sampler, input signature): coefficient table, static buffers, graph."""

    def __init__(self, sampler, x: Tensor, kwargs: dict, graph: bool | None, unroll: int | None) -> None:
        self.sampler = sampler
        self.denoiser = sampler.denoiser
        self.device = x.device
        self.shape = tuple(x.shape)
        self.steps = sampler.steps
        self.table = _table.build(sampler, x.device)

        self.select = getattr(self.denoiser, "output_select", lambda: None)()
        self.in_dtype = get_module_dtype(self.denoiser.backbone) or torch.float32

        # static device state
        self.x = torch.empty_like(x, memory_format=torch.contiguous_format)
        self.x_in = torch.empty(self.shape, dtype=self.in_dtype, device=self.device)
        self.step_idx = torch.zeros((), dtype=torch.int32, device=self.device)
        self.philox = torch.zeros(2, dtype=torch.int64, device=self.device)  # {offset, seed}
        self.time_in = self.table.time[0].clone()
        self.kwargs = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kwargs.items()}

        self.rng_threads, self.offset_inc, self.rng_first = sampler._rng_layout(x.numel())

        if unroll is None:
            unroll = self.steps if (x.numel() <= (1 << 16) and self.steps <= 1024) else 1
        self.unroll = max(1, math.gcd(self.steps, int(unroll)))

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

    # ------------------------------------------------------------------ one step on the stream
    def _step(self) -> None:
        den, tab = self.denoiser, self.table
        out = den.call_backbone(self.x_in, self.time_in, **self.kwargs)
        if out.dtype not in _F_DTYPES:
            out = out.to(torch.float32)
        out = out.contiguous()

        numel = self.x.numel()
        if self.select is None:
            if out.numel() != numel:
                raise RuntimeError(f"backbone output has {out.numel()} elements, expected {numel}")
            n_per, batch, stride = numel, 1, numel
        else:  # the mean is predicted by the first `select` channels of dim 1
            batch = self.shape[0]
            n_per = numel // batch
            stride = out.numel() // batch
        lib = _lib.lib()
        stream = _lib.stream_ptr(self.device)
        _lib.check(
            lib.azb_step_f32(
                self.x.data_ptr(), out.data_ptr(), _lib.DTYPE_CODE[out.dtype], stride, None,
                self.x.data_ptr(), self.x_in.data_ptr(), _lib.DTYPE_CODE[self.in_dtype],
                n_per, batch, tab.coef.data_ptr(), self.step_idx.data_ptr(),
                0, self.philox.data_ptr(), 0, self.rng_threads, self.rng_first, stream,
            ),
            "azb_step_f32",
        )
        row = tab.time[0]
        _lib.check(
            lib.azb_advance(
                self.step_idx.data_ptr(), self.philox.data_ptr(), self.offset_inc,
                tab.time.data_ptr(), self.time_in.data_ptr(), row.element_size(), max(1, row.numel()),
                self.steps, stream,
            ),
            "azb_advance",
        )

    def _reset(self, x: Tensor, kwargs: dict, seed: int, offset: int) -> None:
        self.x.copy_(x)
        self.x_in.copy_(x * self.table.c_in0)  # (c_in * x_t).to(dtype), azula/denoise.py:317
        self.step_idx.zero_()
        # seeds are 64-bit patterns; store them as the int64 with the same bits
        self.philox.copy_(torch.tensor([offset, seed - (1 << 64) if seed >= (1 << 63) else seed], dtype=torch.int64))
        self.time_in.copy_(self.table.time[0])
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                self.kwargs[k].copy_(v)

    def _capture(self) -> None:
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.x.fill_(0.5)  # finite dummy state for the warm-up pass (lazy init of libraries)
            self.x_in.fill_(0.5)
            self.step_idx.zero_()
            self._step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(self.unroll):
                self._step()
        self.graph = graph

    # ------------------------------------------------------------------------------ full loop
    @torch.no_grad()
    def run(self, x: Tensor, kwargs: dict, progress) -> Tensor:
        gen = default_generator(self.device)
        seed, offset = gen.initial_seed(), gen.get_offset()
        self._reset(x, kwargs, seed, offset)

        if self.graph is not None:
            for _ in progress(range(self.steps // self.unroll)):
                self.graph.replay()
        else:
            for _ in progress(range(self.steps)):
                self._step()

        # the reference draws one randn_like per step, whatever eta (azula/sample.py:214,259)
        gen.set_offset(offset + self.steps * self.offset_inc)
        return self.x.clone().reshape(x.shape)


def signature(sampler, x: Tensor, kwargs: dict):
    r"""Cache key of a :class:`FusedLoop`: everything baked into its table, buffers and graph."""
    den = sampler.denoiser
    return (
        tuple(x.shape), x.device, sampler.start, sampler.stop, sampler.steps, sampler._eta(),
        sampler.dtype, sampler.device, sampler.shard, den.training, id(den.schedule), id(den.backbone),
        get_module_dtype(den.backbone),
        tuple(sorted((k, _freeze(v)) for k, v in kwargs.items())),
    )
