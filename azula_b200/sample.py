r"""Reverse diffusion samplers (interface of ``azula/sample.py``).

Starting from :math:`x_1 \sim p(X_1)`, a sampler simulates :math:`T` transitions
:math:`x_s \sim q(X_s \mid x_t)` along a time grid from :py:`start` to :py:`stop`.

Execution model (what differs from the reference): for CUDA float32 inputs and a
:class:`azula_b200.denoise.Preconditioned` denoiser, ``sampler(x)`` does not run a Python
loop of ~86 tiny kernels per step (``azula/sample.py:151-157``); it runs the
:class:`azula_b200.engine.loop.FusedLoop` -- backbone forward plus one hand-written sm_100a
transition kernel per step, captured in a CUDA graph that is replayed for every step.
Anything the fused loop cannot express (subclasses overriding :meth:`Sampler.step`, denoiser
wrappers, inputs that require grad, non-float32 state) goes through :meth:`Sampler.step`,
whose CUDA implementation still performs the whole affine update in that one kernel.  CPU
tensors use plain torch arithmetic; there is no CPU kernel and no silent fallback for CUDA
tensors when ``libazb.so`` is missing (the call raises).
"""

from __future__ import annotations

__all__ = [
    "Sampler",
    "DDPMSampler",
    "DDIMSampler",
    "EulerSampler",
    "HeunSampler",
    "ItoSampler",
    "zABSampler",
    "vABSampler",
    "zEABSampler",
    "xEABSampler",
    "REABSampler",
    "PCSampler",
]

import abc
import inspect
import math
import torch

from collections.abc import Iterable, Sequence
from torch import Tensor
from tqdm import tqdm

from . import _lib
from .denoise import Denoiser
from .engine import loop as _loop
from .engine import table as _table
from .engine.table import transition_scalars


def _unchanged(obj, base: type, *names: str) -> bool:
    r"""Whether ``type(obj)`` still uses ``base``'s definitions of ``names`` (a subclass that overrides one of
    them must see its own code run, i.e. take the generic path)."""
    return all(inspect.getattr_static(type(obj), n) is inspect.getattr_static(base, n) for n in names)


class Sampler(abc.ABC):
    r"""Abstract reverse diffusion sampler (``azula/sample.py:54-176``).

    Arguments:
        start: The starting time :math:`t_T`.
        stop: The stopping time :math:`t_0`.
        steps: The number of discretization steps :math:`T` (constant step size).
        silent: Whether to hide the sampling progress bar or not.
        dtype: The time data type.
        device: The time device.
        graph: Engine knob. :py:`None` captures the fused loop in a CUDA graph when possible,
            :py:`True` insists (errors surface), :py:`False` launches the fused step eagerly.
        unroll: Engine knob, number of steps captured per graph (:py:`None` = automatic).
        shard: Engine knob for batch-sharded sampling, :py:`(rank, world)`: this process holds slice
            :py:`rank` of :py:`world` equal slices (along the first dimension) of a global batch. Noise
            is then addressed by GLOBAL element index, so the concatenation of the shards' results
            equals the single-process result on the global batch bit for bit (CUDA tensors only).
    """

    denoiser: Denoiser

    def __init__(
        self,
        start: float = 1.0,
        stop: float = 0.0,
        steps: int = 64,
        silent: bool = False,
        dtype: torch.dtype | None = None,
        device: torch.device | None = None,
        graph: bool | None = None,
        unroll: int | None = None,
        shard: tuple[int, int] | None = None,
    ) -> None:
        self.start = start
        self.stop = stop
        self.steps = steps
        self.silent = silent

        self.dtype = dtype
        self.device = device

        self.graph = graph
        self.unroll = unroll
        self.shard = shard
        self._loops: dict = {}

    def _rng_layout(self, numel: int) -> tuple[int, int, int]:
        r"""(threads, offset increment, first global element) of this process's noise draws."""
        rank, world = self.shard if self.shard is not None else (0, 1)
        threads, inc = _lib.rng_policy(numel * world)
        return threads, inc, rank * numel

    @property
    def timesteps(self) -> Tensor:
        r"""The :math:`T + 1` grid points :math:`t_T, \dots, t_0`."""
        return torch.linspace(self.start, self.stop, self.steps + 1, dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------------------- init
    @torch.no_grad()
    def init(
        self,
        shape: Sequence[int],
        mean: float | Tensor = 0.0,
        var: float | Tensor = 1.0,
        **kwargs,
    ) -> Tensor:
        r"""Draws :math:`x_{t_T} \sim \mathcal{N}(\alpha_{t_T} \mathbb{E}[X], \alpha_{t_T}^2
        \mathbb{V}[X] + \sigma_{t_T}^2 I)` (``azula/sample.py:96-128``).

        Arguments:
            shape: The shape :math:`(*)` of the tensor.
            mean: The mean of :math:`p(X)`, with shape :math:`()` or :math:`(*)`.
            var: The variance of :math:`p(X)`, with shape :math:`()` or :math:`(*)`.
            kwargs: Keyword arguments passed to :func:`torch.Tensor.to`.
        """
        t_T = self.timesteps[0]

        alpha_T, sigma_T = self.denoiser.schedule(t_T)
        alpha_T, sigma_T = alpha_T.to(**kwargs), sigma_T.to(**kwargs)

        mean_T, std_T = alpha_T * mean, torch.sqrt(alpha_T**2 * var + sigma_T**2)

        scalar = mean_T.ndim == 0 and std_T.ndim == 0 and mean_T.dtype == torch.float32 == std_T.dtype
        numel = math.prod(shape)
        if alpha_T.is_cuda and scalar and numel > 0:
            # one kernel: Philox draw (same bits as randn_like) + affine map
            x = torch.empty(tuple(shape), dtype=torch.float32, device=alpha_T.device)
            with torch.cuda.device(x.device):
                gen = _loop.default_generator(x.device)
                threads, inc, first = self._rng_layout(numel)
                seed, offset = gen.initial_seed(), gen.get_offset()
                _lib.check(
                    _lib.lib().azb_init_noise_f32(
                        x.data_ptr(), numel, float(mean_T), float(std_T), seed, offset, threads, first,
                        _lib.stream_ptr(x.device),
                    ),
                    "azb_init_noise_f32",
                )
                gen.set_offset(offset + inc)
            return x

        mean_T, std_T = mean_T.expand(shape), std_T.expand(shape)
        return mean_T + std_T * torch.randn_like(mean_T)

    def progress_bar(self, it: Iterable) -> Iterable:
        if torch.is_tensor(it):
            it = it.unbind()
        if self.silent:
            return it
        return tqdm(it, miniters=1, unit="step", ncols=79, ascii=True)

    # ------------------------------------------------------------------------------- loop
    def _fusable(self, x: Tensor) -> bool:
        return False

    def _eta(self) -> float | None:
        return None

    def _signature(self) -> tuple:
        r"""The sampler's own hyper-parameters that are frozen into a coefficient table."""
        return ()

    def _table(self, grid):
        r"""The per-stage coefficient table of this sampler (:mod:`azula_b200.engine.table`), or :py:`None`."""
        return None

    def _fused(self, x: Tensor, kwargs: dict) -> Tensor | None:
        r"""Runs the whole loop as graph replays of [backbone + one transition kernel] when the (sampler, denoiser,
        input) triple allows it; :py:`None` tells the caller to run its generic loop."""
        if not (self._fusable(x) and _loop.supports(self, x, kwargs)):
            return None
        with torch.cuda.device(x.device):
            key = _loop.signature(self, x, kwargs)
            loop = self._loops.get(key)
            if loop is None:
                self._loops.clear()  # one live graph per sampler keeps device memory bounded
                loop = _loop.FusedLoop(self, x, kwargs, self.graph, self.unroll)
                if loop.table is None:  # the sampler has no table for these hyper-parameters
                    return None
                self._loops[key] = loop
            return loop.run(x, kwargs, self.progress_bar)

    @torch.no_grad()
    def __call__(self, x: Tensor, **kwargs) -> Tensor:
        r"""Simulates the reverse process from :math:`t_T` to :math:`t_0`
        (``azula/sample.py:139-161``); never mutates :py:`x`."""
        out = self._fused(x, kwargs)
        if out is not None:
            return out

        time_pairs = self.timesteps.unfold(0, 2, 1).to(device=x.device)

        x_t = x
        for t, s in self.progress_bar(time_pairs):
            x_t = self.step(x_t, t, s, **kwargs)
        return x_t

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        r"""Simulates the reverse process from :math:`t` to :math:`s`; returns
        :math:`x_s \sim q(X_s \mid x_t)`."""
        raise NotImplementedError()


class _Ancestral(Sampler):
    r"""Shared implementation of DDPM/DDIM:

    .. math:: x_s = \alpha_s \mu + \sigma_s \sqrt{1 - \tau'} \, \frac{x_t - \alpha_t \mu}{\sigma_t}
        + \sigma_s \sqrt{\tau'} \, \varepsilon \qquad
        \tau = 1 - \frac{\alpha_t^2}{\alpha_s^2} \frac{\sigma_s^2}{\sigma_t^2}

    with :math:`\tau' = \tau` (DDPM) or :math:`\mathrm{clip}(\eta \tau, 0, 1)` (DDIM).
    """

    def __init__(self, denoiser: Denoiser, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser

    def _fusable(self, x: Tensor) -> bool:
        # a subclass that overrides step() (e.g. guidance samplers) must see its own step run
        return type(self).step is _Ancestral.step

    def _signature(self) -> tuple:
        return (self._eta(),)

    def _table(self, grid):
        return _table.ancestral(grid, self._eta())

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        alpha_s, sigma_s = self.denoiser.schedule(s)
        alpha_t, sigma_t = self.denoiser.schedule(t)
        k, n = transition_scalars(alpha_t, sigma_t, alpha_s, sigma_s, self._eta())

        mean = self.denoiser(x_t, t, **kwargs).mean

        if _cuda_step_ok(x_t, mean, alpha_s):
            return _cuda_step(x_t, mean, alpha_s, k, alpha_t, n, self._rng_layout(x_t.numel()))

        x_s = alpha_s * mean
        x_s = x_s + k * (x_t - alpha_t * mean)
        x_s = x_s + n * torch.randn_like(x_t)
        return x_s


class DDPMSampler(_Ancestral):
    r"""DDPM sampler (Ho et al., 2020; ``azula/sample.py:179-216``).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """


class DDIMSampler(_Ancestral):
    r"""DDIM sampler (Song et al., 2021; ``azula/sample.py:219-261``).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        eta: The stochasticity :math:`\eta \in \mathbb{R}_+` (1 = DDPM, 0 = deterministic).
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def __init__(self, denoiser: Denoiser, eta: float = 0.0, **kwargs) -> None:
        super().__init__(denoiser, **kwargs)

        self.eta = eta

    def _eta(self) -> float | None:
        return self.eta


# ------------------------------------------------------------------------- one-step ODE / SDE samplers


def _affine(x_t: Tensor, mean: Tensor, a, k, b, n, sampler: Sampler | None = None) -> Tensor | None:
    r""":math:`x_s = a \mu + k (x_t - b \mu) + n \varepsilon` as ONE ``azb_step_f32`` launch when the tensors
    live on a CUDA device (scalars are 0-d tensors); :py:`None` when the caller must use torch arithmetic."""
    if _cuda_step_ok(x_t, mean, a):
        layout = None if sampler is None else sampler._rng_layout(x_t.numel())  # noise addressed by global index
        return _cuda_step(x_t, mean, a, k, b, n, layout)
    return None


class EulerSampler(Sampler):
    r"""Explicit Euler (1st order) sampler of the probability-flow ODE (``azula/sample.py:264-304``).

    With :math:`z(x_t) = (x_t - \alpha_t \mu) / \sigma_t`,

    .. math:: x_s = \frac{\alpha_s}{\alpha_t} x_t + \alpha_s \left( \frac{\sigma_s}{\alpha_s} -
        \frac{\sigma_t}{\alpha_t} \right) z(x_t)

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def __init__(self, denoiser: Denoiser, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser

    def _fusable(self, x: Tensor) -> bool:
        return _unchanged(self, EulerSampler, "step")

    def _table(self, grid):
        return _table.euler(grid)

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        alpha_s, sigma_s = self.denoiser.schedule(s)
        alpha_t, sigma_t = self.denoiser.schedule(t)

        mean = self.denoiser(x_t, t, **kwargs).mean
        slope = alpha_s * (sigma_s / alpha_s - sigma_t / alpha_t)

        # a mu + k (x - alpha_t mu) with k = alpha_s / alpha_t + slope / sigma_t and a = alpha_s
        fused = _affine(x_t, mean, alpha_s, alpha_s / alpha_t + slope / sigma_t, alpha_t, None)
        if fused is not None:
            return fused

        z_t = (x_t - alpha_t * mean) / sigma_t
        return alpha_s / alpha_t * x_t + slope * z_t


class HeunSampler(Sampler):
    r"""Explicit Heun (2nd order) sampler: an Euler predictor, then the same step with the average of the
    slopes at both ends; two denoiser evaluations per step (``azula/sample.py:306-352``).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def __init__(self, denoiser: Denoiser, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser

    def _fusable(self, x: Tensor) -> bool:
        return _unchanged(self, HeunSampler, "step")

    def _table(self, grid):
        return _table.heun(grid)

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        alpha_s, sigma_s = self.denoiser.schedule(s)
        alpha_t, sigma_t = self.denoiser.schedule(t)
        slope = alpha_s * (sigma_s / alpha_s - sigma_t / alpha_t)

        z_t = (x_t - alpha_t * self.denoiser(x_t, t, **kwargs).mean) / sigma_t
        x_s = alpha_s / alpha_t * x_t + slope * z_t

        z_s = (x_s - alpha_s * self.denoiser(x_s, s, **kwargs).mean) / sigma_s
        return alpha_s / alpha_t * x_t + slope * ((z_t + z_s) / 2)


class ItoSampler(Sampler):
    r"""First-order sampler of the Ito SDE with stochasticity :math:`\eta` and temperature :math:`\tau`
    (``azula/sample.py:355-431``):

    .. math:: x_s = \frac{\alpha_s}{\alpha_t} x_t + \frac{1 + \eta^2}{\tau} \left( \frac{\sigma_s}{\sigma_t}
        - \frac{\alpha_s}{\alpha_t} \right) (x_t - \alpha_t \mu) + \eta \, \alpha_s \sqrt{\left|
        \frac{\sigma_t^2}{\alpha_t^2} - \frac{\sigma_s^2}{\alpha_s^2} \right|} \, \varepsilon

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        eta: The stochasticity parameter :math:`\eta \geq 0`.
        temperature: The temperature parameter :math:`\tau \geq 0`.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def __init__(self, denoiser: Denoiser, eta: float = 1.0, temperature: float = 1.0, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser
        self.eta = eta
        self.temperature = temperature

    def _fusable(self, x: Tensor) -> bool:
        return _unchanged(self, ItoSampler, "step") and self.temperature != 0

    def _signature(self) -> tuple:
        return (self.eta, self.temperature)

    def _table(self, grid):
        return _table.ito(grid, self.eta, self.temperature)

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        alpha_s, sigma_s = self.denoiser.schedule(s)
        alpha_t, sigma_t = self.denoiser.schedule(t)

        mean = self.denoiser(x_t, t, **kwargs).mean

        ratio = alpha_s / alpha_t
        drift = (1 + self.eta**2) / self.temperature * (sigma_s / sigma_t - ratio)
        noise = self.eta * alpha_s * torch.sqrt(torch.abs((sigma_t / alpha_t) ** 2 - (sigma_s / alpha_s) ** 2))

        # ratio x + drift (x - alpha_t mu) + noise eps = alpha_s mu + (ratio + drift) (x - alpha_t mu) + noise eps
        fused = _affine(x_t, mean, alpha_s, ratio + drift, alpha_t, noise, self)
        if fused is not None:
            return fused

        x_s = ratio * x_t
        x_s = x_s + drift * (x_t - alpha_t * mean)
        return x_s + noise * torch.randn_like(x_s)


class PCSampler(Sampler):
    r"""Predictor-corrector sampler: ``corrections`` Langevin-like corrector steps of amplitude
    :math:`\delta` at time :math:`t`, then a deterministic (DDIM) predictor to :math:`s`
    (``azula/sample.py:953-999``).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        corrections: The number of corrector steps for each predictor step.
        delta: The amplitude of corrector steps :math:`\delta \in [0,1]`.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def __init__(self, denoiser: Denoiser, corrections: int = 1, delta: float = 0.01, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser
        self.corrections = corrections
        self.delta = delta

    def _fusable(self, x: Tensor) -> bool:
        return _unchanged(self, PCSampler, "step") and self.corrections >= 0 and 0 <= self.delta <= 1

    def _signature(self) -> tuple:
        return (self.corrections, self.delta)

    def _table(self, grid):
        return _table.predictor_corrector(grid, int(self.corrections), float(self.delta))

    def step(self, x_t: Tensor, t: Tensor, s: Tensor, **kwargs) -> Tensor:
        alpha_s, sigma_s = self.denoiser.schedule(s)
        alpha_t, sigma_t = self.denoiser.schedule(t)
        keep, kick = math.sqrt(1 - self.delta), math.sqrt(self.delta)

        for _ in range(self.corrections):
            mean = self.denoiser(x_t, t, **kwargs).mean
            fused = _affine(x_t, mean, alpha_t, keep * torch.ones_like(alpha_t), alpha_t, kick * sigma_t, self)
            if fused is not None:
                x_t = fused
            else:
                x_t = alpha_t * mean + keep * (x_t - alpha_t * mean) + kick * sigma_t * torch.randn_like(x_t)

        mean = self.denoiser(x_t, t, **kwargs).mean
        fused = _affine(x_t, mean, alpha_s, sigma_s / sigma_t, alpha_t, None)
        if fused is not None:
            return fused
        return alpha_s * mean + sigma_s / sigma_t * (x_t - alpha_t * mean)


# ----------------------------------------------------------------------------------- multi-step samplers


class _Multistep(Sampler):
    r"""Adams-Bashforth-type samplers: in a variable :math:`u(t)` in which the probability-flow ODE is (semi-)
    linear, the integral of the last :math:`n` evaluations' Lagrange interpolant is added at every step,

    .. math:: x_s = r_i \, x_t + g_i \sum_{j=1}^{n} w_{ij} \, h_{t_j}

    where the weights solve the Vandermonde system :math:`V w = b` with :math:`V_{kj} = u_j^k` and
    :math:`b_k = \int_{u_t}^{u_s} \omega(v) \, v^k \, dv` in float64 (``azula/sample.py:485-508``).
    Subclasses define :math:`u`, the moments :math:`b`, the stored quantity :math:`h` and :math:`(r_i, g_i)`.
    """

    def __init__(self, denoiser: Denoiser, order: int = 2, **kwargs) -> None:
        super().__init__(**kwargs)

        self.denoiser = denoiser
        self.order = order

    # -- to be provided
    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        raise NotImplementedError()

    @staticmethod
    def _moments(u: Tensor, i: int, k: Tensor) -> Tensor:
        raise NotImplementedError()

    def _stored(self, x_t: Tensor, mean: Tensor, alpha: Tensor, sigma: Tensor, i: int) -> Tensor:
        raise NotImplementedError()

    def _update(self, x_t: Tensor, integral: Tensor, alpha: Tensor, sigma: Tensor, i: int) -> Tensor:
        raise NotImplementedError()

    # -- the same two hooks as scalar coefficients, for the fused loop's table: h = p x + q mean, x_s = r x + g integral
    def _stored_coef(self, alpha: Tensor, sigma: Tensor) -> tuple[Tensor, Tensor]:
        raise NotImplementedError()

    def _update_coef(self, alpha_t: Tensor, sigma_t: Tensor, alpha_s: Tensor, sigma_s: Tensor) -> tuple[Tensor, Tensor]:
        raise NotImplementedError()

    _HOOKS = ("__call__", "_variable", "_moments", "_stored", "_update", "_weights", "_stored_coef", "_update_coef")

    def _fusable(self, x: Tensor) -> bool:
        base = next((b for b in type(self).__mro__ if b in _MULTISTEP_TYPES), None)
        return base is not None and _unchanged(self, base, *self._HOOKS)

    def _signature(self) -> tuple:
        return (self.order,)

    def _table(self, grid):
        return _table.multistep(grid, self)

    # -- shared
    @classmethod
    def _weights(cls, u: Tensor, i: int, n: int) -> Tensor:
        r"""The :math:`\min(n, i + 1)` weights of step :math:`i`, computed in float64, returned in :py:`u.dtype`."""
        wide = u.to(torch.promote_types(u.dtype, torch.float64))
        n = min(n, i + 1)
        k = torch.arange(n, device=u.device)
        vandermonde = wide[i + 1 - n : i + 1] ** k[:, None]
        return torch.linalg.solve(vandermonde, cls._moments(wide, i, k)).to(u.dtype)

    @torch.no_grad()
    def __call__(self, x: Tensor, **kwargs) -> Tensor:
        out = self._fused(x, kwargs)
        if out is not None:
            return out

        time = self.timesteps.to(device=x.device)
        alpha, sigma = self.denoiser.schedule(time)
        u = self._variable(alpha, sigma)

        x_t, history = x, []
        for i, t in enumerate(self.progress_bar(time[:-1])):
            mean = self.denoiser(x_t, t, **kwargs).mean
            history.append(self._stored(x_t, mean, alpha, sigma, i))
            del history[: -self.order]

            weights = self._weights(u, i, self.order)
            integral = sum(h * w for h, w in zip(history, weights, strict=True))
            x_t = self._update(x_t, integral, alpha, sigma, i)

        return x_t


class zABSampler(_Multistep):
    r"""Adams-Bashforth sampler with noise (:math:`z`) prediction in :math:`u = \sigma / \alpha`
    (``azula/sample.py:434-537``; the LMS sampler of k-diffusion).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        order: The order :math:`n` of the multi-step method.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        return sigma / alpha

    @staticmethod
    def _moments(u: Tensor, i: int, k: Tensor) -> Tensor:
        return u[i + 1] ** (k + 1) / (k + 1) - u[i] ** (k + 1) / (k + 1)

    _adams_bashforth = classmethod(lambda cls, t, i, n: cls._weights(t, i, n))

    def _stored(self, x_t, mean, alpha, sigma, i):
        return (x_t - alpha[i] * mean) / sigma[i]

    def _update(self, x_t, integral, alpha, sigma, i):
        return alpha[i + 1] / alpha[i] * x_t + alpha[i + 1] * integral

    def _stored_coef(self, alpha, sigma):
        return 1 / sigma, -alpha / sigma

    def _update_coef(self, alpha_t, sigma_t, alpha_s, sigma_s):
        return alpha_s / alpha_t, alpha_s


class vABSampler(zABSampler):
    r"""Adams-Bashforth sampler with velocity (:math:`v`) prediction in :math:`u = \sigma / (\alpha +
    \sigma)` (``azula/sample.py:540-593``).

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        order: The order :math:`n` of the multi-step method.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        return sigma / (alpha + sigma)

    def _stored(self, x_t, mean, alpha, sigma, i):
        return 1 / sigma[i] * x_t - (1 + alpha[i] / sigma[i]) * mean

    def _update(self, x_t, integral, alpha, sigma, i):
        total_s, total_t = alpha[i + 1] + sigma[i + 1], alpha[i] + sigma[i]
        return total_s / total_t * x_t + total_s * integral

    def _stored_coef(self, alpha, sigma):
        return 1 / sigma, -(1 + alpha / sigma)

    def _update_coef(self, alpha_t, sigma_t, alpha_s, sigma_s):
        return (alpha_s + sigma_s) / (alpha_t + sigma_t), alpha_s + sigma_s


def _factorials(k: Tensor) -> Tensor:
    return torch.cumprod(torch.clip(k, min=1), dim=0)


class zEABSampler(_Multistep):
    r"""Exponential Adams-Bashforth sampler with noise prediction in :math:`u = \log(\sigma / \alpha)`: a
    multi-step DPM-Solver (``azula/sample.py:596-699``); moments :math:`\int e^v v^k dv`.

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        order: The order :math:`n` of the multi-step method.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        return sigma.log() - alpha.log()

    @staticmethod
    def _moments(u: Tensor, i: int, k: Tensor) -> Tensor:
        fact = _factorials(k)
        hi = torch.exp(u[i + 1]) * torch.cumsum((-u[i + 1]) ** k / fact, dim=0)
        lo = torch.exp(u[i]) * torch.cumsum((-u[i]) ** k / fact, dim=0)
        return (-1) ** k * fact * (hi - lo)

    _exponential_adams_bashforth = classmethod(lambda cls, t, i, n: cls._weights(t, i, n))

    def _stored(self, x_t, mean, alpha, sigma, i):
        return (x_t - alpha[i] * mean) / sigma[i]

    def _update(self, x_t, integral, alpha, sigma, i):
        return alpha[i + 1] / alpha[i] * x_t + alpha[i + 1] * integral

    def _stored_coef(self, alpha, sigma):
        return 1 / sigma, -alpha / sigma

    def _update_coef(self, alpha_t, sigma_t, alpha_s, sigma_s):
        return alpha_s / alpha_t, alpha_s


class xEABSampler(_Multistep):
    r"""Exponential Adams-Bashforth sampler with data (:math:`x`) prediction: a multi-step DPM-Solver++
    (``azula/sample.py:702-801``); moments :math:`\int e^{-v} v^k dv`.

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        order: The order :math:`n` of the multi-step method.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        return sigma.log() - alpha.log()

    @staticmethod
    def _moments(u: Tensor, i: int, k: Tensor) -> Tensor:
        fact = _factorials(k)
        hi = torch.exp(-u[i + 1]) * torch.cumsum(u[i + 1] ** k / fact, dim=0)
        lo = torch.exp(-u[i]) * torch.cumsum(u[i] ** k / fact, dim=0)
        return -fact * (hi - lo)

    _exponential_adams_bashforth = classmethod(lambda cls, t, i, n: cls._weights(t, i, n))

    def _stored(self, x_t, mean, alpha, sigma, i):
        return mean

    def _update(self, x_t, integral, alpha, sigma, i):
        return sigma[i + 1] / sigma[i] * x_t - sigma[i + 1] * integral

    def _stored_coef(self, alpha, sigma):
        return torch.zeros_like(alpha), torch.ones_like(alpha)

    def _update_coef(self, alpha_t, sigma_t, alpha_s, sigma_s):
        return sigma_s / sigma_t, -sigma_s


class REABSampler(_Multistep):
    r"""Rosenbrock-type exponential Adams-Bashforth sampler: a multi-step DPM-Solver-v3
    (``azula/sample.py:804-950``); moments :math:`\int \frac{e^v}{1 + e^{2v}} v^k dv` by the trapezoidal
    rule on 257 points.

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        order: The order :math:`n` of the multi-step method.
        kwargs: Keyword arguments passed to :class:`Sampler`.
    """

    def _variable(self, alpha: Tensor, sigma: Tensor) -> Tensor:
        return sigma.log() - alpha.log()

    @staticmethod
    def _moments(u: Tensor, i: int, k: Tensor) -> Tensor:
        v = torch.linspace(u[i], u[i + 1], steps=256 + 1, dtype=u.dtype, device=u.device)
        y = torch.exp(v) / (1 + torch.exp(2 * v)) * (v ** k[:, None])
        return torch.trapezoid(y, v, dim=-1)

    _exponential_adams_bashforth = classmethod(lambda cls, t, i, n: cls._weights(t, i, n))

    def _stored(self, x_t, mean, alpha, sigma, i):
        a_t = sigma[i] ** 2 / (alpha[i] ** 2 + sigma[i] ** 2)
        b_t = sigma[i] * torch.rsqrt(alpha[i] ** 2 + sigma[i] ** 2)
        return (1 - a_t) / b_t / alpha[i] * x_t - 1 / b_t * mean

    def _update(self, x_t, integral, alpha, sigma, i):
        alpha_t, sigma_t, alpha_s, sigma_s = alpha[i], sigma[i], alpha[i + 1], sigma[i + 1]
        # the reference mixes alpha_s with sigma_t in the second square root (azula/sample.py:944); kept as is
        return (
            torch.sqrt((alpha_s**2 + sigma_s**2) / (alpha_t**2 + sigma_t**2)) * x_t
            + torch.sqrt(alpha_s**2 + sigma_t**2) * integral
        )

    def _stored_coef(self, alpha, sigma):
        a_t = sigma**2 / (alpha**2 + sigma**2)
        b_t = sigma * torch.rsqrt(alpha**2 + sigma**2)
        return (1 - a_t) / b_t / alpha, -1 / b_t

    def _update_coef(self, alpha_t, sigma_t, alpha_s, sigma_s):
        return torch.sqrt((alpha_s**2 + sigma_s**2) / (alpha_t**2 + sigma_t**2)), torch.sqrt(alpha_s**2 + sigma_t**2)


_MULTISTEP_TYPES = (REABSampler, xEABSampler, zEABSampler, vABSampler, zABSampler)


# ---------------------------------------------------------------- eager step on a CUDA device


def _cuda_step_ok(x_t: Tensor, mean: Tensor, alpha_s: Tensor) -> bool:
    from . import engine

    return (
        x_t.is_cuda
        and engine.native_enabled()
        and x_t.dtype == torch.float32
        and mean.dtype == torch.float32
        and mean.shape == x_t.shape
        and alpha_s.ndim == 0
        and x_t.numel() > 0
        and not (torch.is_grad_enabled() and (x_t.requires_grad or mean.requires_grad))
    )


def _cuda_step(x_t: Tensor, mean: Tensor, alpha_s, k, alpha_t, n, layout=None) -> Tensor:
    r"""The affine update of one generic step as a single ``azb_step_f32`` launch
    (row = [0, 1, alpha_s, k, alpha_t, n, 1, inf] so that m = F = the posterior mean).  :py:`n=None`: the
    transition has no noise term, nothing is drawn and the generator does not advance."""
    with torch.cuda.device(x_t.device):
        zero = torch.zeros((), dtype=torch.float32, device=x_t.device)
        draws = n is not None
        n = zero if n is None else n
        cols = [zero, zero + 1, alpha_s, k, alpha_t, n, zero + 1, zero + float("inf")]
        row = torch.stack([c.to(device=x_t.device, dtype=torch.float32).reshape(()) for c in cols])
        idx = torch.zeros((), dtype=torch.int32, device=x_t.device)
        x_c, m_c = x_t.contiguous(), mean.contiguous()
        out = torch.empty_like(x_c)
        gen = _loop.default_generator(x_t.device)
        threads, inc, first = layout if layout is not None else (*_lib.rng_policy(x_c.numel()), 0)
        seed, offset = gen.initial_seed(), gen.get_offset()
        _lib.check(
            _lib.lib().azb_step_f32(
                x_c.data_ptr(), m_c.data_ptr(), _lib.F32, x_c.numel(), None, out.data_ptr(), None, _lib.F32,
                x_c.numel(), 1, row.data_ptr(), idx.data_ptr(), seed, None, offset, threads, first,
                _lib.stream_ptr(x_t.device),
            ),
            "azb_step_f32",
        )
        if draws:
            gen.set_offset(offset + inc)
    return out.reshape(x_t.shape)
