r"""Denoisers :math:`q_\phi(X \mid x_t)` (interface of ``azula/denoise.py``).

Every preconditioned denoiser is described by ONE function, :meth:`Preconditioned.coefficients`,

.. math:: \mu_\phi(x_t) = c_\mathrm{skip}(t) \, x_t + c_\mathrm{out}(t) \,
    b_\phi(c_\mathrm{in}(t) \, x_t, c_\mathrm{time}(t))

which serves both the eager :meth:`forward` (any device, autograd friendly) and the
per-step coefficient table of the fused sampling loop (:mod:`azula_b200.engine.table`), so
the two can never disagree.
"""

from __future__ import annotations

__all__ = [
    "Posterior",
    "DiracPosterior",
    "GaussianPosterior",
    "Denoiser",
    "Preconditioned",
    "SimpleDenoiser",
    "KarrasDenoiser",
]

import abc
import math
import torch
import torch.nn as nn

from dataclasses import dataclass
from torch import Tensor

from .nn.utils import get_module_dtype
from .noise import Schedule


class Posterior(abc.ABC):
    r"""Abstract posterior :math:`q_\phi(X \mid x_t)`; exposes at least :py:`mean`."""

    mean: Tensor


class DiracPosterior(Posterior):
    r"""Dirac delta :math:`\delta(X - \mu)` (``azula/denoise.py:56-66``)."""

    def __init__(self, mean: Tensor) -> None:
        self.mean = mean


class GaussianPosterior(Posterior):
    r"""Diagonal Gaussian :math:`\mathcal{N}(X \mid \mu, \sigma^2)` (``azula/denoise.py:69-94``)."""

    def __init__(self, mean: Tensor, var: Tensor) -> None:
        self.mean = mean
        self.var = var

    def log_prob(self, x: Tensor) -> Tensor:
        r"""Element-wise log-density :math:`\log \mathcal{N}(x \mid \mu, \sigma^2)`."""
        return -((x - self.mean) ** 2 / self.var + torch.log(self.var) + math.log(2 * math.pi)) / 2


class Denoiser(nn.Module):
    r"""Abstract denoiser: an ``nn.Module`` with a :py:`schedule` attribute whose
    :py:`forward(x_t, t, **kwargs)` returns a :class:`Posterior` (``azula/denoise.py:97-114``)."""

    schedule: Schedule

    @abc.abstractmethod
    def forward(self, x_t: Tensor, t: Tensor, **kwargs) -> Posterior:
        pass


@dataclass
class Coefficients:
    r"""Preconditioning scalars at time :math:`t` (tensors shaped like :math:`\alpha_t`)."""

    c_in: Tensor
    c_out: Tensor
    c_skip: Tensor
    c_time: Tensor


def _unsqueeze_like(a: Tensor, ndim: int) -> Tensor:
    return a.reshape(a.shape + (1,) * (ndim - a.ndim)) if a.ndim < ndim else a


class Preconditioned(Denoiser):
    r"""Denoiser of the form :math:`c_\mathrm{skip} x_t + c_\mathrm{out} b_\phi(c_\mathrm{in} x_t,
    c_\mathrm{time})` around a backbone :math:`b_\phi`.

    Subclasses only provide :meth:`coefficients`.
    """

    def __init__(self, backbone: nn.Module, schedule: Schedule) -> None:
        super().__init__()

        self.backbone = backbone
        self.schedule = schedule

    @abc.abstractmethod
    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        pass

    def time_input(self, c_time: Tensor, t: Tensor, dtype: torch.dtype | None) -> Tensor:
        r"""The backbone's second argument: :math:`c_\mathrm{time}` shaped like :py:`t`, in the
        backbone's dtype."""
        return c_time.reshape_as(t).to(dtype)

    def time_rows(self, c_time: Tensor, dtype: torch.dtype | None) -> Tensor:
        r"""Time inputs of a whole grid of steps, one row per step (row shape = shape of a 0-d
        :py:`t` as :meth:`time_input` would return it)."""
        return c_time.to(dtype)

    def call_backbone(self, x_in: Tensor, time: Tensor, **kwargs) -> Tensor:
        r"""How this denoiser invokes its backbone; the fused loop calls it directly."""
        return self.backbone(x_in, time, **kwargs)

    def fusable(self) -> bool:
        r"""Whether :meth:`forward` is the stock one, i.e. fully described by :meth:`coefficients`
        (a subclass overriding :meth:`forward` must see its own code run)."""
        return type(self).forward is Preconditioned.forward

    def forward(self, x_t: Tensor, t: Tensor, **kwargs) -> DiracPosterior:
        alpha_t, sigma_t = self.schedule(t)
        alpha_t, sigma_t = _unsqueeze_like(alpha_t, x_t.ndim), _unsqueeze_like(sigma_t, x_t.ndim)

        c = self.coefficients(alpha_t, sigma_t)
        dtype = get_module_dtype(self.backbone)

        output = self.call_backbone((c.c_in * x_t).to(dtype), self.time_input(c.c_time, t, dtype), **kwargs).to(x_t)

        if c.c_skip is None:
            return DiracPosterior(mean=output)

        return DiracPosterior(mean=c.c_skip * x_t + c.c_out * output)

    def _noised(self, x: Tensor, t: Tensor):
        alpha_t, sigma_t = self.schedule(t)
        alpha_t, sigma_t = _unsqueeze_like(alpha_t, x.ndim), _unsqueeze_like(sigma_t, x.ndim)
        return alpha_t, sigma_t, alpha_t * x + sigma_t * torch.randn_like(x)


class SimpleDenoiser(Preconditioned):
    r"""The backbone predicts the mean directly: :math:`\mu_\phi(x_t) = b_\phi(c_\mathrm{in} x_t,
    c_\mathrm{time})` with :math:`c_\mathrm{in} = (\alpha_t^2 + \sigma_t^2)^{-1/2}` and
    :math:`c_\mathrm{time} = \log(\sigma_t / \alpha_t)` (``azula/denoise.py:177-260``)."""

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        return Coefficients(
            c_in=torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_out=None,
            c_skip=None,
            c_time=torch.log(sigma_t / alpha_t),
        )

    def loss(self, x: Tensor, t: Tensor, max_weight: float = 1e4, **kwargs) -> Tensor:
        r"""SNR-weighted (clipped) mean squared error (training only; plain torch)."""
        alpha_t, sigma_t, x_t = self._noised(x, t)
        w_t = torch.clip((alpha_t / sigma_t) ** 2 + 1, max=max_weight)
        return (w_t * (self(x_t, t, **kwargs).mean - x).square()).mean()


class KarrasDenoiser(Preconditioned):
    r"""EDM-style preconditioning generalised to :math:`\alpha_t \neq 1`
    (``azula/denoise.py:263-353``).

    .. math::
        c_\mathrm{in} = \frac{1}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{out} = \frac{\sigma_t}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{skip} = \frac{\alpha_t}{\alpha_t^2 + \sigma_t^2} \quad
        c_\mathrm{time} = \log \frac{\sigma_t}{\alpha_t}

    Arguments:
        backbone: A noise/time conditional network :math:`b_\phi(x_t, t)`.
        schedule: A noise schedule.
    """

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        return Coefficients(
            c_in=torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_out=sigma_t * torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_skip=alpha_t / (alpha_t**2 + sigma_t**2),
            c_time=torch.log(sigma_t / alpha_t),
        )

    def loss(self, x: Tensor, t: Tensor, **kwargs) -> Tensor:
        r"""SNR-weighted mean squared error (training only; plain torch)."""
        alpha_t, sigma_t, x_t = self._noised(x, t)
        w_t = (alpha_t / sigma_t) ** 2 + 1
        return (w_t * (self(x_t, t, **kwargs).mean - x).square()).mean()
