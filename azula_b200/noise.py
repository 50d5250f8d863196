r"""Noise schedules :math:`t \mapsto (\alpha_t, \sigma_t)` (interface of ``azula/noise.py``).

The perturbation kernel is :math:`p(X_t \mid X) = \mathcal{N}(\alpha_t X, \sigma_t^2 I)`.
Schedules stay plain callables on tensors: the engine evaluates the schedule object itself,
once per sampler, on the time grid to freeze the per-step coefficient table
(:mod:`azula_b200.engine.table`), so any user schedule works with the fused step kernel.
"""

from __future__ import annotations

__all__ = ["Schedule", "VPSchedule", "VESchedule", "CosineSchedule", "RectifiedSchedule", "DecaySchedule"]

import abc
import math
import torch

from torch import Tensor


class Schedule(abc.ABC):
    r"""Abstract noise schedule (``azula/noise.py:49-63``)."""

    @abc.abstractmethod
    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        r"""Returns :math:`(\alpha_t, \sigma_t)`, each with the shape of :py:`t`."""

    def alpha(self, t: Tensor) -> Tensor:
        return self(t)[0]

    def sigma(self, t: Tensor) -> Tensor:
        return self(t)[1]


class VPSchedule(Schedule):
    r"""Variance preserving schedule (``azula/noise.py:99-129``).

    .. math:: \alpha_t = \exp(t^2 \log \alpha_\min) \qquad
        \sigma_t = \sqrt{1 - \alpha_t^2 + \sigma_\min^2}
    """

    def __init__(self, alpha_min: float = 1e-3, sigma_min: float = 1e-3) -> None:
        self.alpha_min = alpha_min
        self.sigma_min = sigma_min

    def alpha(self, t: Tensor) -> Tensor:
        return torch.exp(math.log(self.alpha_min) * t**2)

    def sigma(self, t: Tensor) -> Tensor:
        return torch.sqrt(1 - self.alpha(t) ** 2 + self.sigma_min**2)

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        return self.alpha(t), self.sigma(t)


class VESchedule(Schedule):
    r"""Variance exploding schedule, :math:`\alpha_t = 1`, log-linear :math:`\sigma_t`
    (``azula/noise.py:66-96``)."""

    def __init__(self, sigma_min: float = 1e-3, sigma_max: float = 1e3) -> None:
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        log_sigma = (1 - t) * math.log(self.sigma_min) + t * math.log(self.sigma_max)
        return torch.ones_like(t), torch.exp(log_sigma)


class CosineSchedule(Schedule):
    r"""Cosine schedule, :math:`\alpha_t = \cos(t \arccos \alpha_\min)` (``azula/noise.py:132-155``)."""

    def __init__(self, alpha_min: float = 1e-3, sigma_min: float = 1e-3) -> None:
        self.alpha_min = alpha_min
        self.sigma_min = sigma_min

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        alpha = torch.cos(math.acos(self.alpha_min) * t)
        sigma = torch.sqrt(1 - torch.cos(math.acos(self.alpha_min) * t) ** 2 + self.sigma_min**2)
        return alpha, sigma


def _lerp_pair(u: Tensor, alpha_min: float, sigma_min: float) -> tuple[Tensor, Tensor]:
    return u * alpha_min + (1 - u), u + (1 - u) * sigma_min


class RectifiedSchedule(Schedule):
    r"""Rectified schedule: straight lines from :math:`(1, \sigma_\min)` to
    :math:`(\alpha_\min, 1)` (``azula/noise.py:158-188``)."""

    def __init__(self, alpha_min: float = 1e-3, sigma_min: float = 1e-3) -> None:
        self.alpha_min = alpha_min
        self.sigma_min = sigma_min

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        return _lerp_pair(t, self.alpha_min, self.sigma_min)


class DecaySchedule(Schedule):
    r"""Rectified schedule in the warped time :math:`\tau = (1 - \gamma^t) / (1 - \gamma)`
    (``azula/noise.py:191-231``)."""

    def __init__(self, alpha_min: float = 1e-3, sigma_min: float = 1e-3, gamma: float = 0.1) -> None:
        self.alpha_min = alpha_min
        self.sigma_min = sigma_min
        self.gamma = gamma

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        tau = (1 - self.gamma**t) / (1 - self.gamma)
        return _lerp_pair(tau, self.alpha_min, self.sigma_min)
