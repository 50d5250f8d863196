r"""Elucidated diffusion model (EDM) schedule and preconditioner (interface of
``azula/plugins/edm/__init__.py``).  The pickled NVlabs networks are not re-implemented; any
:py:`backbone(x, sigma, class_labels=...)` module can be wrapped.
"""

from __future__ import annotations

__all__ = ["ElucidatedSchedule", "ElucidatedDenoiser", "load_model"]

import torch
import torch.nn as nn

from torch import Tensor

from ...denoise import Coefficients, Denoiser, Preconditioned
from ...noise import Schedule


class ElucidatedSchedule(Schedule):
    r""":math:`\alpha_t = 1`, :math:`\sigma_t = ((1 - t) \sigma_\min^{1/\rho} + t \sigma_\max^{1/\rho})^\rho`
    (``azula/plugins/edm/__init__.py:45-75``)."""

    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0) -> None:
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.rho = rho

    def alpha(self, t: Tensor) -> Tensor:
        return torch.ones_like(t)

    def sigma(self, t: Tensor) -> Tensor:
        lower = self.sigma_min ** (1 / self.rho)
        upper = self.sigma_max ** (1 / self.rho)
        return torch.pow((1 - t) * lower + t * upper, self.rho)

    def __call__(self, t: Tensor) -> tuple[Tensor, Tensor]:
        return self.alpha(t), self.sigma(t)


class ElucidatedDenoiser(Preconditioned):
    r"""The backbone is itself a denoiser of :math:`x_t / \alpha_t` at noise level
    :math:`\sigma_t / \alpha_t` (``azula/plugins/edm/__init__.py:77-127``)."""

    def __init__(self, backbone: nn.Module, schedule: Schedule | None = None) -> None:
        super().__init__(backbone, ElucidatedSchedule() if schedule is None else schedule)

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        return Coefficients(c_in=1 / alpha_t, c_out=None, c_skip=None, c_time=sigma_t / alpha_t)

    def call_backbone(self, x_in: Tensor, time: Tensor, label: Tensor | None = None, **kwargs) -> Tensor:
        return self.backbone(x_in, time, class_labels=label.to(x_in.dtype), **kwargs)

    def fusable(self) -> bool:
        return type(self).forward is ElucidatedDenoiser.forward


def load_model(name: str) -> Denoiser:
    r"""The pickled NVlabs/edm networks are outside the hot-path scope."""
    raise NotImplementedError(
        "azula_b200 ships the EDM schedule and preconditioner only; wrap your network in ElucidatedDenoiser."
    )
