r"""Velocity diffusion model (VDM) preconditioner (interface of ``azula/plugins/vdm/__init__.py``).

Only the denoiser -- a row of the coefficient table of the fused sampling loop -- is in scope
(SURVEY.md section 8, row f3); the v-diffusion backbones themselves are not re-implemented, any
:py:`backbone(x, t)` module can be wrapped.
"""

from __future__ import annotations

__all__ = ["VelocityDenoiser", "load_model"]

import math
import torch
import torch.nn as nn

from torch import Tensor

from ...denoise import Coefficients, Denoiser, Preconditioned
from ...noise import Schedule, VPSchedule


class VelocityDenoiser(Preconditioned):
    r"""Denoiser around a velocity-prediction network (``azula/plugins/vdm/__init__.py:31-75``).

    .. math:: c_\mathrm{in} = \frac{1}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{out} = -\frac{\sigma_t}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{skip} = \frac{\alpha_t}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{time} = \frac{2}{\pi} \operatorname{atan2}(\sigma_t, \alpha_t)

    Arguments:
        backbone: A time conditional network.
        schedule: A noise schedule. If :py:`None`, :py:`VPSchedule(alpha_min=1e-2, sigma_min=1e-2)`.
    """

    def __init__(self, backbone: nn.Module, schedule: Schedule | None = None) -> None:
        super().__init__(backbone, VPSchedule(alpha_min=1e-2, sigma_min=1e-2) if schedule is None else schedule)

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        return Coefficients(
            c_in=torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_out=-sigma_t * torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_skip=alpha_t * torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_time=torch.atan2(sigma_t, alpha_t).flatten() / math.pi * 2,
        )

    def time_input(self, c_time: Tensor, t: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.to(dtype)  # flattened: (1,) for a 0-d t, (B,) otherwise

    def time_rows(self, c_time: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.to(dtype).reshape(-1, 1)

    def fusable(self) -> bool:
        return type(self).forward is VelocityDenoiser.forward


def load_model(name: str, **kwargs) -> Denoiser:
    r"""The pre-trained v-diffusion backbones (``azula/plugins/vdm/_src``) are outside the hot-path scope."""
    raise NotImplementedError(
        "azula_b200 ships the VDM preconditioner only; build the backbone yourself and wrap it in VelocityDenoiser."
    )
