r"""Just image transformer (JiT) preconditioner (interface of ``azula/plugins/jit/__init__.py``).
The JiT backbones are not re-implemented; any :py:`backbone(x, t, y=label)` module can be wrapped
(SURVEY.md section 8, row f3).
"""

from __future__ import annotations

__all__ = ["JITDenoiser", "load_model"]

import torch
import torch.nn as nn

from torch import Tensor

from ...denoise import Coefficients, Denoiser, Preconditioned
from ...noise import RectifiedSchedule, Schedule


class JITDenoiser(Preconditioned):
    r"""Clean-image prediction on the rectified-flow time axis (``azula/plugins/jit/__init__.py:32-104``):
    :math:`c_\mathrm{in} = 1 / (\alpha_t + \sigma_t)`, :math:`c_\mathrm{time} = \alpha_t / (\alpha_t + \sigma_t)`;
    without a label the null class :py:`num_classes` is used."""

    def __init__(self, backbone: nn.Module, schedule: Schedule | None = None, num_classes: int = 1000) -> None:
        super().__init__(backbone, RectifiedSchedule() if schedule is None else schedule)

        self.num_classes = num_classes

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        return Coefficients(
            c_in=1 / (alpha_t + sigma_t),
            c_out=None,
            c_skip=None,
            c_time=(alpha_t / (alpha_t + sigma_t)).flatten(),
        )

    def time_input(self, c_time: Tensor, t: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.to(dtype)

    def time_rows(self, c_time: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.to(dtype).reshape(-1, 1)

    def call_backbone(self, x_in: Tensor, time: Tensor, label: Tensor | None = None, **kwargs) -> Tensor:
        if label is None:
            label = torch.full((), self.num_classes, dtype=torch.int64, device=x_in.device)  # capture-safe fill
        return self.backbone(x_in, time, y=label.expand(x_in.shape[0]), **kwargs)

    def fusable(self) -> bool:
        return type(self).forward is JITDenoiser.forward


def load_model(name: str, ema: bool = True, **kwargs) -> Denoiser:
    r"""The pre-trained JiT backbones (``azula/plugins/jit/_src``) are outside the hot-path scope."""
    raise NotImplementedError(
        "azula_b200 ships the JiT preconditioner only; build the backbone yourself and wrap it in JITDenoiser."
    )
