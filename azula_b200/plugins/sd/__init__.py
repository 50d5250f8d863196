r"""Stable Diffusion preconditioner (interface of ``azula/plugins/sd/__init__.py``).  The
``diffusers`` U-Net, auto-encoder and text encoder are not re-implemented; any module with the
``diffusers`` call convention can be wrapped.
"""

from __future__ import annotations

__all__ = ["StableDenoiser", "load_model"]

import torch
import torch.nn as nn

from torch import Tensor

from ...denoise import Coefficients, Preconditioned
from ...noise import Schedule, VPSchedule


class StableDenoiser(Preconditioned):
    r"""Denoiser around a latent network trained on a discrete schedule
    (``azula/plugins/sd/__init__.py:140-224``); :py:`prediction` is :py:`"epsilon"`
    (:math:`c_\mathrm{skip} = 1 / \alpha_t`, :math:`c_\mathrm{out} = -\sigma_t / \alpha_t`) or :py:`"velocity"`.

    Arguments:
        backbone: A network called as :py:`backbone(timestep=, sample=, encoder_hidden_states=).sample`.
        sigmas: The discrete noise schedule used during training.
        schedule: A noise schedule. If :py:`None`, the VP schedule spanned by :py:`sigmas`.
        prediction: The backbone prediction type.
    """

    def __init__(self, backbone: nn.Module, sigmas: Tensor, schedule: Schedule | None = None,
                 prediction: str = "epsilon") -> None:
        if schedule is None:
            schedule = VPSchedule(alpha_min=(1 - sigmas[-1].item() ** 2) ** 0.5, sigma_min=sigmas[0].item())
        super().__init__(backbone, schedule)

        if prediction not in ("epsilon", "velocity"):
            raise ValueError(f"Unkown prediction type '{prediction}'.")

        self.prediction = prediction
        self.register_buffer("sigmas", sigmas.to(torch.get_default_dtype()))

    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        if self.prediction == "epsilon":
            c_out = -sigma_t / alpha_t
            c_skip = 1 / alpha_t
        else:
            c_out = -sigma_t * torch.rsqrt(alpha_t**2 + sigma_t**2)
            c_skip = alpha_t * torch.rsqrt(alpha_t**2 + sigma_t**2)
        c_time = sigma_t * torch.rsqrt(alpha_t**2 + sigma_t**2)
        return Coefficients(
            c_in=torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_out=c_out,
            c_skip=c_skip,
            c_time=torch.searchsorted(self.sigmas, c_time.flatten()),
        )

    def time_input(self, c_time: Tensor, t: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time  # int64 indices

    def time_rows(self, c_time: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.reshape(-1, 1)

    def call_backbone(self, x_in: Tensor, time: Tensor, prompt_embeds: Tensor, **kwargs) -> Tensor:
        B = x_in.shape[0]
        _, L, D = prompt_embeds.shape
        return self.backbone(
            timestep=time.expand(B),
            sample=x_in,
            encoder_hidden_states=prompt_embeds.to(x_in.dtype).expand(B, L, D),
            **kwargs,
        ).sample

    def fusable(self) -> bool:
        return type(self).forward is StableDenoiser.forward


def load_model(name: str, **kwargs):
    r"""Needs ``diffusers`` pipelines; outside the hot-path scope."""
    raise NotImplementedError(
        "azula_b200 ships the Stable Diffusion preconditioner only; wrap a diffusers U-Net in StableDenoiser."
    )
