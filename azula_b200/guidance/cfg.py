r"""Classifier-free guidance (interface of ``azula/guidance/cfg.py``).

References:
    | Classifier-Free Diffusion Guidance (Ho et al., 2022)
    | https://arxiv.org/abs/2207.12598
"""

from __future__ import annotations

__all__ = ["CFGDenoiser"]

from torch import Tensor
from typing import Any

from ..denoise import Denoiser, DiracPosterior
from ..noise import Schedule


class CFGDenoiser(Denoiser):
    r"""Wraps a conditional denoiser into its classifier-free-guided version
    (``azula/guidance/cfg.py:19-69``):

    .. math:: \mu = (1 + \omega) \, \mu_\phi(x_t \mid c_+) - \omega \, \mu_\phi(x_t \mid c_-)

    Both evaluations go through the wrapped denoiser, i.e. through its native backbone on a CUDA device; the
    samplers see a plain :class:`Denoiser` and run their generic loop with the fused transition kernel.

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
    """

    def __init__(self, denoiser: Denoiser) -> None:
        super().__init__()

        self.denoiser = denoiser

    @property
    def schedule(self) -> Schedule:
        return self.denoiser.schedule

    def forward(
        self,
        x_t: Tensor,
        t: Tensor,
        positive: dict[str, Any],
        negative: dict[str, Any] = {},  # noqa: B006
        guidance: float | Tensor = 1.0,
        **kwargs,
    ) -> DiracPosterior:
        r"""
        Arguments:
            x_t: A noisy tensor :math:`x_t`, with shape :math:`(B, *)`.
            t: The time :math:`t`, with shape :math:`()` or :math:`(B)`.
            positive: The positive label :math:`c_+` as a dictionary of keyword arguments.
            negative: The negative label :math:`c_-` as a dictionary of keyword arguments.
            guidance: The classifier-free guidance strength :math:`\omega \in \mathbb{R}_+`.
            kwargs: Optional keyword arguments.
        """
        pos = self.denoiser(x_t, t, **positive, **kwargs).mean
        neg = self.denoiser(x_t, t, **negative, **kwargs).mean

        return DiracPosterior(mean=pos + guidance * (pos - neg))
