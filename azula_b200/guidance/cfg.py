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

    Called directly, both evaluations go through the wrapped denoiser as in the reference.  Inside a sampler on a
    CUDA device the wrapper dissolves into the fused loop (:mod:`azula_b200.engine.loop`): when the two branches
    take the same keywords (e.g. two label tensors) they are evaluated by ONE backbone forward over the 2B-sample
    batch :math:`[c_+; c_-]` held in static buffers, otherwise by two forwards of B; in both cases the combination
    above -- each branch's mean clipped on its own, as the wrapped denoiser would -- is computed inside the
    transition kernel from the two output pointers (``azb_step_ex_f32``), so no guided mean is ever materialised.

    Arguments:
        denoiser: A denoiser :math:`q_\phi(X \mid X_t)`.
        batched: Engine knob. :py:`None` batches the two branches into one forward when possible,
            :py:`False` always evaluates them one after the other.
    """

    guided_inner = True  # the fused loop looks through this wrapper (engine/table.py: inner_denoiser)

    def __init__(self, denoiser: Denoiser, batched: bool | None = None) -> None:
        super().__init__()

        self.denoiser = denoiser
        self.batched = batched

    def fusable(self) -> bool:
        r"""Whether :meth:`forward` is the stock one (a subclass overriding it must see its own code run)."""
        return type(self).forward is CFGDenoiser.forward

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
