r"""Per-step coefficient table of a (sampler, denoiser) pair.

The reference re-derives ~40 scalars on 0-d device tensors at every step
(``azula/sample.py:205-208,249-253`` and ``azula/denoise.py:304-312``: the schedule alone is
evaluated three times per step).  They depend only on the time grid, so the engine evaluates
the SAME expressions, in the SAME order, ONCE, vectorised over the grid on the target device
(element-wise CUDA kernels give the same bits for element i of a vector as for a 0-d tensor),
and freezes them into ``float32[steps][8]`` rows that the fused kernel indexes with a device
counter (columns: ``include/azb.h`` ``AZB_C_SKIP`` ...).
"""

from __future__ import annotations

import torch

from dataclasses import dataclass
from torch import Tensor


@dataclass
class StepTable:
    coef: Tensor  # float32 (steps, 8), device
    time: Tensor  # (steps, *time_shape) backbone time inputs, in the dtype the backbone receives
    c_in0: Tensor  # 0-d float32: pre-scale of the very first backbone input
    steps: int


def transition_scalars(alpha_t, sigma_t, alpha_s, sigma_s, eta: float | None):
    r"""(k, n) with :math:`k = \sigma_s \sqrt{1-\tau} / \sigma_t`, :math:`n = \sigma_s \sqrt{\tau}`.

    Operation order of ``azula/sample.py:208,213-214`` (DDPM, :py:`eta=None`) and
    ``:252-253,258-259`` (DDIM).
    """
    tau = 1 - (alpha_t / alpha_s * sigma_s / sigma_t) ** 2
    if eta is not None:
        tau = torch.clip(eta * tau, min=0, max=1)
    k = sigma_s * torch.sqrt(1 - tau) / sigma_t
    n = sigma_s * torch.sqrt(tau)
    return k, n


def build(sampler, device: torch.device) -> StepTable:
    r"""Evaluates the sampler's schedule and the denoiser's preconditioner on the time grid."""
    denoiser = sampler.denoiser
    pairs = sampler.timesteps.unfold(0, 2, 1).to(device=device)
    t, s = pairs[:, 0].contiguous(), pairs[:, 1].contiguous()

    alpha_s, sigma_s = denoiser.schedule(s)
    alpha_t, sigma_t = denoiser.schedule(t)
    k, n = transition_scalars(alpha_t, sigma_t, alpha_s, sigma_s, sampler._eta())

    c = denoiser.coefficients(alpha_t, sigma_t)
    one, zero = torch.ones_like(alpha_t), torch.zeros_like(alpha_t)
    c_skip = zero if c.c_skip is None else c.c_skip
    c_out = one if c.c_out is None else c.c_out
    c_in_next = torch.cat((c.c_in[1:], one[:1]))

    clip = getattr(denoiser, "mean_clip", lambda: None)()
    clip_col = torch.full_like(alpha_t, float("inf") if clip is None else float(clip))

    cols = [c_skip, c_out, alpha_s, k, alpha_t, n, c_in_next, clip_col]
    coef = torch.stack([col.expand_as(alpha_t) for col in cols], dim=-1).to(torch.float32).contiguous()

    from ..nn.utils import get_module_dtype

    dtype = get_module_dtype(denoiser.backbone)
    time = denoiser.time_rows(c.c_time, dtype).contiguous()

    return StepTable(coef=coef, time=time, c_in0=c.c_in[0].to(torch.float32), steps=len(t))
