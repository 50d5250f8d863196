r"""Per-stage coefficient table of a (sampler, denoiser) pair.

The reference re-derives ~40 scalars on 0-d device tensors at every step
(``azula/sample.py:205-208,249-253`` and ``azula/denoise.py:304-312``: the schedule alone is
evaluated three times per step).  They depend only on the time grid, so the engine evaluates
the SAME expressions, in the SAME order, ONCE, vectorised over the grid on the target device
(element-wise CUDA kernels give the same bits for element i of a vector as for a 0-d tensor),
and freezes them into ``float32[stages][32]`` rows that the fused kernel indexes with a device
counter (columns: ``include/azb.h`` ``AZB_C_SKIP`` ... ``AZB_R_W``).

A *stage* is one backbone evaluation followed by one update.  DDPM / DDIM / Euler / Ito and the
Adams-Bashforth family have one stage per sampler step, Heun two (predictor, corrector), the
predictor-corrector sampler ``corrections + 1``.  Every sampler of ``azula/sample.py`` is one of two
row kinds: *affine* (``x_s = a m + k (x - b m) + n eps``) or *history* (``h = p x + q m``,
``x_s = r x_b + sum_j W_j H_j``); see :func:`azb_step_ex_f32`.
"""

from __future__ import annotations

import torch

from dataclasses import dataclass
from torch import Tensor

from .. import _lib

F_XE, F_XB, F_OUT, F_HIST, F_STORE = 1, 2, 4, 8, 256


@dataclass
class StepTable:
    coef: Tensor  # float32 (stages, 32), device
    time: Tensor  # (stages, *time_shape) backbone time inputs, in the dtype the backbone receives
    c_in0: Tensor  # 0-d float32: pre-scale of the very first backbone input
    steps: int  # number of stages (rows)
    per_step: int = 1  # stages per sampler step
    draws: int = 0  # noise draws of the whole loop (each advances the generator by one randn_like)
    slots: int = 0  # history slots the rows address
    alt: bool = False  # whether rows address the second state buffer
    noiseless: bool = False  # no affine row has n != 0: the kernel variant without the in-register generator serves


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


def inner_denoiser(denoiser):
    r"""The preconditioned denoiser whose backbone the loop evaluates (looks through a guidance wrapper)."""
    return denoiser.denoiser if hasattr(denoiser, "guided_inner") else denoiser


class Grid:
    r"""The sampler's time grid evaluated through the denoiser's own schedule on ``device``."""

    def __init__(self, sampler, device: torch.device) -> None:
        self.sampler = sampler
        self.denoiser = inner_denoiser(sampler.denoiser)
        self.device = device
        self.time = sampler.timesteps.to(device=device)
        pairs = self.time.unfold(0, 2, 1)
        self.t, self.s = pairs[:, 0].contiguous(), pairs[:, 1].contiguous()
        # the reference evaluates the schedule at s first, then at t (sample.py:205-206,249-250)
        self.alpha_s, self.sigma_s = self.denoiser.schedule(self.s)
        self.alpha_t, self.sigma_t = self.denoiser.schedule(self.t)
        self.T = len(self.t)

    def finish(self, alpha_e: Tensor, sigma_e: Tensor, *, affine: dict | None = None, hist: dict | None = None,
               draws: Tensor | None = None, per_step: int = 1, slots: int = 0, alt: bool = False) -> StepTable:
        r"""Assembles the rows.  ``alpha_e, sigma_e`` (stages,): where each stage evaluates the denoiser.
        ``affine``: columns a, k, b, n.  ``hist``: columns p, q, r, W (stages, slots), flags (stages,) int."""
        from ..nn.utils import get_module_dtype

        den = self.denoiser
        S = alpha_e.numel()
        c = den.coefficients(alpha_e, sigma_e)
        one, zero = torch.ones_like(alpha_e), torch.zeros_like(alpha_e)
        c_skip = zero if c.c_skip is None else c.c_skip
        c_out = one if c.c_out is None else c.c_out
        c_in = c.c_in.expand_as(alpha_e)
        c_in_next = torch.cat((c_in[1:], one[:1]))
        clip = getattr(den, "mean_clip", lambda: None)()
        clip_col = torch.full_like(alpha_e, float("inf") if clip is None else float(clip))

        coef = torch.zeros(S, _lib.ROW_COLS, dtype=torch.float32, device=self.device)
        a = affine or {}
        cols = [c_skip, c_out, a.get("a", zero), a.get("k", zero), a.get("b", zero), a.get("n", zero), c_in_next, clip_col]
        coef[:, :8] = torch.stack([col.expand_as(alpha_e) for col in cols], dim=-1).to(torch.float32)
        bits = coef.view(torch.int32)
        if hist is not None:
            coef[:, _lib.R_P], coef[:, _lib.R_Q], coef[:, _lib.R_R] = hist["p"], hist["q"], hist["r"]
            W = hist["W"].to(torch.float32)
            coef[:, _lib.R_W : _lib.R_W + W.shape[1]] = W
            bits[:, _lib.R_FLAGS] = hist["flags"].to(device=self.device, dtype=torch.int32)
        if draws is None:
            draws = torch.zeros(S, dtype=torch.int64, device=self.device)
        draws = draws.to(device=self.device, dtype=torch.int64)
        bits[:, _lib.R_DRAW] = (torch.cumsum(draws, 0) - draws).to(torch.int32)

        dtype = get_module_dtype(den.backbone)
        time = den.time_rows(c.c_time, dtype).contiguous()
        noiseless = hist is not None or not bool((coef[:, 5] != 0).any().item())
        return StepTable(coef=coef.contiguous(), time=time, c_in0=c_in[0].to(torch.float32), steps=S, per_step=per_step,
                         draws=int(draws.sum().item()), slots=slots, alt=alt, noiseless=noiseless)


def build(sampler, device: torch.device) -> StepTable:
    r"""Evaluates the sampler's schedule and the denoiser's preconditioner on the time grid."""
    return sampler._table(Grid(sampler, device))


# ---------------------------------------------------------------------------------- row builders


def ancestral(g: Grid, eta: float | None) -> StepTable:
    r"""DDPM (:py:`eta=None`) / DDIM rows; one ``randn_like`` per step whatever eta (sample.py:214,259)."""
    k, n = transition_scalars(g.alpha_t, g.sigma_t, g.alpha_s, g.sigma_s, eta)
    return g.finish(g.alpha_t, g.sigma_t, affine=dict(a=g.alpha_s, k=k, b=g.alpha_t, n=n),
                    draws=torch.ones(g.T, dtype=torch.int64))


def euler(g: Grid) -> StepTable:
    r"""``x_s = alpha_s/alpha_t x + slope (x - alpha_t m)/sigma_t`` (sample.py:297-303) in affine form."""
    slope = g.alpha_s * (g.sigma_s / g.alpha_s - g.sigma_t / g.alpha_t)
    return g.finish(g.alpha_t, g.sigma_t, affine=dict(a=g.alpha_s, k=g.alpha_s / g.alpha_t + slope / g.sigma_t, b=g.alpha_t))


def ito(g: Grid, eta: float, temperature: float) -> StepTable:
    r"""sample.py:414-431."""
    ratio = g.alpha_s / g.alpha_t
    drift = (1 + eta**2) / temperature * (g.sigma_s / g.sigma_t - ratio)
    noise = eta * g.alpha_s * torch.sqrt(torch.abs((g.sigma_t / g.alpha_t) ** 2 - (g.sigma_s / g.alpha_s) ** 2))
    return g.finish(g.alpha_t, g.sigma_t, affine=dict(a=g.alpha_s, k=ratio + drift, b=g.alpha_t, n=noise),
                    draws=torch.ones(g.T, dtype=torch.int64))


def predictor_corrector(g: Grid, corrections: int, delta: float) -> StepTable:
    r"""``corrections`` Langevin-like stages at time t, then the deterministic predictor to s (sample.py:980-999);
    every stage evaluates the denoiser at t."""
    import math

    C = corrections
    keep, kick = math.sqrt(1 - delta), math.sqrt(delta)
    rep = lambda v: v[:, None].expand(g.T, C + 1)  # noqa: E731
    a = rep(g.alpha_t).clone()
    k = torch.full_like(a, keep)
    n = kick * rep(g.sigma_t).clone()
    a[:, C], k[:, C], n[:, C] = g.alpha_s, g.sigma_s / g.sigma_t, 0.0
    draws = torch.ones(g.T, C + 1, dtype=torch.int64)
    draws[:, C] = 0
    flat = lambda v: v.reshape(-1).contiguous()  # noqa: E731
    return g.finish(flat(rep(g.alpha_t)), flat(rep(g.sigma_t)),
                    affine=dict(a=flat(a), k=flat(k), b=flat(rep(g.alpha_t)), n=flat(n)), draws=flat(draws), per_step=C + 1)


def heun(g: Grid) -> StepTable:
    r"""Two stages per step (sample.py:337-352).  A: ``z_t = (x - alpha_t m)/sigma_t`` kept in slot 0, predictor
    ``r x + slope z_t`` into the alternate buffer.  B: ``z_s`` from the predictor at s, ``x_s = r x + slope (z_t + z_s)/2``."""
    slope = g.alpha_s * (g.sigma_s / g.alpha_s - g.sigma_t / g.alpha_t)
    ratio = g.alpha_s / g.alpha_t
    pair = lambda u, v: torch.stack((u, v), dim=1).reshape(-1).contiguous()  # noqa: E731
    W = torch.zeros(g.T, 2, _lib.MAX_SLOTS, dtype=torch.float32, device=g.device)
    W[:, 0, 0] = slope
    W[:, 1, 0] = slope / 2
    W[:, 1, 1] = slope / 2
    flag_a = F_OUT | F_HIST | (0 << 4) | F_STORE | (1 << 12)
    flag_b = F_XE | F_HIST | (1 << 4) | (2 << 12)
    flags = torch.tensor([flag_a, flag_b], dtype=torch.int32).repeat(g.T)
    hist = dict(p=pair(1 / g.sigma_t, 1 / g.sigma_s), q=pair(-g.alpha_t / g.sigma_t, -g.alpha_s / g.sigma_s),
                r=pair(ratio, ratio), W=W.reshape(-1, _lib.MAX_SLOTS), flags=flags)
    return g.finish(pair(g.alpha_t, g.alpha_s), pair(g.sigma_t, g.sigma_s), hist=hist, per_step=2, slots=1, alt=True)


def multistep(g: Grid, sampler) -> StepTable | None:
    r"""Adams-Bashforth family (sample.py:510-537 and siblings): ``h_i = p_i x + q_i m`` enters a ring of ``order``
    slots; ``x_s = r_i x + g_i sum_j w_ij h_j`` with the weights of ``_weights`` (float64 solve, sample.py:487-508)."""
    order = int(sampler.order)
    if order < 1 or order > _lib.MAX_SLOTS:
        return None
    alpha, sigma = g.denoiser.schedule(g.time)
    u = sampler._variable(alpha, sigma)
    p, q = sampler._stored_coef(alpha[:-1], sigma[:-1])
    r, gain = sampler._update_coef(alpha[:-1], sigma[:-1], alpha[1:], sigma[1:])
    W = torch.zeros(g.T, _lib.MAX_SLOTS, dtype=torch.float32)
    gain_host = gain.to("cpu", torch.float64)
    for i in range(g.T):
        w = sampler._weights(u, i, order).to("cpu", torch.float64)  # oldest first: entries i + 1 - n .. i
        n = w.numel()
        for j in range(n):
            W[i, (i + 1 - n + j) % order] = float(gain_host[i] * w[j])
    flags = torch.tensor([F_HIST | ((i % order) << 4) | F_STORE | (order << 12) for i in range(g.T)], dtype=torch.int32)
    one = torch.ones_like(alpha[:-1])
    hist = dict(p=p * one, q=q * one, r=r * one, W=W.to(g.device), flags=flags)
    return g.finish(alpha[:-1].contiguous(), sigma[:-1].contiguous(), hist=hist, slots=order)
