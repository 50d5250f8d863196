"""Oracle (test infrastructure): the remaining samplers of ``azula/sample.py`` as plain functions.

Functional restatement in the reference's own operation order (so a CPU run reproduces the reference fixtures
``tests/golden/samplers.npz`` bit for bit -- ``tests/test_oracle_golden.py`` pins that) of Euler ``:290-303``,
Heun ``:337-352``, Ito ``:417-431``, predictor-corrector ``:980-999`` and the Adams-Bashforth family zAB
``:487-537``, vAB ``:573-598``, zEAB ``:622-699``, xEAB ``:764-801``, REAB ``:884-950``.  ``mean_fn(x, t)`` is the
denoiser's posterior mean, ``schedule(t) -> (alpha, sigma)``, ``noise_fn`` the N(0, I) draw.  Never imported by
product code.
"""

from __future__ import annotations

import math
import torch

from .ref_math import time_grid


def _slope(alpha_t, sigma_t, alpha_s, sigma_s):
    return alpha_s * (sigma_s / alpha_s - sigma_t / alpha_t)


def euler_step(mean_fn, x, t, s, sch, **_):
    (a_s, s_s), (a_t, s_t) = sch(s), sch(t)
    z = (x - a_t * mean_fn(x, t)) / s_t
    return a_s / a_t * x + _slope(a_t, s_t, a_s, s_s) * z


def heun_step(mean_fn, x, t, s, sch, **_):
    (a_s, s_s), (a_t, s_t) = sch(s), sch(t)
    z_t = (x - a_t * mean_fn(x, t)) / s_t
    x_s = a_s / a_t * x + _slope(a_t, s_t, a_s, s_s) * z_t
    z_s = (x_s - a_s * mean_fn(x_s, s)) / s_s
    z_t = (z_t + z_s) / 2
    return a_s / a_t * x + _slope(a_t, s_t, a_s, s_s) * z_t


def ito_step(mean_fn, x, t, s, sch, noise_fn, eta=1.0, temperature=1.0, **_):
    (a_s, s_s), (a_t, s_t) = sch(s), sch(t)
    m = mean_fn(x, t)
    x_s = a_s / a_t * x
    x_s = x_s + (1 + eta**2) / temperature * (s_s / s_t - a_s / a_t) * (x - a_t * m)
    x_s = x_s + eta * a_s * torch.sqrt(torch.abs((s_t / a_t) ** 2 - (s_s / a_s) ** 2)) * noise_fn(x_s)
    return x_s


def pc_step(mean_fn, x, t, s, sch, noise_fn, corrections=1, delta=0.01, **_):
    (a_s, s_s), (a_t, s_t) = sch(s), sch(t)
    for _ in range(corrections):
        m = mean_fn(x, t)
        x = a_t * m + math.sqrt(1 - delta) * (x - a_t * m) + math.sqrt(delta) * s_t * noise_fn(x)
    m = mean_fn(x, t)
    return a_s * m + s_s / s_t * (x - a_t * m)


ONE_STEP = {"EulerSampler": euler_step, "HeunSampler": heun_step, "ItoSampler": ito_step, "PCSampler": pc_step}


def one_step_loop(name, mean_fn, sch, x, steps, noise_fn=torch.randn_like, start=1.0, stop=0.0, **params):
    """``Sampler.__call__`` (sample.py:151-161) around one of the single-step rules."""
    for t, s in time_grid(start, stop, steps).to(x.device).unbind():
        x = ONE_STEP[name](mean_fn, x, t, s, sch, noise_fn=noise_fn, **params)
    return x


# ------------------------------------------------------------------------------------- multi-step family


def _solve(u, i, n, moments):
    """Vandermonde system of the last min(n, i + 1) nodes, in float64 (sample.py:487-508)."""
    u = u.to(torch.float64)
    n = min(n, i + 1)
    k = torch.arange(n, device=u.device)
    V = u[i + 1 - n : i + 1] ** k[:, None]
    return torch.linalg.solve(V, moments(u, i, k))


def _poly(u, i, k):  # int v^k dv
    return u[i + 1] ** (k + 1) / (k + 1) - u[i] ** (k + 1) / (k + 1)


def _exp_plus(u, i, k):  # int e^v v^k dv
    kf = torch.cumprod(torch.clip(k, min=1), dim=0)
    return (-1) ** k * kf * (torch.exp(u[i + 1]) * torch.cumsum((-u[i + 1]) ** k / kf, dim=0)
                             - torch.exp(u[i]) * torch.cumsum((-u[i]) ** k / kf, dim=0))


def _exp_minus(u, i, k):  # int e^-v v^k dv
    kf = torch.cumprod(torch.clip(k, min=1), dim=0)
    return -kf * (torch.exp(-u[i + 1]) * torch.cumsum(u[i + 1] ** k / kf, dim=0)
                  - torch.exp(-u[i]) * torch.cumsum(u[i] ** k / kf, dim=0))


def _rosenbrock(u, i, k):  # int e^v / (1 + e^2v) v^k dv, trapezoidal rule on 257 nodes
    v = torch.linspace(u[i], u[i + 1], steps=256 + 1, dtype=u.dtype, device=u.device)
    return torch.trapezoid(torch.exp(v) / (1 + torch.exp(2 * v)) * (v ** k[:, None]), v, dim=-1)


def _reab_stored(x, m, a, s):
    a_t = s**2 / (a**2 + s**2)
    b_t = s * torch.rsqrt(a**2 + s**2)
    return (1 - a_t) / b_t / a * x - 1 / b_t * m


MULTISTEP = {
    # name: (variable u(alpha, sigma), moments, stored(x, m, alpha_t, sigma_t), update(x, I, a_t, s_t, a_s, s_s))
    "zABSampler": (lambda a, s: s / a, _poly, lambda x, m, a, s: (x - a * m) / s,
                   lambda x, I, a_t, s_t, a_s, s_s: a_s / a_t * x + a_s * I),
    "vABSampler": (lambda a, s: s / (a + s), _poly, lambda x, m, a, s: 1 / s * x - (1 + a / s) * m,
                   lambda x, I, a_t, s_t, a_s, s_s: (a_s + s_s) / (a_t + s_t) * x + (a_s + s_s) * I),
    "zEABSampler": (lambda a, s: s.log() - a.log(), _exp_plus, lambda x, m, a, s: (x - a * m) / s,
                    lambda x, I, a_t, s_t, a_s, s_s: a_s / a_t * x + a_s * I),
    "xEABSampler": (lambda a, s: s.log() - a.log(), _exp_minus, lambda x, m, a, s: m,
                    lambda x, I, a_t, s_t, a_s, s_s: s_s / s_t * x - s_s * I),
    # the second square root mixes alpha_s with sigma_t, as the reference does (sample.py:944)
    "REABSampler": (lambda a, s: s.log() - a.log(), _rosenbrock, _reab_stored,
                    lambda x, I, a_t, s_t, a_s, s_s: torch.sqrt((a_s**2 + s_s**2) / (a_t**2 + s_t**2)) * x
                    + torch.sqrt(a_s**2 + s_t**2) * I),
}


def multistep_loop(name, mean_fn, sch, x, steps, order=2, start=1.0, stop=0.0):
    variable, moments, stored, update = MULTISTEP[name]
    time = torch.linspace(start, stop, steps + 1).to(x.device)
    alpha, sigma = sch(time)
    u = variable(alpha, sigma)
    past = []
    for i, t in enumerate(time[:-1]):
        past.append(stored(x, mean_fn(x, t), alpha[i], sigma[i]))
        past = past[-order:]
        w = _solve(u, i, order, moments).to(u.dtype)
        integral = sum(h * c for h, c in zip(past, w, strict=True))
        x = update(x, integral, alpha[i], sigma[i], alpha[i + 1], sigma[i + 1])
    return x


def sample(name, mean_fn, sch, x, steps, noise_fn=torch.randn_like, **params):
    """Dispatch by the reference's class name."""
    if name in MULTISTEP:
        return multistep_loop(name, mean_fn, sch, x, steps, **params)
    return one_step_loop(name, mean_fn, sch, x, steps, noise_fn=noise_fn, **params)
