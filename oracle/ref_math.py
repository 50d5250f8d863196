"""Oracle (test infrastructure): schedule, preconditioning and DDPM/DDIM transition maths.

Functional restatement in plain torch of the reference arithmetic, in the reference's
own operation order so that results are comparable bit for bit when both sides see the
same backbone output and the same noise.  Citations are ``file:line`` in the reference
checkout (probabilists/azula @ bec12b8).  Never imported by product code.
"""

from __future__ import annotations

import math
import torch

from torch import Tensor
from typing import Callable

# ----------------------------------------------------------------------------- schedule


def vp_alpha_sigma(t: Tensor, alpha_min: float = 1e-3, sigma_min: float = 1e-3):
    """VP schedule: alpha = exp(log(alpha_min) t^2), sigma = sqrt(1 - alpha^2 + sigma_min^2).

    Follows azula/noise.py:122-129 (defaults at :118; ADM uses 1e-2/1e-2,
    azula/plugins/adm/__init__.py:59).
    """
    alpha = torch.exp(math.log(alpha_min) * t**2)
    sigma = torch.sqrt(1 - torch.exp(math.log(alpha_min) * t**2) ** 2 + sigma_min**2)
    return alpha, sigma


def time_grid(start: float, stop: float, steps: int) -> Tensor:
    """The (steps, 2) table of (t, s) pairs. Follows azula/sample.py:86-94,151."""
    ts = torch.linspace(start, stop, steps + 1)
    return torch.stack((ts[:-1], ts[1:]), dim=-1)


# ------------------------------------------------------------------------ preconditioning


def _expand(a: Tensor, ndim: int) -> Tensor:
    while a.ndim < ndim:
        a = a[..., None]
    return a


def karras_coefficients(alpha: Tensor, sigma: Tensor):
    """(c_in, c_out, c_skip, c_time) of the EDM-style preconditioner.

    Follows azula/denoise.py:309-312.
    """
    c_in = torch.rsqrt(alpha**2 + sigma**2)
    c_out = sigma * torch.rsqrt(alpha**2 + sigma**2)
    c_skip = alpha / (alpha**2 + sigma**2)
    c_time = torch.log(sigma / alpha)
    return c_in, c_out, c_skip, c_time


def karras_mean(
    backbone: Callable[..., Tensor],
    schedule: Callable[[Tensor], tuple[Tensor, Tensor]],
    x_t: Tensor,
    t: Tensor,
    backbone_dtype: torch.dtype | None = None,
    **kwargs,
) -> Tensor:
    """Posterior mean of KarrasDenoiser. Follows azula/denoise.py:304-324."""
    alpha, sigma = schedule(t)
    alpha, sigma = _expand(alpha, x_t.ndim), _expand(sigma, x_t.ndim)
    c_in, c_out, c_skip, c_time = karras_coefficients(alpha, sigma)
    c_time = c_time.reshape_as(t)
    dtype = backbone_dtype or x_t.dtype
    out = backbone((c_in * x_t).to(dtype), c_time.to(dtype), **kwargs).to(x_t)
    return c_skip * x_t + c_out * out


def adm_sigmas(discrete_steps: int = 1000, kind: str = "linear", dtype=torch.float32) -> Tensor:
    """Discrete noise levels sqrt(1 - cumprod(1 - beta)) of the ADM plugin.

    Follows azula/plugins/adm/__init__.py:66-84 (fp64 arithmetic, cast at the end).
    """
    if kind == "linear":
        beta = torch.linspace(
            0.1 / discrete_steps, 20.0 / discrete_steps, discrete_steps, dtype=torch.float64
        )
    elif kind == "cosine":
        u = torch.linspace(0, 1, discrete_steps + 1, dtype=torch.float64)
        bar = torch.cos((u + 0.008) / 1.008 * torch.pi / 2) ** 2
        beta = torch.clip(1 - bar[1:] / bar[:-1], max=0.999)
    else:
        raise ValueError(kind)
    return torch.sqrt(1 - torch.cumprod(1 - beta, dim=0)).to(dtype)


def adm_coefficients(alpha: Tensor, sigma: Tensor, sigmas: Tensor):
    """(c_in, c_out, c_skip, c_time:int64, c_var) of AblatedDenoiser.

    Follows azula/plugins/adm/__init__.py:109-114.
    """
    c_in = torch.rsqrt(alpha**2 + sigma**2)
    c_out = -sigma / alpha
    c_skip = 1 / alpha
    c_time = sigma * torch.rsqrt(alpha**2 + sigma**2)
    c_time = torch.searchsorted(sigmas, c_time.flatten())
    c_var = sigma**2 / (alpha**2 + sigma**2)
    return c_in, c_out, c_skip, c_time, c_var


def adm_mean_var(
    backbone: Callable[..., Tensor],
    schedule: Callable[[Tensor], tuple[Tensor, Tensor]],
    sigmas: Tensor,
    x_t: Tensor,
    t: Tensor,
    learn_var: bool = True,
    clip_mean: bool = True,
    label: Tensor | None = None,
    backbone_dtype: torch.dtype | None = None,
):
    """(mean, var) of AblatedDenoiser in eval mode.

    Follows azula/plugins/adm/__init__.py:104-136.
    """
    alpha, sigma = schedule(t)
    alpha, sigma = _expand(alpha, x_t.ndim), _expand(sigma, x_t.ndim)
    c_in, c_out, c_skip, c_time, c_var = adm_coefficients(alpha, sigma, sigmas)
    dtype = backbone_dtype or x_t.dtype
    out = backbone((c_in * x_t).to(dtype), c_time, y=label).to(x_t)
    if learn_var:
        out, log_var = torch.chunk(out, 2, dim=1)
        var = c_var * torch.exp(log_var)
    else:
        var = c_var
    mean = c_skip * x_t + c_out * out
    if clip_mean:
        mean = torch.clip(mean, min=-1.0, max=1.0)
    return mean, var


# ------------------------------------------------------------------------------ transition


def transition_scalars(alpha_t, sigma_t, alpha_s, sigma_s, eta: float | None):
    """tau and the two scalar factors of the DDPM (eta=None) / DDIM update.

    Follows azula/sample.py:208 and :252-253 (operation order preserved).
    Returns (k, n) with k = sigma_s*sqrt(1-tau)/sigma_t and n = sigma_s*sqrt(tau).
    """
    tau = 1 - (alpha_t / alpha_s * sigma_s / sigma_t) ** 2
    if eta is not None:
        tau = torch.clip(eta * tau, min=0, max=1)
    k = sigma_s * torch.sqrt(1 - tau) / sigma_t
    n = sigma_s * torch.sqrt(tau)
    return k, n


def transition(x_t: Tensor, mean: Tensor, eps: Tensor, alpha_t, sigma_t, alpha_s, sigma_s, eta):
    """x_s = alpha_s m + k (x_t - alpha_t m) + n eps.  Follows azula/sample.py:212-214, :257-259."""
    k, n = transition_scalars(alpha_t, sigma_t, alpha_s, sigma_s, eta)
    x_s = alpha_s * mean
    x_s = x_s + k * (x_t - alpha_t * mean)
    x_s = x_s + n * eps
    return x_s


def init_noise(shape, eps: Tensor, alpha_T, sigma_T, mean=0.0, var=1.0) -> Tensor:
    """x_T = alpha_T mean + sqrt(alpha_T^2 var + sigma_T^2) eps. Follows azula/sample.py:120-128."""
    mean_T = alpha_T * mean
    std_T = torch.sqrt(alpha_T**2 * var + sigma_T**2)
    return mean_T.expand(shape) + std_T.expand(shape) * eps


def sample_loop(
    mean_fn: Callable[[Tensor, Tensor], Tensor],
    schedule: Callable[[Tensor], tuple[Tensor, Tensor]],
    x: Tensor,
    steps: int,
    eta: float | None,
    noise_fn: Callable[[Tensor], Tensor] | None = None,
    start: float = 1.0,
    stop: float = 0.0,
    trace: list | None = None,
) -> Tensor:
    """The reverse process from t_T to t_0.  Follows azula/sample.py:151-161.

    ``mean_fn(x_t, t)`` is the posterior mean, ``noise_fn(x_t)`` the N(0, I) draw
    (``torch.randn_like`` in the reference, azula/sample.py:214,259).
    """
    noise_fn = noise_fn or torch.randn_like
    for t, s in time_grid(start, stop, steps).to(x.device).unbind():
        alpha_s, sigma_s = schedule(s)
        alpha_t, sigma_t = schedule(t)
        m = mean_fn(x, t)
        eps = noise_fn(x)
        x = transition(x, m, eps, alpha_t, sigma_t, alpha_s, sigma_s, eta)
        if trace is not None:
            trace.append(x.clone())
    return x


# --------------------------------------------------------------------------- small backbones


def sine_encoding(t: Tensor, features: int, omega: float = 1e4) -> Tensor:
    """Sinusoidal features of azula.nn.layers.SineEncoding. Follows azula/nn/layers.py:286-299."""
    x = t[..., None]
    freqs = torch.linspace(0, 1, features // 2, dtype=x.dtype, device=x.device)
    freqs = torch.exp(math.log(1 / omega) * freqs)
    return torch.cat((torch.sin(x * freqs), torch.cos(x * freqs)), dim=-1)


def mlp_backbone(state: dict[str, Tensor], x: Tensor, t: Tensor, **_) -> Tensor:
    """The 5-feature MLP of BASELINE config 1. Follows tests/test_sample.py:28-51."""
    f = state["l1.weight"].shape[0]
    y = torch.nn.functional.linear(x, state["l1.weight"], state["l1.bias"])
    y = y + sine_encoding(t, f)
    y = torch.relu(y)
    return torch.nn.functional.linear(y, state["l2.weight"], state["l2.bias"])
