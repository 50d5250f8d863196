r"""Ablated diffusion model (ADM) plugin (interface of ``azula/plugins/adm/__init__.py``).

.. code-block:: python

    from azula_b200.plugins import adm

    denoiser = adm.load_model("imagenet_256x256")     # needs the checkpoint in the hub cache
    denoiser = adm.make_model(**adm.cards()["imagenet_256x256"].config)  # random init

References:
    | Diffusion Models Beat GANs on Image Synthesis (Dhariwal et al., 2021)
    | https://arxiv.org/abs/2105.05233
"""

from __future__ import annotations

__all__ = ["AblatedDenoiser", "load_model", "make_model", "cards", "seed_parameters"]

import torch
import torch.nn as nn

from collections.abc import Sequence
from torch import Tensor

from ...denoise import Coefficients, Denoiser, GaussianPosterior, Preconditioned, _unsqueeze_like
from ...hub import download
from ...nn.utils import get_module_dtype, skip_init
from ...noise import Schedule, VPSchedule
from ..utils import load_cards
from . import unet


def _discrete_sigmas(kind: str, steps: int) -> Tensor:
    r""":math:`\sqrt{1 - \bar\alpha_i}` of the discrete training schedule, in float64
    (``azula/plugins/adm/__init__.py:66-84``)."""
    if kind == "linear":
        beta = torch.linspace(0.1 / steps, 20.0 / steps, steps, dtype=torch.float64)
    elif kind == "cosine":
        u = torch.linspace(0, 1, steps + 1, dtype=torch.float64)
        bar = torch.cos((u + 0.008) / 1.008 * torch.pi / 2) ** 2
        beta = torch.clip(1 - bar[1:] / bar[:-1], max=0.999)
    else:
        raise ValueError(f"Unknown discrete schedule '{kind}'.")
    return torch.sqrt(1 - torch.cumprod(1 - beta, dim=0))


class AblatedDenoiser(Preconditioned):
    r"""Denoiser around an :math:`\varepsilon`-prediction network trained on a discrete schedule
    (``azula/plugins/adm/__init__.py:32-136``).

    .. math:: c_\mathrm{in} = \frac{1}{\sqrt{\alpha_t^2 + \sigma_t^2}} \quad
        c_\mathrm{out} = -\frac{\sigma_t}{\alpha_t} \quad c_\mathrm{skip} = \frac{1}{\alpha_t} \quad
        c_\mathrm{time} = \min \{ i : \bar\sigma_i \geq \sigma_t c_\mathrm{in} \}

    Arguments:
        backbone: A time conditional network, called as :py:`backbone(x, timesteps, y=label)`.
        schedule: A noise schedule. If :py:`None`, :py:`VPSchedule(alpha_min=1e-2, sigma_min=1e-2)`.
        clip_mean: Whether the mean is clipped to :math:`[-1, 1]` during evaluation.
        learn_var: Whether the backbone also outputs a log-variance (second half of channels).
        discrete_schedule: The discrete training schedule, :py:`"linear"` or :py:`"cosine"`.
        discrete_steps: The number of discrete training steps.
    """

    def __init__(
        self,
        backbone: nn.Module,
        schedule: Schedule | None = None,
        clip_mean: bool = False,
        learn_var: bool = False,
        discrete_schedule: str = "linear",
        discrete_steps: int = 1000,
    ) -> None:
        super().__init__(backbone, VPSchedule(alpha_min=1e-2, sigma_min=1e-2) if schedule is None else schedule)

        self.clip_mean = clip_mean
        self.learn_var = learn_var

        self.register_buffer("sigmas", _discrete_sigmas(discrete_schedule, discrete_steps).to(torch.get_default_dtype()))

    # ---- what the fused sampling loop needs to know (azula_b200.engine.table / .loop)
    def coefficients(self, alpha_t: Tensor, sigma_t: Tensor) -> Coefficients:
        c_time = sigma_t * torch.rsqrt(alpha_t**2 + sigma_t**2)
        return Coefficients(
            c_in=torch.rsqrt(alpha_t**2 + sigma_t**2),
            c_out=-sigma_t / alpha_t,
            c_skip=1 / alpha_t,
            c_time=torch.searchsorted(self.sigmas, c_time.flatten()),
        )

    def time_input(self, c_time: Tensor, t: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time  # int64 indices, shape (1,) or (B,)

    def time_rows(self, c_time: Tensor, dtype: torch.dtype | None) -> Tensor:
        return c_time.reshape(-1, 1)

    def call_backbone(self, x_in: Tensor, time: Tensor, label: Tensor | None = None, **kwargs) -> Tensor:
        return self.backbone(x_in, time, y=label, **kwargs)

    def output_select(self) -> int | None:
        r"""The posterior mean reads only the first half of the output channels when the variance is learned."""
        return 1 if self.learn_var else None

    def mean_clip(self) -> float | None:
        return 1.0 if (self.clip_mean and not self.training) else None

    def fusable(self) -> bool:
        return type(self).forward is AblatedDenoiser.forward

    # ---- eager interface
    def forward(self, x_t: Tensor, t: Tensor, label: Tensor | None = None, **kwargs) -> GaussianPosterior:
        r"""
        Arguments:
            x_t: A noisy tensor :math:`x_t`, with shape :math:`(B, 3, H, W)`.
            t: The time :math:`t`, with shape :math:`()` or :math:`(B)`.
            label: The class label :math:`c` as an integer, with shape :math:`(B)`.

        Returns:
            The Gaussian :math:`\mathcal{N}(X \mid \mu_\phi(x_t \mid c), \sigma^2_\phi(x_t \mid c))`.
        """
        alpha_t, sigma_t = self.schedule(t)
        alpha_t, sigma_t = _unsqueeze_like(alpha_t, x_t.ndim), _unsqueeze_like(sigma_t, x_t.ndim)

        c = self.coefficients(alpha_t, sigma_t)
        c_var = sigma_t**2 / (alpha_t**2 + sigma_t**2)

        dtype = get_module_dtype(self.backbone)
        output = self.call_backbone((c.c_in * x_t).to(dtype), c.c_time, label=label, **kwargs).to(x_t)

        if self.learn_var:
            output, log_var = torch.chunk(output, 2, dim=1)
            var = c_var * torch.exp(log_var)
        else:
            var = c_var
        mean = c.c_skip * x_t + c.c_out * output

        if not self.training and self.clip_mean:
            mean = torch.clip(mean, min=-1.0, max=1.0)

        return GaussianPosterior(mean=mean, var=var)


def cards():
    r"""The model cards of this plugin (``cards.yaml``)."""
    return load_cards(__name__)


def load_model(name: str, **kwargs) -> Denoiser:
    r"""Loads a pre-trained ADM denoiser (``azula/plugins/adm/__init__.py:139-161``).

    Arguments:
        name: The pre-trained model name.
        kwargs: Keyword arguments passed to :func:`torch.load`.
    """
    kwargs.setdefault("map_location", "cpu")
    kwargs.setdefault("weights_only", True)

    card = load_cards(__name__)[name]
    state = torch.load(download(card.url, hash_prefix=card.hash), **kwargs)

    with skip_init():
        denoiser = make_model(**card.config)

    denoiser.backbone.load_state_dict(state)

    return denoiser.eval()


def make_model(
    # Denoiser
    clip_mean: bool = True,
    learn_var: bool = True,
    # Discrete schedule
    discrete_schedule: str = "linear",
    discrete_steps: int = 1000,
    # Data
    image_channels: int = 3,
    image_size: int = 64,
    # Backbone
    attention_resolutions: Sequence[int] = (32, 16, 8),
    channel_mult: Sequence[int] = (1, 2, 3, 4),
    num_channels: int = 128,
    num_classes: int | None = None,
    **kwargs,
) -> Denoiser:
    r"""Initializes an ADM denoiser (``azula/plugins/adm/__init__.py:164-202``): attention is
    placed at the downsampling rates ``image_size // r``."""
    backbone = unet.UNetModel(
        image_size=image_size,
        in_channels=image_channels,
        out_channels=2 * image_channels if learn_var else image_channels,
        model_channels=num_channels,
        channel_mult=channel_mult,
        num_classes=num_classes,
        attention_resolutions={image_size // r for r in attention_resolutions},
        **kwargs,
    )

    return AblatedDenoiser(
        backbone,
        clip_mean=clip_mean,
        learn_var=learn_var,
        discrete_schedule=discrete_schedule,
        discrete_steps=discrete_steps,
    )


@torch.no_grad()
def seed_parameters(backbone: nn.Module, seed: int = 1234) -> nn.Module:
    r"""Overwrites EVERY parameter from a seeded generator on the parameter's own device.

    Pre-trained weights need a network; a default-initialised ADM outputs exactly zero (its last
    convolutions are zero-initialised) and ``skip_init`` leaves memory uninitialised, so synthetic
    benchmarks and tests must overwrite all of them: matrices and filters
    :math:`\mathcal{N}(0, 0.7^2 / \mathrm{fan\_in})`, norm gains :math:`1 + \mathcal{N}(0, 0.1^2)`,
    biases :math:`\mathcal{N}(0, 0.02^2)`.
    """
    gens: dict = {}
    for name, p in sorted(backbone.named_parameters()):
        g = gens.get(p.device)
        if g is None:
            g = gens[p.device] = torch.Generator(device=p.device).manual_seed(seed)
        r = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
        if p.ndim >= 2:
            r *= 0.7 / p[0].numel() ** 0.5
        elif name.endswith("weight"):
            r = 1 + 0.1 * r
        else:
            r *= 0.02
        p.copy_(r)
    return backbone
