"""CPU tests of the host-side mirror of the reference interface (no GPU, no kernels).

Modelled on the reference's tests/test_sample.py, tests/test_noise.py and tests/test_denoise.py,
plus bit-exact comparisons with the reference-generated fixtures in tests/golden/.
"""

import pytest
import torch

from conftest import load_golden
from functools import partial

from azula_b200.denoise import DiracPosterior, GaussianPosterior, KarrasDenoiser, SimpleDenoiser
from azula_b200.noise import (
    CosineSchedule,
    DecaySchedule,
    RectifiedSchedule,
    Schedule,
    VESchedule,
    VPSchedule,
)
from azula_b200.sample import DDIMSampler, DDPMSampler


class Mlp(torch.nn.Module):
    """The backbone of BASELINE config 1 (reference tests/test_sample.py:28-51)."""

    def __init__(self, features=5, with_label=False):
        super().__init__()
        self.with_label = with_label
        self.l1 = torch.nn.Linear(features, 64)
        self.l2 = torch.nn.Linear(64, features)

    def forward(self, x, t, label=None):
        freqs = torch.exp(torch.linspace(0, 1, 32, dtype=t.dtype, device=t.device) * -9.210340371976184)
        enc = torch.cat((torch.sin(t[..., None] * freqs), torch.cos(t[..., None] * freqs)), dim=-1)
        if self.with_label:
            assert isinstance(label, str)
        else:
            assert label is None
        return self.l2(torch.relu(self.l1(x) + enc))


@pytest.mark.parametrize("with_label", [False, True])
@pytest.mark.parametrize("batch", [(), (64,)])
def test_samplers_shapes(with_label, batch):
    den = KarrasDenoiser(Mlp(5, with_label), VPSchedule())
    for S in (DDPMSampler, partial(DDIMSampler, eta=0.0), partial(DDIMSampler, eta=1.0)):
        smp = S(den, steps=64, silent=True)
        x1 = smp.init((*batch, 5))
        assert x1.shape == (*batch, 5) and torch.isfinite(x1).all()
        x0 = smp(x1, label="cat") if with_label else smp(x1)
        assert x0.shape == (*batch, 5) and torch.isfinite(x0).all()


@pytest.mark.parametrize("S", [VPSchedule, VESchedule, CosineSchedule, RectifiedSchedule, DecaySchedule])
def test_schedules(S):
    sch = S()
    assert isinstance(sch, Schedule)
    t = torch.rand(17).sort().values
    a, s = sch(t)
    assert a.shape == t.shape and s.shape == t.shape
    assert (a > 0).all() and (s > 0).all()
    assert ((a[:-1] / s[:-1]) >= (a[1:] / s[1:])).all()
    assert sch(torch.tensor(0.0))[0] == 1.0


def test_vp_matches_reference_bits():
    g = load_golden("schedule")
    a, s = VPSchedule()(g["t"])
    assert torch.equal(a, g["alpha_default"]) and torch.equal(s, g["sigma_default"])
    a, s = VPSchedule(1e-2, 1e-2)(g["t"])
    assert torch.equal(a, g["alpha_adm"]) and torch.equal(s, g["sigma_adm"])


def _golden_denoiser(g):
    net = Mlp()
    net.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("w_")})
    return KarrasDenoiser(net, VPSchedule()).eval()


def test_karras_forward_matches_reference_bits():
    g = load_golden("mlp_karras")
    den = _golden_denoiser(g)
    with torch.no_grad():
        q = den(g["x"], g["t"])
        assert isinstance(q, DiracPosterior)
        assert torch.equal(q.mean, g["mean_batched_t"])
        assert torch.equal(den(g["x"], torch.tensor(0.37)).mean, g["mean_scalar_t"])


def test_cpu_sampling_matches_reference_bits():
    """BASELINE config 1 (DDPM-1000 on the 5-feature MLP, CPU) and DDIM variants."""
    g = load_golden("mlp_karras")
    den = _golden_denoiser(g)
    cases = {
        "ddpm1000": DDPMSampler(den, steps=1000, silent=True),
        "ddim64_eta0": DDIMSampler(den, steps=64, eta=0.0, silent=True),
        "ddim64_eta1": DDIMSampler(den, steps=64, eta=1.0, silent=True),
        "ddim16_eta05_partial": DDIMSampler(den, steps=16, eta=0.5, start=0.8, stop=0.1, silent=True),
    }
    for name, smp in cases.items():
        torch.manual_seed(0)
        x1 = smp.init((64, 5))
        assert torch.equal(x1, g[f"{name}_x1"]), name
        keep = x1.clone()
        x0 = smp(x1)
        assert torch.equal(x1, keep)  # the caller's tensor is never mutated
        assert torch.equal(x0, g[f"{name}_x0"]), name


@pytest.mark.parametrize("D", [SimpleDenoiser, KarrasDenoiser])
@pytest.mark.parametrize("S", [VPSchedule, RectifiedSchedule])
def test_reschedule_invariance(D, S):
    """reference tests/test_denoise.py:135-143: mean is invariant to (alpha, sigma) -> (1, sigma/alpha)."""

    class Rescaled(Schedule):
        def __init__(self, inner):
            self.inner = inner

        def __call__(self, t):
            a, s = self.inner(t)
            return torch.ones_like(a), s / a

    net = Mlp()
    sch = S()
    x, t = torch.randn(32, 5), torch.rand(32)
    a, _ = sch(t)
    with torch.no_grad():
        m1 = D(net, sch)(x, t).mean
        m2 = D(net, Rescaled(sch))(x / a[:, None], t).mean
    assert torch.allclose(m1, m2, atol=1e-6 if D is SimpleDenoiser else 1e-5)


def test_loss_differentiable():
    den = KarrasDenoiser(Mlp(), VPSchedule())
    loss = den.loss(torch.randn(16, 5), torch.rand(16))
    assert loss.shape == ()
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in den.parameters())


def test_gaussian_posterior_log_prob():
    m, v, x = torch.randn(7), torch.rand(7) + 0.1, torch.randn(7)
    ref = torch.distributions.Normal(m, v.sqrt()).log_prob(x)
    assert torch.allclose(GaussianPosterior(m, v).log_prob(x), ref, atol=1e-6)


def test_overridden_step_is_respected():
    calls = []

    class Custom(DDIMSampler):
        def step(self, x_t, t, s, **kw):
            calls.append(float(t))
            return super().step(x_t, t, s, **kw)

    den = KarrasDenoiser(Mlp(), VPSchedule())
    Custom(den, steps=8, silent=True)(torch.randn(4, 5))
    assert len(calls) == 8
