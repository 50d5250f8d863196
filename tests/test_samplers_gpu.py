"""The widened sampler family on the GPU: every transition that is affine in
(x_t, mean, eps) runs as one azb_step_f32 launch; results are held to the north-star fp32 tolerance
(rtol 1e-3 / atol 1e-5, BASELINE.json) against (i) the reference fixtures for deterministic samplers and (ii) the
plain torch execution model on the same device with the same seed for the stochastic ones (the kernel's Philox
stream equals torch.randn_like's, so the noise bits are identical)."""

import pytest
import torch

from conftest import close, load_golden
from oracle.gen_golden_cfg import SAMPLER_CASES, LabelMlp
from test_samplers_cpu import denoiser

import azula_b200.sample as S
from azula_b200 import engine
from azula_b200.denoise import KarrasDenoiser
from azula_b200.guidance.cfg import CFGDenoiser
from azula_b200.noise import VPSchedule

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        yield


@pytest.mark.parametrize("tag", list(SAMPLER_CASES))
def test_sampler_on_gpu(tag):
    g = load_golden("samplers")
    name, kw = SAMPLER_CASES[tag]
    smp = getattr(S, name)(denoiser(g, DEV), silent=True, **kw)
    x1 = g[f"{tag}_x1"].to(DEV)
    torch.manual_seed(1)
    x0 = smp(x1)
    assert x0.shape == x1.shape and torch.isfinite(x0).all()
    with engine.eager_torch():  # same device, same seed, reference execution model
        torch.manual_seed(1)
        eager = smp(x1)
    assert close(x0, eager, rtol=1e-3, atol=2e-5), (x0 - eager).abs().max()
    if tag not in ("ito", "pc"):  # deterministic: comparable with the CPU fixture of the reference
        ref = g[f"{tag}_x0"].to(DEV)
        assert close(x0, ref, rtol=1e-3, atol=1e-4), (x0 - ref).abs().max()


def test_cfg_on_gpu():
    g = load_golden("samplers")
    net = LabelMlp(torch.nn.Module, torch)
    net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("cfg_w_")})
    den = CFGDenoiser(KarrasDenoiser(net.to(DEV), VPSchedule())).eval()
    pos = {"label": (torch.arange(8) % 3).to(DEV)}
    neg = {"label": torch.zeros(8, dtype=torch.long, device=DEV)}
    x0 = S.DDIMSampler(den, steps=8, silent=True)(g["cfg_x"].to(DEV), positive=pos, negative=neg, guidance=1.5)
    assert close(x0, g["cfg_ddim_x0"].to(DEV), rtol=1e-3, atol=1e-4)
