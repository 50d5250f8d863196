"""Every sampler of azula/sample.py and the classifier-free-guidance wrapper: the host
mirror on CPU against fixtures produced by the unmodified reference (tests/golden/samplers.npz,
oracle/gen_golden_samplers.py)."""

import pytest
import torch

from conftest import load_golden
from oracle.gen_golden_cfg import SAMPLER_CASES, LabelMlp

import azula_b200.sample as S
from azula_b200.denoise import KarrasDenoiser
from azula_b200.guidance.cfg import CFGDenoiser
from azula_b200.nn.layers import SineEncoding
from azula_b200.noise import VPSchedule


class Mlp(torch.nn.Module):
    """The README-style backbone of the fixtures (reference tests/test_sample.py:28-51)."""

    def __init__(self, features=5):
        super().__init__()
        self.l1 = torch.nn.Linear(features, 64)
        self.l2 = torch.nn.Linear(64, features)
        self.enc = SineEncoding(64)

    def forward(self, x, t):
        return self.l2(torch.relu(self.l1(x) + self.enc(t)))


def denoiser(g, device="cpu"):
    net = Mlp()
    net.load_state_dict({k[2:]: v for k, v in g.items() if k.startswith("w_")})
    return KarrasDenoiser(net.to(device), VPSchedule()).eval()


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.mark.parametrize("tag", list(SAMPLER_CASES))
def test_sampler_matches_reference_bit_for_bit(tag):
    g = load_golden("samplers")
    name, kw = SAMPLER_CASES[tag]
    smp = getattr(S, name)(denoiser(g), silent=True, **kw)
    torch.manual_seed(0)
    x1 = smp.init((16, 5))
    assert torch.equal(x1, g[f"{tag}_x1"])
    torch.manual_seed(1)
    x0 = smp(x1)
    assert torch.equal(x0, g[f"{tag}_x0"]), (x0 - g[f"{tag}_x0"]).abs().max()


def test_multistep_weights():
    u = torch.linspace(2.0, 0.1, 9)
    w = S.zABSampler._weights(u, 0, 3)
    assert w.shape == (1,) and torch.allclose(w, u[1] - u[0])  # first step = Euler
    w = S.zABSampler._weights(u, 4, 2)  # uniform grid: the classic AB2 weights (-1/2, 3/2) h
    h = u[5] - u[4]
    assert torch.allclose(w, torch.stack((-0.5 * h, 1.5 * h)), atol=1e-6)
    assert S.zEABSampler._weights(u, 5, 3).dtype == torch.float32


def test_cfg_matches_reference():
    g = load_golden("samplers")
    net = LabelMlp(torch.nn.Module, torch)
    net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("cfg_w_")})
    den = CFGDenoiser(KarrasDenoiser(net, VPSchedule())).eval()
    assert den.schedule is den.denoiser.schedule
    pos, neg = {"label": torch.arange(8) % 3}, {"label": torch.zeros(8, dtype=torch.long)}
    mean = den(g["cfg_x"], torch.tensor(0.4), positive=pos, negative=neg, guidance=2.5).mean
    assert torch.equal(mean, g["cfg_mean"])
    x0 = S.DDIMSampler(den, steps=8, silent=True)(g["cfg_x"], positive=pos, negative=neg, guidance=1.5)
    assert torch.equal(x0, g["cfg_ddim_x0"])
