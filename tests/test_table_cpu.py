"""The per-stage coefficient tables of EVERY sampler and of classifier-free guidance (f2),
executed by the CPU restatement of the transition kernel (tests/rowsim.py), against fixtures produced by the
unmodified reference (tests/golden/samplers.npz): DDPM / DDIM bit for bit (the rows keep the reference's operation
order), the collapsed affine / history forms within the north-star fp32 tolerance (rtol 1e-3, atol 1e-5)."""

import pytest
import torch

from conftest import close, load_golden
from oracle.gen_golden_cfg import SAMPLER_CASES, LabelMlp
from rowsim import run_table
from test_samplers_cpu import denoiser

import azula_b200.sample as S
from azula_b200 import _lib
from azula_b200.denoise import KarrasDenoiser
from azula_b200.engine import table as T
from azula_b200.guidance.cfg import CFGDenoiser
from azula_b200.noise import VPSchedule


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.mark.parametrize("tag", list(SAMPLER_CASES))
def test_table_reproduces_reference_sampler(tag):
    g = load_golden("samplers")
    name, kw = SAMPLER_CASES[tag]
    smp = getattr(S, name)(denoiser(g), silent=True, **kw)
    assert smp._fusable(g[f"{tag}_x1"]), "every stock sampler has a table"
    torch.manual_seed(1)
    x0 = run_table(smp, g[f"{tag}_x1"])
    ref = g[f"{tag}_x0"]
    if name in ("DDPMSampler", "DDIMSampler"):
        assert torch.equal(x0, ref), (x0 - ref).abs().max()
    else:
        assert close(x0, ref, rtol=1e-3, atol=1e-5), (x0 - ref).abs().max()


def test_stage_structure():
    g = load_golden("samplers")
    den = denoiser(g)
    grid = lambda smp: T.Grid(smp, torch.device("cpu"))  # noqa: E731
    tab = S.HeunSampler(den, steps=6)._table(grid(S.HeunSampler(den, steps=6)))
    assert (tab.steps, tab.per_step, tab.slots, tab.alt, tab.draws) == (12, 2, 1, True, 0)
    smp = S.PCSampler(den, steps=5, corrections=2)
    tab = smp._table(grid(smp))
    assert (tab.steps, tab.per_step, tab.draws) == (15, 3, 10)
    draw = tab.coef.view(torch.int32)[:, _lib.R_DRAW].tolist()
    assert draw[:7] == [0, 1, 2, 2, 3, 4, 4]  # the predictor stages draw nothing
    smp = S.zABSampler(den, steps=5, order=3)
    tab = smp._table(grid(smp))
    flags = tab.coef.view(torch.int32)[:, _lib.R_FLAGS].tolist()
    assert [(f >> 4) & 15 for f in flags] == [0, 1, 2, 0, 1] and tab.slots == 3
    W = tab.coef[:, _lib.R_W : _lib.R_W + 3]
    assert (W[0, 1:] == 0).all() and (W[1, 2] == 0) and (W[3] != 0).all()  # warm-up orders 1, 2, then 3
    assert S.zABSampler(den, steps=5, order=9)._table(grid(smp)) is None  # more slots than the kernel has


def test_user_subclasses_take_the_generic_path():
    g = load_golden("samplers")
    den = denoiser(g)

    class MyEuler(S.EulerSampler):
        def step(self, x_t, t, s, **kwargs):
            return super().step(x_t, t, s, **kwargs)

    class MyAB(S.zABSampler):
        def _stored(self, x_t, mean, alpha, sigma, i):
            return super()._stored(x_t, mean, alpha, sigma, i)

    class Harmless(S.vABSampler):
        pass

    x = g["euler_x1"]
    assert not MyEuler(den)._fusable(x) and not MyAB(den)._fusable(x) and Harmless(den)._fusable(x)


def test_cfg_table_matches_reference():
    g = load_golden("samplers")
    net = LabelMlp(torch.nn.Module, torch)
    net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("cfg_w_")})
    den = CFGDenoiser(KarrasDenoiser(net, VPSchedule())).eval()
    assert den.fusable() and T.inner_denoiser(den) is den.denoiser
    pos, neg = {"label": torch.arange(8) % 3}, {"label": torch.zeros(8, dtype=torch.long)}
    smp = S.DDIMSampler(den, steps=8, silent=True)
    x0 = run_table(smp, g["cfg_x"], guidance=1.5, positive=pos, negative=neg)
    assert torch.equal(x0, g["cfg_ddim_x0"]), (x0 - g["cfg_ddim_x0"]).abs().max()


def test_loop_signature_follows_weights_schedule_and_hyperparameters():
    """the key of a cached fused loop must change whenever something its graph or table froze may
    have changed -- parameters (in place or replaced), buffers, schedule attributes, sampler hyper-parameters --
    and must NOT change otherwise (no re-capture per call, none for another guidance strength)."""
    from azula_b200.engine import loop as L

    g = load_golden("samplers")
    den = denoiser(g)
    smp = S.ItoSampler(den, steps=8, eta=0.5)
    x = g["ito_x1"]
    key = L.signature(smp, x, {})
    assert key == L.signature(smp, x.clone(), {}) and hash(key) is not None
    den.backbone.l1.weight.data.add_(1.0)  # does not bump _version, but ...
    with torch.no_grad():
        den.backbone.l1.weight.add_(1.0)  # ... an optimiser step does
    k2 = L.signature(smp, x, {})
    assert k2 != key
    den.backbone.load_state_dict({k: v.clone() for k, v in den.backbone.state_dict().items()})
    k3 = L.signature(smp, x, {})
    assert k3 != k2
    den.schedule.sigma_min = 5e-3
    k4 = L.signature(smp, x, {})
    assert k4 != k3
    smp.eta = 0.6
    assert L.signature(smp, x, {}) != k4
    den.eval() if den.training else den.train()
    assert L.signature(smp, x, {}) != k4

    net = LabelMlp(torch.nn.Module, torch)
    cfg = CFGDenoiser(KarrasDenoiser(net, VPSchedule())).eval()
    smp = S.DDIMSampler(cfg, steps=4)
    pos, neg = {"label": torch.arange(8) % 3}, {"label": torch.zeros(8, dtype=torch.long)}
    xs = g["cfg_x"]
    a = L.signature(smp, xs, dict(positive=pos, negative=neg, guidance=1.5))
    assert a == L.signature(smp, xs, dict(positive=neg, negative=pos, guidance=7.0))
    assert a != L.signature(smp, xs, dict(positive=pos, guidance=1.5))
