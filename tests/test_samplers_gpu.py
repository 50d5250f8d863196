"""The widened sampler family on the GPU.  EVERY sampler of azula/sample.py runs as graph
replays of [backbone + ONE azb_step_ex_f32 launch + azb_advance] per stage -- no per-step ATen dispatch -- and is held
to the north-star fp32 tolerance (rtol 1e-3 / atol 1e-5 ... 1e-4, BASELINE.json) against

  (i)  the ORACLE (oracle/ref_samplers.py, itself pinned bit for bit to the unmodified reference on CPU) executed on
       the same device with the same generator state: the kernel's Philox stream equals torch.randn_like's, so the
       stochastic samplers (Ito, predictor-corrector) see identical noise bits;
  (ii) the reference fixtures (CPU run of the unmodified reference) for the deterministic ones.
"""

import pytest
import torch

from conftest import close, load_golden
from oracle import ref_math as RM
from oracle import ref_samplers as RS
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


def _oracle_mean(g):
    sd = {k[2:]: v.to(DEV) for k, v in g.items() if k.startswith("w_")}
    net = lambda x, t: RM.mlp_backbone(sd, x, t)  # noqa: E731
    return lambda x, t: RM.karras_mean(net, RM.vp_alpha_sigma, x, t)  # noqa: E731


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("tag", list(SAMPLER_CASES))
def test_sampler_on_gpu(tag, graph):
    g = load_golden("samplers")
    name, kw = SAMPLER_CASES[tag]
    smp = getattr(S, name)(denoiser(g, DEV), silent=True, graph=graph, **kw)
    x1 = g[f"{tag}_x1"].to(DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(1)
    x0 = smp(x1)
    after = gen.get_offset()
    assert x0.shape == x1.shape and torch.isfinite(x0).all()
    # the fused loop ran: one table-driven loop object, graph-captured when asked for
    assert len(smp._loops) == 1, "sampler fell back to the generic Python loop"
    loop = next(iter(smp._loops.values()))
    assert (loop.graph is not None) == graph, loop.graph_error
    torch.manual_seed(1)
    ref = RS.sample(name, _oracle_mean(g), RM.vp_alpha_sigma, x1, **kw)
    assert gen.get_offset() == after, "the fused loop must leave the generator where the reference's draws leave it"
    assert close(x0, ref, rtol=1e-3, atol=2e-5), (x0 - ref).abs().max()
    if tag not in ("ito", "pc"):  # deterministic: comparable with the CPU fixture of the reference
        fix = g[f"{tag}_x0"].to(DEV)
        assert close(x0, fix, rtol=1e-3, atol=1e-4), (x0 - fix).abs().max()
    # the opt-out path (plain torch execution model on the device) agrees too
    with engine.eager_torch():
        torch.manual_seed(1)
        eager = smp(x1)
    assert close(x0, eager, rtol=1e-3, atol=2e-5), (x0 - eager).abs().max()


def test_fused_loop_is_reused_and_reproducible():
    g = load_golden("samplers")
    smp = S.zABSampler(denoiser(g, DEV), steps=12, order=3, silent=True, graph=True)
    x1 = g["zab3_x1"].to(DEV)
    a = smp(x1)
    loop = next(iter(smp._loops.values()))
    b = smp(x1)
    assert next(iter(smp._loops.values())) is loop and torch.equal(a, b)
    c = smp(2 * x1)  # same signature, other data: same graph
    assert next(iter(smp._loops.values())) is loop and not torch.equal(a, c)


def _cfg_setup(batched=None):
    g = load_golden("samplers")
    net = LabelMlp(torch.nn.Module, torch)
    net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("cfg_w_")})
    den = CFGDenoiser(KarrasDenoiser(net.to(DEV), VPSchedule()), batched=batched).eval()
    pos = {"label": (torch.arange(8) % 3).to(DEV)}
    neg = {"label": torch.zeros(8, dtype=torch.long, device=DEV)}
    return g, den, pos, neg


@pytest.mark.parametrize("batched", [None, False])
def test_cfg_on_gpu(batched):
    """Classifier-free guidance inside the fused loop: ONE forward over the 2B batch [c+; c-] (or two of B), the
    combine m+ + w (m+ - m-) inside the transition kernel; vs the reference fixture."""
    g, den, pos, neg = _cfg_setup(batched)
    smp = S.DDIMSampler(den, steps=8, silent=True, graph=True)
    x0 = smp(g["cfg_x"].to(DEV), positive=pos, negative=neg, guidance=1.5)
    loop = next(iter(smp._loops.values()))
    assert loop.guided and loop.graph is not None and loop.batched == (batched is None)
    assert close(x0, g["cfg_ddim_x0"].to(DEV), rtol=1e-3, atol=1e-4), (x0 - g["cfg_ddim_x0"].to(DEV)).abs().max()
    # another guidance strength / other labels: same graph (both live in device memory), other result
    with engine.eager_torch():
        ref = S.DDIMSampler(den, steps=8, silent=True)(g["cfg_x"].to(DEV), positive=neg, negative=pos, guidance=0.3)
    x0b = smp(g["cfg_x"].to(DEV), positive=neg, negative=pos, guidance=0.3)
    assert next(iter(smp._loops.values())) is loop
    assert close(x0b, ref, rtol=1e-3, atol=1e-4) and not close(x0b, x0, rtol=1e-3, atol=1e-4)


def test_cfg_with_multistep_sampler_and_tensor_guidance():
    g, den, pos, neg = _cfg_setup()
    w = torch.tensor(2.0, device=DEV)
    smp = S.zEABSampler(den, steps=8, order=2, silent=True, graph=True)
    x = g["cfg_x"].to(DEV)
    x0 = smp(x, positive=pos, negative=neg, guidance=w)
    assert len(smp._loops) == 1 and next(iter(smp._loops.values())).guided
    with engine.eager_torch():
        ref = S.zEABSampler(den, steps=8, order=2, silent=True)(x, positive=pos, negative=neg, guidance=w)
    assert close(x0, ref, rtol=1e-3, atol=1e-4), (x0 - ref).abs().max()


def test_cfg_negative_without_keywords_takes_two_forwards():
    """negative = {} (the reference's default): the branches cannot share a batch, the loop runs two forwards."""

    class Optional(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = LabelMlp(torch.nn.Module, torch)

        def forward(self, x, t, label=None):
            return self.net(x, t, torch.zeros(x.shape[0], dtype=torch.long, device=x.device) if label is None else label)

    g = load_golden("samplers")
    net = Optional()
    net.net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("cfg_w_")})
    den = CFGDenoiser(KarrasDenoiser(net.to(DEV), VPSchedule())).eval()
    pos = {"label": (torch.arange(8) % 3).to(DEV)}
    smp = S.DDIMSampler(den, steps=8, silent=True, graph=True)
    x0 = smp(g["cfg_x"].to(DEV), positive=pos, guidance=1.5)
    loop = next(iter(smp._loops.values()))
    assert loop.guided and not loop.batched
    assert close(x0, g["cfg_ddim_x0"].to(DEV), rtol=1e-3, atol=1e-4)
