"""Preconditioners of the vdm / edm / jit / sd plugins: host mirror on CPU against fixtures
produced by the unmodified reference (tests/golden/precond.npz, oracle/gen_golden_precond.py)."""

import importlib
import pytest
import torch

from conftest import load_golden
from oracle.gen_golden_cfg import precond_backbone, precond_cases

from azula_b200.sample import DDIMSampler

CASES = precond_cases(torch)


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def build(tag, g, device="cpu"):
    plugin, cls, ctor, call = CASES[tag]
    mod = importlib.import_module(f"azula_b200.plugins.{plugin}")
    net = precond_backbone(plugin, torch)
    prefix = f"{tag}_w_"
    net.load_state_dict({k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)})
    den = getattr(mod, cls)(net, **ctor).eval().to(device)
    return den, {k: v.to(device) for k, v in call.items()}


@pytest.mark.parametrize("tag", list(CASES))
def test_posterior_mean_matches_reference_bit_for_bit(tag):
    g = load_golden("precond")
    den, call = build(tag, g)
    assert torch.equal(den(g[f"{tag}_x"], torch.tensor(0.6), **call).mean, g[f"{tag}_mean0"])
    assert torch.equal(den(g[f"{tag}_x"], torch.tensor([0.9, 0.5, 0.2, 0.05]), **call).mean, g[f"{tag}_meanB"])


@pytest.mark.parametrize("tag", list(CASES))
def test_ddim_matches_reference_bit_for_bit(tag):
    g = load_golden("precond")
    den, call = build(tag, g)
    smp = DDIMSampler(den, steps=6, eta=0.3, silent=True)
    torch.manual_seed(3)
    x1 = smp.init((4, 3, 8, 8))
    assert torch.equal(x1, g[f"{tag}_x1"])
    torch.manual_seed(4)
    assert torch.equal(smp(x1, **call), g[f"{tag}_x0"])


def test_elucidated_schedule():
    from azula_b200.plugins.edm import ElucidatedSchedule

    g = load_golden("precond")
    a, s = ElucidatedSchedule()(g["edm_sched_t"])
    assert torch.equal(a, g["edm_sched_alpha"]) and torch.equal(s, g["edm_sched_sigma"])


def test_load_model_is_out_of_scope():
    for plugin in ("vdm", "edm", "jit", "sd"):
        with pytest.raises(NotImplementedError):
            importlib.import_module(f"azula_b200.plugins.{plugin}").load_model("x")
