"""Preconditioners of the vdm / edm / jit / sd plugins through the fused graph loop on the GPU (SURVEY section 8
f3: rows of the coefficient table, no new kernel): the fused loop must have taken the native path, and its result
is held to the north-star fp32 tolerance against the plain torch execution model with the same seed (identical
Philox bits) and, for eta = 0, loosely (cuDNN vs CPU conv rounding through 6 steps) against the CPU host mirror."""

import pytest
import torch

from conftest import close, load_golden
from test_precond_cpu import CASES, build

from azula_b200 import engine
from azula_b200.sample import DDIMSampler

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        yield


@pytest.mark.parametrize("tag", list(CASES))
def test_fused_loop_runs_plugin_preconditioners(tag):
    g = load_golden("precond")
    den, call = build(tag, g, DEV)
    smp = DDIMSampler(den, steps=6, eta=0.3, silent=True)
    x1 = g[f"{tag}_x1"].to(DEV)
    assert engine.loop.supports(smp, x1), "the fused path must accept this denoiser"
    torch.manual_seed(4)
    x0 = smp(x1, **call)
    assert smp._loops, "fused loop was not used"
    with engine.eager_torch():
        torch.manual_seed(4)
        eager = smp(x1, **call)
    assert close(x0, eager, rtol=1e-3, atol=2e-5), (x0 - eager).abs().max()
    # deterministic (eta = 0) run against the host mirror on CPU, which the CPU tests hold bit-exact to the reference
    det = DDIMSampler(den, steps=6, silent=True)(x1, **call)
    den_cpu, call_cpu = build(tag, g)
    want = DDIMSampler(den_cpu, steps=6, silent=True)(x1.cpu(), **call_cpu)
    assert (det.cpu() - want).abs().max() < 1e-3
    mean = den(g[f"{tag}_x"].to(DEV), torch.tensor(0.6, device=DEV), **call).mean
    assert close(mean, g[f"{tag}_mean0"].to(DEV), rtol=1e-3, atol=1e-4)
