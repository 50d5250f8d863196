"""Generates ``tests/golden/precond.npz``: the preconditioners of the reference's vdm / edm / jit / sd plugins
(``VelocityDenoiser``, ``ElucidatedDenoiser`` + ``ElucidatedSchedule``, ``JITDenoiser``, ``StableDenoiser``), run
UNMODIFIED (read-only import) around tiny seeded backbones: posterior means at a scalar and a batched time, and a
6-step DDIM sampling each.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_precond.py

Test infrastructure; never imported by product code.
"""

from __future__ import annotations

import importlib
import os
import sys
import torch
import types

REF = os.environ.get("AZULA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.modules.setdefault("gdown", types.ModuleType("gdown"))

import azula  # noqa: E402
from azula.sample import DDIMSampler  # noqa: E402

from oracle.gen_golden import save  # noqa: E402
from oracle.gen_golden_cfg import precond_backbone, precond_cases  # noqa: E402

assert azula.__file__.startswith(REF), azula.__file__


def main():
    out = {}
    for tag, (plugin, cls, ctor, call) in precond_cases(torch).items():
        mod = importlib.import_module(f"azula.plugins.{plugin}")
        torch.manual_seed(len(tag) + 17)
        net = precond_backbone(plugin, torch)
        den = getattr(mod, cls)(net, **ctor).eval()
        out.update({f"{tag}_w_{k}": v.detach().clone() for k, v in net.state_dict().items()})
        x = torch.randn(4, 3, 8, 8)
        out[f"{tag}_x"] = x
        out[f"{tag}_mean0"] = den(x, torch.tensor(0.6), **call).mean
        out[f"{tag}_meanB"] = den(x, torch.tensor([0.9, 0.5, 0.2, 0.05]), **call).mean
        smp = DDIMSampler(den, steps=6, eta=0.3, silent=True)
        torch.manual_seed(3)
        x1 = smp.init((4, 3, 8, 8))
        torch.manual_seed(4)
        out[f"{tag}_x1"], out[f"{tag}_x0"] = x1, smp(x1, **call)
    t = torch.linspace(0, 1, 33)
    a, s = importlib.import_module("azula.plugins.edm").ElucidatedSchedule()(t)
    out["edm_sched_t"], out["edm_sched_alpha"], out["edm_sched_sigma"] = t, a, s
    save("precond", **out)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
