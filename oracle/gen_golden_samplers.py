"""Generates ``tests/golden/samplers.npz``: every sampler of the reference's ``azula/sample.py`` and the
classifier-free-guidance wrapper, run UNMODIFIED (read-only import) on the README-style MLP denoiser.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_samplers.py

Test infrastructure; never imported by product code.
"""

from __future__ import annotations

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
import azula.sample as RS  # noqa: E402
from azula.denoise import KarrasDenoiser  # noqa: E402
from azula.guidance.cfg import CFGDenoiser  # noqa: E402
from azula.noise import VPSchedule  # noqa: E402

from oracle.gen_golden import Mlp, save  # noqa: E402
from oracle.gen_golden_cfg import SAMPLER_CASES, LabelMlp  # noqa: E402

assert azula.__file__.startswith(REF), azula.__file__


def main():
    torch.manual_seed(7)
    net = Mlp()
    den = KarrasDenoiser(net, VPSchedule()).eval()
    out = {f"w_{k}": v.detach().clone() for k, v in net.state_dict().items()}
    for tag, (name, kw) in SAMPLER_CASES.items():
        smp = getattr(RS, name)(den, silent=True, **kw)
        torch.manual_seed(0)
        x1 = smp.init((16, 5))
        torch.manual_seed(1)
        out[f"{tag}_x1"], out[f"{tag}_x0"] = x1, smp(x1)
    # classifier-free guidance around a label-conditional MLP
    torch.manual_seed(11)
    lnet = LabelMlp(torch.nn.Module, torch)
    lden = CFGDenoiser(KarrasDenoiser(lnet, VPSchedule())).eval()
    out.update({f"cfg_w_{k}": v.detach().clone() for k, v in lnet.state_dict().items()})
    x = torch.randn(8, 5)
    pos, neg = {"label": torch.arange(8) % 3}, {"label": torch.zeros(8, dtype=torch.long)}
    out["cfg_x"] = x
    out["cfg_mean"] = lden(x, torch.tensor(0.4), positive=pos, negative=neg, guidance=2.5).mean
    smp = RS.DDIMSampler(lden, steps=8, silent=True)
    out["cfg_ddim_x0"] = smp(x, positive=pos, negative=neg, guidance=1.5)
    save("samplers", **out)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
