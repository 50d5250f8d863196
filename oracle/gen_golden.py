"""Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference (read-only import).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference has no golden vectors for the generation path, so these
fixtures -- outputs of the reference itself on seeded inputs -- are what pins the oracle.
Test infrastructure; never imported by product code.
"""

from __future__ import annotations

import numpy as np
import os
import sys
import torch
import types

REF = os.environ.get("AZULA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.modules.setdefault("gdown", types.ModuleType("gdown"))  # azula/hub.py:9 (no network here)

import azula  # noqa: E402
from azula.denoise import KarrasDenoiser  # noqa: E402
from azula.nn.layers import SineEncoding  # noqa: E402
from azula.noise import VPSchedule  # noqa: E402
from azula.plugins import adm  # noqa: E402
from azula.sample import DDIMSampler, DDPMSampler  # noqa: E402

from oracle.adm_unet import seeded_state  # noqa: E402

assert azula.__file__.startswith(REF), azula.__file__

from oracle.gen_golden_cfg import MID_ADM, TINY_ADM  # noqa: E402


class Mlp(torch.nn.Module):
    """tests/test_sample.py:28-51 of the reference (BASELINE config 1 backbone)."""

    def __init__(self, features=5):
        super().__init__()
        self.l1 = torch.nn.Linear(features, 64)
        self.l2 = torch.nn.Linear(64, features)
        self.enc = SineEncoding(64)

    def forward(self, x, t):
        return self.l2(torch.relu(self.l1(x) + self.enc(t)))


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_schedule():
    t = torch.linspace(0, 1, 65)
    out = {"t": t}
    for tag, (am, sm) in {"default": (1e-3, 1e-3), "adm": (1e-2, 1e-2)}.items():
        a, s = VPSchedule(alpha_min=am, sigma_min=sm)(t)
        out[f"alpha_{tag}"], out[f"sigma_{tag}"] = a, s
    den = adm.AblatedDenoiser(torch.nn.Identity())
    out["adm_sigmas"] = den.sigmas
    a, s = den.schedule(t)
    c_time = torch.searchsorted(den.sigmas, (s * torch.rsqrt(a**2 + s**2)).flatten())
    out["adm_c_time"] = c_time
    save("schedule", **out)


def gen_mlp():
    torch.manual_seed(7)
    net = Mlp()
    den = KarrasDenoiser(net, VPSchedule()).eval()
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out = {f"w_{k}": v for k, v in state.items()}
    x = torch.randn(64, 5)
    t = torch.rand(64)
    out["x"], out["t"] = x, t
    out["mean_batched_t"] = den(x, t).mean.detach()
    out["mean_scalar_t"] = den(x, torch.tensor(0.37)).mean.detach()
    for name, smp in {
        "ddpm1000": DDPMSampler(den, steps=1000, silent=True),
        "ddim64_eta0": DDIMSampler(den, steps=64, eta=0.0, silent=True),
        "ddim64_eta1": DDIMSampler(den, steps=64, eta=1.0, silent=True),
        "ddim16_eta05_partial": DDIMSampler(den, steps=16, eta=0.5, start=0.8, stop=0.1, silent=True),
    }.items():
        torch.manual_seed(0)
        x1 = smp.init((64, 5))
        x0 = smp(x1)
        out[f"{name}_x1"], out[f"{name}_x0"] = x1, x0
    save("mlp_karras", **out)


def gen_adm(tag, cfg, batch, steps):
    den = adm.make_model(**cfg).eval()
    sd = seeded_state(den.backbone.state_dict(), seed=1234)
    den.backbone.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    size = cfg["image_size"]
    x = torch.randn(batch, 3, size, size, generator=g)
    out = {"x": x}
    with torch.no_grad():
        for i, tstep in enumerate((3, 500, 999)):
            ts = torch.full((batch,), tstep, dtype=torch.int64)
            out[f"unet_t{tstep}"] = den.backbone(x, ts)
        taps = {}
        hooks = []
        for name, mod in den.backbone.named_modules():
            if name.count(".") == 2 or name in ("input_blocks.0.0",):
                hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o)))
        den.backbone(x, torch.tensor([500]))
        for h in hooks:
            h.remove()
        for name in list(taps)[:: max(1, len(taps) // 6)]:
            o = taps[name]
            out[f"tap_{name}"] = torch.stack((o.mean(), o.std(), o.flatten()[:: max(1, o.numel() // 97)].sum()))
        q = den(x, torch.tensor(0.6))
        out["den_mean_t06"], out["den_var_t06"] = q.mean, q.var
        q = den(x, torch.tensor([0.05, 0.95][:batch] if batch <= 2 else torch.linspace(0.05, 0.95, batch).tolist()))
        out["den_mean_tb"] = q.mean
        smp = DDIMSampler(den, steps=steps, silent=True)
        torch.manual_seed(0)
        x1 = smp.init((batch, 3, size, size))
        out["ddim_x1"], out["ddim_x0"] = x1, smp(x1)
        smp = DDPMSampler(den, steps=steps, silent=True)
        torch.manual_seed(0)
        x1 = smp.init((batch, 3, size, size))
        out["ddpm_x1"], out["ddpm_x0"] = x1, smp(x1)
    save(tag, **out)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    gen_schedule()
    gen_mlp()
    gen_adm("adm_tiny", TINY_ADM, batch=2, steps=4)
    gen_adm("adm_mid", MID_ADM, batch=1, steps=2)
