"""Generates ``tests/golden/nn_*.npz`` by running the UNMODIFIED reference backbones
(``azula.nn.unet.UNet``, ``azula.nn.vit.ViT``, ``azula.nn.dit.DiT``; read-only import).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_nn.py

Weights are not stored: every fixture uses ``seeded_state(module.state_dict(), seed)``, which depends only
on parameter names and shapes, so the tests rebuild identical weights from the host mirror's own
``state_dict`` (and thereby also check that the mirror's parameter names equal the reference's).
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
from azula.denoise import KarrasDenoiser  # noqa: E402
from azula.nn.dit import DiT  # noqa: E402
from azula.nn.unet import UNet  # noqa: E402
from azula.nn.vit import ViT  # noqa: E402
from azula.noise import VPSchedule  # noqa: E402
from azula.sample import DDIMSampler, DDPMSampler  # noqa: E402

from oracle.adm_unet import seeded_state  # noqa: E402
from oracle.gen_golden import save  # noqa: E402
from oracle.gen_golden_cfg import DIT_CASE, UNET_CASES, VIT_CASES, time_wrapper  # noqa: E402

assert azula.__file__.startswith(REF), azula.__file__


def _inputs(shape, mod_features, cond_channels=0, seed=21):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g)
    out = {"x": x}
    if mod_features:
        out["mod1"] = torch.randn(mod_features, generator=g)
        out["modB"] = torch.randn(shape[0], mod_features, generator=g)
    if cond_channels:
        out["cond"] = torch.randn(shape[0], cond_channels, *shape[2:], generator=g)
    return out


def gen_backbones(name, cls, cases):
    out = {}
    for tag, (kw, shape) in cases.items():
        net = cls(**kw).eval()
        net.load_state_dict(seeded_state(net.state_dict(), seed=77))
        ins = _inputs(shape, kw.get("mod_features", 0), kw.get("cond_channels", 0))
        cond = ins.get("cond")
        for k, v in ins.items():
            out[f"{tag}_{k}"] = v
        extra = {} if cond is None else {"cond": cond}
        if "mod1" in ins:
            out[f"{tag}_y_mod1"] = net(ins["x"], ins["mod1"], **extra)
            out[f"{tag}_y_modB"] = net(ins["x"], ins["modB"], **extra)
        else:
            out[f"{tag}_y"] = net(ins["x"], **extra)
        out[f"{tag}_keys"] = torch.tensor([len(net.state_dict())])
    save(name, **out)


def gen_dit():
    kw, shape = DIT_CASE
    net = DiT(**kw).eval()
    net.load_state_dict(seeded_state(net.state_dict(), seed=77))
    ins = _inputs(shape, kw["mod_features"])
    out = dict(ins)
    out["y_mod1"] = net(ins["x"], ins["mod1"])
    out["y_modB"] = net(ins["x"], ins["modB"])
    save("nn_dit", **out)


def gen_samplers():
    """KarrasDenoiser(Wrapper(UNet | ViT)) + DDIM / DDPM: BASELINE configs 2 and 4 at fixture size."""
    out = {}
    for tag, cls, kw, shape in (
        ("unet", UNet, dict(in_channels=3, out_channels=3, hid_channels=(16, 32), hid_blocks=(1, 1)), (2, 3, 8, 8)),
        ("vit", ViT, dict(in_channels=4, out_channels=4, hid_channels=64, hid_blocks=2, attention_heads=1, patch_size=2), (2, 4, 8, 8)),
    ):
        net = time_wrapper(cls, 32, **kw).eval()
        net.load_state_dict(seeded_state(net.state_dict(), seed=99))
        den = KarrasDenoiser(net, VPSchedule()).eval()
        g = torch.Generator().manual_seed(5)
        x = torch.randn(*shape, generator=g)
        out[f"{tag}_x"] = x
        out[f"{tag}_mean_t05"] = den(x, torch.tensor(0.5)).mean
        for sname, smp in (("ddim4", DDIMSampler(den, steps=4, silent=True)), ("ddpm4", DDPMSampler(den, steps=4, silent=True))):
            torch.manual_seed(0)
            x1 = smp.init(shape)
            out[f"{tag}_{sname}_x1"], out[f"{tag}_{sname}_x0"] = x1, smp(x1)
    save("nn_samplers", **out)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    gen_backbones("nn_unet", UNet, UNET_CASES)
    gen_backbones("nn_vit", ViT, VIT_CASES)
    gen_dit()
    gen_samplers()
