"""In-repo backbones on CPU: the host mirror (azula_b200.nn) and the oracle restatement (oracle/nn_backbones.py)
against fixtures produced by the unmodified reference (tests/golden/nn_*.npz, oracle/gen_golden_nn.py)."""

import pytest
import torch

from conftest import close, load_golden
from oracle import nn_backbones as NB
from oracle.adm_unet import seeded_state
from oracle.gen_golden_cfg import DIT_CASE, UNET_CASES, VIT_CASES, time_wrapper

from azula_b200.denoise import KarrasDenoiser
from azula_b200.nn.dit import DiT
from azula_b200.nn.layers import LayerNorm, Patchify, RMSNorm, SineEncoding, Unpatchify
from azula_b200.nn.unet import UNet
from azula_b200.nn.vit import ViT
from azula_b200.noise import VPSchedule
from azula_b200.sample import DDIMSampler, DDPMSampler


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _seeded(net, seed=77):
    sd = seeded_state(net.state_dict(), seed=seed)
    net.load_state_dict(sd)
    return sd


def _calls(g, tag):
    extra = {"cond": g[f"{tag}_cond"]} if f"{tag}_cond" in g else {}
    if f"{tag}_mod1" in g:
        return [((g[f"{tag}_x"], g[f"{tag}_mod1"]), extra, g[f"{tag}_y_mod1"]),
                ((g[f"{tag}_x"], g[f"{tag}_modB"]), extra, g[f"{tag}_y_modB"])]
    return [((g[f"{tag}_x"], None), extra, g[f"{tag}_y"])]


@pytest.mark.parametrize("tag", list(UNET_CASES))
def test_unet_mirror_and_oracle_match_reference(tag):
    g = load_golden("nn_unet")
    kw, _ = UNET_CASES[tag]
    net = UNet(**kw).eval()
    assert len(net.state_dict()) == int(g[f"{tag}_keys"])  # same parameter names as the reference
    sd = _seeded(net)
    for (x, mod), extra, want in _calls(g, tag):
        got = net(x, mod, **extra)
        assert got.shape == want.shape and close(got, want, rtol=1e-5, atol=1e-6)
        xin = torch.cat((x, extra["cond"]), dim=1) if extra else x
        ora = NB.unet_forward(sd, xin, mod, hid_blocks=kw["hid_blocks"], norm=kw.get("norm", "layer"), groups=kw.get("groups", 16))
        assert close(ora, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", list(VIT_CASES))
def test_vit_mirror_and_oracle_match_reference(tag):
    g = load_golden("nn_vit")
    kw, _ = VIT_CASES[tag]
    net = ViT(**kw).eval()
    assert len(net.state_dict()) == int(g[f"{tag}_keys"])
    sd = _seeded(net)
    for (x, mod), extra, want in _calls(g, tag):
        got = net(x, mod, **extra)
        assert got.shape == want.shape and close(got, want, rtol=1e-5, atol=1e-6)
        if isinstance(kw["patch_size"], int):
            ora = NB.vit_forward(sd, x, mod, kw["patch_size"], kw["hid_blocks"], kw["attention_heads"],
                                 kw.get("qk_norm", True), kw.get("ffn_activation", "silu"), cond=extra.get("cond"))
            assert close(ora, want, rtol=1e-5, atol=1e-6)


def test_dit_tokens_match_reference():
    g = load_golden("nn_dit")
    kw, _ = DIT_CASE
    net = DiT(**kw).eval()
    sd = _seeded(net)
    pos = torch.arange(g["x"].shape[1], dtype=torch.float32)[:, None]
    for mod, want in ((g["mod1"], g["y_mod1"]), (g["modB"], g["y_modB"])):
        assert close(net(g["x"], mod), want, rtol=1e-5, atol=1e-6)
        ora = NB.dit_forward(sd, g["x"], mod, pos, kw["hid_blocks"], kw["attention_heads"])
        assert close(ora, want, rtol=1e-5, atol=1e-6)


def test_samplers_with_in_repo_backbones_match_reference():
    """BASELINE configs 2 and 4 at fixture size: KarrasDenoiser(Wrapper(UNet | ViT)), DDIM-4 / DDPM-4."""
    g = load_golden("nn_samplers")
    for tag, cls, kw in (
        ("unet", UNet, dict(in_channels=3, out_channels=3, hid_channels=(16, 32), hid_blocks=(1, 1))),
        ("vit", ViT, dict(in_channels=4, out_channels=4, hid_channels=64, hid_blocks=2, attention_heads=1, patch_size=2)),
    ):
        net = time_wrapper(cls, 32, **kw).eval()
        _seeded(net, seed=99)
        den = KarrasDenoiser(net, VPSchedule()).eval()
        assert close(den(g[f"{tag}_x"], torch.tensor(0.5)).mean, g[f"{tag}_mean_t05"], rtol=1e-5, atol=1e-6)
        for sname, S in (("ddim4", DDIMSampler), ("ddpm4", DDPMSampler)):
            smp = S(den, steps=4, silent=True)
            torch.manual_seed(0)
            x1 = smp.init(tuple(g[f"{tag}_x"].shape))
            assert torch.equal(x1, g[f"{tag}_{sname}_x1"])
            assert close(smp(x1), g[f"{tag}_{sname}_x0"], rtol=1e-4, atol=1e-5)


def test_layers():
    x = torch.randn(3, 6, 4, 10)
    p = Patchify((2, 5))
    q = Unpatchify((2, 5))
    assert p(x).shape == (3, 60, 2, 2) and torch.equal(q(p(x)), x)
    pl, ql = Patchify((2, 5), channel_last=True), Unpatchify((2, 5), channel_last=True)
    t = pl(x)
    assert t.shape == (3, 2, 2, 60) and torch.equal(ql(t), x)
    # channel index = (z, a, b) row-major; token (i, j)
    assert t[1, 1, 0, (4 * 2 + 1) * 5 + 3] == x[1, 4, 2 + 1, 0 + 3]
    assert torch.equal(p(x).movedim(1, -1), t)
    h = torch.randn(5, 7, 3)
    ln = LayerNorm(dim=1)(h)
    v, m = torch.var_mean(h, dim=1, keepdim=True)
    assert torch.allclose(ln, (h - m) * torch.rsqrt(v + 1e-5))
    assert torch.allclose(RMSNorm(dim=-1)(h), NB.rms_norm(h))
    e = SineEncoding(8, omega=1e2)(torch.tensor([0.0, 2.0]))
    assert e.shape == (2, 8) and torch.allclose(e, NB.sine_encoding(torch.tensor([0.0, 2.0]), 8, 1e2))
    assert torch.equal(e[0], torch.tensor([0.0] * 4 + [1.0] * 4))


def test_odd_sizes_and_grad_use_the_torch_path():
    net = UNet(3, 3, hid_channels=(8, 16), hid_blocks=(1, 1), mod_features=8)
    x = torch.randn(1, 3, 7, 9)
    with torch.enable_grad():
        y = net(x.requires_grad_(), torch.randn(8))
        y.sum().backward()
    assert y.shape == x.shape and x.grad is not None and torch.isfinite(x.grad).all()
