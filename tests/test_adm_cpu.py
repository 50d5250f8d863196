"""CPU tests of the ADM plugin mirror (azula_b200.plugins.adm): bit-exact against fixtures that
the unmodified reference produced (tests/golden/adm_*.npz, made by oracle/gen_golden.py)."""

import pytest
import torch

from conftest import load_golden
from oracle import adm_unet as AU
from oracle.gen_golden_cfg import IMAGENET_256, MID_ADM, TINY_ADM

from azula_b200.denoise import GaussianPosterior
from azula_b200.plugins import adm
from azula_b200.plugins.adm import unet
from azula_b200.sample import DDIMSampler, DDPMSampler


def _seeded(cfg):
    den = adm.make_model(**cfg).eval()
    den.backbone.load_state_dict(AU.seeded_state(den.backbone.state_dict(), seed=1234))
    return den


@pytest.mark.parametrize("tag,cfg,steps", [("adm_tiny", TINY_ADM, 4), ("adm_mid", MID_ADM, 2)])
def test_plugin_matches_reference_bits(tag, cfg, steps):
    g = load_golden(tag)
    den = _seeded(cfg)
    x = g["x"]
    with torch.no_grad():
        for tstep in (3, 500, 999):
            ts = torch.full((x.shape[0],), tstep, dtype=torch.int64)
            assert torch.equal(den.backbone(x, ts), g[f"unet_t{tstep}"]), (tag, tstep)
        q = den(x, torch.tensor(0.6))
        assert isinstance(q, GaussianPosterior)
        assert torch.equal(q.mean, g["den_mean_t06"]) and torch.equal(q.var, g["den_var_t06"])
        assert q.mean.abs().max() <= 1.0  # clip_mean in eval mode
        for kind, S in (("ddim", DDIMSampler), ("ddpm", DDPMSampler)):
            smp = S(den, steps=steps, silent=True)
            torch.manual_seed(0)
            x1 = smp.init(tuple(x.shape))
            assert torch.equal(x1, g[f"{kind}_x1"])
            assert torch.equal(smp(x1), g[f"{kind}_x0"]), (tag, kind)


def test_state_dict_names_are_the_checkpoint_names():
    """A guided-diffusion state_dict must load unchanged: same keys and shapes as the oracle derives
    from the reference constructor (oracle/adm_unet.py: state_shapes)."""
    for cfg in (TINY_ADM, MID_ADM):
        net = adm.make_model(**cfg).backbone
        want = AU.state_shapes(AU.block_table(**cfg))
        have = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert have == want
    cond = dict(TINY_ADM, num_classes=10, use_new_attention_order=True)
    net = adm.make_model(**cond).backbone
    assert tuple(net.state_dict()["label_emb.weight"].shape) == (10, 128)


def test_imagenet256_card_layout():
    card = adm.cards()["imagenet_256x256"]
    assert {k: (tuple(v) if isinstance(v, list) else v) for k, v in card.config.items()} == IMAGENET_256
    lay = unet.make_layout(3, 256, 6, 2, {256 // r for r in (32, 16, 8)}, channel_mult=(1, 1, 2, 2, 4, 4),
                           num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True)
    units = list(lay.units())
    assert sum(u.kind == "res" for u in units) == 42  # SURVEY.md section 8(a) row A5 census
    assert sum(u.kind == "attn" for u in units) == 16
    assert len(lay.encoder) == 18 and len(lay.decoder) == 18
    shapes = AU.state_shapes(AU.block_table(**{k: v for k, v in IMAGENET_256.items() if not k.startswith("discrete")}))
    assert sum(int(torch.Size(s).numel()) for s in shapes.values()) == 552_814_086


def test_default_init_is_zero_output():
    """zero_module sites (reference _src/unet.py:207,285,602): a default-initialised network outputs 0."""
    den = adm.make_model(**TINY_ADM).eval()
    with torch.no_grad():
        out = den.backbone(torch.randn(1, 3, 16, 16), torch.tensor([7]))
    assert out.shape == (1, 6, 16, 16) and (out == 0).all()


def test_class_conditional_and_errors():
    cond = dict(TINY_ADM, num_classes=10)
    den = adm.make_model(**cond).eval()
    den.backbone.load_state_dict(AU.seeded_state(den.backbone.state_dict(), seed=5))
    x = torch.randn(2, 3, 16, 16)
    with torch.no_grad():
        a = den(x, torch.tensor(0.5), label=torch.tensor([1, 2])).mean
        b = den(x, torch.tensor(0.5), label=torch.tensor([3, 2])).mean
    assert not torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    with pytest.raises(ValueError):
        den.backbone(x, torch.tensor([1]))
    with pytest.raises(NotImplementedError):
        adm.make_model(**dict(TINY_ADM, resblock_updown=False))


def test_gradients_flow_through_torch_path():
    den = _seeded(TINY_ADM)
    x = torch.randn(1, 3, 16, 16, requires_grad=True)
    den(x, torch.tensor(0.3)).mean.sum().backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0


@pytest.mark.parametrize("name", ["imagenet_64x64_cond", "imagenet_128x128_cond", "ffhq_256x256"])
def test_cards_match_reference_on_cpu(name):
    """The cards the headline does not use (VERDICT r1 missing #6): 192-channel / cosine-schedule / new attention order
    (64x64), num_heads=4 with head widths 128 / 192 / 256 (128x128), one res-block and a single attention level (ffhq);
    host mirror AND oracle against the unmodified reference at a reduced spatial size (tests/golden/adm_cards.npz,
    oracle/gen_golden_cards.py).  The other three cards run in the GPU suite."""
    from oracle.gen_golden_cards_cfg import SIZE, STRIDE, card_inputs

    g = load_golden("adm_cards")
    config = adm.cards()[name].config
    with torch.no_grad():
        den = adm.make_model(**config).eval()
        sd = AU.seeded_state(den.backbone.state_dict(), seed=1234)
        den.backbone.load_state_dict(sd)
        x, ts, y = card_inputs(name, config)
        assert x.shape[-1] == SIZE
        u = den.backbone(x, ts, y=y)[..., ::STRIDE, ::STRIDE]
        assert torch.allclose(u, g[f"{name}_unet"], rtol=1e-4, atol=1e-5), (u - g[f"{name}_unet"]).abs().max()
        m = den(x, torch.tensor([0.3, 0.8]), label=y).mean[..., ::STRIDE, ::STRIDE]
        assert torch.allclose(m, g[f"{name}_mean"], rtol=1e-4, atol=1e-5)
        cfg = {k: v for k, v in config.items() if not k.startswith("discrete")}
        tab = AU.block_table(**cfg)
        o = AU.forward(sd, tab, x, ts, y)[..., ::STRIDE, ::STRIDE]
        assert torch.allclose(o, g[f"{name}_unet"], rtol=1e-4, atol=1e-5), (o - g[f"{name}_unet"]).abs().max()
