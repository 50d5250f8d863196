"""Fixture configurations shared by oracle/gen_golden.py and the tests (test infrastructure)."""

TINY_ADM = dict(
    image_size=16,
    num_channels=32,
    channel_mult=(1, 2),
    num_res_blocks=1,
    attention_resolutions=(8,),
    num_head_channels=16,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# The imagenet_256x256 card's structure (reference azula/plugins/adm/cards.yaml:36-50) at a
# CPU-friendly width: same depth / multipliers / attention placement / head width, model
# channels 64, 64x64 input.
MID_ADM = dict(
    image_size=64,
    num_channels=64,
    channel_mult=(1, 1, 2, 2, 4, 4),
    num_res_blocks=2,
    attention_resolutions=(8, 4, 2),
    num_head_channels=64,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# The card's width (256 channels: GroupNorm groups of 8+ channels) at a small spatial size: the configuration class in
# which GroupNorm + SiLU ride on the halo tiles of the 3 x 3 convolutions (no reference fixture: oracle-checked on GPU).
WIDE_ADM = dict(
    image_size=32,
    num_channels=256,
    channel_mult=(1, 2),
    num_res_blocks=1,
    attention_resolutions=(2,),
    num_head_channels=64,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# The card itself (cards.yaml:36-50): 552.8 M parameters.
IMAGENET_256 = dict(
    discrete_schedule="linear",
    discrete_steps=1000,
    attention_resolutions=(32, 16, 8),
    channel_mult=(1, 1, 2, 2, 4, 4),
    image_size=256,
    num_channels=256,
    num_classes=None,
    num_head_channels=64,
    num_res_blocks=2,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# ---- in-repo backbones (azula/nn/unet.py, azula/nn/vit.py): small CPU-friendly fixtures
UNET_CASES = {
    # tag: (constructor kwargs, input (B, C, H, W), modulation rows: None | 1 | B, cond channels)
    "layer": (dict(in_channels=3, out_channels=3, hid_channels=(16, 32, 64), hid_blocks=(2, 2, 1), mod_features=32), (2, 3, 16, 16)),
    "rms_free": (dict(in_channels=2, out_channels=5, hid_channels=(16, 32), hid_blocks=(1, 2), norm="rms", ffn_factor=2), (3, 2, 8, 12)),
    "group_cond": (dict(in_channels=3, out_channels=3, cond_channels=1, hid_channels=(32, 64), hid_blocks=(1, 1), mod_features=16,
                        norm="group", groups=4), (2, 3, 8, 8)),
}

VIT_CASES = {
    "dit_b2_small": (dict(in_channels=4, out_channels=4, mod_features=64, hid_channels=128, hid_blocks=2, attention_heads=2,
                          patch_size=2), (2, 4, 8, 8)),
    "relu2_free": (dict(in_channels=3, out_channels=2, hid_channels=64, hid_blocks=1, attention_heads=2, patch_size=(2, 1),
                        ffn_activation="relu2", qk_norm=False, ffn_factor=2), (2, 3, 4, 6)),
    # rotary positional embedding (azula/nn/attention.py:97-100,111-156).  (No condition image: the reference's
    # ViT.forward concatenates a 4-d patchified cond to 3-d tokens and raises, azula/nn/vit.py:97-104.)
    "rope_mod": (dict(in_channels=3, out_channels=3, mod_features=32, hid_channels=128, hid_blocks=2,
                      attention_heads=2, patch_size=2, rope=True), (2, 3, 8, 8)),
    "rope_nonorm": (dict(in_channels=4, out_channels=4, hid_channels=64, hid_blocks=1, attention_heads=4, patch_size=2, rope=True,
                         qk_norm=False), (2, 4, 8, 4)),
}

DIT_CASE = (dict(in_channels=6, out_channels=3, mod_features=32, hid_channels=64, hid_blocks=2, attention_heads=4), (2, 10, 6))


def time_wrapper(net_cls, features: int, **kwargs):
    """The tutorial pattern around an in-repo backbone (reference docs/tutorials/mnist.ipynb cell 8, minus the
    label embedding): mod = MLP(log_snr[..., None]); BASELINE configs 2 and 4 use it."""
    import torch

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = net_cls(mod_features=features, **kwargs)
            self.time_embedding = torch.nn.Sequential(
                torch.nn.Linear(1, features), torch.nn.SiLU(), torch.nn.Linear(features, features)
            )

        def forward(self, x_t, log_snr_t):
            return self.net(x_t, self.time_embedding(log_snr_t[..., None]))

    return Wrapper()


# ---- every sampler of azula/sample.py on the README-style MLP denoiser: tag -> (class name, kwargs)
SAMPLER_CASES = {
    "euler": ("EulerSampler", dict(steps=16)),
    "heun": ("HeunSampler", dict(steps=8)),
    "ito": ("ItoSampler", dict(steps=16, eta=0.7, temperature=0.9)),
    "ito_ode": ("ItoSampler", dict(steps=16, eta=0.0)),
    "pc": ("PCSampler", dict(steps=8, corrections=2, delta=0.05)),
    "zab2": ("zABSampler", dict(steps=12, order=2)),
    "zab3": ("zABSampler", dict(steps=12, order=3)),
    "vab": ("vABSampler", dict(steps=12, order=2)),
    "zeab": ("zEABSampler", dict(steps=12, order=3)),
    "xeab": ("xEABSampler", dict(steps=12, order=2)),
    "reab": ("REABSampler", dict(steps=12, order=2, stop=0.02)),
}


def LabelMlp(module_cls, torch):
    """A tiny label-conditional backbone b(x, t, label) for the classifier-free-guidance fixtures."""

    class Net(module_cls):
        def __init__(self):
            super().__init__()
            self.l1 = torch.nn.Linear(5 + 1, 32)
            self.l2 = torch.nn.Linear(32, 5)
            self.emb = torch.nn.Embedding(3, 32)

        def forward(self, x, t, label):
            h = self.l1(torch.cat((x, t.expand(x.shape[:-1])[..., None]), dim=-1)) + self.emb(label)
            return self.l2(torch.tanh(h))

    return Net()


# ---- preconditioners of the other plugins: tag -> (plugin module, class, ctor kwargs, call kwargs)
def precond_backbone(kind: str, torch):
    """Tiny image backbones with the call conventions of the vdm / edm / jit / sd plugins."""
    nn = torch.nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 8, 3, padding=1)
            self.conv2 = nn.Conv2d(8, 3, 3, padding=1)
            self.time = nn.Linear(1, 8)
            self.cond = nn.Linear(4, 8) if kind in ("edm", "sd") else nn.Embedding(5, 8)

        def body(self, x, t, c):
            t = t.to(x.dtype).reshape(-1, 1).expand(x.shape[0], 1)
            h = self.conv1(x) + (self.time(t) + c)[:, :, None, None]
            return self.conv2(torch.tanh(h))

        def forward(self, *args, **kwargs):
            if kind == "vdm":
                x, t = args
                return self.body(x, t, 0.0)
            if kind == "edm":
                x, t = args
                return self.body(x, t, self.cond(kwargs["class_labels"]))
            if kind == "jit":
                x, t = args
                return self.body(x, t, self.cond(kwargs["y"]))
            from types import SimpleNamespace

            c = self.cond(kwargs["encoder_hidden_states"]).mean(dim=1)
            return SimpleNamespace(sample=self.body(kwargs["sample"], kwargs["timestep"] / 1000.0, c))

    return Net()


def precond_cases(torch):
    sig = torch.linspace(0.03, 0.995, 1000)
    return {
        "vdm": ("vdm", "VelocityDenoiser", {}, {}),
        "edm": ("edm", "ElucidatedDenoiser", {}, {"label": torch.eye(4)[[0, 2, 1, 3]]}),
        "jit": ("jit", "JITDenoiser", {"num_classes": 4}, {"label": torch.tensor([1, 0, 3, 2])}),
        "jit_null": ("jit", "JITDenoiser", {"num_classes": 4}, {}),
        "sd_eps": ("sd", "StableDenoiser", {"sigmas": sig}, {"prompt_embeds": torch.linspace(-1, 1, 24).reshape(1, 6, 4)}),
        "sd_v": ("sd", "StableDenoiser", {"sigmas": sig, "prediction": "velocity"},
                 {"prompt_embeds": torch.linspace(-1, 1, 96).reshape(4, 6, 4)}),
    }
