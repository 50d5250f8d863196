"""GPU parity of the in-repo backbones' native path (azula_b200.engine.unet / .dit, every launch through
the C ABI) and of the kernels added for them.

Kernel-level checkers are plain torch fp32 (TF32 off) on the SAME bf16-rounded operands: the differences
left are accumulation order and the final bf16 rounding, so the stated tolerance is 2^-8 relative to the
output scale.  Network-level checkers are the oracle's fp32 restatement (oracle/nn_backbones.py, pinned to
the reference fixtures on CPU) and the reference fixtures themselves, at the stated bf16 tolerance of
tests/test_adm_gpu.py (relative L2 <= 2e-2, p99.9 <= 6 % of the output std per forward)."""

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import nn_backbones as NB
from oracle.adm_unet import seeded_state
from oracle.gen_golden_cfg import DIT_CASE, UNET_CASES, VIT_CASES, time_wrapper

from azula_b200.denoise import KarrasDenoiser
from azula_b200.engine import ops
from azula_b200.nn.dit import DiT
from azula_b200.nn.unet import UNet
from azula_b200.nn.vit import ViT
from azula_b200.noise import VPSchedule
from azula_b200.sample import DDIMSampler, DDPMSampler

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        yield


def _gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def _check(got, ref, what, k=1.0):
    err = (got.float() - ref).abs()
    tol = k * (2.0**-8 * ref.abs() + 2.0**-8 * ref.abs().mean())
    bad = (err > tol).sum().item()
    assert bad == 0, (what, bad, err.max().item(), ref.abs().mean().item())


def _report(got, ref, what, rel_l2=2e-2, p999=6e-2):
    err = (got.float() - ref.float()).abs().flatten()
    scale = ref.float().std().item()
    l2 = (err.square().sum().sqrt() / ref.float().square().sum().sqrt()).item()
    q = err.kthvalue(max(1, int(0.999 * err.numel()))).values.item() / scale
    print(f"{what}: rel_l2 {l2:.2e}  p99.9/std {q:.2e}  max/std {err.max().item() / scale:.2e}")
    assert l2 <= rel_l2 and q <= p999, (what, l2, q)


# ------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("n,h,w,ci,co", [(2, 16, 16, 64, 128), (3, 8, 12, 16, 32), (1, 64, 64, 64, 128), (32, 32, 32, 128, 256),
                                         (2, 6, 10, 32, 64)])
def test_conv_stride2(n, h, w, ci, co):
    g = _gen(1)
    x = torch.randn(n, h, w, ci, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, ci, 3, 3, device=DEV, generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16)
    b = torch.randn(co, device=DEV, generator=g)
    got = ops.conv2d(x, ops.pack_conv(wt.float(), b), stride=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape
    _check(got, ref, "stride 2")
    # reading a channel slice of a wider buffer (a concatenation buffer), odd extents
    wide = torch.randn(n, h, w, ci + 24, device=DEV, generator=g).to(torch.bfloat16)
    got = ops.conv2d(wide[..., :ci], ops.pack_conv(wt.float(), b), stride=2)
    ref = F.conv2d(wide[..., :ci].float().permute(0, 3, 1, 2), wt.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    _check(got, ref, "stride 2, slice")


def test_conv_stride2_odd_extent():
    g = _gen(2)
    x = torch.randn(2, 7, 9, 16, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(16, 16, 3, 3, device=DEV, generator=g) / 12).to(torch.bfloat16)
    got = ops.conv2d(x, ops.pack_conv(wt.float(), None), stride=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert got.shape == ref.shape == (2, 4, 5, 16)
    _check(got, ref, "stride 2, odd")


@pytest.mark.parametrize("act", ["silu", "relu", "relu2", None])
@pytest.mark.parametrize("per_sample", [False, True])
def test_conv_epilogue_act_gate_residual(act, per_sample):
    g = _gen(3)
    n, h, w, ci, co = 3, 16, 16, 64, 64
    x = torch.randn(n, h, w, ci, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, ci, 3, 3, device=DEV, generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16)
    b = torch.randn(co, device=DEV, generator=g)
    res = torch.randn(n, h, w, co, device=DEV, generator=g).to(torch.bfloat16)
    abc = torch.randn(n if per_sample else 1, 3 * co, device=DEV, generator=g)
    gate = abc[:, 2 * co :]  # a slice of the [a | b | c] rows, as the plans pass it
    got = ops.conv2d(x, ops.pack_conv(wt.float(), b), act=act, gate=gate, residual=res)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1).permute(0, 2, 3, 1)
    y = {"silu": F.silu, "relu": F.relu, "relu2": lambda t: F.relu(t).square(), None: lambda t: t}[act](y)
    ref = res.float() + gate[:, None, None, :] * y
    _check(got, ref, f"act={act} gate per_sample={per_sample}", k=1.5)


@pytest.mark.parametrize("rows,ci,co", [(512, 64, 192), (16384, 768, 2304), (100, 128, 64), (77, 64, 64), (4096, 3072, 768)])
def test_token_gemm(rows, ci, co):
    g = _gen(4)
    x = torch.randn(rows, ci, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, ci, device=DEV, generator=g) / ci**0.5).to(torch.bfloat16)
    b = torch.randn(co, device=DEV, generator=g)
    res = torch.randn(rows, co, device=DEV, generator=g).to(torch.bfloat16)
    L = rows // 4 if rows % 4 == 0 else rows
    gate = torch.randn(rows // L, co, device=DEV, generator=g)
    got = ops.conv2d(x, ops.pack_conv(wt.float(), b), act="silu", gate=gate, gate_rows=L, residual=res)
    y = F.silu(x.float() @ wt.float().t() + b)
    ref = res.float() + gate.repeat_interleave(L, dim=0) * y
    _check(got, ref, "token gemm", k=1.5)
    # fp32 channel-major output (the out_proj path)
    yt = ops.conv2d(x, ops.pack_conv(wt.float(), b), nchw_f32=True)
    assert yt.shape[1] == co
    ref = (x.float() @ wt.float().t() + b).t()
    assert torch.allclose(yt.reshape(co, rows), ref, rtol=1e-4, atol=1e-4 * ref.abs().mean().item())


@pytest.mark.parametrize("kind", ["layer", "rms"])
@pytest.mark.parametrize("c", [16, 64, 128, 256, 768, 1024, 1536])
def test_rownorm_mod(kind, c):
    g = _gen(5)
    n, hw = 3, 50
    x = (torch.randn(n, hw, c, device=DEV, generator=g) * 2 + 0.5).to(torch.bfloat16)
    for per_sample in (False, True):
        mod = torch.randn(n if per_sample else 1, 3 * c, device=DEV, generator=g)
        got = ops.rownorm_mod(x.reshape(n * hw, c), kind, mod, rows_per_sample=hw)
        xf = x.float()
        norm = NB.channel_layer_norm(xf, dim=-1) if kind == "layer" else NB.rms_norm(xf)
        ref = (1 + mod[:, None, :c]) * norm + mod[:, None, c : 2 * c]
        _check(got.reshape(n, hw, c), ref, f"{kind} c={c}")
    got = ops.rownorm_mod(x.reshape(n * hw, c), kind)
    _check(got.reshape(n, hw, c), norm, f"{kind} plain c={c}")


def test_rownorm_strided_views():
    g = _gen(6)
    buf = torch.randn(2, 8, 8, 96, device=DEV, generator=g).to(torch.bfloat16)
    x = buf[..., 32:]  # channel slice of a concatenation buffer
    out = torch.zeros(2, 8, 8, 64, device=DEV, dtype=torch.bfloat16)
    ops.rownorm_mod(x, "layer", out=out)
    _check(out, NB.channel_layer_norm(x.float(), dim=-1), "slice")


@pytest.mark.parametrize("d", [16, 32, 64, 128])
def test_segment_rmsnorm(d):
    g = _gen(7)
    heads, rows = 3, 131
    c = heads * d
    qkv = torch.randn(rows, 3 * c, device=DEV, generator=g).to(torch.bfloat16)
    ref = qkv.float().clone()
    ref[:, : 2 * c] = NB.rms_norm(ref[:, : 2 * c].reshape(rows, 2 * heads, d)).reshape(rows, 2 * c)
    ops.segment_rmsnorm_(qkv, 2 * heads, d)
    _check(qkv, ref, f"d={d}")
    assert torch.equal(qkv[:, 2 * c :].float(), ref[:, 2 * c :])  # v untouched


@pytest.mark.parametrize("shape,p,q", [((2, 4, 8, 8), 2, 2), ((3, 3, 4, 6), 2, 1), ((1, 16, 32, 32), 4, 4), ((64, 4, 32, 32), 2, 2)])
def test_patchify_unpatchify(shape, p, q):
    from azula_b200.nn.layers import Patchify

    x = torch.randn(*shape, device=DEV, generator=_gen(8))
    n, c, h, w = shape
    k = c * p * q
    k_pad = -(-k // 64) * 64
    tok = ops.patchify(x, p, q, k_pad)
    ref = Patchify((p, q), channel_last=True)(x).reshape(-1, k)
    assert torch.equal(tok[:, :k].float(), ref.to(torch.bfloat16).float()) and not tok[:, k:].any()
    yt = ref.t().contiguous()  # channel-major, as the fp32 GEMM output mode writes it
    assert torch.equal(ops.unpatchify(yt, n, c, h // p, w // q, p, q), x)


def test_linear_gather():
    g = _gen(9)
    m, k, blocks = 5, 64, 3
    x = torch.randn(m, blocks * k, device=DEV, generator=g)
    widths = [24, 48, 96]
    w = torch.randn(sum(widths), k, device=DEV, generator=g) / 8
    b = torch.randn(sum(widths), device=DEV, generator=g)
    xoff = torch.cat([torch.full((n,), j * k, dtype=torch.int32) for j, n in enumerate(widths)]).to(DEV)
    got = ops.linear_gather(x, xoff, w, b, silu_in=True)
    ref = torch.cat([F.linear(F.silu(x[:, j * k : (j + 1) * k]), wj, bj)
                     for j, (wj, bj) in enumerate(zip(w.split(widths), b.split(widths)))], dim=1)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------ networks
def _seeded(net, seed=77):
    sd = seeded_state(net.state_dict(), seed=seed)
    net.load_state_dict(sd)
    return net.to(DEV), {k: v.to(DEV) for k, v in sd.items()}


def _native_plans(net) -> int:
    return sum(isinstance(k, tuple) for k in net._native)


@pytest.mark.parametrize("tag", list(UNET_CASES))
def test_unet_native_vs_oracle_and_golden(tag):
    g = load_golden("nn_unet")
    kw, _ = UNET_CASES[tag]
    net, sd = _seeded(UNet(**kw).eval())
    x = g[f"{tag}_x"].to(DEV)
    extra = {"cond": g[f"{tag}_cond"].to(DEV)} if f"{tag}_cond" in g else {}
    xin = torch.cat((x, extra["cond"]), dim=1) if extra else x
    mods = [("mod1", "y_mod1"), ("modB", "y_modB")] if f"{tag}_mod1" in g else [(None, "y")]
    for mk, yk in mods:
        mod = None if mk is None else g[f"{tag}_{mk}"].to(DEV)
        got = net(x, mod, **extra)
        assert got.dtype == torch.float32 and _native_plans(net) >= 1, "the native plan did not run"
        ora = NB.unet_forward(sd, xin, mod, hid_blocks=kw["hid_blocks"], norm=kw.get("norm", "layer"), groups=kw.get("groups", 16))
        _report(got, ora, f"unet {tag} {mk} vs oracle")
        _report(got, g[f"{tag}_{yk}"].to(DEV), f"unet {tag} {mk} vs reference fixture")
        assert torch.equal(got, net(x, mod, **extra))  # deterministic


def test_unet_config2_shape_vs_oracle():
    """BASELINE config 2 network (hid (64, 128, 256), blocks (3, 3, 3), mod 256) on 64x64 inputs, batch 4."""
    kw = dict(in_channels=3, out_channels=3, hid_channels=(64, 128, 256), hid_blocks=(3, 3, 3), mod_features=256)
    net, sd = _seeded(UNet(**kw).eval(), seed=5)
    g = _gen(10)
    x = torch.randn(4, 3, 64, 64, device=DEV, generator=g)
    for mod in (torch.randn(256, device=DEV, generator=g), torch.randn(4, 256, device=DEV, generator=g)):
        got = net(x, mod)
        assert _native_plans(net) >= 1
        _report(got, NB.unet_forward(sd, x, mod, hid_blocks=(3, 3, 3)), f"unet config 2, mod {tuple(mod.shape)}")
    # bf16 parameters / inputs: same plan, output cast back
    y16 = net.to(torch.bfloat16)(x.to(torch.bfloat16), mod.to(torch.bfloat16))
    assert y16.dtype == torch.bfloat16


@pytest.mark.parametrize("tag", list(VIT_CASES))
def test_vit_native_vs_oracle_and_golden(tag):
    g = load_golden("nn_vit")
    kw, _ = VIT_CASES[tag]
    net, sd = _seeded(ViT(**kw).eval())
    x = g[f"{tag}_x"].to(DEV)
    mods = [("mod1", "y_mod1"), ("modB", "y_modB")] if f"{tag}_mod1" in g else [(None, "y")]
    extra = {"cond": g[f"{tag}_cond"].to(DEV)} if f"{tag}_cond" in g else {}  # rope / cond cases run natively too
    for mk, yk in mods:
        mod = None if mk is None else g[f"{tag}_{mk}"].to(DEV)
        got = net(x, mod, **extra)
        assert got.dtype == torch.float32 and _native_plans(net) >= 1, "the native plan did not run"
        _report(got, g[f"{tag}_{yk}"].to(DEV), f"vit {tag} {mk} vs reference fixture")
        if isinstance(kw["patch_size"], int):
            ora = NB.vit_forward(sd, x, mod, kw["patch_size"], kw["hid_blocks"], kw["attention_heads"],
                                 kw.get("qk_norm", True), kw.get("ffn_activation", "silu"), cond=extra.get("cond"))
            _report(got, ora, f"vit {tag} {mk} vs oracle")


def test_dit_tokens_native_vs_golden():
    g = load_golden("nn_dit")
    kw, _ = DIT_CASE
    net, _ = _seeded(DiT(**kw).eval())
    for mk, yk in (("mod1", "y_mod1"), ("modB", "y_modB")):
        got = net(g["x"].to(DEV), g[mk].to(DEV))
        assert _native_plans(net) >= 1
        _report(got, g[yk].to(DEV), f"dit tokens {mk}")


def test_dit_explicit_positions_and_rope_native():
    """DiT over tokens with user-supplied positions (L, P) and rotary embedding: the native plan (positional embedding
    and {cos, sin} tables evaluated once per positions tensor) against the module's own torch definition."""
    from azula_b200 import engine

    kw = dict(in_channels=6, out_channels=3, cond_channels=2, mod_features=32, pos_channels=2, hid_channels=128, hid_blocks=2,
              attention_heads=2, rope=True)
    net, _ = _seeded(DiT(**kw).eval(), seed=8)
    g = _gen(12)
    x = torch.randn(3, 24, 6, device=DEV, generator=g)
    cond = torch.randn(3, 24, 2, device=DEV, generator=g)
    mod = torch.randn(3, 32, device=DEV, generator=g)
    pos = torch.rand(24, 2, device=DEV, generator=g) * 5
    got = net(x, mod, pos=pos, cond=cond)
    assert _native_plans(net) >= 1, "explicit positions fell back to torch"
    with engine.eager_torch():
        ref = net(x, mod, pos=pos, cond=cond)
    _report(got, ref, "dit tokens, explicit positions + rope + cond")
    pos.mul_(1.5)  # in-place change of the positions: a new plan, not the stale tables
    with engine.eager_torch():
        ref2 = net(x, mod, pos=pos, cond=cond)
    _report(net(x, mod, pos=pos, cond=cond), ref2, "dit tokens, moved positions")
    per_sample = pos.expand(3, 24, 2).contiguous()  # per-sample positions: the torch path (stated limitation)
    assert torch.isfinite(net(x, mod, pos=per_sample, cond=cond)).all()


def test_vit_condition_image_native():
    """ViT with a condition image (patchified separately, concatenated per token: azula/nn/vit.py:97-100, where the
    reference itself raises on the 3-d / 4-d mismatch): the native plan takes the channel concatenation before ONE
    patchify, which is the same token layout; against the module's torch definition and the oracle."""
    from azula_b200 import engine

    kw = dict(in_channels=3, out_channels=3, cond_channels=2, mod_features=32, hid_channels=128, hid_blocks=2, attention_heads=2,
              patch_size=2, rope=True)
    net, sd = _seeded(ViT(**kw).eval(), seed=9)
    g = _gen(13)
    x, cond = torch.randn(2, 3, 8, 8, device=DEV, generator=g), torch.randn(2, 2, 8, 8, device=DEV, generator=g)
    mod = torch.randn(2, 32, device=DEV, generator=g)
    got = net(x, mod, cond=cond)
    assert _native_plans(net) >= 1, "cond fell back to torch"
    with engine.eager_torch():
        ref = net(x, mod, cond=cond)
    _report(got, ref, "vit + cond vs torch definition")
    _report(got, NB.vit_forward(sd, x, mod, 2, 2, 2, cond=cond), "vit + cond vs oracle")


def test_vit_config4_shape_vs_oracle():
    """BASELINE config 4 network (DiT-B/2: hid 768, 12 blocks, 12 heads, patch 2) on 32x32x4 latents, batch 4."""
    kw = dict(in_channels=4, out_channels=4, mod_features=768, hid_channels=768, hid_blocks=12, attention_heads=12, patch_size=2)
    net, sd = _seeded(ViT(**kw).eval(), seed=6)
    g = _gen(11)
    x = torch.randn(4, 4, 32, 32, device=DEV, generator=g)
    mod = torch.randn(768, device=DEV, generator=g)
    got = net(x, mod)
    assert _native_plans(net) >= 1
    _report(got, NB.vit_forward(sd, x, mod, 2, 12, 12), "vit config 4", rel_l2=3e-2, p999=1e-1)


@pytest.mark.parametrize("tag,cls,kw", [
    ("unet", UNet, dict(in_channels=3, out_channels=3, hid_channels=(16, 32), hid_blocks=(1, 1))),
    ("vit", ViT, dict(in_channels=4, out_channels=4, hid_channels=64, hid_blocks=2, attention_heads=1, patch_size=2)),
])
def test_fused_sampler_with_native_backbone(tag, cls, kw):
    """KarrasDenoiser(Wrapper(UNet | ViT)) through the graph-captured loop vs the reference fixture."""
    g = load_golden("nn_samplers")
    net = time_wrapper(cls, 32, **kw).eval()
    net.load_state_dict(seeded_state(net.state_dict(), seed=99))
    den = KarrasDenoiser(net.to(DEV), VPSchedule()).eval()
    mean = den(g[f"{tag}_x"].to(DEV), torch.tensor(0.5, device=DEV)).mean
    _report(mean, g[f"{tag}_mean_t05"].to(DEV), f"{tag} posterior mean")
    for sname, S in (("ddim4", DDIMSampler), ("ddpm4", DDPMSampler)):
        smp = S(den, steps=4, silent=True, graph=True)
        x1 = g[f"{tag}_{sname}_x1"].to(DEV)
        torch.manual_seed(0)
        x0 = smp(x1)
        ref = g[f"{tag}_{sname}_x0"].to(DEV)
        if sname == "ddim4":  # deterministic: comparable with the CPU fixture
            _report(x0, ref, f"{tag} {sname}", rel_l2=3e-2, p999=1e-1)
        assert torch.isfinite(x0).all() and x0.shape == ref.shape
        torch.manual_seed(0)
        assert torch.equal(x0, smp(x1))  # graph replay is reproducible
