"""GPU parity of the halo-tile convolution kernels and of the GroupNorm + SiLU transform fused into their A path.

Checkers: (1) torch fp32 convolution (TF32 off) of the same bf16 operands, 2^-8 of the output scale;
(2) the tap-wise kernel (AZB_CONV_KNOB_HALO = 0): same products, different fp32 summation order;
(3) for the fused transform: azb_gn_apply_acc_bf16 followed by the same halo kernel -- the transform warps use the
same coefficients and the same arithmetic as the stand-alone pass, so the two must agree BIT FOR BIT.
"""

import pytest
import torch
import torch.nn.functional as F

from ctypes import byref

from azula_b200 import _lib
from azula_b200.engine import ops

from test_conv_gpu import _acc_to_sums, _check, _mk, _ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # halo tiles wherever the shape allows: by default a layer WITHOUT an input transform takes them only when every SM gets
    # several tiles, and these small test problems would compare a halo launch with a tap-wise one (another K order)
    ops.conv_tuning(ops.KNOB_HALO, 1)
    yield
    for knob in (ops.KNOB_HALO, ops.KNOB_PAIR, ops.KNOB_BLOCKN):
        ops.conv_tuning(knob, -1)


def _wide_tiles(co):
    """Small test problems would get narrow N tiles (to fill the SMs) and, without an input transform, the tap-wise
    kernel (few tiles per SM); the halo kernels exist for N >= 128 -- force both."""
    ops.conv_tuning(ops.KNOB_BLOCKN, 256 if co % 256 == 0 else 128)
    ops.conv_tuning(ops.KNOB_HALO, 1)


def _choice(x, pc, out, **kw):
    return ops.conv_choice(ops.conv_desc(x, pc, out, **kw))


HALO_SHAPES = [
    # n, h, w, c_in, c_out: 3 x 3, stride 1
    (2, 16, 16, 64, 128),     # 2 x 2 tiles per image, one channel block
    (1, 32, 32, 128, 256),    # N = 256
    (2, 32, 32, 256, 256),    # 4 channel blocks: the A ring wraps
    (1, 64, 64, 320, 512),    # two N tiles, 5 channel blocks
    (3, 16, 8, 64, 128),      # ONE tile per image (odd tile count: single-CTA kernel)
    (1, 20, 24, 192, 128),    # extents that are no multiples of the patch: partial tiles
    (3, 48, 40, 192, 256),
    (1, 128, 128, 256, 256),
    (1, 16, 16, 768, 512),
]


@pytest.mark.parametrize("shape", HALO_SHAPES)
@pytest.mark.parametrize("pair", [0, 1])
def test_halo_conv_matches_torch_and_tapwise(shape, pair):
    n, h, w, ci, co = shape
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=11)
    res = torch.randn(n, h, w, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    _wide_tiles(co)
    out = torch.empty(n, h, w, co, dtype=torch.bfloat16, device=DEV)
    ch = _choice(x, pc, out, residual=res)
    assert ch.halo == 1, (shape, ch.halo, ch.block_n)
    got, acc = ops.conv_acc(x, pc, residual=res)
    ops.conv_tuning(ops.KNOB_HALO, 0)
    assert _choice(x, pc, out, residual=res).halo == 0
    tap, acc_t = ops.conv_acc(x, pc, residual=res)
    torch.cuda.synchronize()
    ref = _ref(x, wt, b, res)
    _check(got, ref, ("halo", shape, pair))
    _check(tap, ref, ("tap-wise", shape))
    # same products, another summation order: differences are single bf16 roundings of nearly tied values
    d = (got.float() - tap.float()).abs()
    assert (d > 0).float().mean().item() < 0.02 and d.max().item() <= 2.0**-7 * ref.abs().max().item()
    assert torch.allclose(_acc_to_sums(acc), _acc_to_sums(acc_t), rtol=1e-3, atol=0.5)
    # statistics are those of the stored values, and reproducible
    o = got.double().reshape(n, h * w, co // 8, 8)
    want = torch.stack((o.sum(dim=(1, 3)), o.square().sum(dim=(1, 3))), dim=-1)
    assert torch.allclose(_acc_to_sums(acc), want, rtol=1e-5, atol=1e-3)
    ops.conv_tuning(ops.KNOB_HALO, 1)
    got2, acc2 = ops.conv_acc(x, pc, residual=res)
    assert torch.equal(got, got2) and torch.equal(acc, acc2)


def _normalised_input(n, h, w, c, split=None, seed=0):
    """A tensor with exact GroupNorm accumulators (as a producing convolution leaves them) + affine + scale/shift."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    parts, outs = [], []
    for cc in ([c] if split is None else [split, c - split]):
        x, wt, b = _mk(n, h, w, 64, cc, 3, seed=seed + cc)
        out, acc = ops.conv_acc(x, ops.pack_conv(wt.float() * 3, b))
        parts.append((acc, cc)), outs.append(out)
    t = torch.cat(outs, dim=-1).contiguous()
    gamma, beta = 1 + 0.1 * torch.randn(c, device=DEV, generator=g), 0.1 * torch.randn(c, device=DEV, generator=g)
    ss = 0.2 * torch.randn(n, 2 * c, device=DEV, generator=g)
    return t, parts, gamma, beta, ss


FUSED_SHAPES = [
    # n, h, w, c_in, c_out, split of the input into two producers, scale/shift, silu
    # (32 groups with one accumulator entry per 8 channels: C_in is a multiple of 256)
    (2, 16, 16, 256, 128, None, False, True),
    (2, 32, 32, 256, 256, None, True, True),
    (3, 16, 8, 256, 128, None, True, True),      # one tile per image, single-CTA kernel
    (1, 64, 64, 768, 256, 512, False, True),     # decoder concatenation: group 21 straddles the two producers
    (2, 20, 24, 512, 128, None, True, False),    # partial tiles, no activation (attention-style norm)
    (1, 128, 128, 256, 256, None, True, True),
]


@pytest.mark.parametrize("shape", FUSED_SHAPES)
@pytest.mark.parametrize("pair", [0, 1])
def test_fused_groupnorm_silu_conv_is_bit_exact_with_two_passes(shape, pair):
    n, h, w, ci, co, split, use_ss, silu = shape
    t, parts, gamma, beta, ss = _normalised_input(n, h, w, ci, split, seed=17)
    ss = ss if use_ss else None
    _, wt, b = _mk(n, h, w, ci, co, 3, seed=23)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    _wide_tiles(co)
    # two passes: normalise to HBM, then convolve
    y = ops.gn_apply_acc(t, parts, gamma, beta, scale_shift=ss, silu=silu)
    two, acc_two = ops.conv_acc(y, pc)
    # fused: coefficients (N x C values), then ONE convolution that reads the raw tensor
    coef = ops.gn_coef(n, h, w, parts, gamma, beta, scale_shift=ss, silu=silu)
    one, acc_one = ops.conv_acc(t, pc, in_coef=coef, in_silu=silu)
    torch.cuda.synchronize()
    assert torch.equal(one, two), (one.float() - two.float()).abs().max().item()
    assert torch.equal(acc_one, acc_two)
    # and against torch: group_norm -> modulation -> SiLU -> (bf16) -> conv
    ref = F.group_norm(t.float().permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5)
    if ss is not None:
        ref = ref * (1 + ss[:, :ci, None, None]) + ss[:, ci:, None, None]
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    _check(y, ref, "normalised input")
    _check(one, _ref(y, wt, b), ("fused conv", shape))


def test_fused_transform_with_skip_operand_and_residual():
    """The ResBlock tail at an ADM decoder shape: conv3x3(SiLU(GN(h) (1 + scale) + shift)) + conv1x1(x) as one GEMM
    (K = [9 taps | skip]); the 1 x 1 operand passes the transform warps untouched."""
    n, h, w, ci, co, skip = 2, 32, 32, 256, 256, 512
    t, parts, gamma, beta, ss = _normalised_input(n, h, w, ci, None, seed=29)
    _, wt, b = _mk(n, h, w, ci, co, 3, seed=31)
    g = torch.Generator(device=DEV).manual_seed(5)
    x2 = torch.randn(n, h, w, skip, device=DEV, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(co, skip, 1, 1, device=DEV, generator=g) / skip**0.5).to(torch.bfloat16)
    b2 = torch.randn(co, device=DEV, generator=g)
    pc = ops.pack_conv_skip(ops.pack_conv(wt.float(), b), ops.pack_conv(w2.float(), b2))
    y = ops.gn_apply_acc(t, parts, gamma, beta, scale_shift=ss)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta, scale_shift=ss)
    _wide_tiles(co)
    for pair in (1, 0):
        ops.conv_tuning(ops.KNOB_PAIR, pair)
        two, acc_two = ops.conv_acc(y, pc, x2=x2)
        one, acc_one = ops.conv_acc(t, pc, x2=x2, in_coef=coef, in_silu=True)
        assert torch.equal(one, two) and torch.equal(acc_one, acc_two)
        _check(one, _ref(y, wt, b) + _ref(x2, w2, b2), ("fused conv + skip", pair))


def test_fused_transform_into_network_output():
    """The head of the network (out.0 GroupNorm + SiLU, out.2 conv to 6 channels, fp32 NCHW, _src/unet.py:599-602)."""
    n, h, w, ci, co = 2, 32, 32, 256, 6
    t, parts, gamma, beta, _ = _normalised_input(n, h, w, ci, None, seed=37)
    _, wt, b = _mk(n, h, w, ci, co, 3, seed=41)
    pc = ops.pack_conv(wt.float(), b)
    y = ops.gn_apply_acc(t, parts, gamma, beta)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta)
    out2 = torch.empty(n, co, h, w, device=DEV)
    out1 = torch.empty(n, co, h, w, device=DEV)
    s = _lib.stream_ptr(t.device)
    d2 = ops.conv_desc(y, pc, out2, nchw_f32=True)
    d1 = ops.conv_desc(t, pc, out1, nchw_f32=True, in_coef=coef, in_silu=True)
    assert ops.conv_choice(d1).halo == 1 and ops.conv_choice(d1).block_n == 16
    _lib.check(_lib.lib().azb_conv_bf16(byref(d2), s), "azb_conv_bf16")
    _lib.check(_lib.lib().azb_conv_bf16(byref(d1), s), "azb_conv_bf16")
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    ref = _ref(y, wt, b).permute(0, 3, 1, 2)
    assert torch.allclose(out1, ref, rtol=1e-3, atol=1e-3 * ref.abs().mean().item())


def test_fused_transform_refused_where_no_halo_kernel_exists():
    """8 x 8 maps (and 1 x 1 / strided layers) keep the tap-wise kernel: in_coef must be rejected, not ignored."""
    n, h, w, ci, co = 2, 8, 8, 256, 128
    t, parts, gamma, beta, _ = _normalised_input(n, h, w, ci, None, seed=43)
    _, wt, b = _mk(n, h, w, ci, co, 3, seed=47)
    pc = ops.pack_conv(wt.float(), b)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta)
    out = torch.empty(n, h, w, co, dtype=torch.bfloat16, device=DEV)
    d = ops.conv_desc(t, pc, out, in_coef=coef, in_silu=True)
    assert _lib.lib().azb_conv_bf16(byref(d), _lib.stream_ptr(t.device)) == -6  # AZB_E_UNSUPPORTED
    d0 = ops.conv_desc(t, pc, out)
    assert ops.conv_choice(d0).halo == 0


@pytest.mark.parametrize("shape", [(2, 32, 32, 256, 256), (3, 16, 16, 128, 128), (2, 8, 8, 64, 64)])
def test_residual_read_through_nearest_upsampling(shape):
    """Upsampling ResBlock tail (_src/unet.py:231-233,247): out = up(x) + conv(h) with up(x) never stored -- the epilogue
    reads the half-resolution tensor at (h / 2, w / 2).  Bit-equal to adding the materialised upsampled tensor."""
    n, h, w, ci, co = shape
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=51)
    low = torch.randn(n, h // 2, w // 2, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6)).to(torch.bfloat16)
    up = low.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()
    pc = ops.pack_conv(wt.float(), b)
    want, acc_w = ops.conv_acc(x, pc, residual=up, gran=8 if co % 256 == 0 else 1)
    got, acc_g = ops.conv_acc(x, pc, residual=low, res_up=True, gran=8 if co % 256 == 0 else 1)
    assert torch.equal(got, want) and torch.equal(acc_g, acc_w)
    _check(got, _ref(x, wt, b, up), ("res_up", shape))


def test_pooling_both_branches_in_one_pass():
    """Downsampling ResBlock head (_src/unet.py:229-233): pool(SiLU(GN(x))) and pool(x) from one read of x, bit-equal
    to the two single-output passes."""
    from azula_b200 import _lib as L

    n, h, w, c = 2, 32, 32, 256
    t, parts, gamma, beta, _ = _normalised_input(n, h, w, c, None, seed=53)
    a = ops.gn_apply_acc(t, parts, gamma, beta, mode=2)
    raw = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=DEV)
    L.check(L.lib().azb_gn_apply_bf16(t.data_ptr(), c, raw.data_ptr(), c, n, h, w, c, 32, None, None, None, None, 0, None, 0,
                                      0, 2, L.stream_ptr(t.device)), "azb_gn_apply_bf16")
    a2, raw2 = torch.empty_like(a), torch.empty_like(raw)
    (acc, ca) = parts[0]
    L.check(L.lib().azb_gn_pool_acc_bf16(t.data_ptr(), c, a2.data_ptr(), c, raw2.data_ptr(), c, n, h, w, c, 32, acc.data_ptr(),
                                         ca, None, 0, 8, 1e-5, gamma.data_ptr(), beta.data_ptr(), None, 0, 1,
                                         L.stream_ptr(t.device)), "azb_gn_pool_acc_bf16")
    torch.cuda.synchronize()
    assert torch.equal(a2, a) and torch.equal(raw2, raw)
    ref = F.avg_pool2d(t.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    _check(raw2, ref, "pooled raw branch")


@pytest.mark.parametrize("shape", [(2, 16, 16, 256, 256, True), (3, 8, 8, 256, 128, False), (1, 32, 16, 512, 256, True),
                                   (2, 24, 20, 256, 128, True)])
@pytest.mark.parametrize("pair", [0, 1])
def test_upsampling_on_load_is_bit_exact_with_materialised_upsampling(shape, pair):
    """conv(up(SiLU(GN(x) ...))) of an upsampling ResBlock (_src/unet.py:101-109,229-233): the halo tiles come from the
    HALF-resolution tensor through a tensor map with zero-stride replication dimensions and are normalised in place;
    bit-equal to normalise + upsample to HBM (azb_gn_apply_acc_bf16 mode 1) followed by the same halo kernel."""
    n, h, w, ci, co, use_ss = shape  # (h, w): the half resolution
    t, parts, gamma, beta, ss = _normalised_input(n, h, w, ci, None, seed=61)
    ss = ss if use_ss else None
    _, wt, b = _mk(n, 2 * h, 2 * w, ci, co, 3, seed=67)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    _wide_tiles(co)
    y = ops.gn_apply_acc(t, parts, gamma, beta, scale_shift=ss, mode=1)
    assert y.shape == (n, 2 * h, 2 * w, ci)
    two, acc_two = ops.conv_acc(y, pc)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta, scale_shift=ss)
    one, acc_one = ops.conv_acc(t, pc, in_coef=coef, in_silu=True, in_up=True)
    torch.cuda.synchronize()
    assert one.shape == two.shape
    assert torch.equal(one, two), (one.float() - two.float()).abs().max().item()
    assert torch.equal(acc_one, acc_two)
    _check(one, _ref(y, wt, b), ("upsampling on load", shape))


@pytest.mark.parametrize("shape", [(2, 16, 16, 256, 256, True), (1, 32, 16, 512, 256, True), (2, 24, 20, 256, 128, False),
                                   (3, 16, 8, 256, 128, True)])
@pytest.mark.parametrize("pair", [0, 1])
def test_phase_decomposed_upsampling_convolution(shape, pair):
    """conv3x3(up2(z)) as four 2 x 2 convolutions of z (azb.h: AzbConv::in_up = 2, weights from pack_conv_up): the same
    function with 2.25 x fewer multiply-adds.  The taps that fall on one half-resolution pixel are summed in fp32 before
    the bf16 rounding of the weights, so the result is held to the fp32 convolution of the upsampled tensor (with the
    ORIGINAL fp32 weights) within the bf16 tolerance of weights and outputs, and to the 9-tap kernel within twice that."""
    n, h, w, ci, co, use_ss = shape  # (h, w): the half resolution
    t, parts, gamma, beta, ss = _normalised_input(n, h, w, ci, None, seed=71)
    ss = ss if use_ss else None
    g = torch.Generator(device=DEV).manual_seed(73)
    wt = torch.randn(co, ci, 3, 3, device=DEV, generator=g) / (9 * ci) ** 0.5  # fp32 weights
    b = torch.randn(co, device=DEV, generator=g)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    _wide_tiles(co)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta, scale_shift=ss)
    y = ops.gn_apply_acc(t, parts, gamma, beta, scale_shift=ss, mode=1)  # bf16 up(SiLU(GN(t))): what both kernels convolve
    nine, _ = ops.conv_acc(t, ops.pack_conv(wt, b), in_coef=coef, in_silu=True, in_up=True)
    pc_up = ops.pack_conv_up(wt, b)
    assert pc_up.taps == 16
    four, acc = ops.conv_acc(t, pc_up, in_coef=coef, in_silu=True, in_up=True)
    torch.cuda.synchronize()
    assert four.shape == (n, 2 * h, 2 * w, co)
    ref = F.conv2d(y.float().permute(0, 3, 1, 2), wt, b, padding=1).permute(0, 2, 3, 1)
    tol = 2.0**-7 * ref.abs() + 2.0**-7 * ref.abs().mean()  # bf16 weights (2^-9 each, ~sqrt(K) of them) + bf16 output
    assert ((four.float() - ref).abs() <= tol).all(), (four.float() - ref).abs().max().item()
    assert ((nine.float() - ref).abs() <= tol).all()
    # statistics of the stored values, per image and 8-channel block
    o = four.double().reshape(n, 4 * h * w, co // 8, 8)
    want = torch.stack((o.sum(dim=(1, 3)), o.square().sum(dim=(1, 3))), dim=-1)
    assert torch.allclose(_acc_to_sums(acc), want, rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------------------------------------------------------
# Per-pixel normalisation fused into the input transform (UNetBlock of azula/nn/unet.py:97-107) and the per-pixel sums
# the producer writes for it (AzbConv.rowstat / in_norm)

def _launch(d):
    _lib.check(_lib.lib().azb_conv_bf16(byref(d), _lib.stream_ptr(torch.device(DEV))), "azb_conv_bf16")


PIXNORM_SHAPES = [
    # n, h, w, c (the block keeps the channel count), kind
    (2, 16, 16, 64, "layer"),     # 64-column tiles: the halo kernel's N = 64 instantiation
    (3, 32, 32, 64, "rms"),
    (2, 32, 32, 128, "layer"),    # two 64-channel blocks per pixel: sums from two epilogue warps
    (1, 16, 24, 256, "layer"),    # four blocks; partial tiles
    (2, 16, 16, 256, "rms"),
    (5, 64, 64, 64, "layer"),     # several tiles per CTA
]


@pytest.mark.parametrize("shape", PIXNORM_SHAPES)
def test_pixel_norm_fused_into_the_convolution(shape):
    n, h, w, c, kind = shape
    g = torch.Generator(device=DEV).manual_seed(c + h)
    # producer: gated residual convolution (the tail of the previous block), row-domain epilogue + per-pixel sums
    x0, wt0, b0 = _mk(n, h, w, c, c, 3, seed=5)
    res = (torch.randn(n, h, w, c, device=DEV, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gate = torch.randn(n, c, device=DEV, generator=g)
    pc0 = ops.pack_conv(wt0.float(), b0)
    x = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
    stat = torch.full((n * h * w, c // 64, 2), float("nan"), device=DEV)
    d0 = ops.conv_desc(x0, pc0, x, gate=gate.data_ptr(), gate_ld=gate.stride(0), gate_rows=h * w, residual=res, rowstat=stat)
    ch0 = ops.conv_choice(d0)
    assert ch0.epi == 2, (shape, ch0.epi, ch0.block_n)
    _launch(d0)
    ref0 = _ref(x0, wt0, b0) * gate[:, None, None, :] + res.float()
    _check(x, ref0, ("producer", shape))
    blocks = x.float().reshape(n * h * w, c // 64, 64)
    assert torch.allclose(stat[..., 0], blocks.sum(-1), rtol=1e-5, atol=1e-4), ("sums", shape)
    assert torch.allclose(stat[..., 1], blocks.square().sum(-1), rtol=1e-5, atol=1e-4), ("sums of squares", shape)

    # consumer: SiLU(conv3x3((1 + a) * norm(x) + b)) with the normalisation in the input transform
    _, wt1, b1 = _mk(n, h, w, c, c, 3, seed=6)
    pc1 = ops.pack_conv(wt1.float(), b1)
    mod = 0.3 * torch.randn(n, 3 * c, device=DEV, generator=g)
    out = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
    d1 = ops.conv_desc(x, pc1, out, act=ops.ACT["silu"], in_norm=1 if kind == "layer" else 2, in_eps=1e-5, in_rowstat=stat,
                       in_mod=mod.data_ptr(), in_mod_ld=mod.stride(0))
    ch1 = ops.conv_choice(d1)
    assert ch1.halo == 1 and ch1.epi == 2, (shape, ch1.halo, ch1.epi, ch1.block_n)
    _launch(d1)
    xf = x.float()
    if kind == "layer":
        v, m = torch.var_mean(xf, dim=-1, keepdim=True)  # unbiased, as azula.nn.layers.layer_norm
        nx = (xf - m) * torch.rsqrt(v + 1e-5)
    else:
        nx = xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + 1e-5)
    y = ((1 + mod[:, None, None, :c]) * nx + mod[:, None, None, c : 2 * c]).to(torch.bfloat16)
    ref1 = F.silu(_ref(y, wt1, b1))
    # the transform evaluates y with two packed bf16 multiply-adds (two roundings, bf16 row scales) from statistics summed
    # in another order than torch's: y is within ~1 bf16 ulp of torch's (0.5 ulp), and the convolution sums 9 C of them
    err = (out.float() - ref1).abs()
    tol = 2.0**-5 * ref1.abs() + 2.0**-5 * ref1.abs().mean()
    assert (err.square().sum() / ref1.square().sum()).sqrt().item() < 6e-3, (shape, "relative L2")
    assert (err > tol).sum().item() == 0, (shape, err.max().item(), ref1.abs().mean().item())
    # ... and agrees with the two-launch route (azb_rownorm_mod_bf16, then the same convolution) to the same bar
    y2 = ops.rownorm_mod(x, kind=kind, mod=mod, rows_per_sample=h * w)
    two = ops.conv2d(y2, pc1, act="silu")
    err2 = (out.float() - two.float()).abs()
    assert (err2 > tol).sum().item() == 0, (shape, "vs two launches", err2.max().item())
    out2 = torch.empty_like(out)
    d1.out = out2.data_ptr()
    _launch(d1)
    assert torch.equal(out, out2), "not reproducible"


def test_pixel_norm_needs_the_halo_kernel():
    """Shapes the halo kernels do not serve report AZB_E_UNSUPPORTED for in_norm (the plan then keeps the separate pass)."""
    n, h, w, c = 2, 8, 4, 64  # smaller than one 8 x 16 patch
    x, wt, b = _mk(n, h, w, c, c, 3)
    pc = ops.pack_conv(wt.float(), b)
    stat = torch.zeros(n * h * w, 1, 2, device=DEV)
    mod = torch.zeros(n, 3 * c, device=DEV)
    out = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
    d = ops.conv_desc(x, pc, out, in_norm=1, in_rowstat=stat, in_mod=mod.data_ptr(), in_mod_ld=mod.stride(0))
    c_ = ops.AzbConvChoice()
    assert _lib.lib().azb_conv_choice(byref(d), byref(c_)) == -6  # AZB_E_UNSUPPORTED
