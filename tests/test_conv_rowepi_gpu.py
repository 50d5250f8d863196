"""The row-domain epilogue with TMA stores (conv_gemm_kernel<..., EPI = 2>, csrc/conv_gemm.cu) against the epilogue it
replaces (AZB_CONV_KNOB_ROWEPI = 0: shared-memory transpose + per-lane stores) on the SAME launches: both apply
bias -> activation -> gate -> residual -> bf16 rounding in the same order to the same fp32 accumulators, so the stored
tensors must agree BIT FOR BIT -- through partial tiles, channel tails, channel-slice outputs, CTA pairs, halo tiles,
the residual read through an upsampling and the phase scatter of an upsampling convolution -- and the exact GroupNorm
accumulators must describe the stored values (they are sums of the same numbers in another order).  Every case is also
held to a torch fp32 reference.
"""

import pytest
import torch
import torch.nn.functional as F

from azula_b200.engine import ops

from test_conv_gpu import _acc_to_sums, _check, _mk, _ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    for knob in (ops.KNOB_HALO, ops.KNOB_PAIR, ops.KNOB_BLOCKN, ops.KNOB_ROWEPI):
        ops.conv_tuning(knob, -1)


def _both(fn):
    """Runs fn() with the row-domain epilogue (forced: also for activation / gate epilogues, which the automatic choice
    leaves to the old one) and with the old one; returns the two results."""
    ops.conv_tuning(ops.KNOB_ROWEPI, 1)
    new = fn()
    ops.conv_tuning(ops.KNOB_ROWEPI, 0)
    old = fn()
    ops.conv_tuning(ops.KNOB_ROWEPI, -1)
    torch.cuda.synchronize()
    return new, old


SHAPES = [
    # n, h, w, c_in, c_out, k, forced N tile (0 = automatic)
    (2, 16, 16, 64, 128, 3, 0),
    (1, 64, 64, 64, 64, 3, 0),        # N tile 64: the two warps of a lane quarter share a 64-column store block
    (2, 32, 32, 256, 256, 3, 256),    # halo, pairs
    (1, 20, 24, 192, 128, 3, 128),    # partial tiles in both directions
    (3, 4, 4, 64, 64, 3, 0),          # tiles that span images (8 pixels per image and 32-row store block)
    (5, 2, 2, 128, 256, 3, 0),
    (2, 12, 12, 64, 128, 3, 0),       # patch taller than the image: rows beyond H are clipped by the map
    (2, 16, 16, 128, 72, 1, 64),      # channel tail: 72 = 64 + 8 valid columns in the last store block
    (1, 32, 32, 512, 512, 1, 256),
    (2, 16, 16, 128, 384, 1, 128),
    (1, 48, 40, 192, 320, 3, 64),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("pair", [0, 1])
def test_rowepi_equals_old_epilogue_bits(shape, pair):
    n, h, w, ci, co, k, bn = shape
    x, wt, b = _mk(n, h, w, ci, co, k, seed=31)
    res = torch.randn(n, h, w, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    if bn:
        ops.conv_tuning(ops.KNOB_BLOCKN, bn)
    if bn == 64 and k == 3 and ci > 64:
        # 64-column tiles take halo tiles only together with the row-domain epilogue, and halo tiles traverse K in another
        # order (channel block outer, tap inner): keep both launches tap-wise so that only the epilogue differs
        ops.conv_tuning(ops.KNOB_HALO, 0)
    # into a channel slice of a wider buffer: the tensor map must not touch the neighbours
    def run():
        wide = torch.full((n, h, w, co + 64), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.conv(x, pc, out=wide[..., 64:], residual=res)
        return wide
    new, old = _both(run)
    assert torch.equal(new, old), (new.float() - old.float()).abs().max().item()
    assert (new[..., :64] == 7.0).all()
    _check(new[..., 64:], _ref(x, wt, b, res), shape)


@pytest.mark.parametrize("shape", [(2, 32, 32, 256, 256, 3), (3, 16, 8, 64, 128, 3), (1, 20, 24, 192, 256, 3), (2, 16, 16, 128, 256, 1),
                                   (1, 64, 64, 320, 512, 3), (16, 8, 8, 256, 256, 3)])
@pytest.mark.parametrize("pair", [0, 1])
def test_rowepi_groupnorm_accumulators(shape, pair):
    n, h, w, ci, co, k = shape
    x, wt, b = _mk(n, h, w, ci, co, k, seed=37)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_PAIR, pair)
    ops.conv_tuning(ops.KNOB_BLOCKN, 256 if co % 256 == 0 else 128)
    (new, acc_new), (old, acc_old) = _both(lambda: ops.conv_acc(x, pc))
    assert torch.equal(new, old)
    o = new.double().reshape(n, h * w, co // 8, 8)
    want = torch.stack((o.sum(dim=(1, 3)), o.square().sum(dim=(1, 3))), dim=-1)
    assert torch.allclose(_acc_to_sums(acc_new), want, rtol=1e-5, atol=1e-3)
    assert torch.allclose(_acc_to_sums(acc_new), _acc_to_sums(acc_old), rtol=1e-5, atol=1e-3)
    again, acc_again = ops.conv_acc(x, pc)
    assert torch.equal(again, new) and torch.equal(acc_again, acc_new)  # integer accumulators: order independent


@pytest.mark.parametrize("act", [None, "silu", "relu", "relu2"])
@pytest.mark.parametrize("shape", [(2, 32, 32, 128, 128, 3, 1), (4, 16, 16, 256, 256, 3, 1), (2, 64, 64, 64, 64, 3, 1), (2, 32, 32, 64, 128, 3, 2)])
def test_rowepi_activation_gate_residual(act, shape):
    """The in-repo U-Net block's two convolutions (azula/nn/unet.py:97-107): SiLU(conv(y) + b) and x + c * (conv(h) + b)
    with a per-sample gate row; also a strided convolution."""
    n, h, w, ci, co, k, stride = shape
    x, wt, b = _mk(n, h, w, ci, co, k, seed=41)
    ho, wo = -(-h // stride), -(-w // stride)
    g = torch.Generator(device=DEV).manual_seed(7)
    gate = torch.randn(n, 3 * co, device=DEV, generator=g)[:, 2 * co :]  # a slice of a wider [a | b | c] row
    res = torch.randn(n, ho, wo, co, device=DEV, generator=g).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_HALO, 0)  # the old epilogue has no halo kernel with an activation: compare tap-wise kernels
    new, old = _both(lambda: ops.conv2d(x, pc, stride=stride, act=act, gate=gate, residual=res))
    assert torch.equal(new, old), (new.float() - old.float()).abs().max().item()
    ops.conv_tuning(ops.KNOB_HALO, 1)
    if stride == 1 and co >= 128:  # ... and the halo kernel with the row-domain epilogue against torch
        ops.conv_tuning(ops.KNOB_BLOCKN, 256 if co % 256 == 0 else 128)
        out = torch.empty(n, ho, wo, co, dtype=torch.bfloat16, device=DEV)
        d = ops.conv_desc(x, pc, out, act=ops.ACT[act], gate=gate.data_ptr(), gate_ld=gate.stride(0), gate_rows=ho * wo, residual=res)
        ops.conv_tuning(ops.KNOB_ROWEPI, 1)
        assert ops.conv_choice(d).halo == 1
        new = ops.conv2d(x, pc, stride=stride, act=act, gate=gate, residual=res)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1, stride=stride).permute(0, 2, 3, 1)
    y = {None: lambda t: t, "silu": F.silu, "relu": F.relu, "relu2": lambda t: F.relu(t) ** 2}[act](y)
    y = res.float() + gate[:, None, None, :] * y
    err = (new.float() - y).abs()
    assert (err <= 2.0**-7 * y.abs() + 2.0**-7 * y.abs().mean()).all(), err.max().item()


@pytest.mark.parametrize("rows,ci,co", [(16384, 768, 2304), (16384, 768, 768), (4096, 3072, 768), (300, 64, 192), (128, 256, 1024)])
def test_rowepi_token_gemm(rows, ci, co):
    """DiT projections as 1-tap GEMMs over (rows, C) token matrices, with activation / gate / residual epilogues."""
    g = torch.Generator(device=DEV).manual_seed(rows)
    x = torch.randn(rows, ci, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, ci, device=DEV, generator=g) / ci**0.5).to(torch.bfloat16)
    b = torch.randn(co, device=DEV, generator=g)
    res = torch.randn(rows, co, device=DEV, generator=g).to(torch.bfloat16)
    samples = 4 if rows % 4 == 0 else 1
    gate = torch.randn(samples, co, device=DEV, generator=g)
    pc = ops.pack_conv(wt.float(), b)
    new, old = _both(lambda: ops.conv2d(x, pc, act="silu", gate=gate, gate_rows=rows // samples, residual=res))
    assert torch.equal(new, old)
    y = F.silu(F.linear(x.float(), wt.float(), b))
    y = res.float() + gate.repeat_interleave(rows // samples, dim=0) * y
    err = (new.float() - y).abs()
    assert (err <= 2.0**-7 * y.abs() + 2.0**-7 * y.abs().mean()).all(), err.max().item()


def test_rowepi_residual_through_upsampling_and_phase_scatter():
    """An upsampling ResBlock (_src/unet.py:229-233): conv1(up(SiLU(GN(x)))) phase-decomposed -- the store goes through
    the 5-d (channel + dx ld, w, dy, h, n) map -- and conv2 + up(x) with the residual read at (h / 2, w / 2)."""
    from test_conv_halo_gpu import _normalised_input

    n, h, w, c = 2, 16, 16, 256
    t, parts, gamma, beta, _ = _normalised_input(n, h, w, c, seed=43)
    _, wt, b = _mk(n, h, w, c, c, 3, seed=47)
    coef = ops.gn_coef(n, h, w, parts, gamma, beta, silu=True)
    ops.conv_tuning(ops.KNOB_BLOCKN, 256)
    pcu = ops.pack_conv_up(wt.float(), b)
    (new, acc_new), (old, acc_old) = _both(lambda: ops.conv_acc(t, pcu, in_coef=coef, in_silu=True, in_up=True))
    assert new.shape == (n, 2 * h, 2 * w, c) and torch.equal(new, old)
    assert torch.allclose(_acc_to_sums(acc_new), _acc_to_sums(acc_old), rtol=1e-5, atol=1e-3)
    # into a channel slice as well (decoder concatenation buffers): dx * ld addressing of the 5-d map
    wide = torch.full((n, 2 * h, 2 * w, c + 128), 3.0, device=DEV, dtype=torch.bfloat16)
    ops.conv_acc(t, pcu, out=wide[..., :c], in_coef=coef, in_silu=True, in_up=True)
    assert torch.equal(wide[..., :c], new) and (wide[..., c:] == 3.0).all()
    # conv2: residual through the upsampling
    x2, wt2, b2 = _mk(n, 2 * h, 2 * w, c, c, 3, seed=53)
    pc2 = ops.pack_conv(wt2.float(), b2)
    (a, _), (o, _) = _both(lambda: ops.conv_acc(x2, pc2, residual=t, res_up=True))
    assert torch.equal(a, o)
    up = t.float().repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    _check(a, _ref(x2, wt2, b2, up), "res_up")


@pytest.mark.parametrize("shape", [(2, 16, 16, 256, 256, 3), (3, 32, 32, 128, 128, 3), (2, 16, 24, 128, 256, 1)])
def test_output_stored_through_nearest_upsampling(shape):
    """AzbConv.out_up: the row-domain epilogue stores every staged block four times, to (2 h + dy, 2 w + dx) of a channel
    slice of a wider (n, 2 h, 2 w, .) buffer -- the nn.Upsample(2, nearest) after the last block of an ascent level of the
    in-repo U-Net (azula/nn/unet.py:186-190) without its own pass; gate + residual epilogue as in that block."""
    from ctypes import byref

    from azula_b200 import _lib

    n, h, w, ci, co, k = shape
    x, wt, b = _mk(n, h, w, ci, co, k, seed=51)
    g = torch.Generator(device=DEV).manual_seed(9)
    gate = torch.randn(n, co, device=DEV, generator=g)
    res = torch.randn(n, h, w, co, device=DEV, generator=g).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    ops.conv_tuning(ops.KNOB_BLOCKN, 256 if co % 256 == 0 else 128)  # (small test problems would get 64-column tiles)
    wide = torch.full((n, 2 * h, 2 * w, co + 64), 7.0, device=DEV, dtype=torch.bfloat16)
    d = ops.conv_desc(x, pc, wide[..., :co], gate=gate.data_ptr(), gate_ld=gate.stride(0), gate_rows=h * w, residual=res, out_up=True)
    ch = ops.conv_choice(d)
    assert ch.epi == 2 and ch.block_n >= 128
    _lib.check(_lib.lib().azb_conv_bf16(byref(d), _lib.stream_ptr(torch.device(DEV))), "azb_conv_bf16")
    plain = ops.conv2d(x, pc, gate=gate, residual=res)
    up = plain.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    assert torch.equal(wide[..., :co], up), (wide[..., :co].float() - up.float()).abs().max().item()
    assert (wide[..., co:] == 7.0).all()
