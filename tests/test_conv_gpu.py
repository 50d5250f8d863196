"""GPU parity of the tcgen05 implicit-GEMM convolution (azb_conv_gemm_bf16) through the C ABI.

Checker: torch fp32 convolution (TF32 off) of the SAME bf16-rounded operands -- the only
differences left are fp32 accumulation order and the final bf16 rounding of the output, so the
stated tolerance is 2^-8 relative to the output scale (bf16 has 8 significand bits).
"""

import pytest
import torch
import torch.nn.functional as F

from azula_b200.engine import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _mk(n, h, w, ci, co, k, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(n, h, w, ci, device=DEV, generator=g).to(torch.bfloat16)
    wt = (torch.randn(co, ci, k, k, device=DEV, generator=g) / (ci * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(co, device=DEV, generator=g)
    return x, wt, b


def _ref(x, wt, b, residual=None):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=wt.shape[-1] // 2)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    return y


def _check(got, ref, what):
    err = (got.float() - ref).abs()
    tol = 2.0**-8 * ref.abs() + 2.0**-8 * ref.abs().mean()
    bad = (err > tol).sum().item()
    assert bad == 0, (what, bad, err.max().item(), ref.abs().mean().item())


SHAPES = [
    # n, h, w, c_in, c_out, k
    (2, 16, 16, 64, 128, 3),
    (1, 8, 8, 128, 128, 3),      # patch spans 2 images, batch 1 -> TMA out-of-bounds rows
    (2, 32, 32, 256, 256, 3),
    (1, 64, 64, 64, 64, 3),      # N tile 64
    (2, 16, 16, 32, 32, 3),      # C_in < 64: zero-padded K, N tile 32
    (2, 16, 16, 64, 16, 3),      # N tile 16
    (2, 16, 16, 64, 8, 1),       # fewer output channels than the narrowest tile
    (3, 4, 4, 64, 64, 3),
    (5, 2, 2, 128, 256, 3),
    (2, 12, 12, 64, 128, 3),     # extent not a multiple of the patch
    (1, 20, 24, 192, 128, 3),
    (2, 16, 16, 128, 384, 1),    # 1x1 (attention qkv / skip connection)
    (1, 32, 32, 512, 512, 1),
    (1, 128, 128, 256, 256, 3),
    (1, 16, 16, 768, 512, 3),    # concatenated input
]


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_matches_torch(shape):
    n, h, w, ci, co, k = shape
    x, wt, b = _mk(*shape)
    pc = ops.pack_conv(wt.float(), b)
    got = ops.conv(x, pc)
    torch.cuda.synchronize()
    assert got.shape == (n, h, w, co)
    _check(got, _ref(x, wt, b), shape)


def test_conv_residual_and_strided_output():
    """Epilogue adds a residual and writes into a channel slice of a wider (concat) buffer."""
    n, h, w, ci, co = 2, 16, 16, 128, 128
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=3)
    res = torch.randn(n, h, w, co, device=DEV).to(torch.bfloat16)
    wide = torch.zeros(n, h, w, co + 64, device=DEV, dtype=torch.bfloat16)
    view = wide[..., 64:]
    ops.conv(x, ops.pack_conv(wt.float(), b), out=view, residual=res)
    torch.cuda.synchronize()
    _check(view, _ref(x, wt, b, res), "residual+slice")
    assert (wide[..., :64] == 0).all()
    # strided input: read the convolution input from a channel slice as well
    xin = torch.zeros(n, h, w, ci + 64, device=DEV, dtype=torch.bfloat16)
    xin[..., :ci] = x
    got = ops.conv(xin[..., :ci], ops.pack_conv(wt.float(), b))
    _check(got, _ref(x, wt, b), "strided input")


def test_conv_network_output_nchw_fp32():
    """The UNet's last conv (256 -> 6, learn_var) written as fp32 NCHW, N tile 16."""
    n, h, w, ci, co = 2, 32, 32, 256, 6
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=5)
    got = ops.conv(x, ops.pack_conv(wt.float(), b), nchw_f32=True)
    torch.cuda.synchronize()
    ref = _ref(x, wt, b).permute(0, 3, 1, 2)
    assert got.shape == (n, co, h, w) and got.dtype == torch.float32
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got - ref).abs().max().item()


def test_linear_layer_rows():
    """nn.Linear as a 1-tap GEMM over rows (time-embedding MLP, attention projections)."""
    for rows, ci, co in ((1, 256, 1024), (16, 1024, 512), (300, 64, 192)):
        g = torch.Generator(device=DEV).manual_seed(rows)
        x = torch.randn(rows, ci, device=DEV, generator=g).to(torch.bfloat16)
        wt = (torch.randn(co, ci, device=DEV, generator=g) / ci**0.5).to(torch.bfloat16)
        b = torch.randn(co, device=DEV, generator=g)
        got = ops.conv(x, ops.pack_conv(wt.float(), b))
        torch.cuda.synchronize()
        _check(got, F.linear(x.float(), wt.float(), b), (rows, ci, co))


def _group_stats(t, groups=32, eps=1e-5):
    """(N, groups, 2) = (mean, rstd) of an NHWC tensor, as torch.nn.GroupNorm defines them."""
    n, c = t.shape[0], t.shape[-1]
    grp = t.float().permute(0, 3, 1, 2).reshape(n, groups, -1)
    return torch.stack((grp.mean(-1), torch.rsqrt(grp.var(-1, unbiased=False) + eps)), dim=-1)


@pytest.mark.parametrize("shape", [
    (2, 16, 16, 64, 128, 3), (3, 8, 8, 128, 256, 3), (1, 64, 64, 64, 64, 3), (2, 32, 32, 256, 256, 3),
    (2, 12, 12, 64, 128, 3), (1, 20, 24, 192, 128, 3), (2, 16, 16, 128, 384, 1), (5, 8, 8, 64, 512, 1),
    (16, 8, 8, 256, 1024, 3), (2, 16, 16, 32, 32, 3),
])
@pytest.mark.parametrize("gran", [1, 8])
def test_conv_epilogue_statistics(shape, gran):
    """azb_conv_gemm_stats_bf16 + azb_gn_finalize_f32 == GroupNorm statistics of the stored output,
    and the output itself equals the plain convolution's (bit for bit when both run the same kernel class)."""
    n, h, w, ci, co, k = shape
    if gran == 8 and (co // 32) % 8:
        pytest.skip("GroupNorm groups are not multiples of 8 channels")
    x, wt, b = _mk(*shape, seed=7)
    pc = ops.pack_conv(wt.float(), b)
    res = torch.randn(n, h, w, co, device=DEV).to(torch.bfloat16)
    rows, ok = ops.colsum_rows(n, h, w)
    assert ok
    colsum = torch.full((rows, co // gran, 2), float("nan"), device=DEV)
    plain = ops.conv(x, pc, residual=res)
    got = ops.conv(x, pc, residual=res, colsum=colsum)
    if not torch.equal(got, plain):
        # the plain launch may take halo tiles (another K order than the tap-wise kernel that serves `colsum`): same
        # products, fp32 summation order differs -> isolated single roundings
        d = (got.float() - plain.float()).abs()
        assert (d > 0).float().mean().item() < 0.02 and d.max().item() <= 2.0**-7 * plain.float().abs().max().item(), shape
    stats = ops.gn_finalize([(colsum, co)], n, h, w)
    ref = _group_stats(got)
    assert torch.allclose(stats[..., 0], ref[..., 0], atol=2e-5, rtol=1e-4), (stats[..., 0] - ref[..., 0]).abs().max()
    assert torch.allclose(stats[..., 1], ref[..., 1], rtol=1e-4)
    assert torch.equal(stats, ops.gn_finalize([(colsum, co)], n, h, w))  # deterministic


def test_statistics_of_a_concatenation():
    """Two producers write the halves of a decoder concat buffer; groups straddle the boundary
    (1024 + 512 channels -> 48 channels per group)."""
    n, h, w = 2, 16, 16
    ca, cb = 1024, 512
    wide = torch.empty(n, h, w, ca + cb, device=DEV, dtype=torch.bfloat16)
    rows, _ = ops.colsum_rows(n, h, w)
    xa, wa, ba = _mk(n, h, w, 128, ca, 1, seed=1)
    xb, wb, bb = _mk(n, h, w, 64, cb, 3, seed=2)
    ref = None
    for ga, gb in ((1, 1), (8, 8), (8, 1)):
        sa, sb = torch.empty(rows, ca // ga, 2, device=DEV), torch.empty(rows, cb // gb, 2, device=DEV)
        ops.conv(xa, ops.pack_conv(wa.float(), ba), out=wide[..., :ca], colsum=sa)
        ops.conv(xb, ops.pack_conv(wb.float(), bb), out=wide[..., ca:], colsum=sb)
        stats = ops.gn_finalize([(sa, ca), (sb, cb)], n, h, w)
        ref = _group_stats(wide) if ref is None else ref
        assert torch.allclose(stats[..., 0], ref[..., 0], atol=2e-5, rtol=1e-4), (ga, gb)
        assert torch.allclose(stats[..., 1], ref[..., 1], rtol=1e-4), (ga, gb)


def test_conv_throughput_report():
    """Not a pass/fail perf gate: prints achieved TFLOP/s of the heaviest ADM shapes."""
    for (n, h, w, ci, co) in ((4, 256, 256, 256, 256), (16, 64, 64, 512, 512), (16, 32, 32, 512, 512),
                              (16, 16, 16, 1024, 1024), (16, 8, 8, 1024, 1024)):
        x, wt, b = _mk(n, h, w, ci, co, 3)
        pc = ops.pack_conv(wt.float(), b)
        out = torch.empty(n, h, w, co, device=DEV, dtype=torch.bfloat16)
        for _ in range(3):
            ops.conv(x, pc, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv(x, pc, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * n * h * w * co * ci * 9 / (ms * 1e-3) / 1e12
        print(f"conv3x3 {n}x{h}x{w} {ci}->{co}: {ms:.3f} ms, {tf:.0f} TFLOP/s")


@pytest.mark.parametrize("n,h,w,ci,ci2,co", [(2, 16, 16, 64, 128, 64), (1, 32, 32, 256, 512, 256), (3, 8, 8, 128, 192, 128),
                                            (2, 16, 16, 32, 64, 32)])
def test_conv_with_fused_1x1_skip(n, h, w, ci, ci2, co):
    """out = conv3x3(h) + conv1x1(x) + biases as one GEMM (the ResBlock tail, _src/unet.py:243-247), with the
    GroupNorm column sums of the result."""
    g = torch.Generator(device=DEV).manual_seed(21)
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=5)
    x2 = torch.randn(n, h, w, ci2, device=DEV, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(co, ci2, 1, 1, device=DEV, generator=g) / ci2**0.5).to(torch.bfloat16)
    b2 = torch.randn(co, device=DEV, generator=g)
    pc = ops.pack_conv_skip(ops.pack_conv(wt.float(), b), ops.pack_conv(w2.float(), b2))
    ref = _ref(x, wt, b) + _ref(x2, w2, b2)
    got = ops.conv_skip(x, x2, pc)
    _check(got, ref, "conv + skip")
    rows, ok = ops.colsum_rows(n, h, w)
    if ok:
        colsum = torch.zeros(rows, co, 2, device=DEV)
        got2 = ops.conv_skip(x, x2, pc, colsum=colsum)
        assert torch.equal(got, got2)
        tot = colsum.sum(0)
        gf = got.float().reshape(-1, co)
        assert torch.allclose(tot[:, 0], gf.sum(0), rtol=1e-4, atol=1e-2)
        assert torch.allclose(tot[:, 1], gf.square().sum(0), rtol=1e-4, atol=1e-2)


def _acc_to_sums(acc):
    """int64 (N, blocks, 4) fixed point -> float64 (N, blocks, 2) = (sum, sum of squares)."""
    a = acc.to(torch.float64)
    return torch.stack(((a[..., 0] * 2.0**32 + a[..., 1]) / 2.0**40, (a[..., 2] * 2.0**32 + a[..., 3]) / 2.0**40), dim=-1)


@pytest.mark.parametrize("n,h,w,ci,co,gran", [(2, 16, 16, 64, 128, 8), (3, 32, 32, 128, 256, 8), (2, 8, 8, 32, 32, 1),
                                             (16, 8, 8, 256, 512, 8), (1, 64, 64, 64, 64, 8)])
def test_conv_exact_groupnorm_accumulators(n, h, w, ci, co, gran):
    """The epilogue's integer-atomic {sum, sumsq} per (image, channel block): equal to the sums of the stored
    values, and bit-identical from run to run (integer additions commute)."""
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=9)
    res = torch.randn(n, h, w, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2)).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    out, acc = ops.conv_acc(x, pc, residual=res, gran=gran)
    _check(out, _ref(x, wt, b, res), "conv with accumulators")
    got = _acc_to_sums(acc)
    o = out.double().reshape(n, h * w, co // gran, gran)
    want = torch.stack((o.sum(dim=(1, 3)), o.square().sum(dim=(1, 3))), dim=-1)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-3), (got - want).abs().max()
    for _ in range(3):
        out2, acc2 = ops.conv_acc(x, pc, residual=res, gran=gran)
        assert torch.equal(acc, acc2) and torch.equal(out, out2)


@pytest.mark.parametrize("split", [None, 512, 256])
def test_gn_apply_from_accumulators(split):
    """GroupNorm(32) + SiLU with statistics folded from one or two producers' accumulators == torch group_norm
    (768 channels: groups of 24, so with the 512 | 256 split group 21 straddles the two producers)."""
    n, h, w, c = 3, 16, 16, 768
    g = torch.Generator(device=DEV).manual_seed(4)
    parts, outs = [], []
    for cc in ([c] if split is None else [split, c - split]):
        x, wt, b = _mk(n, h, w, 64, cc, 3, seed=cc)
        out, acc = ops.conv_acc(x, ops.pack_conv(wt.float() * 3, b))
        parts.append((acc, cc)), outs.append(out)
    t = torch.cat(outs, dim=-1).contiguous()
    gamma, beta = 1 + 0.1 * torch.randn(c, device=DEV, generator=g), 0.1 * torch.randn(c, device=DEV, generator=g)
    ss = 0.2 * torch.randn(n, 2 * c, device=DEV, generator=g)
    got = ops.gn_apply_acc(t, parts, gamma, beta, scale_shift=ss)
    ref = F.group_norm(t.float().permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5)
    ref = F.silu(ref * (1 + ss[:, :c, None, None]) + ss[:, c:, None, None]).permute(0, 2, 3, 1)
    _check(got, ref, f"gn_apply_acc split={split}")


@pytest.mark.parametrize("n,h,w,ci,co,skip", [(16, 8, 8, 1024, 1024, 0), (16, 8, 8, 2048, 1024, 0), (4, 8, 8, 512, 512, 0),
                                             (16, 8, 8, 1024, 1024, 1536), (2, 16, 16, 1024, 1024, 0)])
def test_conv_split_k(n, h, w, ci, co, skip):
    """Small maps with long reductions: with a workspace the kernel splits K over CTAs; same result as the unsplit
    kernel up to fp32 summation order, accumulators included, and reproducible."""
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=31)
    res = torch.randn(n, h, w, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    x2 = None
    ref = _ref(x, wt, b, None if skip else res)
    if skip:
        g = torch.Generator(device=DEV).manual_seed(5)
        x2 = torch.randn(n, h, w, skip, device=DEV, generator=g).to(torch.bfloat16)
        w2 = (torch.randn(co, skip, 1, 1, device=DEV, generator=g) / skip**0.5).to(torch.bfloat16)
        b2 = torch.randn(co, device=DEV, generator=g)
        pc = ops.pack_conv_skip(pc, ops.pack_conv(w2.float(), b2))
        ref = ref + _ref(x2, w2, b2)
    ws = ops.splitk_workspace(DEV)
    kw = dict(residual=None if skip else res, x2=x2)
    plain, acc0 = ops.conv_acc(x, pc, **kw)
    split, acc1 = ops.conv_acc(x, pc, workspace=ws, **kw)
    _check(plain, ref, "unsplit")
    _check(split, ref, "split-K")
    assert not ws[:256].any(), "flags must be left zeroed"
    s0, s1 = _acc_to_sums(acc0), _acc_to_sums(acc1)
    assert torch.allclose(s0, s1, rtol=2e-2, atol=2.0)
    for _ in range(3):
        again, acc2 = ops.conv_acc(x, pc, workspace=ws, **kw)
        assert torch.equal(again, split) and torch.equal(acc1, acc2)


@pytest.fixture
def force_pairs():
    ops.conv_tuning(ops.KNOB_PAIR, 1)
    yield
    ops.conv_tuning(ops.KNOB_PAIR, -1)


PAIR_SHAPES = [
    # n, h, w, c_in, c_out, k: an even number of 128-pixel tiles and C_out a multiple of 128
    (2, 16, 16, 64, 128, 3),     # 4 M tiles, one pair-unit per N tile: the smallest case
    (1, 32, 32, 128, 256, 3),    # N = 256 across the pair
    (2, 32, 32, 256, 256, 3),
    (4, 8, 8, 128, 128, 3),      # patch spans 2 images
    (1, 64, 64, 320, 512, 3),    # 2 N tiles of 256, many units per pair (persistent loop, both accumulator stages)
    (1, 32, 32, 512, 384, 1),    # 1x1, N tile 128
    (3, 48, 40, 192, 256, 3),    # extents not multiples of the patch: out-of-bounds rows in both CTAs
]


@pytest.mark.parametrize("shape", PAIR_SHAPES)
def test_conv_cta_pairs_match_single_cta(shape, force_pairs):
    """tcgen05.mma.cta_group::2 path: same accumulation order per output element as the single-CTA kernel (K is
    traversed identically), so the two must agree BIT FOR BIT -- output, residual add and GroupNorm accumulators."""
    n, h, w, ci, co, k = shape
    x, wt, b = _mk(*shape, seed=41)
    res = torch.randn(n, h, w, co, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9)).to(torch.bfloat16)
    pc = ops.pack_conv(wt.float(), b)
    pair, acc_p = ops.conv_acc(x, pc, residual=res, gran=1)
    ops.conv_tuning(ops.KNOB_PAIR, 0)
    single, acc_s = ops.conv_acc(x, pc, residual=res, gran=1)
    torch.cuda.synchronize()
    _check(single, _ref(x, wt, b, res), ("single", shape))
    _check(pair, _ref(x, wt, b, res), ("pair", shape))
    assert torch.equal(pair, single)
    assert torch.equal(acc_p, acc_s)


def test_conv_cta_pairs_with_fused_skip_and_chunked_stats(force_pairs):
    """The ADM decoder shape class: 3x3 + fused 1x1 skip operand, 8-channel-block sums carried across tiles."""
    n, h, w, ci, co, skip = 2, 64, 64, 256, 256, 512
    x, wt, b = _mk(n, h, w, ci, co, 3, seed=43)
    g = torch.Generator(device=DEV).manual_seed(5)
    x2 = torch.randn(n, h, w, skip, device=DEV, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(co, skip, 1, 1, device=DEV, generator=g) / skip**0.5).to(torch.bfloat16)
    b2 = torch.randn(co, device=DEV, generator=g)
    pc = ops.pack_conv_skip(ops.pack_conv(wt.float(), b), ops.pack_conv(w2.float(), b2))
    pair, acc_p = ops.conv_acc(x, pc, x2=x2)
    ops.conv_tuning(ops.KNOB_PAIR, 0)
    single, acc_s = ops.conv_acc(x, pc, x2=x2)
    _check(pair, _ref(x, wt, b) + _ref(x2, w2, b2), "pair + skip")
    assert torch.equal(pair, single)
    # carried fp32 partial sums depend on which tiles a CTA owns: equal up to fp32 rounding of the partial sums
    assert torch.allclose(_acc_to_sums(acc_p), _acc_to_sums(acc_s), rtol=1e-4, atol=1e-2)
