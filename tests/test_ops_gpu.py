"""GPU parity of the non-GEMM kernels of the native ADM backbone, through the C ABI.

Checker: the torch ops the reference calls at the cited sites, evaluated in fp32 on the same
bf16-rounded inputs; tolerance = bf16 output rounding (2^-8 relative to the output scale).
"""

import math

import pytest
import torch
import torch.nn.functional as F

from azula_b200.engine import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, what, rel=2.0**-7):
    err = (got.float() - ref).abs()
    tol = rel * ref.abs() + rel * ref.abs().mean() + 1e-6
    bad = (err > tol).sum().item()
    assert bad == 0, (what, bad, err.max().item(), ref.abs().mean().item())


@pytest.fixture(scope="module")
def scratch():
    return ops.GroupNormScratch(DEV)


@pytest.mark.parametrize("shape", [(2, 16, 16, 256), (1, 64, 64, 512), (3, 8, 8, 1024), (2, 4, 4, 2048),
                                   (2, 16, 16, 32), (2, 8, 8, 96), (1, 32, 32, 768), (2, 2, 2, 64), (1, 128, 128, 256)])
def test_group_norm_stats_and_apply(shape, scratch):
    n, h, w, c = shape
    g = torch.Generator(device=DEV).manual_seed(c)
    x = (torch.randn(shape, device=DEV, generator=g) * 1.7 + 0.6).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(c, device=DEV, generator=g)
    beta = 0.1 * torch.randn(c, device=DEV, generator=g)
    stats = ops.gn_stats(x, scratch)
    xf = x.float().permute(0, 3, 1, 2)
    grp = xf.reshape(n, 32, -1)
    assert torch.allclose(stats[..., 0], grp.mean(-1), atol=1e-4, rtol=1e-4)
    assert torch.allclose(stats[..., 1], torch.rsqrt(grp.var(-1, unbiased=False) + 1e-5), rtol=1e-3)
    # deterministic
    assert torch.equal(stats, ops.gn_stats(x, scratch))

    ref = F.silu(F.group_norm(xf, 32, gamma, beta, 1e-5))
    got = ops.gn_apply(x, stats=stats, gamma=gamma, beta=beta, silu=True)
    _close(got, ref.permute(0, 2, 3, 1), ("silu(gn)", shape))

    # scale/shift conditioning (ResBlock.out_layers with use_scale_shift_norm): shared row and per-image rows
    for rows in (1, n):
        ss = 0.3 * torch.randn(rows, 2 * c, device=DEV, generator=g)
        sc, sh = ss[:, :c, None, None], ss[:, c:, None, None]
        ref = F.silu(F.group_norm(xf, 32, gamma, beta, 1e-5) * (1 + sc) + sh)
        got = ops.gn_apply(x, stats=stats, gamma=gamma, beta=beta, scale_shift=ss, silu=True)
        _close(got, ref.permute(0, 2, 3, 1), ("scale_shift", shape, rows))

    # resampling variants (ResBlock up/down: h_upd on the activated tensor, x_upd on the raw one)
    act = F.silu(F.group_norm(xf, 32, gamma, beta, 1e-5))
    got = ops.gn_apply(x, stats=stats, gamma=gamma, beta=beta, silu=True, mode=1)
    _close(got, F.interpolate(act, scale_factor=2, mode="nearest").permute(0, 2, 3, 1), ("up", shape))
    got = ops.gn_apply(x, stats=stats, gamma=gamma, beta=beta, silu=True, mode=2)
    _close(got, F.avg_pool2d(act, 2, 2).permute(0, 2, 3, 1), ("down", shape))
    got = ops.gn_apply(x, silu=False, mode=2)
    _close(got, F.avg_pool2d(xf, 2, 2).permute(0, 2, 3, 1), ("raw down", shape))
    got = ops.gn_apply(x, silu=False, mode=1)
    assert torch.equal(got, F.interpolate(xf, scale_factor=2, mode="nearest").permute(0, 2, 3, 1).to(torch.bfloat16))


def test_group_norm_on_channel_slices(scratch):
    """Statistics and apply read / write channel slices of wider (concatenation) buffers."""
    n, h, w, c = 2, 8, 8, 256
    wide = torch.randn(n, h, w, c + 128, device=DEV).to(torch.bfloat16)
    x = wide[..., 128:]
    gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    stats = ops.gn_stats(x, scratch)
    out_wide = torch.zeros(n, h, w, c + 64, device=DEV, dtype=torch.bfloat16)
    ops.gn_apply(x, out=out_wide[..., :c], stats=stats, gamma=gamma, beta=beta, silu=False)
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    _close(out_wide[..., :c], ref, "slice")
    assert (out_wide[..., c:] == 0).all()


def test_per_step_scale_shift_table(scratch):
    n, h, w, c, steps = 1, 8, 8, 64, 5
    x = torch.randn(n, h, w, c, device=DEV).to(torch.bfloat16)
    gamma, beta = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
    table = 0.5 * torch.randn(steps, 2 * c + 10, device=DEV)  # row stride larger than 2C: a slice of a bigger table
    stats = ops.gn_stats(x, scratch)
    idx = torch.zeros((), dtype=torch.int32, device=DEV)
    for s in (0, 3, 4):
        idx.fill_(s)
        row = table[s, : 2 * c].contiguous()[None]
        ref = ops.gn_apply(x, stats=stats, gamma=gamma, beta=beta, scale_shift=row)
        view = table[:, : 2 * c]
        out = torch.empty_like(ref)
        from azula_b200 import _lib
        _lib.check(_lib.lib().azb_gn_apply_bf16(
            x.data_ptr(), c, out.data_ptr(), c, n, h, w, c, 32, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
            view.data_ptr(), 0, idx.data_ptr(), table.stride(0), 1, 0, _lib.stream_ptr()))
        assert torch.equal(out, ref), s


@pytest.mark.parametrize("case", [(2, 64, 4, 16), (1, 256, 8, 64), (2, 1024, 8, 64), (3, 16, 16, 64), (2, 4, 2, 32),
                                  (1, 100, 2, 128), (2, 300, 3, 64), (1, 129, 1, 64), (5, 640, 2, 64), (1, 4096, 2, 64)])
@pytest.mark.parametrize("new_order", [False, True])
def test_attention_matches_reference_formula(case, new_order):
    n, t, heads, d = case
    c = heads * d
    g = torch.Generator(device=DEV).manual_seed(t + d)
    qkv = torch.randn(n, t, 3 * c, device=DEV, generator=g).to(torch.bfloat16)
    got = ops.attention(qkv, heads, new_order)
    # reference formula (QKVAttentionLegacy / QKVAttention) on (N, 3C, T) fp32
    x = qkv.float().permute(0, 2, 1)
    if new_order:
        q, k, v = (z.reshape(n * heads, d, t) for z in x.chunk(3, dim=1))
    else:
        q, k, v = x.reshape(n * heads, 3 * d, t).split(d, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, v).reshape(n, c, t).permute(0, 2, 1)
    _close(got, ref, (case, new_order), rel=2.0**-6)
    if d == 64:  # the tcgen05 kernel ran above; the warp-level kernel must agree with the same formula
        _close(ops.attention(qkv, heads, new_order, kernel="mma"), ref, (case, new_order, "mma"), rel=2.0**-6)
    # large logits: the softmax must stay finite and exact about its maximum
    big = (qkv.float() * 6).to(torch.bfloat16)
    assert torch.isfinite(ops.attention(big, heads, new_order).float()).all()


def test_first_conv_via_im2col():
    n, c, h, w, co = 2, 3, 32, 32, 256
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(n, c, h, w, device=DEV, generator=g)
    wt = torch.randn(co, c, 3, 3, device=DEV, generator=g) / 27**0.5
    b = torch.randn(co, device=DEV, generator=g)
    patches = ops.im2col3x3(x)
    got = ops.conv(patches, ops.pack_first_conv(wt, b))
    ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), b, padding=1).permute(0, 2, 3, 1)
    _close(got, ref, "first conv")


def test_time_embedding_path():
    torch.backends.cuda.matmul.allow_tf32 = False
    ts = torch.tensor([0, 3, 500, 999], dtype=torch.int64, device=DEV)
    dim = 256
    feats = ops.timestep_features(ts, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(DEV)
    ang = ts[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
    assert torch.allclose(feats, ref, atol=2e-4), (feats - ref).abs().max()
    g = torch.Generator(device=DEV).manual_seed(2)
    w1 = torch.randn(1024, dim, device=DEV, generator=g) / dim**0.5
    b1 = torch.randn(1024, device=DEV, generator=g)
    y = ops.linear_f32(feats, w1, b1)
    assert torch.allclose(y, F.linear(feats, w1, b1), atol=1e-4, rtol=1e-4)
    w2 = torch.randn(512, 1024, device=DEV, generator=g) / 32
    y2 = ops.linear_f32(y, w2, None, silu_in=True)
    ref2 = F.linear(F.silu(y), w2)
    assert torch.allclose(y2, ref2, atol=1e-3, rtol=1e-4)
    table = torch.randn(10, 1024, device=DEV, generator=g)
    idx = torch.tensor([1, 9, 0, 1], device=DEV)
    ref3 = y + table[idx]
    assert torch.allclose(ops.add_rows(y, table, idx), ref3)


@pytest.mark.parametrize("case", [(3, 256, 12, 64), (2, 64, 4, 64), (2, 100, 2, 64), (1, 200, 3, 64), (4, 16, 1, 64),
                                  (40, 256, 12, 64), (33, 100, 11, 64), (70, 64, 9, 64), (50, 129, 7, 64)])
def test_attention_with_folded_qk_norm(case):
    """azb_attention_qknorm_bf16 (persistent short-sequence kernel: logits and probabilities resident in tensor memory,
    q / k RMS-normalised in place in shared memory) == RMS-normalise q and k per head, then attention
    (azula/nn/attention.py:103,110-116), and the same kernel without the normalisation == plain attention; the
    two-launch route (azb_segment_rmsnorm_bf16 first) rounds q and k to bf16 like the fused one.  The large cases give
    every CTA several (image, head) items: the stage rings and the event-driven MMA issue order are exercised."""
    n, t, heads, d = case
    c = heads * d
    g = torch.Generator(device=DEV).manual_seed(t)
    qkv = (torch.randn(n, t, 3 * c, device=DEV, generator=g) * 2).to(torch.bfloat16)
    got = ops.attention_qknorm(qkv, heads, eps=1e-5)
    q, k, v = (z.reshape(n, t, heads, d).transpose(1, 2) for z in qkv.float().chunk(3, dim=-1))
    rms = lambda z: (z * torch.rsqrt(z.square().mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16).float()  # noqa: E731
    ref = (torch.softmax(rms(q) @ rms(k).transpose(-1, -2) / d**0.5, dim=-1) @ v).transpose(1, 2).reshape(n, t, c)
    _close(got, ref, (case, "folded"), rel=2.0**-6)
    two = qkv.clone()
    ops.segment_rmsnorm_(two.reshape(n * t, 3 * c), 2 * heads, d)
    _close(ops.attention(two, heads, True), ref, (case, "two launches"), rel=2.0**-6)
    plain = (torch.softmax(q @ k.transpose(-1, -2) / d**0.5, dim=-1) @ v).transpose(1, 2).reshape(n, t, c)
    _close(ops.attention(qkv, heads, True), plain, (case, "no normalisation"), rel=2.0**-6)
    assert torch.equal(ops.attention_qknorm(qkv, heads, eps=1e-5), got), "not reproducible"
