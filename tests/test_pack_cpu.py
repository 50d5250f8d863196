"""Host-side weight packing, checked on CPU with plain torch (no kernel involved).

`ops.pack_conv_up` turns conv3x3(upsample2x(z)) (the conv1 of an upsampling ResBlock, azula/plugins/adm/_src/unet.py:
101-109,229-233) into four 2 x 2 convolutions of z, one per output phase (dy, dx), by summing the taps that fall on the
same half-resolution pixel.  The packed layout is what `azb_conv_bf16` with `AzbConv::in_up = 2` consumes."""

import torch
import torch.nn.functional as F

from azula_b200.engine import ops


def test_phase_weights_reproduce_the_upsampled_convolution():
    g = torch.Generator().manual_seed(0)
    n, ci, co, h, w = 2, 64, 16, 6, 5
    z = torch.randn(n, ci, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(co, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(z, scale_factor=2, mode="nearest"), wt, b, padding=1)

    pc = ops.pack_conv_up(wt.float(), b.float())
    assert pc.taps == 16 and pc.w.shape == (16, 16, 64) and pc.w.dtype == torch.bfloat16
    # the same sums in float64 (the packed tensor holds their bf16 roundings)
    sel = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    zp = F.pad(z, (1, 1, 1, 1))
    out = torch.zeros(n, co, 2 * h, 2 * w, dtype=torch.float64)
    for dy in (0, 1):
        for dx in (0, 1):
            k = torch.zeros(co, ci, 2, 2, dtype=torch.float64)
            for a in (0, 1):
                for bb in (0, 1):
                    k[:, :, a, bb] = sum(wt[:, :, i, j] for i in sel[dy][a] for j in sel[dx][bb])
                    q = (dy * 2 + dx) * 4 + a * 2 + bb
                    assert torch.allclose(pc.w[:co, q, :ci].double(), k[:, :, a, bb], rtol=2.0**-8, atol=1e-6)
            # output pixel (2 i + dy, 2 j + dx) reads half-resolution rows i - 1 + dy, i + dy and columns j - 1 + dx, j + dx
            y = F.conv2d(zp[:, :, dy : dy + h + 1, dx : dx + w + 1], k, b)
            out[:, :, dy::2, dx::2] = y
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-10)


def test_pack_conv_layouts():
    g = torch.Generator().manual_seed(1)
    wt = torch.randn(24, 40, 3, 3, generator=g)
    pc = ops.pack_conv(wt, None)
    assert pc.taps == 9 and pc.k_per_tap == 64 and pc.c_out_rows == 32 and pc.bias is None
    assert torch.equal(pc.w[:24, 4, :40], wt[:, :, 1, 1].to(torch.bfloat16)) and not pc.w[24:].any() and not pc.w[:, :, 40:].any()
    skip = ops.pack_conv(torch.randn(24, 72, 1, 1, generator=g), torch.zeros(24))
    both = ops.pack_conv_skip(ops.pack_conv(wt, torch.ones(24)), skip)
    assert both.w.shape == (32, 9 * 64 + 128) and both.k2 == 128 and both.c_in2 == 72 and torch.equal(both.bias, torch.ones(24))
