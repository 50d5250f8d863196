"""A/B of the convolution variants on the in-repo U-Net's block shapes (event-timed, 20 launches back to back):
tap-wise / halo tiles x transposing / row-domain epilogue x plain / +act / +gate+res / +sums / norm fused.
python scripts/unet_conv_ab.py [--c 64 --hw 64 --n 32]"""
import argparse
import os
import sys
from ctypes import byref

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azula_b200 import _lib  # noqa: E402
from azula_b200.engine import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c", type=int, default=64)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--only", default="", help="one variant, launched twice (for ncu): name,halo,rowepi e.g. plain,-1,-1")
    a = ap.parse_args()
    dev = "cuda"
    n, h, c = a.n, a.hw, a.c
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
    res = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
    pc = ops.pack_conv(torch.randn(c, c, 3, 3, device=dev, generator=g) / (9 * c) ** 0.5, torch.randn(c, device=dev, generator=g))
    out = torch.empty(n, h, h, c, dtype=torch.bfloat16, device=dev)
    gate = torch.randn(n, c, device=dev, generator=g)
    mod = torch.randn(n, 3 * c, device=dev, generator=g) * 0.1
    stat_in = torch.rand(n * h * h, c // 64, 2, device=dev) + torch.tensor([0.0, 64.0], device=dev)
    stat_out = torch.empty(n * h * h, c // 64, 2, device=dev)
    s = _lib.stream_ptr(torch.device(dev))
    variants = {
        "plain": {},
        "+act": dict(act=ops.ACT["silu"]),
        "+gate+res": dict(gate=gate.data_ptr(), gate_ld=gate.stride(0), gate_rows=h * h, residual=res),
        "+gate+res+sums": dict(gate=gate.data_ptr(), gate_ld=gate.stride(0), gate_rows=h * h, residual=res, rowstat=stat_out),
        "norm+act": dict(act=ops.ACT["silu"], in_norm=1, in_rowstat=stat_in, in_mod=mod.data_ptr(), in_mod_ld=mod.stride(0)),
    }
    if a.only:
        name, halo, rowepi = a.only.split(",")
        ops.conv_tuning(ops.KNOB_HALO, int(halo))
        ops.conv_tuning(ops.KNOB_ROWEPI, int(rowepi))
        d = ops.conv_desc(x, pc, out, **variants[name])
        for _ in range(2):
            _lib.check(_lib.lib().azb_conv_bf16(byref(d), s), "conv")
        torch.cuda.synchronize()
        return
    for halo in (0, -1):
        for rowepi in (-1, 1):
            ops.conv_tuning(ops.KNOB_HALO, halo)
            ops.conv_tuning(ops.KNOB_ROWEPI, rowepi)
            for name, kw in variants.items():
                d = ops.conv_desc(x, pc, out, **kw)
                ch = ops.AzbConvChoice()
                if _lib.lib().azb_conv_choice(byref(d), byref(ch)) != 0:
                    continue
                for _ in range(3):
                    _lib.check(_lib.lib().azb_conv_bf16(byref(d), s), "conv")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    _lib.lib().azb_conv_bf16(byref(d), s)
                e1.record()
                torch.cuda.synchronize()
                us = 1e3 * e0.elapsed_time(e1) / 20
                print(f"c {c} hw {h} knobs halo {halo:2d} rowepi {rowepi:2d}  {name:16s} -> halo {ch.halo} pair {ch.pair} N {ch.block_n:3d} epi {ch.epi}: "
                      f"{us:6.1f} us {2.0 * n * h * h * c * c * 9 / us / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
