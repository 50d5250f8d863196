"""One convolution shape under a chosen launcher configuration (for ncu captures).

    ncu --set full -k regex:conv_gemm -c 2 python scripts/conv_one.py --hw 256 --ci 512 --co 256 --pair 1
"""

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--ci", type=int, default=256)
    ap.add_argument("--co", type=int, default=256)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--pair", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--halo", type=int, default=-1)
    ap.add_argument("--fused", type=int, default=0, help="apply act(a x + b) to the halo tiles (in_coef)")
    ap.add_argument("--skip", type=int, default=0, help="channels of a fused 1x1 operand")
    ap.add_argument("--sa", type=int, default=-1)
    ap.add_argument("--ahead", type=int, default=-1, help="0: fused 1x1 blocks after the last halo item (default: spread)")
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(a.batch, a.hw, a.hw, a.ci, device=dev, generator=g).to(torch.bfloat16)
    pc = ops.pack_conv(torch.randn(a.co, a.ci, a.k, a.k, device=dev, generator=g) / (a.ci * a.k * a.k) ** 0.5,
                       torch.randn(a.co, device=dev, generator=g))
    out = torch.empty(a.batch, a.hw, a.hw, a.co, dtype=torch.bfloat16, device=dev)
    x2 = None
    if a.skip:
        x2 = torch.randn(a.batch, a.hw, a.hw, a.skip, device=dev, generator=g).to(torch.bfloat16)
        pc = ops.pack_conv_skip(pc, ops.pack_conv(torch.randn(a.co, a.skip, 1, 1, device=dev, generator=g) / a.skip**0.5,
                                                  torch.randn(a.co, device=dev, generator=g)))
    ops.conv_tuning(ops.KNOB_PAIR, a.pair)
    ops.conv_tuning(ops.KNOB_HALO, a.halo)
    ops.conv_tuning(ops.KNOB_HALO_SA, a.sa)
    ops.conv_tuning(ops.KNOB_HALO_SPREAD, a.ahead)
    coef = None
    if a.fused:
        coef = torch.stack((torch.full((a.batch, a.ci), 0.5, device=dev), torch.zeros(a.batch, a.ci, device=dev)), dim=-1).contiguous()
    for _ in range(a.reps):
        ops.conv_acc(x, pc, out=out, x2=x2, in_coef=coef, in_silu=True)
    torch.cuda.synchronize()
    print("ok", float(out.float().abs().mean()))


if __name__ == "__main__":
    main()
