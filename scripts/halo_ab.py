"""A/B of the 3 x 3 convolution variants on ADM shapes (one process, CUDA events, median of reps):

    tap      tap-wise TMA loads (AZB_CONV_KNOB_HALO = 0), normalised input read from HBM
    halo     halo tiles, normalised input read from HBM
    fused    halo tiles + GroupNorm/SiLU transform in the A path (raw input + coefficients)

and of the halo kernels' ring split (A slots / early A loads).  python scripts/halo_ab.py [--reps 20]
"""

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import ops  # noqa: E402

DEV = "cuda"

SHAPES = [
    # n, hw, c_in, c_out, skip channels (fused 1x1 operand), residual
    (16, 256, 256, 256, 0, False),
    (16, 256, 256, 256, 0, True),
    (16, 256, 512, 256, 0, False),
    (16, 256, 256, 256, 512, False),
    (16, 128, 512, 512, 0, True),
    (16, 128, 256, 256, 512, False),
    (16, 64, 512, 512, 0, True),
    (16, 32, 1024, 1024, 0, True),
    (16, 16, 1024, 1024, 0, True),
]


def timed(fn, reps):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--configs", default="3:0,4:0,4:1,2:0")
    a = ap.parse_args()
    g = torch.Generator(device=DEV).manual_seed(0)
    configs = [tuple(int(v) for v in c.split(":")) for c in a.configs.split(",")]
    print("shape".ljust(44) + "tap".rjust(9) + "".join(f"{'halo' + str(c):>14}{'fused' + str(c):>14}" for c in configs) + "   (TFLOP/s)")
    for n, hw, ci, co, skip, res in SHAPES:
        x = torch.randn(n, hw, hw, ci, device=DEV, generator=g).to(torch.bfloat16)
        wt = torch.randn(co, ci, 3, 3, device=DEV, generator=g) / (9 * ci) ** 0.5
        pc = ops.pack_conv(wt, torch.randn(co, device=DEV, generator=g))
        x2 = None
        if skip:
            x2 = torch.randn(n, hw, hw, skip, device=DEV, generator=g).to(torch.bfloat16)
            pc = ops.pack_conv_skip(pc, ops.pack_conv(torch.randn(co, skip, 1, 1, device=DEV, generator=g) / skip**0.5,
                                                      torch.randn(co, device=DEV, generator=g)))
        r = torch.randn(n, hw, hw, co, device=DEV, generator=g).to(torch.bfloat16) if res else None
        out = torch.empty(n, hw, hw, co, dtype=torch.bfloat16, device=DEV)
        coef = torch.stack((torch.full((n, ci), 0.5, device=DEV), torch.zeros(n, ci, device=DEV)), dim=-1).contiguous()
        flops = 2.0 * n * hw * hw * co * (9 * ci + skip)
        run = lambda c=None: ops.conv_acc(x, pc, out=out, residual=r, x2=x2, in_coef=c, in_silu=True)  # noqa: E731
        ops.conv_tuning(ops.KNOB_HALO, 0)
        row = [flops / timed(run, a.reps) * 1e-9]
        ops.conv_tuning(ops.KNOB_HALO, -1)
        for sa, ahead in configs:
            ops.conv_tuning(ops.KNOB_HALO_SA, sa)
            ops.conv_tuning(ops.KNOB_HALO_AHEAD, ahead)
            row.append(flops / timed(run, a.reps) * 1e-9)
            row.append(flops / timed(lambda: run(coef), a.reps) * 1e-9)
        ops.conv_tuning(ops.KNOB_HALO_SA, -1)
        ops.conv_tuning(ops.KNOB_HALO_AHEAD, -1)
        name = f"{n}x{hw}x{hw} {ci}->{co}" + (f" +skip{skip}" if skip else "") + (" +res" if res else "")
        print(name.ljust(44) + f"{row[0]:9.0f}" + "".join(f"{v:14.0f}" for v in row[1:]))
        del x, x2, r, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
