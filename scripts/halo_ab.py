"""A/B of the 3 x 3 convolution variants on ADM shapes (one process, CUDA events, median of reps):

    tap      tap-wise TMA loads (AZB_CONV_KNOB_HALO = 0), normalised input read from HBM
    halo     halo tiles, normalised input read from HBM
    fused    halo tiles + GroupNorm/SiLU transform in the A path (raw input + coefficients)

and of the halo kernels' ring split: --configs "sa:spread,..." = A slots : 1x1 blocks spread between the halo items.
python scripts/halo_ab.py [--reps 20]
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


def timed_interleaved(variants, reps):
    """variants: list of callables; one launch of each per round, round after round (robust against clock drift);
    returns the median milliseconds of each."""
    for fn in variants:
        for _ in range(2):
            fn()
    times = [[] for _ in variants]
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(variants) + 1)]
        ev[0].record()
        for i, fn in enumerate(variants):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        for i in range(len(variants)):
            times[i].append(ev[i].elapsed_time(ev[i + 1]))
    return [sorted(t)[len(t) // 2] for t in times]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=15)
    ap.add_argument("--configs", default="3:1,4:2,4:1,3:2")
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
        def variant(halo, sa, ahead, c):
            def fn():
                ops.conv_tuning(ops.KNOB_HALO, halo)
                ops.conv_tuning(ops.KNOB_HALO_SA, sa)
                ops.conv_tuning(ops.KNOB_HALO_SPREAD, ahead)
                ops.conv_acc(x, pc, out=out, residual=r, x2=x2, in_coef=c, in_silu=True)
            return fn

        # (the first launch after a synchronize runs ~5 % slow: a throw-away launch leads every round)
        variants = [variant(0, -1, -1, None), variant(0, -1, -1, None)]
        for sa, ahead in configs:
            variants += [variant(1, sa, ahead, None), variant(1, sa, ahead, coef)]
        row = [flops / t * 1e-9 for t in timed_interleaved(variants, a.reps)][1:]
        for k in (ops.KNOB_HALO, ops.KNOB_HALO_SA, ops.KNOB_HALO_SPREAD):
            ops.conv_tuning(k, -1)
        name = f"{n}x{hw}x{hw} {ci}->{co}" + (f" +skip{skip}" if skip else "") + (" +res" if res else "")
        print(name.ljust(44) + f"{row[0]:9.0f}" + "".join(f"{v:14.0f}" for v in row[1:]))
        del x, x2, r, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
