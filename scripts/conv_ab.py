"""A/B timing of the convolution launcher's choices on the ADM-256 layer shapes (batch 16), same process, same box,
configurations interleaved: single CTA vs CTA pairs (tcgen05 cta_group::2), weight prefetch, split-K.

    python scripts/conv_ab.py [--batch 16] [--reps 5]
"""

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import ops  # noqa: E402

# (h = w, c_in, c_out, taps, skip channels of the fused 1x1 operand)
SHAPES = [
    (256, 256, 256, 9, 0), (256, 512, 256, 9, 0), (256, 256, 256, 9, 512), (256, 64, 256, 1, 0),
    (128, 256, 256, 9, 0), (128, 512, 512, 9, 0), (128, 512, 256, 9, 0),
    (64, 512, 512, 9, 0), (64, 1024, 512, 9, 0), (64, 512, 512, 9, 1024),
    (32, 512, 512, 9, 0), (32, 1024, 1024, 9, 0), (32, 512, 1536, 1, 0),
    (16, 1024, 1024, 9, 0), (16, 2048, 1024, 9, 0), (16, 1024, 3072, 1, 0),
    (8, 1024, 1024, 9, 0), (8, 2048, 1024, 9, 0),
    (64, 256, 256, 9, 0), (32, 1024, 512, 9, 0), (32, 512, 512, 9, 1024), (16, 1024, 1024, 9, 2048), (16, 512, 1024, 9, 0),
    (16, 1024, 1024, 1, 0), (32, 512, 512, 1, 0), (8, 1024, 3072, 1, 0), (8, 1024, 1024, 1, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--configs", default="single,pair")
    ap.add_argument("--max-hw", type=int, default=256)
    args = ap.parse_args()
    dev = "cuda"
    configs = {
        "single": {ops.KNOB_PAIR: 0},
        "pair": {ops.KNOB_PAIR: 1},
        "auto": {},
        "noprefetch": {ops.KNOB_PREFETCH: 0},
        "nosplit": {ops.KNOB_SPLITK: 0},
        "p256": {ops.KNOB_PAIR: 1, ops.KNOB_BLOCKN: 256, ops.KNOB_SPLITK: 0},
        "p128": {ops.KNOB_PAIR: 1, ops.KNOB_BLOCKN: 128, ops.KNOB_SPLITK: 0},
        "s256": {ops.KNOB_PAIR: 0, ops.KNOB_BLOCKN: 256, ops.KNOB_SPLITK: 0},
        "s128": {ops.KNOB_PAIR: 0, ops.KNOB_BLOCKN: 128, ops.KNOB_SPLITK: 0},
        "s64": {ops.KNOB_PAIR: 0, ops.KNOB_BLOCKN: 64, ops.KNOB_SPLITK: 0},
        "generic": {ops.KNOB_LEAN: 0},
        "tap": {ops.KNOB_HALO: 0},  # tap-wise TMA loads instead of halo tiles
        "tap_p128": {ops.KNOB_HALO: 0, ops.KNOB_PAIR: 1, ops.KNOB_BLOCKN: 128, ops.KNOB_SPLITK: 0},
    }
    names = args.configs.split(",")
    ws = ops.splitk_workspace(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    out_rows = []
    for hw, ci, co, taps, skip in SHAPES:
        if hw > args.max_hw:
            continue
        n = args.batch
        x = torch.randn(n, hw, hw, ci, device=dev, generator=g).to(torch.bfloat16)
        k = 3 if taps == 9 else 1
        pc = ops.pack_conv(torch.randn(co, ci, k, k, device=dev, generator=g) / (ci * taps) ** 0.5, torch.randn(co, device=dev, generator=g))
        x2 = None
        if skip:
            x2 = torch.randn(n, hw, hw, skip, device=dev, generator=g).to(torch.bfloat16)
            pc = ops.pack_conv_skip(pc, ops.pack_conv(torch.randn(co, skip, 1, 1, device=dev, generator=g) / skip**0.5,
                                                      torch.randn(co, device=dev, generator=g)))
        out = torch.empty(n, hw, hw, co, dtype=torch.bfloat16, device=dev)
        flop = 2.0 * n * hw * hw * co * (ci * taps + skip)
        inner = 3 if flop > 2e11 else 10
        best = {c: float("inf") for c in names}
        ref = None
        for rep in range(args.reps + 1):
            for c in names:
                for kn in (ops.KNOB_PAIR, ops.KNOB_PREFETCH, ops.KNOB_SPLITK, ops.KNOB_BLOCKN, ops.KNOB_LEAN, ops.KNOB_HALO):
                    ops.conv_tuning(kn, configs[c].get(kn, -1))
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(inner):
                    _, acc = ops.conv_acc(x, pc, out=out, x2=x2, workspace=ws)
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    best[c] = min(best[c], e0.elapsed_time(e1) / inner)
                elif ref is None:
                    ref = out.clone()
                else:
                    assert torch.allclose(ref.float(), out.float(), rtol=2e-2, atol=2e-2), (c, hw, ci, co)
        row = {"shape": f"{n}x{hw}x{hw} {ci}->{co} k{k}" + (f" +skip{skip}" if skip else ""), "gflop": flop / 1e9}
        row.update({c: round(best[c], 4) for c in names})
        row.update({f"{c}_tflops": round(flop / best[c] / 1e9, 1) for c in names})
        out_rows.append(row)
        print("  ".join(f"{k}={v}" for k, v in row.items()), flush=True)
        del x, x2, out, pc
    for kn in (ops.KNOB_PAIR, ops.KNOB_PREFETCH, ops.KNOB_SPLITK, ops.KNOB_BLOCKN, ops.KNOB_LEAN, ops.KNOB_HALO):
        ops.conv_tuning(kn, -1)
    print(json.dumps(out_rows))


if __name__ == "__main__":
    main()
