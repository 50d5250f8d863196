"""A/B timing of the GroupNorm-apply launch policy (AZB_GN_KNOB_WAVE) on the ADM-256 shapes, configurations
interleaved in one process.

    python scripts/gn_ab.py [--waves 0,2,4,8]
"""

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import ops  # noqa: E402

# (h = w, channels, mode)   mode 0 same size, 1 nearest x2, 2 average pool
SHAPES = [(256, 256, 0), (256, 512, 0), (256, 256, 2), (128, 256, 0), (128, 512, 0), (128, 256, 1), (128, 256, 2),
          (64, 512, 0), (64, 1024, 0), (64, 512, 1), (32, 512, 0), (32, 1024, 0), (16, 1024, 0), (16, 2048, 0), (8, 1024, 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--waves", default="0,2,4,8")
    args = ap.parse_args()
    dev = "cuda"
    waves = [int(v) for v in args.waves.split(",")]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for hw, c, mode in SHAPES:
        n = args.batch
        x = torch.randn(n, hw, hw, c, device=dev, generator=g).to(torch.bfloat16)
        acc = torch.zeros(n, c // 8, 4, dtype=torch.int64, device=dev)
        xs = x.float().reshape(n, hw * hw, c // 8, 8)
        acc[..., 1] = (xs.sum(dim=(1, 3)).double() * 2.0**40).round().long()  # lo word only (hi = 0): good enough to time
        acc[..., 3] = (xs.square().sum(dim=(1, 3)).double() * 2.0**40).round().long()
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        ss = torch.randn(n, 2 * c, device=dev, generator=g) * 0.1
        ho = hw * 2 if mode == 1 else hw // 2 if mode == 2 else hw
        out = torch.empty(n, ho, ho, c, dtype=torch.bfloat16, device=dev)
        nbytes = 2.0 * (x.numel() + out.numel())
        inner = 3 if nbytes > 2e8 else 10
        best = {w: float("inf") for w in waves}
        ref = None
        for rep in range(args.reps + 1):
            for w in waves:
                ops.conv_tuning(ops.KNOB_GN_WAVE, w)
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(inner):
                    ops.gn_apply_acc(x, [(acc, c)], gamma, beta, out=out, scale_shift=ss, mode=mode)
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    best[w] = min(best[w], e0.elapsed_time(e1) / inner)
                elif ref is None:
                    ref = out.clone()
                else:
                    assert torch.equal(ref, out), (hw, c, mode, w)
        row = {"shape": f"{n}x{hw}x{hw}x{c} mode{mode}"}
        row.update({f"w{w}_us": round(1e3 * best[w], 1) for w in waves})
        row.update({f"w{w}_gbs": round(nbytes / best[w] / 1e6) for w in waves})
        rows.append(row)
        print("  ".join(f"{k}={v}" for k, v in row.items()), flush=True)
    ops.conv_tuning(ops.KNOB_GN_WAVE, -1)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
