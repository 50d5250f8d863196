#!/usr/bin/env python
r"""Per-launch table of a native launch plan (CUDA events around every launch, second pass): what each kernel of a
forward costs and at which rate it runs.  python scripts/plan_detail.py --config unet64|dit_b2|adm [--rowepi 0|1]"""

from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="unet64", choices=list(bench.WORKLOADS))
    ap.add_argument("--rowepi", type=int, default=-1)
    ap.add_argument("--top", type=int, default=0)
    args = ap.parse_args()
    from azula_b200.engine import ops

    ops.conv_tuning(ops.KNOB_ROWEPI, args.rowepi)
    device = torch.device("cuda", 0)
    wl = bench.WORKLOADS[args.config]
    with torch.no_grad():
        den = bench.build_denoiser("azula_b200", args.config, device)
        bench.seed_backbone(den)
        smp = bench.sampler_of("azula_b200", args.config, den, graph=True)
        smp.steps = 2
        x1 = smp.init(wl["shape"], device=device)
        smp(x1)
        loop = next(iter(smp._loops.values()))
        plan = loop.pinned[0][2]
        detail: list = []
        for _ in range(3):
            detail.clear()
            table = plan.profile(detail)
    total = sum(r["ms"] for r in table.values())
    print(f"# {args.config} rowepi={args.rowepi}: {len(detail)} launches, {total:.3f} ms per forward (event-timed one by one)")
    rows = sorted(detail, key=lambda d: -d[2]) if args.top else detail
    for kind, desc, ms, flops, nbytes in rows[: args.top or None]:
        rate = f"{flops / ms / 1e9:8.1f} TFLOP/s" if flops else f"{nbytes / ms / 1e6:8.1f} GB/s   "
        print(f"{kind:10s} {1e3 * ms:8.1f} us  {rate}  {desc}")
    for k, r in table.items():
        print(f"## {k:10s} launches {r['launches']:3d}  {r['ms']:.3f} ms")


if __name__ == "__main__":
    main()
