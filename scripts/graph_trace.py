#!/usr/bin/env python
r"""In-graph timeline of the convolution / GEMM launches of a sampler step (azb_debug_trace: %globaltimer stamps written by
the kernels themselves while the captured graph replays): per launch the earliest CTA entry, the earliest CTA start after
the programmatic-launch wait, the latest CTA end -- i.e. kernel durations and the gaps between them as they are inside the
graph, which event timing of isolated launches and ncu cannot show.  python scripts/graph_trace.py --config unet64"""

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
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    from azula_b200 import _lib
    from azula_b200.engine import ops

    device = torch.device("cuda", 0)
    cap = 8192
    buf = torch.zeros(8 + 4 * cap, dtype=torch.int64, device=device)

    def reset():
        buf.zero_()
        v = buf[8:].view(cap, 4)
        v[:, 0] = torch.iinfo(torch.int64).max
        v[:, 1] = torch.iinfo(torch.int64).max

    reset()
    _lib.check(_lib.lib().azb_debug_trace(buf.data_ptr()), "azb_debug_trace")
    wl = bench.WORKLOADS[args.config]
    with torch.no_grad():
        den = bench.build_denoiser("azula_b200", args.config, device)
        bench.seed_backbone(den)
        smp = bench.sampler_of("azula_b200", args.config, den, graph=True)
        smp.steps = args.steps
        x1 = smp.init(wl["shape"], device=device)
        smp(x1)  # builds the plan, captures the graph
        torch.cuda.synchronize()
        reset()
        smp(x1)
        torch.cuda.synchronize()
        n = int(buf[0].item())
        rec = buf[8 : 8 + 4 * n].view(n, 4).cpu().numpy()
        loop = next(iter(smp._loops.values()))
        plan = loop.pinned[0][2]
        detail: list = []
        plan.profile(detail)
        torch.cuda.synchronize()
    _lib.lib().azb_debug_trace(None)
    convs = [(k, d) for k, d, *_ in detail if k in ("gemm", "conv3x3", "conv1x1", "linear")]
    per_fwd = n // args.steps
    print(f"# {args.config}: {n} conv launches traced over {args.steps} steps ({per_fwd} per forward; plan lists {len(convs)})")
    # the LAST step (steady state)
    r = rec[(args.steps - 1) * per_fwd : args.steps * per_fwd]
    t0 = r[0, 1]
    prev_end = None
    busy = 0.0
    for i, (enter, start, end, _) in enumerate(r):
        desc = convs[i][1] if i < len(convs) else "(not in the plan list: output conv)"
        gap = "" if prev_end is None else f"gap {1e-3 * (start - prev_end):6.1f}"
        print(f"{i:3d} start {1e-3 * (start - t0):8.1f}  run {1e-3 * (end - start):6.1f} us  resident-before-start {1e-3 * (start - enter):5.1f}  {gap:12s} {desc}")
        busy += end - start
        prev_end = end
    print(f"## forward span {1e-3 * (r[-1, 2] - r[0, 1]):.1f} us, conv kernels busy {1e-3 * busy:.1f} us")


if __name__ == "__main__":
    main()
