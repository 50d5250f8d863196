"""Per-kernel timing of one native ADM forward (CUDA events around every launch of the plan).

    python scripts/adm_profile.py [--batch 16] [--size 256] [--card imagenet_256x256]
"""

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.plugins import adm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--card", default="imagenet_256x256")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    torch.manual_seed(0)
    cfg = adm.cards()[args.card].config
    t0 = time.time()
    den = adm.make_model(**cfg).eval()
    den = den.cuda()
    adm.seed_parameters(den.backbone, seed=1234)
    print(f"model built in {time.time() - t0:.1f}s", file=sys.stderr)
    x = torch.randn(args.batch, 3, args.size, args.size, device="cuda")
    ts = torch.tensor([500], device="cuda")
    with torch.no_grad():
        out = den.backbone(x, ts)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            den.backbone(x, ts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        plan = next(v for k, v in den.backbone._native.items() if k != "packed")
        detail = []
        table = plan.profile(detail)
    total = sum(r["ms"] for r in table.values())
    flops = sum(r["flops"] for r in table.values())
    print(f"forward {ms:.2f} ms eager ({plan.launches} launches), sum of kernels {total:.2f} ms, "
          f"{flops / 1e12:.1f} TFLOP -> {flops / ms / 1e9:.0f} TFLOP/s; scratch {plan.scratch_bytes / 2**30:.2f} GiB")
    for kind, r in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {kind:10s} n={r['launches']:4d}  {r['ms']:8.2f} ms  {100 * r['ms'] / total:5.1f}%  "
              f"{r['flops'] / r['ms'] / 1e9:8.0f} TFLOP/s  {r['bytes'] / r['ms'] / 1e6:8.0f} GB/s")
    groups = {}
    for kind, desc, t, fl, by in detail:
        g = groups.setdefault((kind, desc), [0, 0.0, 0.0, 0.0])
        g[0] += 1; g[1] += t; g[2] += fl; g[3] += by
    print("  -- by shape --")
    for (kind, desc), g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:90]:
        print(f"  {kind:10s} {desc:34s} n={g[0]:3d} {g[1]:7.3f} ms  {g[2] / g[1] / 1e9:6.0f} TFLOP/s {g[3] / g[1] / 1e6:6.0f} GB/s")
    print(json.dumps({"forward_ms": ms, "kernels": table}))


if __name__ == "__main__":
    main()
