"""Ablation of the plan-level design decisions on the full ADM-256 forward (batch 16), measured in ONE process with the
variants interleaved round after round, so that clock / thermal drift of the power-capped part hits all of them alike:

    python scripts/ablation.py [--rounds 8]

Each variant is a complete launch plan (own buffers); a forward is timed with CUDA events around Plan.run.
"""

import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import adm as engine  # noqa: E402
from azula_b200.engine import ops  # noqa: E402
from azula_b200.plugins import adm  # noqa: E402

VARIANTS = [
    # name, module flags, launcher knobs
    ("all on (default)", {}, {}),
    ("no phase-decomposed upsampling convs", {"UP_PHASES": False}, {}),
    ("... and resampled tensors materialised", {"UP_PHASES": False, "RESAMPLE_SHORTCUTS": False}, {}),
    ("... and GroupNorm as a separate pass (halo kernels)", {"UP_PHASES": False, "RESAMPLE_SHORTCUTS": False, "FUSE_NORM": False}, {}),
    ("... and tap-wise convolutions (round start)", {"UP_PHASES": False, "RESAMPLE_SHORTCUTS": False, "FUSE_NORM": False},
     {ops.KNOB_HALO: 0}),
    ("all on, plain stream-ordered launches (no PDL)", {}, {ops.KNOB_PDL: 0}),
]


KNOB_VARIANTS = [
    ("all on (default)", {}, {}),
    ("halo kernels: 2 A slots", {}, {ops.KNOB_HALO_SA: 2}),
    ("halo kernels: 3 A slots", {}, {ops.KNOB_HALO_SA: 3}),
    ("halo kernels: 4 A slots", {}, {ops.KNOB_HALO_SA: 4}),
    ("halo kernels: 4 A slots, again", {}, {ops.KNOB_HALO_SA: 4}),
    ("halo kernels: fused 1x1 blocks after the last halo item", {}, {ops.KNOB_HALO_SPREAD: 0}),
    ("CTA pairs wherever the shape allows", {}, {ops.KNOB_PAIR: 1}),
    ("no weight prefetch into L2 on small maps", {}, {ops.KNOB_PREFETCH: 0}),
    ("GroupNorm pass: short CTAs (no single wave)", {}, {ops.KNOB_GN_WAVE: 0}),
    ("all on (default), again", {}, {}),
]
ALL_KNOBS = (ops.KNOB_HALO, ops.KNOB_PDL, ops.KNOB_HALO_SA, ops.KNOB_HALO_SPREAD, ops.KNOB_PAIR, ops.KNOB_PREFETCH,
             ops.KNOB_GN_WAVE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--knobs", action="store_true", help="launcher knobs instead of plan-level decisions")
    ap.add_argument("--rounds", type=int, default=8)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    cfg = adm.cards()["imagenet_256x256"].config
    den = adm.make_model(**cfg).eval().cuda()
    adm.seed_parameters(den.backbone, seed=1234)
    x = torch.randn(args.batch, 3, 256, 256, device="cuda")
    ts = torch.tensor([500], device="cuda")
    out = torch.empty(args.batch, 6, 256, 256, device="cuda")
    plans = []
    defaults = {k: getattr(engine, k) for k in ("UP_PHASES", "RESAMPLE_SHORTCUTS", "FUSE_NORM")}
    with torch.no_grad():
        for name, flags, knobs in (KNOB_VARIANTS if args.knobs else VARIANTS):
            for k, v in {**defaults, **flags}.items():
                setattr(engine, k, v)
            for k in ALL_KNOBS:
                ops.conv_tuning(k, knobs.get(k, -1))
            den.backbone._native.clear()  # weights are repacked too (the phase weights depend on UP_PHASES)
            den.backbone(x, ts)
            packed = den.backbone._native["packed"]
            plan = next(v for k, v in den.backbone._native.items() if k != "packed")
            plans.append((name, knobs, plan, packed))
        torch.cuda.synchronize()
        times = [[] for _ in plans]
        import random

        rng = random.Random(0)
        for rnd in range(args.rounds + 1):
            order = list(range(len(plans)))
            rng.shuffle(order)  # a fixed position in the round is worth +-2 % (what ran just before sets the clocks)
            for i in order:
                name, knobs, plan, _ = plans[i]
                for k in ALL_KNOBS:
                    ops.conv_tuning(k, knobs.get(k, -1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.run(x, ts, None, out)
                plan.run(x, ts, None, out)
                e1.record()
                torch.cuda.synchronize()
                if rnd:
                    times[i].append(e0.elapsed_time(e1) / 2)
    for k in ALL_KNOBS:
        ops.conv_tuning(k, -1)
    base = statistics.median(times[0])
    print(f"{'variant':62s} {'ms / forward':>12s} {'vs default':>10s} {'kernels':>8s} {'scratch GiB':>11s}")
    for (name, _, plan, _), t in zip(plans, times):
        m = statistics.median(t)
        print(f"{name:62s} {m:12.2f} {m / base:10.3f} {plan.launches:8d} {plan.scratch_bytes / 2**30:11.2f}")


if __name__ == "__main__":
    main()
