#!/usr/bin/env python
r"""Headline benchmark: images/sec of ADM 256x256 DDIM-64 sampling (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One bench "step" = one pass of the hot path over one batch: ``DDIMSampler(steps=64)`` called on a
(16, 3, 256, 256) batch per GPU with the ``imagenet_256x256`` ADM card (BASELINE.json configs[2],
the configuration the metric is quoted on; random-init seeded weights, synthetic x1).  Inside a
step the device runs 64 x [native U-Net forward (~370 launches) + fused transition + advance]
as CUDA-graph replays.  Ranks are independent replicas of the sampler over disjoint slices of the
global batch (weak scaling); the only collective is the weight broadcast at init.

Printed JSON (rank 0, one line): see the bench contract in the task statement.  Extra keys:
``roofline`` (dominant kernel = tcgen05 convolution, tensor bound), ``roofline_e2e`` (whole
sampler vs the tensor roofline), ``step_kernel`` (HBM GB/s of the transition kernel, the second
half of BASELINE.json's metric), ``cpu_baseline``, ``clocks``, ``e2e``, ``gpu_launches``.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec ADM 256x256 DDIM-64"
UNIT = "images/s"
CARD = "imagenet_256x256"
BATCH = 16  # per GPU
SAMPLER_STEPS = 64
SIZE = 256
FLOP_PER_IMAGE_FORWARD = 2239.7e9  # FlopCounterMode on the reference module


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------ clocks


class ClockSampler:
    r"""Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                power = max(power, float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "power_w_max": power or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- engine arm


def build_denoiser(device, rank: int, world: int):
    from azula_b200.nn.utils import skip_init
    from azula_b200.plugins import adm

    config = adm.cards()[CARD].config
    with torch.device(device), skip_init():
        den = adm.make_model(**config).eval()
    if rank == 0:
        adm.seed_parameters(den.backbone, seed=1234)
    if world > 1:  # the ONE collective of the path: weights from rank 0 at init (no per-step collective)
        from azula_b200 import parallel

        parallel.broadcast_parameters(den.backbone, src=0)
    return den


def step_kernel_bandwidth(device) -> dict:
    r"""HBM GB/s of the fused transition at a footprint far above L2 (3 x 512 MiB): 12 B/element."""
    from azula_b200 import _lib

    n = 128 * 1024 * 1024
    x = torch.randn(n, device=device)
    f = torch.randn(n, device=device)
    out = torch.empty_like(x)
    row = torch.tensor([[0.9, -0.4, 0.8, 0.5, 0.7, 0.0, 1.0, float("inf")]], device=device)
    idx = torch.zeros((), dtype=torch.int32, device=device)
    lib, s = _lib.lib(), _lib.stream_ptr(device)
    T, _ = _lib.rng_policy(n)

    def launch():
        _lib.check(lib.azb_step_f32(x.data_ptr(), f.data_ptr(), _lib.F32, n, None, out.data_ptr(), None, _lib.F32, n, 1,
                                    row.data_ptr(), idx.data_ptr(), 0, None, 0, T, 0, s), "azb_step_f32")

    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    gbs = 12.0 * n / ms / 1e6
    pk = peaks()
    return {"bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": round(gbs / pk["hbm_gbs"], 4), "bytes_per_launch": 12 * n, "us_per_launch": round(1e3 * ms, 1),
            "note": "azb_step_f32 alone, fp32 x_t + F -> x_s, deterministic (n=0) row, 1.5 GiB footprint >> L2"}


def cpu_baseline_sample(threads: int, forwards: int = 2) -> dict:
    r"""The oracle's fp32 restatement of the reference path on the host cores: ADM-256, batch 1,
    ``forwards`` DDIM steps out of 64, extrapolated linearly to images/sec."""
    from oracle import adm_unet as AU
    from oracle import ref_math as RM
    from oracle.gen_golden_cfg import IMAGENET_256

    torch.set_num_threads(threads)
    cfg = {k: v for k, v in IMAGENET_256.items() if not k.startswith("discrete")}
    tab = AU.block_table(**cfg)
    sd = AU.seeded_state({k: torch.empty(s) for k, s in AU.state_shapes(tab).items()}, seed=1234)
    sched = lambda t: RM.vp_alpha_sigma(t, 1e-2, 1e-2)  # noqa: E731
    sig = RM.adm_sigmas()
    net = lambda x, t, y=None: AU.forward(sd, tab, x, t)  # noqa: E731
    mean = lambda x, t: RM.adm_mean_var(net, sched, sig, x, t)[0]  # noqa: E731
    x = torch.randn(1, 3, SIZE, SIZE, generator=torch.Generator().manual_seed(0))
    pairs = RM.time_grid(1.0, 0.0, SAMPLER_STEPS)
    times = []
    with torch.no_grad():
        for i in range(forwards):
            t, s = pairs[i]
            t0 = time.perf_counter()
            a_s, s_s = sched(s)
            a_t, s_t = sched(t)
            m = mean(x, t)
            x = RM.transition(x, m, torch.randn_like(x), a_t, s_t, a_s, s_s, 0.0)
            times.append(time.perf_counter() - t0)
    return {"step_s": times}


def run_reference(args) -> None:
    r"""``--impl reference``: the reference's own CPU path (the oracle port: plain fp32 torch on the
    host cores, all threads), one DDIM step of one ADM-256 image per bench step, extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    out = cpu_baseline_sample(threads, forwards=args.warmup + args.steps)["step_s"][args.warmup:]
    per_step = sum(out) / len(out)
    value = 1.0 / (SAMPLER_STEPS * per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ADM {CARD} DDIM-{SAMPLER_STEPS}, CPU sample: batch 1, one DDIM step per bench step, "
                               f"images/s = 1 / ({SAMPLER_STEPS} x step time)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} DDIM steps of 1 image (of {SAMPLER_STEPS}), extrapolated linearly"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_engine(args) -> None:
    import torch.distributed as dist

    from azula_b200.sample import DDIMSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.gpus != world:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    den = build_denoiser(device, rank, world)
    # rank r owns samples [r*B, (r+1)*B) of the global batch; noise is addressed by global element index,
    # so the N-GPU run reproduces the one-GPU run on the global batch bit for bit
    sampler = DDIMSampler(den, steps=SAMPLER_STEPS, silent=True, graph=True, shard=(rank, world))
    shape = (args.batch, 3, SIZE, SIZE)
    torch.manual_seed(1000)
    x1 = sampler.init(shape, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # 2 x L2

    # ---- warm-up (first call builds the plan, the coefficient table and captures the graph)
    for _ in range(max(args.warmup, 1)):
        x0 = sampler(x1)
    assert torch.isfinite(x0).all(), "non-finite sample"

    # ---- timed region: K full samplings, inputs resident in HBM
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile_range:  # ncu --profile-from-start off: only the timed region is captured
        torch.cuda.profiler.start()
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations
            x0 = sampler(x1)
        e1.record()
        barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    images = world * args.batch * args.steps
    value = images / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    host_in = x1.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        xd = host_in.to(device, non_blocking=True)
        host_out.copy_(sampler(xd), non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = images / e2e_s

    line = None
    if rank == 0:
        loop = next(iter(sampler._loops.values()))
        plan = next(v for k, v in den.backbone._native.items() if k != "packed")
        launches_per_sampler_step = plan.launches + 2
        detail: list = []
        for _ in range(2):  # second pass: warm instruction caches / clocks as inside the loop
            detail.clear()
            table = plan.profile(detail)
        total_ms = sum(r["ms"] for r in table.values())
        conv = table["conv3x3"]
        pk = peaks()
        conv_tflops = conv["flops"] / conv["ms"] / 1e9
        # the dominant kernel class: 3x3 convolution 256 -> 256 at 256 x 256 (6 launches per forward, each 1.24 TFLOP
        # at batch 16; 5 of them with the GroupNorm + SiLU of their input fused into the halo tiles); its DRAM traffic
        # comes from the committed ncu --set full capture of exactly this launch
        dom = [d for d in detail if d[0] == "conv3x3" and f"x{SIZE}x{SIZE} 256->256" in d[1] and "skip" not in d[1]]
        dom_ms = sum(d[2] for d in dom) / max(len(dom), 1)
        dom_flop = dom[0][3] if dom else 0.0
        dom_tflops = dom_flop / dom_ms / 1e9 if dom else 0.0
        traffic = 1099858432 if args.batch == 16 else None  # profiles/r1_ncu_full_conv_fused_v5.csv: dram read + write
        e2e_tflops = value / world * SAMPLER_STEPS * FLOP_PER_IMAGE_FORWARD / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ADM {CARD} (552.8M params, seeded random init), DDIMSampler(steps={SAMPLER_STEPS}, eta=0), "
                                   f"batch {args.batch}/GPU x {world} GPU, 3x{SIZE}x{SIZE} fp32 state, bf16 backbone",
                       "global_batch": world * args.batch, "parallelism": f"replicas x{world} (batch-sharded, no per-step collective)",
                       "graph": loop.graph is not None, "l2": "256 MiB flush write between timed iterations; working set ~8 GiB >> L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": host_in.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": args.steps * SAMPLER_STEPS * launches_per_sampler_step,
            "roofline": {"bound": "tensor",
                         "kernel": f"conv_gemm_kernel<256, pair, lean, halo>: GroupNorm+SiLU+3x3 conv 256->256 on "
                                   f"{args.batch}x{SIZE}x{SIZE} (halo-tile implicit GEMM, tcgen05 cta_group::2)",
                         "achieved": round(dom_tflops, 1), "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": round(dom_tflops / pk["tflops"], 4), "traffic": traffic,
                         "flop_per_launch": dom_flop, "us_per_launch": round(1e3 * dom_ms, 1), "launches_per_forward": len(dom),
                         "algorithmic_bytes_per_launch": 2 * 2 * args.batch * SIZE * SIZE * 256 + 2 * 9 * 256 * 256,
                         "peak_source": pk["source"] + ", sustained bf16 (the kernel is timed inside a full forward pass)",
                         "peak_burst": pk["tflops_burst"], "frac_of_burst": round(dom_tflops / pk["tflops_burst"], 4),
                         "all_conv3x3": {"achieved": round(conv_tflops, 1), "frac": round(conv_tflops / pk["tflops"], 4),
                                         "launches_per_forward": conv["launches"], "flop_per_forward": conv["flops"],
                                         "share_of_forward": round(conv["ms"] / total_ms, 4)}},
            "roofline_e2e": {"bound": "tensor", "achieved": round(e2e_tflops, 1), "peak": pk["tflops"], "unit": "TFLOP/s",
                             "frac": round(e2e_tflops / pk["tflops"], 4),
                             "note": "images/s/GPU x 64 x 2239.7 GFLOP (algorithmic FLOPs of the reference forward)"},
            "forward_kernels": {k: {"launches": r["launches"], "ms": round(r["ms"], 3)} for k, r in table.items()},
            "clocks": clocks.summary(),
        }
        line["step_kernel"] = step_kernel_bandwidth(device)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = cpu_baseline_sample(threads, forwards=3)["step_s"][1:]
            per = sum(sample) / len(sample)
            line["cpu_baseline"] = {"value": 1.0 / (SAMPLER_STEPS * per), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"oracle (fp32 torch restatement) on host cores: 2 DDIM steps of 1 ADM-256 image "
                                              f"after 1 warm-up, {per:.2f} s/step, extrapolated x{SAMPLER_STEPS}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
