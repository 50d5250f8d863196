#!/usr/bin/env python
r"""Secondary measurements (not the bench contract line): BASELINE.json configs 2 and 4 through the
fused sampler, per-kernel tables of their launch plans, and the reference's execution model (plain
torch eager: Python loop, ATen / cuDNN kernels, fp32 -- ``azula_b200.engine.eager_torch()``) timed on
the same GPU for configs 2, 3 and 4 (the "reference PyTorch-eager sampler on 1xB200" of the north star;
the reference package itself cannot travel to the GPU box, the host mirror's torch path is the same
arithmetic, see tests/test_nn_cpu.py and tests/test_adm_cpu.py).

    python scripts/nn_bench.py [--configs 2,3,4] [--eager] [--out gpurun_out/nn_bench.jsonl]
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from azula_b200 import engine  # noqa: E402
from azula_b200.denoise import KarrasDenoiser  # noqa: E402
from azula_b200.noise import VPSchedule  # noqa: E402
from azula_b200.sample import DDIMSampler, DDPMSampler  # noqa: E402


class TimeWrapper(torch.nn.Module):
    r"""The tutorial pattern (reference docs/tutorials/mnist.ipynb cell 8): mod = MLP(log_snr)."""

    def __init__(self, net: torch.nn.Module, features: int) -> None:
        super().__init__()
        self.net = net
        self.time_embedding = torch.nn.Sequential(
            torch.nn.Linear(1, features), torch.nn.SiLU(), torch.nn.Linear(features, features)
        )

    def forward(self, x_t, log_snr_t):
        return self.net(x_t, self.time_embedding(log_snr_t[..., None]))


def seed_all(module: torch.nn.Module, seed: int) -> None:
    from azula_b200.plugins.adm import seed_parameters

    seed_parameters(module, seed=seed)


def config(idx: int, device):
    if idx == 2:
        from azula_b200.nn.unet import UNet

        net = TimeWrapper(UNet(3, 3, hid_channels=(64, 128, 256), hid_blocks=(3, 3, 3), mod_features=256), 256)
        seed_all(net, 2)
        den = KarrasDenoiser(net.to(device), VPSchedule()).eval()
        return dict(name="azula.nn.unet UNet 64x64x3, DDIM-50, batch 32", den=den, sampler=DDIMSampler, steps=50,
                    shape=(32, 3, 64, 64), flop_per_image=13.02e9, plan_of=lambda: net.net)
    if idx == 3:
        from azula_b200.nn.utils import skip_init
        from azula_b200.plugins import adm

        with torch.device(device), skip_init():
            den = adm.make_model(**adm.cards()["imagenet_256x256"].config).eval()
        adm.seed_parameters(den.backbone, seed=1234)
        return dict(name="ADM imagenet_256x256, DDIM-64, batch 16", den=den, sampler=DDIMSampler, steps=64,
                    shape=(16, 3, 256, 256), flop_per_image=2239.7e9, plan_of=lambda: den.backbone)
    if idx == 4:
        from azula_b200.nn.vit import ViT

        net = TimeWrapper(ViT(4, 4, mod_features=768, hid_channels=768, hid_blocks=12, attention_heads=12, patch_size=2), 768)
        seed_all(net, 4)
        den = KarrasDenoiser(net.to(device), VPSchedule()).eval()
        return dict(name="DiT-B/2 (ViT hid 768 x 12 blocks x 12 heads, patch 2) on 32x32x4, DDPM-250, batch 64", den=den,
                    sampler=DDPMSampler, steps=250, shape=(64, 4, 32, 32), flop_per_image=46.58e9, plan_of=lambda: net.net)
    raise ValueError(idx)


def time_calls(fn, reps: int, device) -> float:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / reps


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,4")
    ap.add_argument("--eager", action="store_true", help="also time the plain torch eager execution model")
    ap.add_argument("--eager-steps", type=int, default=0, help="sampler steps of the eager run (0 = the config's)")
    ap.add_argument("--no-tf32", action="store_true", help="eager run with TF32 off (strict fp32 convolutions / matmuls)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "nn_bench.jsonl"))
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    peak = 1353.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["bf16_tflops_sustained"]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as fout, torch.no_grad():
        for idx in (int(c) for c in args.configs.split(",")):
            cfg = config(idx, device)
            smp = cfg["sampler"](cfg["den"], steps=cfg["steps"], silent=True, graph=True)
            torch.manual_seed(0)
            x1 = smp.init(cfg["shape"], device=device)
            x0 = smp(x1)
            assert torch.isfinite(x0).all()
            ms = time_calls(lambda: smp(x1), args.reps, device)
            batch = cfg["shape"][0]
            ips = batch / (ms / 1e3)
            tflops = ips * cfg["steps"] * cfg["flop_per_image"] / 1e12
            model = cfg["plan_of"]()
            plan = next(v for k, v in model._native.items() if isinstance(k, tuple))
            table = plan.profile()
            line = {
                "config": idx, "workload": cfg["name"], "images_per_s": round(ips, 2), "ms_per_sampling": round(ms, 3),
                "ms_per_sampler_step": round(ms / cfg["steps"], 4), "tflops": round(tflops, 1),
                "frac_of_sustained_bf16": round(tflops / peak, 4), "launches_per_forward": plan.launches,
                "graph": next(iter(smp._loops.values())).graph is not None,
                "forward_kernels": {k: {"launches": r["launches"], "ms": round(r["ms"], 4),
                                        "tflops": round(r["flops"] / r["ms"] / 1e9, 1) if r["flops"] else None,
                                        "gbs": round(r["bytes"] / r["ms"] / 1e6, 1)} for k, r in table.items()},
            }
            if args.eager:
                steps = args.eager_steps or cfg["steps"]
                if args.no_tf32:
                    torch.backends.cudnn.allow_tf32 = False
                    torch.backends.cuda.matmul.allow_tf32 = False
                with engine.eager_torch():
                    esmp = cfg["sampler"](cfg["den"], steps=steps, silent=True)
                    esmp(x1)  # warm-up (cuDNN autotune, lazy init)
                    ems = time_calls(lambda: esmp(x1), 1 if idx == 3 else 2, device) * cfg["steps"] / steps
                line["eager_torch"] = {"images_per_s": round(batch / (ems / 1e3), 3), "ms_per_sampling": round(ems, 2),
                                       "steps_timed": steps, "tf32": bool(torch.backends.cudnn.allow_tf32),
                                       "speedup": round(ems / ms, 2),
                                       "what": "plain torch fp32 eager (reference execution model) on the same GPU"}
            print(json.dumps(line), flush=True)
            fout.write(json.dumps(line) + "\n")
            del smp, cfg, plan, model
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
