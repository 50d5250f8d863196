"""One token GEMM (rows x K -> N) with an epilogue of choice under the launcher knobs; event-timed, or --once for ncu.

    python scripts/gemm_one.py --rows 16384 --k 768 --n 3072 --act silu --rowepi 1
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azula_b200.engine import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16384)
    ap.add_argument("--k", type=int, default=768)
    ap.add_argument("--n", type=int, default=3072)
    ap.add_argument("--act", default="none")
    ap.add_argument("--gate", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--rowepi", type=int, default=-1)
    ap.add_argument("--pair", type=int, default=-1)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(a.rows, a.k, device=dev, generator=g).to(torch.bfloat16)
    pc = ops.pack_conv(torch.randn(a.n, a.k, 1, 1, device=dev, generator=g) / a.k**0.5, torch.randn(a.n, device=dev, generator=g))
    out = torch.empty(a.rows, a.n, dtype=torch.bfloat16, device=dev)
    gate = torch.randn(64, a.n, device=dev, generator=g) if a.gate else None
    res = torch.randn(a.rows, a.n, device=dev, generator=g).to(torch.bfloat16) if a.res else None
    ops.conv_tuning(ops.KNOB_ROWEPI, a.rowepi)
    ops.conv_tuning(ops.KNOB_PAIR, a.pair)
    run = lambda: ops.conv2d(x, pc, out=out, act=None if a.act == "none" else a.act, gate=gate, gate_rows=a.rows // 64, residual=res)  # noqa: E731
    run()
    if a.once:
        torch.cuda.synchronize()
        return
    ref = torch.nn.functional.linear(x.float(), pc.w.float().reshape(a.n, -1)[:, : a.k], pc.bias)
    if a.act == "silu":
        ref = torch.nn.functional.silu(ref)
    if gate is not None:
        ref = ref * gate.repeat_interleave(a.rows // 64, 0)
    if res is not None:
        ref = ref + res.float()
    err = (out.float() - ref).abs().max().item()
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    print(f"rows {a.rows} K {a.k} N {a.n} act {a.act} gate {a.gate} res {a.res} rowepi {a.rowepi} pair {a.pair}: {us:7.1f} us "
          f"{2.0 * a.rows * a.k * a.n / us / 1e6:7.1f} TFLOP/s  max|err| {err:.3g}")


if __name__ == "__main__":
    main()
