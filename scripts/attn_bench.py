"""Times azb_attention_bf16 (tcgen05, d = 64) against the mma.sync kernel on the ADM / DiT shapes.

    python scripts/attn_bench.py [--once]      (--once: one launch per shape, for ncu)
"""

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from azula_b200.engine import ops  # noqa: E402

SHAPES = [  # (images, tokens, heads, new_order, what)
    (16, 1024, 8, False, "ADM-256 32x32, 512 ch"),
    (16, 256, 16, False, "ADM-256 16x16, 1024 ch"),
    (16, 64, 16, False, "ADM-256 8x8, 1024 ch"),
    (64, 256, 12, True, "DiT-B/2 32x32 latents"),
]


def main():
    once = "--once" in sys.argv
    for n, t, heads, new_order, what in SHAPES:
        qkv = torch.randn(n, t, 3 * heads * 64, device="cuda").to(torch.bfloat16)
        flops = 4.0 * n * heads * t * t * 64
        for kernel in ("auto", "mma"):
            out = ops.attention(qkv, heads, new_order, kernel=kernel)
            if once:
                continue
            for _ in range(3):
                ops.attention(qkv, heads, new_order, out=out, kernel=kernel)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ops.attention(qkv, heads, new_order, out=out, kernel=kernel)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / reps
            print(f"{what:28s} {'tcgen05' if kernel == 'auto' else 'mma.sync':9s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")
        if t <= 256 and not once:  # the q / k RMS normalisation folded into the short-sequence kernel vs the separate pass
            for what2, fn in (("fused qk-norm", lambda: ops.attention_qknorm(qkv, heads, out=out)),
                              ("rmsnorm pass + attention", lambda: (ops.segment_rmsnorm_(qkv.reshape(n * t, -1), 2 * heads, 64), ops.attention(qkv, heads, True, out=out)))):
                if not new_order:
                    continue
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                print(f"{what:28s} {what2:26s} {1e3 * e0.elapsed_time(e1) / 20:8.1f} us")
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
