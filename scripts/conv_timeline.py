"""Tile timeline of CTA 0 of one convolution launch (diagnostics stamps of conv_gemm_kernel): MMA issuer and first epilogue
warp.  Needs a diagnostics build: AZB_NVCC_EXTRA=-DAZB_TIMELINE python -m azula_b200.csrc.build --force
python scripts/conv_timeline.py --only plain,-1,-1 [--c 64 --hw 64]"""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azula_b200 import _lib  # noqa: E402
from azula_b200.engine import ops  # noqa: E402

buf = torch.zeros((1 << 16) + 1024, dtype=torch.int64, device="cuda")
buf[8 : 8 + 4 * 64].view(64, 4)[:, :2] = torch.iinfo(torch.int64).max
_lib.check(_lib.lib().azb_debug_trace(buf.data_ptr()), "trace")
import unet_conv_ab  # noqa: E402

sys.argv = [sys.argv[0]] + sys.argv[1:]
unet_conv_ab.main()
torch.cuda.synchronize()
_lib.lib().azb_debug_trace(None)
tl = buf[(1 << 16) : (1 << 16) + 256].view(16, 16).cpu().numpy()
t0 = tl[0, 0]
names = ["mma: tile", "acc free", "A ready", "issued", "epi: wait", "acc full", "done", "c0 ld", "c0 bias+act", "c0 gate+res+pack", "c0 sums", "c1 ld",
         "c1 bias+act", "c1 gate+res+pack", "c1 sums"]
for i in range(10):
    if tl[i, 0] == 0:
        break
    print(f"tile {i}: " + "  ".join(f"{n}={int(v - t0):6d}" for n, v in zip(names, tl[i]) if v > 0) + " (ns)")
