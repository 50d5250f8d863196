"""Timeline of the short-sequence attention kernel (AZB_ATTN_TRACE): clock64 stamps of both softmax groups and their MMA
issuers for the first items of a few CTAs, relative to the CTA's first stamp.  python scripts/attn_trace.py [--qknorm]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azula_b200.engine import ops  # noqa: E402

n, t, heads = 64, 256, 12
qkv = torch.randn(n, t, 3 * heads * 64, device="cuda").to(torch.bfloat16)
fn = (lambda: ops.attention_qknorm(qkv, heads)) if "--qknorm" in sys.argv else (lambda: ops.attention(qkv, heads, True))
for _ in range(3):
    fn()
torch.cuda.synchronize()
path = "/tmp/attn_trace.bin"
os.environ["AZB_ATTN_TRACE"] = path
fn()
torch.cuda.synchronize()
del os.environ["AZB_ATTN_TRACE"]
tr = np.fromfile(path, dtype=np.int64).reshape(-1, 4, 8, 8)
names_g = ["wait S", "S ready", "pass1 done", "P half", "P full", "O ready", "drained"]
names_m = ["sfree ok", "qk ok", "v ok", "p half ok", "p full ok", "PV issued"]
for cta in (0, 77, 147):
    t0 = tr[cta][tr[cta] > 0].min()
    print(f"== CTA {cta}")
    for role in range(4):
        for j in range(6):
            ev = tr[cta, role, j]
            names = names_g if role < 2 else names_m
            s = " ".join(f"{nm}={int(v - t0):6d}" for nm, v in zip(names, ev) if v > 0)
            if s:
                print(f"  {'group' if role < 2 else 'mma  '} {role & 1} item {j}: {s}")
