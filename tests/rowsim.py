"""CPU restatement of what ``azb_step_ex_f32`` does with one row of the coefficient table (csrc/step.cu), driving a
whole fused loop in plain torch.  Test infrastructure: lets the ``-m "not gpu"`` suite hold the table builders of
every sampler (azula_b200/engine/table.py) to the reference fixtures without a GPU; the ``-m gpu`` suite then holds
the kernel to the same fixtures."""

from __future__ import annotations

import torch

from azula_b200 import _lib
from azula_b200.engine import table as T


def run_table(sampler, x, noise_fn=torch.randn_like, guidance=None, positive=None, negative=None, **kwargs):
    """Executes the sampler's coefficient table stage by stage; returns x_0."""
    tab = T.build(sampler, x.device)
    assert tab is not None
    den = T.inner_denoiser(sampler.denoiser)
    coef = tab.coef
    bits = coef.view(torch.int32)
    src = [x.clone(), x.clone()]
    hist = [None] * _lib.MAX_SLOTS
    x_in = (x * tab.c_in0).to(x.dtype)
    for j in range(tab.steps):
        row = coef[j]
        c_skip, c_out, a, k, b, n, c_in_next, clip = (row[i] for i in range(8))
        flags = int(bits[j, _lib.R_FLAGS])
        xe, xb, out_sel = src[flags & 1], src[(flags >> 1) & 1], (flags >> 2) & 1

        def mean_of(**kw):
            f = den.call_backbone(x_in, tab.time[j], **kw, **kwargs)
            sel = getattr(den, "output_select", lambda: None)()
            if sel is not None:
                f = f[:, : x.shape[1]]
            m = c_skip * xe + c_out * f
            return torch.clamp(m, -clip, clip) if torch.isfinite(clip) else m

        if positive is None:
            m = mean_of()
        else:
            mp, mn = mean_of(**positive), mean_of(**(negative or {}))
            m = mp + guidance * (mp - mn)

        if flags & T.F_HIST:
            wslot, nslots = (flags >> 4) & 15, (flags >> 12) & 15
            h = row[_lib.R_P] * xe + row[_lib.R_Q] * m
            new = row[_lib.R_R] * xb
            for s in range(nslots):
                w = row[_lib.R_W + s]
                if s == wslot:
                    new = new + w * h
                elif w != 0:
                    new = new + w * hist[s]
            if flags & T.F_STORE:
                hist[wslot] = h
        else:
            new = a * m
            new = new + k * (xe - b * m)
            drew = j + 1 < tab.steps and int(bits[j + 1, _lib.R_DRAW]) > int(bits[j, _lib.R_DRAW])
            drew = drew or (j + 1 == tab.steps and tab.draws > int(bits[j, _lib.R_DRAW]))
            if drew:
                new = new + n * noise_fn(xe)
        src[out_sel] = new
        x_in = c_in_next * new
    return src[0]
