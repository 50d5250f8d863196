"""Generates ``tests/golden/adm_cards.npz``: every card of the reference's ``azula/plugins/adm/cards.yaml``
(imagenet 64 / 128 / 256 / 256-cond / 512, ffhq) built by the UNMODIFIED reference (``adm.make_model``), every
parameter overwritten from a seed, evaluated once at a reduced spatial size (64 x 64: all cards' layer stacks, widths,
head counts, attention placements, label embeddings at a CPU-friendly cost).  Stored: input, timestep, label, a strided
subsample of the U-Net output and of the posterior mean.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_cards.py

Test infrastructure; never imported by product code.
"""

from __future__ import annotations

import os
import sys
import torch
import types

REF = os.environ.get("AZULA_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.modules.setdefault("gdown", types.ModuleType("gdown"))

import azula  # noqa: E402
from azula.plugins import adm  # noqa: E402
from azula.plugins.utils import load_cards  # noqa: E402

from oracle.adm_unet import seeded_state  # noqa: E402
from oracle.gen_golden import save  # noqa: E402

assert azula.__file__.startswith(REF), azula.__file__

from oracle.gen_golden_cards_cfg import STRIDE, card_inputs  # noqa: E402


def main():
    out = {}
    for name, card in load_cards(adm.__name__).items():
        den = adm.make_model(**card.config).eval()
        den.backbone.load_state_dict(seeded_state(den.backbone.state_dict(), seed=1234))
        x, ts, y = card_inputs(name, card.config)
        u = den.backbone(x, ts, y=y)
        q = den(x, torch.tensor([0.3, 0.8]), label=y)
        out[f"{name}_unet"] = u[..., ::STRIDE, ::STRIDE].contiguous()
        out[f"{name}_mean"] = q.mean[..., ::STRIDE, ::STRIDE].contiguous()
        print(name, tuple(u.shape), float(u.std()), sum(p.numel() for p in den.parameters()))
        del den
    save("adm_cards", **out)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
