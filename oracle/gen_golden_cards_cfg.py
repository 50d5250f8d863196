"""Inputs of the per-card fixtures (shared by oracle/gen_golden_cards.py and the tests; test infrastructure)."""

import torch

SIZE, BATCH, STRIDE = 64, 2, 4


def card_inputs(name: str, config: dict):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(BATCH, 3, SIZE, SIZE, generator=g)
    ts = torch.tensor([37, 911])
    y = torch.tensor([3, 998]) if config.get("num_classes") else None
    return x, ts, y
