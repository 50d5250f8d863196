r"""Plugin helpers (interface of ``azula/plugins/utils.py:29-60``)."""

from __future__ import annotations

__all__ = ["load_cards"]

import os
import sys
import torch
import yaml

from types import ModuleType, SimpleNamespace


def _dtype(name: str | None) -> torch.dtype | None:
    if name is None:
        return None
    dtype = getattr(torch, name, None)
    if not isinstance(dtype, torch.dtype):
        raise ValueError(f"Unknown data type '{name}'.")
    return dtype


def load_cards(plugin: ModuleType | str) -> dict[str, SimpleNamespace]:
    r"""Returns the ``name -> card`` mapping of a plugin's ``cards.yaml`` (url, hash, config
    and optional ``dtype_map`` with torch dtypes resolved)."""
    module = sys.modules[plugin] if isinstance(plugin, str) else plugin
    path = os.path.join(os.path.dirname(module.__file__), "cards.yaml")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{module} is not a plugin (no cards.yaml)")
    with open(path) as f:
        cards = yaml.safe_load(f)
    out = {}
    for name, card in cards.items():
        if "dtype_map" in card:
            card["dtype_map"] = {k: _dtype(v) for k, v in card["dtype_map"].items()}
        out[name] = SimpleNamespace(**card)
    return out
