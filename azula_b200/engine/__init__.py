r"""Execution engine behind the Azula-compatible surface: coefficient tables, the fused
graph-captured sampling loop and the native sm_100a backbones (ADM U-Net, in-repo U-Net, DiT / ViT)."""

from __future__ import annotations

import contextlib

from . import ops  # noqa: F401,E402  (registers the engine entry points with the ctypes loader)

_NATIVE = True


def native_enabled() -> bool:
    r"""Whether CUDA inputs under ``torch.no_grad()`` run the native launch plans (the default)."""
    return _NATIVE


@contextlib.contextmanager
def eager_torch():
    r"""Explicit opt-out for measurements and debugging: inside this context the samplers and backbones
    execute their plain torch definitions -- the reference's execution model (Python loop, ATen / cuDNN
    kernels, fp32) -- on whatever device the tensors live.  Nothing selects this path implicitly."""
    global _NATIVE
    previous, _NATIVE = _NATIVE, False
    try:
        yield
    finally:
        _NATIVE = previous


def set_precision(module, precision: str):
    r"""Selects the arithmetic of the native ADM backbone(s) inside ``module``: :py:`"bf16"` (default: bf16 activations
    and tensor-core operands, the fast path with a stated bf16 tolerance) or :py:`"tf32"` (the reference-numerics
    path: fp32 activations in HBM, ``tcgen05.mma.kind::tf32`` contractions, fp32 GroupNorm / SiLU / residuals -- what
    the reference's fp32 modules compute under PyTorch's default flags).  Returns ``module``."""
    if precision not in ("bf16", "tf32"):
        raise ValueError(f"unknown precision {precision!r} (expected 'bf16' or 'tf32')")
    from ..plugins.adm.unet import UNetModel

    found = False
    for m in module.modules():
        if isinstance(m, UNetModel):
            m.precision, found = precision, True
    if not found:
        raise ValueError("no ADM U-Net inside the module: the precision switch exists for plugins.adm backbones")
    return module
