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
