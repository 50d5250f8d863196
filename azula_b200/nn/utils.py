r"""Module helpers used by the denoisers (interface of ``azula/nn/utils.py:24-42,172-188``)."""

from __future__ import annotations

__all__ = ["get_module_dtype", "get_module_device", "skip_init"]

import itertools
import torch


def _first(module: torch.nn.Module, attr: str, only_float: bool):
    for tensor in itertools.chain(module.parameters(), module.buffers()):
        if not only_float or tensor.is_floating_point():
            return getattr(tensor, attr)
    return None


def get_module_dtype(module: torch.nn.Module) -> torch.dtype | None:
    r"""Returns the first floating-point dtype among parameters, then buffers (else ``None``).

    This dtype decides the cast of the backbone input (reference ``azula/denoise.py:314``).
    """
    for group in (module.parameters(), module.buffers()):
        for tensor in group:
            if tensor.is_floating_point():
                return tensor.dtype
    return None


def get_module_device(module: torch.nn.Module) -> torch.device | None:
    r"""Returns the device of the first parameter or buffer (else ``None``)."""
    return _first(module, "device", only_float=False)


class skip_init(torch.overrides.TorchFunctionMode):
    r"""Context in which ``torch.nn.init.*`` calls leave their tensor untouched.

    Used by ``load_model`` so that weights about to be overwritten are not initialised twice
    (reference ``azula/nn/utils.py:172-188``).
    """

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if getattr(func, "__module__", None) == "torch.nn.init":
            return kwargs["tensor"] if "tensor" in kwargs else args[0]
        return func(*args, **kwargs)
