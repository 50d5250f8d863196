r"""Module helpers used by the denoisers (interface of ``azula/nn/utils.py:24-42,172-188``)."""

from __future__ import annotations

__all__ = ["NativeCache", "checkpoint", "get_module_dtype", "get_module_device", "promote_dtype", "skip_init"]

import functools
import itertools
import torch
import torch.utils.checkpoint


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


def checkpoint(f, reentrant: bool = False):
    r"""Activation checkpointing of a function (interface of ``azula/nn/utils.py:109-161``).

    Training-side utility, outside the generation path: without autograd the function is simply
    called; with autograd the inputs are stored and the graph recomputed during the backward pass
    (:func:`torch.utils.checkpoint.checkpoint`, non-reentrant, so that implicit inputs such as module
    parameters receive gradients in both modes).
    """

    @functools.wraps(f)
    def g(*args, **kwargs):
        if not torch.is_grad_enabled():
            return f(*args, **kwargs)
        return torch.utils.checkpoint.checkpoint(f, *args, use_reentrant=False, **kwargs)

    return g


def promote_dtype(f, min_dtype: torch.dtype = torch.float32):
    r"""Runs a function of tensors in at least ``min_dtype`` and casts the result(s) back to the
    promoted input dtype (interface of ``azula/nn/utils.py:191-221``)."""

    @functools.wraps(f)
    def g(*args, **kwargs):
        dtype = functools.reduce(torch.promote_types, [a.dtype for a in args])
        outs = f(*(a.to(torch.promote_types(a.dtype, min_dtype)) for a in args), **kwargs)
        if torch.is_tensor(outs):
            return outs.to(dtype)
        return tuple(o.to(dtype) for o in outs)

    return g


class NativeCache:
    r"""Mixin of the backbones that own a native launch-plan cache (``self._native``): the cache is
    dropped when the module moves / changes dtype and is never copied or pickled with the module."""

    def _apply(self, fn, *args, **kwargs):
        self._native.clear()  # packed weights and plans belong to the old device / dtype
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_native"] = {}
        return state
