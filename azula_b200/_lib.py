"""ctypes binding of libazb.so (the C ABI declared in include/azb.h).

There is no fallback: on a CUDA device every hot-path call goes through this library, and a
missing or stale library raises instead of silently running something else.
"""

from __future__ import annotations

import ctypes
import os
import torch

from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libazb.so")
if os.environ.get("AZB_LIBRARY"):  # A/B measurements of two builds on one box (scripts/build_ab.sh)
    LIB_PATH = os.environ["AZB_LIBRARY"]
ABI_VERSION = 4

F32, BF16, F16, I64 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.int64: I64}

_lib = None

ROW_COLS = 32  # floats per extended coefficient row (include/azb.h AZB_ROW_COLS)
R_P, R_Q, R_R, R_FLAGS, R_DRAW, R_W = 8, 9, 10, 12, 13, 16
MAX_SLOTS = 8


class AzbStep(ctypes.Structure):
    """Mirror of ``struct AzbStep`` (include/azb.h)."""

    _fields_ = [
        ("src", c_void_p * 2), ("dst", c_void_p * 2), ("f", c_void_p), ("f_neg", c_void_p), ("guidance", c_void_p),
        ("eps", c_void_p), ("x_in_next", c_void_p), ("hist", c_void_p), ("table", c_void_p), ("step_idx", c_void_p),
        ("philox_state", c_void_p),
        ("f_batch_stride", c_int64), ("n_per_sample", c_int64), ("batch", c_int64), ("hist_stride", c_int64),
        ("offset_host", c_int64), ("offset_inc", c_int64), ("rng_threads", c_int64), ("rng_elem_offset", c_int64),
        ("seed", c_uint64),
        ("f_dtype", c_int32), ("in_dtype", c_int32), ("row_floats", c_int32), ("x_in_copies", c_int32),
        ("noise_hint", c_int32), ("reserved_", c_int32),
    ]


_SIGNATURES = {
    "azb_version": (c_int, []),
    "azb_strerror": (c_char_p, [c_int]),
    "azb_rng_policy": (c_int, [c_int64, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    "azb_step_f32": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
         c_void_p, c_void_p, c_uint64, c_void_p, c_int64, c_int64, c_int64, c_void_p],
    ),
    "azb_step_ex_f32": (c_int, [ctypes.POINTER(AzbStep), c_void_p]),
    "azb_advance": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int32, c_void_p]),
    "azb_init_noise_f32": (c_int, [c_void_p, c_int64, c_float, c_float, c_uint64, c_int64, c_int64, c_int64, c_void_p]),
}


class AzbError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads libazb.so once; raises AzbError when it is absent (build with __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AzbError(
                f"{LIB_PATH} is missing: the sm_100a extension is not built "
                "(run `python -m azula_b200.csrc.build`); azula_b200 has no CPU/eager fallback for CUDA tensors"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:  # stale build
                raise AzbError(f"libazb.so lacks symbol {name}; rebuild it") from e
            fn.restype, fn.argtypes = res, args
        if handle.azb_version() != ABI_VERSION:
            raise AzbError(f"libazb.so ABI {handle.azb_version()} != expected {ABI_VERSION}; rebuild it")
        _lib = handle
    return _lib


def register(signatures: dict) -> None:
    """Lets engine modules add the signatures of further entry points before the first load."""
    _SIGNATURES.update(signatures)
    global _lib
    if _lib is not None:
        for name, (res, args) in signatures.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().azb_strerror(code).decode()
        raise AzbError(f"{what or 'libazb'} failed with code {code}: {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def rng_policy(numel: int) -> tuple[int, int]:
    """(T, offset_inc) of the randn launch ATen would use for numel elements on the current device."""
    T, inc = c_int64(0), c_int64(0)
    check(lib().azb_rng_policy(numel, ctypes.byref(T), ctypes.byref(inc)), "azb_rng_policy")
    return T.value, inc.value
