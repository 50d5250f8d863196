"""The C-ABI library loads without a GPU and exports every symbol include/azb.h declares."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from azula_b200.csrc.build import build

    return build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "azb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(azb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib_path):
    handle = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 6
    for name in names:
        assert hasattr(handle, name), name


def test_version_and_strerror(lib_path):
    from azula_b200 import _lib

    handle = _lib.lib()
    assert handle.azb_version() == _lib.ABI_VERSION
    assert b"NULL" in handle.azb_strerror(-1)
    assert handle.azb_strerror(0) == b"ok"


def test_python_signatures_cover_header(lib_path):
    from azula_b200 import _lib
    import azula_b200.engine  # noqa: F401  (registers engine entry points)

    _lib.lib()
    missing = [n for n in declared_symbols() if n not in _lib._SIGNATURES]
    assert not missing, missing


def test_argument_errors_without_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    from azula_b200 import _lib

    handle = _lib.lib()
    rc = handle.azb_step_f32(None, None, 0, 0, None, None, None, 0, 4, 1, None, None, 0, None, 0, 1024, 0, None)
    assert rc == -1
    with pytest.raises(_lib.AzbError):
        _lib.check(rc, "azb_step_f32")
