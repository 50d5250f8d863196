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


def test_descriptor_structs_match_the_header(tmp_path):
    """The ctypes mirrors of the plain-C descriptors (AzbConv, AzbConvChoice, AzbStep) have the size and the field
    offsets the C compiler gives the structs of include/azb.h: a field added on one side only would shift every later
    argument silently."""
    import ctypes
    import shutil
    import subprocess

    from azula_b200 import _lib as L
    from azula_b200.engine import ops

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"AzbConv": ops.AzbConv, "AzbConvChoice": ops.AzbConvChoice, "AzbStep": L.AzbStep}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "azb.h")}"', "int main(void) {"]
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, *_ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == ctypes.sizeof(cls), (name, out[name], ctypes.sizeof(cls))
        for field, *_ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)
