r"""Imports the UNMODIFIED reference package (``azula``) from ``baseline/_ref`` -- measurement infrastructure only.

``scripts/fetch_ref.sh`` copies ``/root/reference/azula`` there (git-ignored, shipped to the GPU box with the
snapshot).  Only ``bench.py`` (reference arm, ``cpu_baseline``, ``eager_gpu``) and ``scripts/`` use this module;
nothing under ``azula_b200/`` imports it or the reference.
"""

from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "azula", "__init__.py"))


def load():
    r"""Returns the reference's top-level module.  ``gdown`` (Google-Drive downloads, ``azula/hub.py:9,78-79``) is
    not installed in this image and is never reached without a network: an empty stub lets ``azula.hub`` import."""
    if not available():
        raise ImportError(f"{REF_DIR}/azula is missing: run scripts/fetch_ref.sh in the build container")
    if "gdown" not in sys.modules:
        try:
            import gdown  # noqa: F401
        except ImportError:
            sys.modules["gdown"] = types.ModuleType("gdown")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    import azula

    assert os.path.abspath(azula.__file__).startswith(REF_DIR), f"imported azula from {azula.__file__}, not {REF_DIR}"
    return azula
