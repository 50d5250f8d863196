r"""Weight cache (interface of ``azula/hub.py``): same directory and file naming as the
reference so an existing ``~/.cache/azula/hub`` is reused as is.

Network download is the reference's job (``azula/hub.py:40-125``) and out of scope for the
engine: files must already be in the cache (or be fetched with ``torch.hub`` when a network
exists).
"""

from __future__ import annotations

__all__ = ["get_hub_dir", "set_hub_dir", "download"]

import hashlib
import os
import re
import shutil
import sys
import tarfile
import tempfile
import torch
import zipfile

_HUB = os.path.expanduser("~/.cache/azula/hub")


def get_hub_dir() -> str:
    return _HUB


def set_hub_dir(cache_dir: str) -> None:
    global _HUB
    _HUB = os.path.abspath(os.path.expanduser(cache_dir))


def cache_path(url: str) -> str:
    r"""The cache file of a URL: the URL with every run of non ``[a-zA-Z0-9_]`` replaced by a dot."""
    return os.path.join(get_hub_dir(), re.sub(r"[^a-zA-Z0-9_]+", ".", url))


def download(url: str, filename: str | None = None, hash_prefix: str | None = None, extract: bool = False,
             quiet: bool = False) -> str:
    r"""Returns the local path of a (cached) file, fetching it with :mod:`torch.hub` if absent;
    verifies an ``"alg:prefix"`` hash when given.  With :py:`extract=True` the file is a tar or zip
    archive and the directory ``<file>+x`` holding its contents is returned (``azula/hub.py:40-125``)."""
    path = cache_path(url) if filename is None else os.path.abspath(os.path.expanduser(filename))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    fetched = False
    if not os.path.exists(path):
        if not quiet:
            print(f"Downloading {url} to {path}", file=sys.stderr)
        if "drive.google" in url:
            # Google Drive answers plain HTTP clients with an HTML interstitial: the reference goes through gdown
            # (azula/hub.py:78-79); without it, fail loudly instead of caching that page as a checkpoint
            try:
                import gdown
            except ImportError as e:
                raise RuntimeError(
                    f"{url} is a Google Drive link and needs the `gdown` package (pip install gdown), "
                    f"or place the file at {path} by hand"
                ) from e
            gdown.download(url, path, quiet=quiet)
        else:
            torch.hub.download_url_to_file(url, path, progress=not quiet)
        fetched = True
    elif not quiet:
        print(f"Loading from {path}", file=sys.stderr)
    if hash_prefix is not None:
        alg, prefix = hash_prefix.split(":")
        digest = hashlib.new(alg)
        with open(path, "rb") as f:
            for block in iter(lambda: f.read(1 << 20), b""):
                digest.update(block)
        if not digest.hexdigest().startswith(prefix):
            if fetched:  # never leave a bad download in the cache, where the next call would trust it
                os.remove(path)
            raise AssertionError(f"hash of {path} ({alg}:{digest.hexdigest()}) does not start with {alg}:{prefix}")
    if extract:
        xd = f"{path}+x"
        if os.path.exists(xd):
            return xd
        if not quiet:
            print(f"Extracting to {xd}", file=sys.stderr)
        with tempfile.TemporaryDirectory() as td:
            if tarfile.is_tarfile(path):
                with tarfile.open(path, "r") as f:
                    f.extractall(td, filter="data")  # no absolute paths, links out of the tree or device nodes
            elif zipfile.is_zipfile(path):
                with zipfile.ZipFile(path, "r") as f:
                    f.extractall(td)
            else:
                raise ValueError("Unknown archive format.")
            shutil.move(td, xd)
            os.makedirs(td, exist_ok=True)  # TemporaryDirectory cleans up the (now empty) original path
        return xd
    return path
