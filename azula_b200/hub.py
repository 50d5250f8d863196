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
import sys
import torch

_HUB = os.path.expanduser("~/.cache/azula/hub")


def get_hub_dir() -> str:
    return _HUB


def set_hub_dir(cache_dir: str) -> None:
    global _HUB
    _HUB = os.path.abspath(os.path.expanduser(cache_dir))


def cache_path(url: str) -> str:
    r"""The cache file of a URL: the URL with every run of non ``[a-zA-Z0-9_]`` replaced by a dot."""
    return os.path.join(get_hub_dir(), re.sub(r"[^a-zA-Z0-9_]+", ".", url))


def download(url: str, filename: str | None = None, hash_prefix: str | None = None, quiet: bool = False) -> str:
    r"""Returns the local path of a (cached) file, fetching it with :mod:`torch.hub` if absent;
    verifies an ``"alg:prefix"`` hash when given."""
    path = cache_path(url) if filename is None else os.path.abspath(os.path.expanduser(filename))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if not os.path.exists(path):
        if not quiet:
            print(f"Downloading {url} to {path}", file=sys.stderr)
        torch.hub.download_url_to_file(url, path, progress=not quiet)
    elif not quiet:
        print(f"Loading from {path}", file=sys.stderr)
    if hash_prefix is not None:
        alg, prefix = hash_prefix.split(":")
        digest = hashlib.new(alg)
        with open(path, "rb") as f:
            for block in iter(lambda: f.read(1 << 20), b""):
                digest.update(block)
        if not digest.hexdigest().startswith(prefix):
            raise AssertionError(f"hash of {path} ({alg}:{digest.hexdigest()}) does not start with {alg}:{prefix}")
    return path
