"""Builds libazb.so (sm_100a) in-tree with nvcc.  ``python -m azula_b200.csrc.build [--force]``."""

from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libazb.so")
OBJ = os.path.join(HERE, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


# extra compile flags for diagnostics builds, e.g. AZB_NVCC_EXTRA=-DAZB_TIMELINE (per-tile stamps for scripts/conv_timeline.py)
NVCC_FLAGS += os.environ.get("AZB_NVCC_EXTRA", "").split()


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libazb cannot be built")
    return exe


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every .cu of csrc/ for sm_100a and links azula_b200/libazb.so."""
    headers = glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    os.makedirs(OBJ, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
