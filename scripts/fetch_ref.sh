#!/usr/bin/env bash
# Ships the UNMODIFIED reference package to the GPU box: copies /root/reference/azula (pure Python, v0.11.1) into
# baseline/_ref/azula.  baseline/_ref/ is git-ignored but NOT gpurun-ignored, so it travels with the snapshot like
# libazb.so does.  bench.py (--impl reference, cpu_baseline, eager_gpu) imports the reference from there; nothing
# under azula_b200/ ever does.  Run here (the build container); the GPU box has no /root/reference.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/azula" ]; then
    echo "fetch_ref: $SRC/azula not found (nothing copied)" >&2
    exit 0
fi
rm -rf "$DST/azula"
mkdir -p "$DST"
cp -r "$SRC/azula" "$DST/azula"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find azula -name '*.py' -o -name '*.yaml' | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "fetch_ref: copied $(find "$DST/azula" -type f | wc -l) files to $DST/azula"
