#!/usr/bin/env bash
# Install the UNMODIFIED reference (chzhang18/RAG, pure Python, no setup.py -> nothing to pip-install)
# at baseline/_ref/ so that it travels to the GPU box like the built .so files do.
#   baseline/_ref/ is git-ignored (the reference's sources never enter this repo's history)
#   and NOT gpurun-ignored.  Only src/ is needed: models/, automl/, approaches/, utils.py, utilstool/.
# Used by: tests/test_reference_network_gpu.py (drop-in proven on the real Network / BasicNetwork),
#          bench.py --impl reference and bench.py's cpu_baseline / torch_cuda_reference legs.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${RAG_REFERENCE:-/root/reference}"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/src/models" ]; then
  echo "install_ref: $SRC/src not found (the reference only exists in the build container)" >&2
  exit 3
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$DST/src"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
# record what was installed: file list + sha256, so a test can prove the copy is unmodified
( cd "$DST/src" && find . -type f -name '*.py' | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "installed $(wc -l < "$DST/MANIFEST.sha256") reference files into $DST"
