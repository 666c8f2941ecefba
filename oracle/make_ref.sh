#!/bin/sh
# Build oracle/_ref: the UNMODIFIED reference package, copied from the read-only checkout so that it
# travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, NOT gpurun-ignored).
#
# `pip install --no-index --no-build-isolation --target ... /root/reference` cannot do this here: the
# project's build backend (uv_build, pyproject.toml:1-3) has no wheel in /opt/wheelhouse and there is no
# network.  The reference is pure Python, so an install is exactly this copy of
# src/configurable_spectrograms.  Its third-party imports that are absent on the box (cdflib, matplotlib)
# are served by oracle/stubs.py at run time.  Nothing is edited; nothing is committed.
set -eu
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/src/configurable_spectrograms" ]; then
  echo "make_ref.sh: $SRC/src/configurable_spectrograms not found (no reference checkout on this machine)" >&2
  exit 3
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src/configurable_spectrograms" "$DST/configurable_spectrograms"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "no-git" ) > "$DST/SOURCE_REV"
( cd "$DST" && find configurable_spectrograms -name '*.py' | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "oracle/_ref: $(find "$DST/configurable_spectrograms" -name '*.py' | wc -l) files copied from $SRC"
