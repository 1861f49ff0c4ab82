#!/usr/bin/env bash
# Installs the UNMODIFIED reference into the git-ignored baseline/_ref/ (it travels to the GPU box with the
# gpurun snapshot; it never enters the history).  Run in the build container only:
#
#     tools/install_reference.sh [/root/reference]
#
# The reference is a script tree without packaging metadata (no setup.py / pyproject.toml), so
#     pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# stops with "Neither 'setup.py' nor 'pyproject.toml' found".  As the contract allows, the install is done from
# a copy under /tmp that gets a five-line setup.py naming the reference's own packages (quantization,
# quantization.adaround, utils, models) and its main.py -- the sources themselves are byte-identical
# (checked below with diff -r).  --no-deps: its pinned dependencies (torch 1.4, transformers 4.1, ...) are
# not in the wheelhouse; the image's torch / transformers are used with tests/hf41_shim.py.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
TMP="$(mktemp -d /tmp/tq_ref_XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"
cat > "$TMP/src/setup.py" <<'EOF'
from setuptools import setup
setup(name='transformer_quantization_reference', version='0.0',
      packages=['quantization', 'quantization.adaround', 'utils', 'models'], py_modules=['main'])
EOF
rm -rf "$DST"
mkdir -p "$DST"
python -m pip install --no-index --no-build-isolation --no-deps --no-compile --find-links /opt/wheelhouse \
    --target "$DST" "$TMP/src" > "$TMP/pip.log" 2>&1 || { cat "$TMP/pip.log"; exit 1; }
for d in quantization utils models; do
    diff -r -x __pycache__ "$REF/$d" "$DST/$d" > /dev/null || { echo "installed $d differs from the reference"; exit 1; }
done
cmp "$REF/main.py" "$DST/main.py"
echo "reference installed unmodified under $DST"
