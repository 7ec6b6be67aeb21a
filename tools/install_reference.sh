#!/bin/sh
# Install the UNMODIFIED reference into baseline/_ref/ (git-ignored, but it travels to the GPU box with gpurun) so that its own
# Triton kernels can be timed on the same B200 (tools/ref_triton_bench.py, bench.py's `reference_triton` record).
# The reference ships no setup.py / pyproject.toml, so `pip install /root/reference` has nothing to build: the three
# packages are copied as they are.  Nothing under baseline/_ref is imported by the product or by the parity tests.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
rm -rf "$DST"
mkdir -p "$DST"
for pkg in ecc_codecs kv_cache evaluation; do cp -r "$SRC/$pkg" "$DST/$pkg"; done
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "installed reference packages into $DST"
