#!/bin/sh
# Build libeccq_sm100a.so in-tree (sm_100a only). Usage: sh build.sh [extra nvcc flags]
# One plain nvcc run.  nvcc's split compile can come out in a 4-16 % slower code generation for identical sources (DESIGN.md,
# "Build reproducibility"): `python -c "import __graft_entry__ as g; g.build(force=True)"` checks for it and recompiles.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/lib"
exec nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -split-compile=8 -Xcompiler -fPIC -shared \
  -I"$HERE/../include" -I"$HERE/csrc" "$@" "$HERE/csrc/eccq_lib.cu" -o "$HERE/lib/libeccq_sm100a.so"
