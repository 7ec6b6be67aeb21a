#!/bin/sh
# Build libeccq_sm100a.so in-tree (sm_100a only).  Usage: sh build.sh [extra nvcc flags]
# Every translation unit is compiled on its own (in parallel, plain nvcc, no split compile: the output is reproducible),
# then linked.  __graft_entry__.build() drives the same commands and only recompiles units whose sources changed.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OBJ="$HERE/lib/obj"
mkdir -p "$OBJ"
UNITS="eccq_lib eccq_codecs eccq_cache_write eccq_cache_read eccq_paged_raw eccq_attention eccq_attention_tc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
pids=""
for u in $UNITS; do
  nvcc $FLAGS -I"$HERE/../include" -I"$HERE/csrc" "$@" -c "$HERE/csrc/$u.cu" -o "$OBJ/$u.o" &
  pids="$pids $!"
done
for p in $pids; do wait "$p"; done
objs=""
for u in $UNITS; do objs="$objs $OBJ/$u.o"; done
exec nvcc -gencode arch=compute_100a,code=sm_100a -shared $objs -o "$HERE/lib/libeccq_sm100a.so"
