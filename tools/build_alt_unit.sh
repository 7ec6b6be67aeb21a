#!/bin/sh
# A/B helper: rebuild ONE translation unit with extra -D flags and link lib/<name>.so next to the main library (selected at run
# time with ECCQ_LIB=<path>).  Usage: sh tools/build_alt_unit.sh <unit, e.g. eccq_cache_write> <name, e.g. libeccq_w8> [nvcc flags...]
set -e
P="$(cd "$(dirname "$0")/.." && pwd)/quantized-kv-cache-ecc-protection_b200"
UNIT=$1; NAME=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$P/../include" -I"$P/csrc" "$@" -c "$P/csrc/$UNIT.cu" -o "$P/lib/obj/${NAME}_$UNIT.o"
objs=""
for u in eccq_lib eccq_codecs eccq_cache_write eccq_cache_read eccq_paged_raw eccq_attention eccq_attention_tc; do
  if [ "$u" = "$UNIT" ]; then objs="$objs $P/lib/obj/${NAME}_$UNIT.o"; else objs="$objs $P/lib/obj/$u.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared $objs -o "$P/lib/$NAME.so"
echo "built $P/lib/$NAME.so ($UNIT $*)"
