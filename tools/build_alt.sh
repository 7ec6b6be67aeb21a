#!/bin/sh
# A/B helper: build lib/libeccq_alt.so with a different eccq_attention_tc.cu (a git revision or extra -D flags); select it
# at run time with ECCQ_LIB=$PWD/quantized-kv-cache-ecc-protection_b200/lib/libeccq_alt.so.
# Usage: sh tools/build_alt.sh [git-rev | -] [extra nvcc flags...]
set -e
P="$(cd "$(dirname "$0")/.." && pwd)/quantized-kv-cache-ecc-protection_b200"
REV=${1:--}; shift || true
SRC="$P/csrc/eccq_attention_tc.cu"
if [ "$REV" != "-" ]; then git -C "$P/.." show "$REV:quantized-kv-cache-ecc-protection_b200/csrc/eccq_attention_tc.cu" > "$P/csrc/_alt_tc.cu"; SRC="$P/csrc/_alt_tc.cu"; fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$P/../include" -I"$P/csrc" "$@" -c "$SRC" -o "$P/lib/obj/${ALT_NAME:-alt}_tc.o"
rm -f "$P/csrc/_alt_tc.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -shared "$P/lib/obj/eccq_lib.o" "$P/lib/obj/eccq_codecs.o" "$P/lib/obj/eccq_cache_write.o" "$P/lib/obj/eccq_cache_read.o" "$P/lib/obj/eccq_paged_raw.o" \
  "$P/lib/obj/eccq_attention.o" "$P/lib/obj/${ALT_NAME:-alt}_tc.o" -o "$P/lib/${ALT_NAME:-libeccq_alt}.so"
echo "built $P/lib/${ALT_NAME:-libeccq_alt}.so from $REV $*"
