#!/bin/bash
# Build libprd_sm100.so in-tree (sm_100a only).  Usage: csrc/build.sh [--clean] [extra nvcc flags]
#
# Objects are reused only when the CONTENT hash of (nvcc flags, the .cu file, every header) matches the hash stored next
# to the object -- never by mtime, so an object that travelled with a snapshot cannot be linked against newer sources.
# The library embeds the hash of all sources (prd_source_hash()); _lib.load() refuses a library whose hash differs from
# the sources beside it.
set -euo pipefail
export LC_ALL=C
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libprd_sm100.so"
INC="$HERE/../../include/prd_denoiser.h"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [[ "${1:-}" == "--clean" ]]; then shift; rm -rf "$HERE/_obj" "$OUT"; fi
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v "$@")
mkdir -p "$HERE/_obj"
HDR_HASH="$(cat "$HERE"/*.cuh "$HERE"/*.h "$INC" | sha256sum | cut -c1-32)"
SRC_HASH="$(cat "$HERE"/*.cu "$HERE"/*.cuh "$HERE"/*.h "$INC" | sha256sum | cut -c1-16)"
pids=()
compile() {  # $1 = source, $2 = object
  local want; want="$( (echo "${FLAGS[*]}" "$HDR_HASH"; cat "$1") | sha256sum | cut -c1-32)"
  if [[ ! -f "$2" || ! -f "$2.hash" || "$(cat "$2.hash")" != "$want" ]]; then
    rm -f "$2" "$2.hash"
    ( "$NVCC" "${FLAGS[@]}" -c "$1" -o "$2" > "$2.log" 2>&1 && echo "$want" > "$2.hash" ) &
    pids+=($!)
  fi
}
for f in "$HERE"/*.cu; do compile "$f" "$HERE/_obj/$(basename "${f%.cu}").o"; done
# drop objects whose source is gone
for o in "$HERE"/_obj/*.o; do
  b="$(basename "${o%.o}")"
  [[ "$b" == "prd_buildinfo" || -f "$HERE/$b.cu" ]] || rm -f "$o" "$o.hash" "$o.log"
done
printf 'extern "C" const char* prd_source_hash(void) { return "%s"; }\n' "$SRC_HASH" > "$HERE/_obj/prd_buildinfo.cu"
compile "$HERE/_obj/prd_buildinfo.cu" "$HERE/_obj/prd_buildinfo.o"
fail=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || fail=1; }; done
if [[ $fail -ne 0 ]]; then cat "$HERE"/_obj/*.log | grep -iE "error|fatal" -A3 | head -80; exit 1; fi
"$NVCC" -shared -o "$OUT" "$HERE"/_obj/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT (sources $SRC_HASH)"
