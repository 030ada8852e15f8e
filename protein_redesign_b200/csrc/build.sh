#!/bin/bash
# Build libprd_sm100.so in-tree (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libprd_sm100.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v "$@")
mkdir -p "$HERE/_obj"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/_obj/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$HERE" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" -print -quit)" || "$HERE/../../include/prd_denoiser.h" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && { wait "$p" || fail=1; }; done
if [[ $fail -ne 0 ]]; then cat "$HERE"/_obj/*.log | grep -iE "error|fatal" -A3 | head -80; exit 1; fi
"$NVCC" -shared -o "$OUT" "$HERE"/_obj/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
