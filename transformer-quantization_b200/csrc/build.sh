#!/usr/bin/env bash
# Builds libtq_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
#   transformer-quantization_b200/csrc/build.sh [extra nvcc flags]
# Translation units are compiled in parallel (objects under lib/obj/, rebuilt only when a source or
# header is newer) and linked into one shared library.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
OBJ="$OUT/obj"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
UNITS=(tq_abi tq_qdq tq_minmax tq_mse tq_linear tq_fused tq_qat)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false
       -Xcompiler -fPIC -Xcompiler -O2 -Xptxas -v "$@")
HDRS=("$HERE"/tq_common.cuh "$HERE"/../../include/tq_b200.h "$HERE"/build.sh)
pids=()
for u in "${UNITS[@]}"; do
    src="$HERE/$u.cu"; obj="$OBJ/$u.o"
    [ -f "$src" ] || continue
    stale=0
    [ -f "$obj" ] || stale=1
    for d in "$src" "${HDRS[@]}"; do [ "$d" -nt "$obj" ] && stale=1; done
    [ -n "$*" ] && stale=1
    if [ "$stale" = 1 ]; then
        ( "$NVCC" "${FLAGS[@]}" -c -o "$obj" "$src" 2> "$OBJ/$u.log" || { cat "$OBJ/$u.log"; rm -f "$obj"; exit 1; } ) &
        pids+=($!)
    fi
done
fail=0
for p in "${pids[@]:-}"; do [ -n "$p" ] && { wait "$p" || fail=1; }; done
[ "$fail" = 0 ] || { echo "build failed"; exit 1; }
OBJS=()
: > "$OUT/ptxas.log"
for u in "${UNITS[@]}"; do
    [ -f "$OBJ/$u.o" ] && OBJS+=("$OBJ/$u.o")
    [ -f "$OBJ/$u.log" ] && cat "$OBJ/$u.log" >> "$OUT/ptxas.log"
done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libtq_b200.so" "${OBJS[@]}"
grep -E "error|warning" "$OUT/ptxas.log" | grep -v "ptxas info" || true
echo "built $OUT/libtq_b200.so"
