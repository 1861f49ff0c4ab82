#!/usr/bin/env bash
# Builds libtq_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
#   transformer-quantization_b200/csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=("$HERE"/tq_abi.cu "$HERE"/tq_qdq.cu "$HERE"/tq_minmax.cu "$HERE"/tq_mse.cu "$HERE"/tq_linear.cu "$HERE"/tq_fused.cu)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false
       -Xcompiler -fPIC -Xcompiler -O2 -shared -Xptxas -v "$@")
SRC_EXIST=()
for s in "${SRCS[@]}"; do [ -f "$s" ] && SRC_EXIST+=("$s"); done
"$NVCC" "${FLAGS[@]}" -o "$OUT/libtq_b200.so" "${SRC_EXIST[@]}" 2> "$OUT/ptxas.log" || { cat "$OUT/ptxas.log"; exit 1; }
grep -E "error|warning" "$OUT/ptxas.log" | grep -v "ptxas info" || true
echo "built $OUT/libtq_b200.so"
