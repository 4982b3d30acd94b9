#!/bin/bash
# Build libtdsa.so in-tree for sm_100a. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${TDSA_OUT:-$(dirname "$HERE")/libtdsa.so}"
BUILD="${TDSA_BUILD_DIR:-$HERE/build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@")
mkdir -p "$BUILD"
pids=()
for f in tdsa_fft_f32 tdsa_fft_f64 tdsa_wl_f32 tdsa_wl_f64 tdsa_api; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$BUILD/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT" "$BUILD/tdsa_fft_f32.o" "$BUILD/tdsa_fft_f64.o" "$BUILD/tdsa_wl_f32.o" "$BUILD/tdsa_wl_f64.o" "$BUILD/tdsa_api.o"
echo "built $OUT"
