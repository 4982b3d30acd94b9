#!/bin/bash
# float32 one-engine kernel at three CTAs per SM (80 registers, one stage) against the default
for lib in "" variants/libtdsa_f32c3.so; do
  if [ -n "$lib" ]; then export TDSA_LIB=$PWD/$lib; fi
  echo "== ${lib:-default}"
  timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "N=  4096 f32|time N=4096 B=8192 f32"
done
