#!/bin/bash
for lib in "" variants/libtdsa_winreload.so; do
  if [ -n "$lib" ]; then export TDSA_LIB=$PWD/$lib; fi
  echo "== ${lib:-default}"
  timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "N=  4096 f64|time N=4096 B=8192 f64"
done
