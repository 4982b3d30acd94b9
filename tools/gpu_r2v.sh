#!/bin/bash
# float64 headline kernel: integer-pipe widening of the staged samples (variant) against the default library
for lib in "" variants/libtdsa_widen.so; do
  if [ -n "$lib" ]; then export TDSA_LIB=$PWD/$lib; fi
  echo "== ${lib:-default}"
  timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "N=  4096|time N=4096|time N=1024"
done
TDSA_LIB=$PWD/variants/libtdsa_widen.so timeout 600 python -m pytest tests/test_gpu_kernel1.py -q -x 2>&1 | tail -2
