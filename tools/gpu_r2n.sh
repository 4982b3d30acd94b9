#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096 f64|^time N=(4096)|FAILED|Error" | cut -c 1-110
for lib in variants/libtdsa_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096 f64|^time N=(4096)|FAILED|Error" | cut -c 1-110
done
