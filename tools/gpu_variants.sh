#!/bin/bash
mkdir -p gpurun_out
echo "== WL (default)"; timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
for lib in variants/libtdsa_v*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
done
timeout -s KILL 600 python -m pytest tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -5
