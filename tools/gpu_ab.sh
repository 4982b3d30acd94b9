#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096|1024)|FAILED|Error"
for lib in variants/libtdsa_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096|1024)|FAILED|Error"
done
timeout 300 python tools/e2e_chunks.py
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -3
