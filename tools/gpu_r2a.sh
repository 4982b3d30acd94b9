#!/bin/bash
# round 2, run A: integer-pipe conversions (variants) + SHFL/LDS microbenchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== default"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=|^time N=(4096|1024)|FAILED|Error"
for lib in variants/libtdsa_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=|^time N=(4096|1024)|FAILED|Error"
done
timeout 120 tools/bin/ubench 2>&1 | grep -E "^G "
TDSA_LIB=$PWD/variants/libtdsa_both.so timeout 600 python -m pytest tests/test_gpu_kernel1.py tests/test_gpu_wl_kernel.py -m gpu -x -q 2>&1 | tail -4
