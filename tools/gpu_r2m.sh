#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
echo "== default"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-110
for lib in variants/libtdsa_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-110
  TDSA_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_wl_kernel.py tests/test_gpu_kernel1.py -m gpu -q -x 2>&1 | tail -2
done
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_reference_app.py -m gpu -q -x -k "sweep or reference or set_source" 2>&1 | tail -4
