#!/bin/bash
# A/B: run the quick timing for the default library and every variants/libtdsa_*.so
mkdir -p gpurun_out
echo "== default"; timeout 300 python tools/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=4096|^time N=1024|FAILED|Error"
for lib in variants/libtdsa_*.so; do
  echo "== $lib"; TDSA_LIB=$PWD/$lib timeout 300 python tools/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=4096|^time N=1024|FAILED|Error"
done
