#!/bin/bash
mkdir -p gpurun_out
echo "== WL1 + dynamic old kernels"; timeout -s KILL 300 python tests/dev/gpu_dev.py 2>&1 | grep -E "^time|FAILED|Error|error" | cut -c 1-120
echo "== TDSA_WL=0 (old kernel, dynamic)"; TDSA_WL=0 timeout -s KILL 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time|FAILED|Error|error" | cut -c 1-120
echo "== TDSA_WL=0 TDSA_DYNAMIC=0 (old kernel, static)"; TDSA_WL=0 TDSA_DYNAMIC=0 timeout -s KILL 300 python tests/dev/gpu_dev.py 2>&1 | grep -E "^time|FAILED|Error|error" | cut -c 1-120
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
