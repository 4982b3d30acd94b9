#!/bin/bash
mkdir -p gpurun_out
./tools/bin/ubench 2>&1 | tee gpurun_out/ubench.log | grep "^C "
timeout -s KILL 300 python tests/dev/gpu_dev.py 2>&1 | grep -E "^time|FAILED|Error|error" | cut -c 1-120
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
