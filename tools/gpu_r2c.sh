#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=|FAILED|Error"
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_wl_kernel.py tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -5
