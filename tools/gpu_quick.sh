#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/dev/gpu_dev.py > gpurun_out/dev.log 2>&1; echo "dev rc=$?"; grep -E "^N=|^time|FAILED|Error" gpurun_out/dev.log | cut -c 1-150
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TDSA_LOGR_F32=3 TDSA_LOGR_F64=3 timeout 600 python -m pytest tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -3
