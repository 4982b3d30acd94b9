#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_dev.py quick > gpurun_out/dev.log 2>&1; echo "dev rc=$?"; tail -14 gpurun_out/dev.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
