#!/bin/bash
# round 2, run B: TMEM probe, new kernels (accumulating epilogue, two-engine 8192), full GPU test-suite
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 60 tools/bin/tmem_test
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=|^time N=|FAILED|Error"
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_round2.py 2>&1 | tail -15
