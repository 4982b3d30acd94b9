#!/bin/bash
# diagnostics: pipe microbenchmarks + the fused kernel at 1 CTA/SM vs 2 CTAs/SM
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./tools/bin/ubench 2>&1 | tee gpurun_out/ubench.log
echo "== default (2 CTAs/SM)"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096|1024)|FAILED|Error"
echo "== 1 CTA/SM"; TDSA_DEBUG_EXTRA_SMEM=100000 timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096|1024)|FAILED|Error"
