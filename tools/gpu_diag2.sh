#!/bin/bash
for st in 0 500 1000 1500 2500 4000; do
echo "== stagger $st"; TDSA_DEBUG_STAGGER=$st timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
done
