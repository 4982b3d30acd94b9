#!/bin/bash
# round 2, run E: whole GPU test-suite, configs 3/4/5 numbers
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 900 python tools/configs_bench.py > gpurun_out/configs_e.jsonl 2> gpurun_out/configs_e.err; echo "configs rc=$?"; cut -c 1-900 gpurun_out/configs_e.jsonl; tail -3 gpurun_out/configs_e.err
