#!/bin/bash
# round 2, run D: whole GPU test-suite (full-size configs included), bench (both arms), configs 3/4/5 numbers
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; cut -c 1-6000 gpurun_out/bench_d.json; tail -5 gpurun_out/bench_d.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 | cut -c 1-600
timeout 900 python tools/configs_bench.py > gpurun_out/configs_d.jsonl 2> gpurun_out/configs_d.err; echo "configs rc=$?"; cut -c 1-700 gpurun_out/configs_d.jsonl; tail -3 gpurun_out/configs_d.err
