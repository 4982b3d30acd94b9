#!/bin/bash
# round 2: cfg3 through the head + warp-local-tail path: parity (few segments, full size, all three paths), timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -k "cfg3 or welch" -x 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_state.py -q -k welch 2>&1 | tail -3
timeout 600 python tools/configs_bench.py 2>gpurun_out/configs_p.err | tee gpurun_out/configs_p.jsonl | cut -c 1-330
tail -3 gpurun_out/configs_p.err
