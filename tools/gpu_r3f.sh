#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_state.py -q -x 2>&1 | tail -2
timeout 600 python tools/acc_bench.py 2>&1 | grep -E "fused running average|fused max/min|Welch 4096" | cut -c 1-140
