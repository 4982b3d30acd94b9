#!/bin/bash
# N = 8192 float64 rows on the two-engine kernel: parity + timing
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernel1.py tests/test_gpu_state.py tests/test_gpu_wl_kernel.py -q -x 2>&1 | tail -3
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "N=  8192|time N=8192"
