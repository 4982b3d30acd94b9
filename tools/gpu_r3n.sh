#!/bin/bash
# split second barrier in the two-engine kernel too: parity + timing
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernel1.py tests/test_gpu_wl_kernel.py tests/test_gpu_state.py -q -x 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_full_size.py -q -x -k "cfg4 or cfg3" 2>&1 | tail -2
timeout 300 python tools/tg8192.py
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "time N=(4096|8192)"
