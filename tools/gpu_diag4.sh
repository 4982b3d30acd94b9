#!/bin/bash
mkdir -p gpurun_out
echo "== WL (default)"; timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
echo "== old kernel"; TDSA_WL=0 timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-110
timeout 600 python -m pytest tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -5
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing
timeout 300 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
python tools/phase_timing.py gpurun_out/timing_wl_f64_g296.bin gpurun_out/timing_wl_f32_g296.bin
rm -f gpurun_out/timing_*_g4.bin
