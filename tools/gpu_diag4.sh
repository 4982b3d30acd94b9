#!/bin/bash
mkdir -p gpurun_out
echo "== WL (default)"; timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^N=  4096|^time N=(4096)|FAILED|Error|error" | cut -c 1-200
for st in 400 800 1600; do echo "== WL skew $st"; TDSA_DEBUG_STAGGER=$st timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error|error" | cut -c 1-110; done
timeout -s KILL 600 python -m pytest tests/test_gpu_kernel1.py -m gpu -x -q 2>&1 | tail -5
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing
timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
python tools/phase_timing.py gpurun_out/timing_wl_f64_g148.bin gpurun_out/timing_wl_f32_g296.bin
rm -f gpurun_out/timing_*_g4.bin
