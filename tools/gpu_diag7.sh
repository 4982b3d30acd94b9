#!/bin/bash
mkdir -p gpurun_out
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing
timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
python tools/phase_timing.py gpurun_out/timing_pp_f64_g148.bin gpurun_out/timing_pp_f32_g148.bin
rm -f gpurun_out/timing_*_g2.bin gpurun_out/timing_*_g4.bin
