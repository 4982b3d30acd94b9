#!/bin/bash
# phase time stamps: needs variants/libtdsa_timing.so built with
#   TDSA_OUT=$PWD/variants/libtdsa_timing.so TDSA_BUILD_DIR=/tmp/build_timing bash topdogspectrumanalyser_b200/csrc/build.sh -DTDSA_DEBUG_TIMING
mkdir -p gpurun_out
export TDSA_LIB=$PWD/variants/libtdsa_timing.so TDSA_DEBUG_TIMING_OUT=$PWD/gpurun_out/timing
timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
TDSA_WL=0 timeout -s KILL 120 python tests/dev/gpu_dev.py quick 2>&1 | grep -E "^time N=(4096)|FAILED|Error" | cut -c 1-100
rm -f gpurun_out/timing_*_g4.bin gpurun_out/timing_*_g2.bin
python tools/phase_timing.py gpurun_out/timing_*.bin
